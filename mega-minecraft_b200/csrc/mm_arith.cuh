// Device arithmetic core of the chunk-generation path: integer hash + minstd engines, GLM-equivalent
// simplex 2-D/3-D and fbm, written as explicit fp32 operation sequences.
//
// What is computed follows /root/reference/src/util/rng.hpp:69-96,166-191 and the vendored GLM
// simplex (/root/reference/external/include/glm/gtc/noise.inl:591-720). HOW it rounds follows the
// reference's own sm_100 build: the world is a chaotic function of these values (a 1-ulp change in
// a biome-noise argument moves heights by tenths of a block), so the placement of every FMA is
// part of the result. This translation unit is compiled with -fmad=false: a*b+c written with
// plain operators is two roundings, fmaf() is one, and nothing else is contracted.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// The simplex bodies are ~250 (2-D) / ~450 (3-D) instructions. Inlined at every call site the
// cave and fill kernels grow to > 500 KB of SASS and stall on instruction fetch (ncu:
// stalled_no_instruction 13.7 per issue, profiles/r01); as real functions each kernel holds one copy
// per skew variant and fits the instruction cache. Same operations either way: results are identical.
#ifndef MMG_NOISE_INLINE
#define MMG_NOISE_INLINE __noinline__
#endif

namespace mmg {

// ---------------------------------------------------------------- integer hash + minstd
// rng.hpp:69-78
__device__ __forceinline__ uint32_t hash_u32(uint32_t a)
{
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}

// thrust::minstd_rand: x <- 48271 x mod (2^31-1); seed s -> s mod m, 0 -> 1.
// uniform_real_distribution<float>(a,b): (x - 1) / (1 + float(max - min)) * (b - a) + a with
// min = 1, max = 2^31-2, i.e. divisor 2147483648.f (thrust/random/detail/uniform_real_distribution.inl).
struct Minstd
{
    uint32_t x;
    // reduction mod the Mersenne prime 2^31 - 1: v = hi * 2^31 + lo == hi + lo (mod m), one conditional subtract
    static __device__ __forceinline__ uint32_t mod_m(uint64_t v)
    {
        uint32_t r = (uint32_t)(v & 0x7fffffffu) + (uint32_t)(v >> 31);     // v < 2^47: r < 2^31 + 2^16
        return r >= 2147483647u ? r - 2147483647u : r;
    }
    __device__ __forceinline__ explicit Minstd(uint32_t seed)
    {
        x = mod_m(seed);
        if (x == 0) x = 1;
    }
    // resume from a state produced by an earlier Minstd (x is already in [1, m))
    static __device__ __forceinline__ Minstd from_state(uint32_t state)
    {
        Minstd r(1u);
        r.x = state;
        return r;
    }
    __device__ __forceinline__ uint32_t next()
    {
        x = mod_m((uint64_t)x * 48271u);
        return x;
    }
    __device__ __forceinline__ float u01() { return (float)(next() - 1u) / 2147483648.0f; }
    // thrust evaluates (x-1)/2^31 * (b-a) + a; for (-1,1): * 2 then + (-1). The mul by 2 is exact.
    __device__ __forceinline__ float u11() { return ((float)(next() - 1u) / 2147483648.0f) * 2.0f + -1.0f; }
};

// rng.hpp:80-96 (int arithmetic on purpose: negative coordinates wrap as in the reference)
__device__ __forceinline__ Minstd make_rng1(int x) { return Minstd(hash_u32((uint32_t)x)); }
__device__ __forceinline__ Minstd make_rng3(int x, int y, int z)
{
    uint32_t h = hash_u32((uint32_t)((1 << 31) | (x << 22) | y)) ^ hash_u32((uint32_t)z);
    return Minstd(h);
}
__device__ __forceinline__ Minstd make_rng4(int x, int y, int z, int w)
{
    uint32_t h = hash_u32((uint32_t)((1 << 31) | (x << 22) | (y << 11) | w)) ^ hash_u32((uint32_t)z);
    return Minstd(h);
}

// ---------------------------------------------------------------- GLM helpers
__device__ __forceinline__ float g_fract(float x) { return x - floorf(x); }                 // func_common.inl:185-190
__device__ __forceinline__ float g_clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } // :244
// smoothstep (func_common.inl:564-570): true division, then t*t*(3-2t); 3-2t is one FFMA on the GPU
__device__ __forceinline__ float g_smoothstep(float e0, float e1, float x)
{
    float t = g_clamp01((x - e0) / (e1 - e0));
    return (t * t) * fmaf(t, -2.0f, 3.0f);
}

// ---------------------------------------------------------------- simplex lattice tables
// GLM's simplex hashes every lattice corner with permute(x) = mod289((34x+1)x) chains and decodes a
// gradient from the final value p. All of that is a function of small integers: permute() only ever sees
// 0..579 and the normalised gradient depends on p in 0..289 alone. The reference recomputes both per
// corner in floating point (~45 instructions, 8 of them FRND on the quarter-rate XU pipe, per corner);
// here they are tabulated ONCE per library init by the reference's own fp32 formulas (k_init_noise_tables,
// same -fmad=false code as before, so the quirks of the float mod289 are kept bit for bit) and every CTA
// stages the 10 KB of tables in shared memory.
struct NoiseTab
{
    float4 grad3[290];            // normalised 3-D gradient (ax, ay, az) of p   (noise.inl:676-711)
    float4 grad2[290];            // 2-D: (a0, h, norm factor) of p              (noise.inl:622-637)
    unsigned short perm[584];     // 2 * permute(x), x = 0..579: a byte offset into this table (and 1/8 of one into grad*),
                                  // so that every step of a permute chain is one 3-input add + one load
};
constexpr int kNoiseSmemBytes = (int)sizeof(NoiseTab);
__device__ NoiseTab g_noiseTab;  // filled by k_init_noise_tables

extern __shared__ float4 mmg_dyn_smem[];
__device__ __forceinline__ const NoiseTab* noise_tab() { return reinterpret_cast<const NoiseTab*>(mmg_dyn_smem); }

// every kernel that evaluates simplex noise is launched with kNoiseSmemBytes of dynamic shared memory
// and calls this once (all threads of the CTA) before the first evaluation
__device__ __forceinline__ void noise_tab_stage()
{
    const float4* src = reinterpret_cast<const float4*>(&g_noiseTab);
    for (int i = threadIdx.x + blockDim.x * threadIdx.y; i < kNoiseSmemBytes / 16; i += blockDim.x * blockDim.y) mmg_dyn_smem[i] = src[i];
    __syncthreads();
}
// (Reading the tables where they lie - global memory through L1 - instead of staging them per CTA was measured in round 2:
// k_caves 79.2 ms against 70.9, k_fill_rock 34.2 against 28.6 per 128x128 region; profiles/r02_variants.txt.)

// permute(x) = mod289((34x+1)x) on small non-negative integers: every step is exact in fp32,
// so contraction cannot change it (34*580+1 and its product with 580 are < 2^24).
__device__ __forceinline__ float sx_mod289(float x) { return x - floorf(x * (1.0f / 289.0f)) * 289.0f; }
__device__ __forceinline__ float sx_permute(float x) { return sx_mod289(fmaf(x, 34.0f, 1.0f) * x); }

__global__ void k_init_noise_tables()
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 584) g_noiseTab.perm[i] = (unsigned short)(2 * (int)sx_permute((float)i));      // stored doubled: see NoiseTab
    if (i < 290)
    {
        const float p = (float)i;
        {
            // 3-D gradient of p (noise.inl:676-711)
            const float NZ = 0.142857142857f, NX = NZ * 2.0f, NY = NZ * 0.5f - 1.0f;
            const float j = p - 49.0f * floorf((p * NZ) * NZ);
            const float x_ = floorf(j * NZ);
            const float y_ = floorf(j - 7.0f * x_);
            const float gx = fmaf(x_, NX, NY), gy = fmaf(y_, NX, NY);
            const float h = (1.0f - fabsf(gx)) - fabsf(gy);
            const float sh = (h <= 0.0f) ? -1.0f : 0.0f;
            float ax = fmaf(fmaf(floorf(gx), 2.0f, 1.0f), sh, gx);
            float ay = fmaf(fmaf(floorf(gy), 2.0f, 1.0f), sh, gy);
            float az = h;
            const float nrm = fmaf(fmaf(az, az, fmaf(ax, ax, ay * ay)), -0.85373472095314f, 1.79284291400159f);
            g_noiseTab.grad3[i] = make_float4(ax * nrm, ay * nrm, az * nrm, 0.0f);
        }
        {
            // 2-D: x = 2*fract(p*C.w) - 1 ; h = |x| - 0.5 ; a0 = x - floor(x + 0.5) ; norm = 1.79 - 0.85*(a0*a0 + h*h)
            const float C3 = 0.024390243902439f;
            const float q = p * C3;
            const float gx = fmaf(q - floorf(q), 2.0f, -1.0f);
            const float h = fabsf(gx) - 0.5f;
            const float a0 = gx - floorf(gx + 0.5f);
            g_noiseTab.grad2[i] = make_float4(a0, h, fmaf(fmaf(h, h, a0 * a0), -0.85373472095314f, 1.79284291400159f), 0.0f);
        }
    }
}

// ---------------------------------------------------------------- simplex 2-D
// returns dot(m, g) BEFORE the final *130: callers that add to the result fuse that multiply
// (fma(raw, 130, c)); everything else uses 130*raw rounded.
// SKEW_X selects which product of the skew dot(v, C.yy) the reference build fused: most inlined
// copies compute fma(v.y, C1, v.x*C1) (false); the copies listed in DESIGN.md (direct simplex()
// calls hoisted in front of the biome loop) compute fma(v.x, C1, v.y*C1) (true). The value only
// feeds floor(), so the two differ only when v + s lands within an ulp of an integer.
template <bool SKEW_X = false>
__device__ MMG_NOISE_INLINE float simplex2_raw(float vx, float vy)
{
    const NoiseTab* T = noise_tab();
    const float C0 = 0.211324865405187f, C1 = 0.366025403784439f, C2 = -0.577350269189626f;
    // i = floor(v + dot(v, C.yy)); dot = fma(v.y, C1, v.x*C1)
    float s = SKEW_X ? fmaf(vx, C1, vy * C1) : fmaf(vy, C1, vx * C1);
    float ix = floorf(vx + s), iy = floorf(vy + s);
    // x0 = v - i + dot(i, C.xx); dot = fma(i.x, C0, i.y*C0)
    float t = fmaf(ix, C0, iy * C0);
    float x0x = (vx - ix) + t, x0y = (vy - iy) + t;
    const bool gt = x0x > x0y;
    float i1x = gt ? 1.0f : 0.0f, i1y = gt ? 0.0f : 1.0f;
    float x1x = (x0x + C0) - i1x, x1y = (x0y + C0) - i1y;
    float x2x = x0x + C2, x2y = x0y + C2;
    // mod(i, 289) = i - 289*floor(i/289) with a true fp32 division (GLM mod(vec2, float)). For integers
    // |i| < 2^22 the rounded quotient cannot reach the next integer (its error is < 2^-11, the distance is
    // >= 1/289), so floor() of it is the exact integer quotient and the result is i mod 289 in [0, 288]:
    // computed in integers there (an fp32 division is ~15 instructions), in floats elsewhere.
    int jx, jy;
    if (fmaxf(fabsf(ix), fabsf(iy)) < 4194304.0f)
    {
        jx = (int)ix % 289; jx += jx < 0 ? 289 : 0;
        jy = (int)iy % 289; jy += jy < 0 ? 289 : 0;
    }
    else
    {
        jx = (int)(ix - 289.0f * floorf(ix / 289.0f));
        jy = (int)(iy - 289.0f * floorf(iy / 289.0f));
    }
    // permute chains in byte offsets (perm[] holds doubled values): P2(off) = 2 * permute(off / 2)
    const char* PB = reinterpret_cast<const char*>(T->perm);
    const char* GB = reinterpret_cast<const char*>(T->grad2);
    const int jx2 = jx + jx, jy2 = jy + jy;
    const int e1x = gt ? 2 : 0;
#define MMG_P2(off) ((int)*reinterpret_cast<const unsigned short*>(PB + (off)))
    // the middle corner's first permute is one of the two the outer corners load anyway
    const int py0 = MMG_P2(jy2), py1 = MMG_P2(jy2 + 2);
    const float4 G0 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2(py0 + jx2));
    const float4 G1 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2((gt ? py0 : py1) + jx2 + e1x));
    const float4 G2 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2(py1 + jx2 + 2));
    float m0 = fmaxf(0.5f - fmaf(x0x, x0x, x0y * x0y), 0.0f);
    float m1 = fmaxf(0.5f - fmaf(x1x, x1x, x1y * x1y), 0.0f);
    float m2 = fmaxf(0.5f - fmaf(x2x, x2x, x2y * x2y), 0.0f);
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    // m *= 1.79284291400159 - 0.85373472095314 * (a0*a0 + h*h)   (tabulated: G.z)
    m0 = m0 * G0.z; m1 = m1 * G1.z; m2 = m2 * G2.z;
    // g = a0*x.x + h*x.y  -> fma(x.y, h, x.x*a0)
    float g0 = fmaf(x0y, G0.y, x0x * G0.x);
    float g1 = fmaf(x1y, G1.y, x1x * G1.x);
    float g2 = fmaf(x2y, G2.y, x2x * G2.x);
    // 130 * dot(m, g): mul on .y, fma .x, fma .z
    return fmaf(g2, m2, fmaf(g0, m0, g1 * m1));
}
template <bool SKEW_X = false>
__device__ __forceinline__ float simplex2(float vx, float vy) { return 130.0f * simplex2_raw<SKEW_X>(vx, vy); }

// ---------------------------------------------------------------- simplex 3-D
// returns the dot BEFORE the final *42 (same reason as simplex2_raw)
// SKEW_Y: which product of the skew dot(v, C.yyy) the reference build left unfused: copies inlined
// through fbm<> compute fma(v.z, C, fma(v.y, C, v.x*C)) (false); direct simplex(vec3) calls compute
// fma(v.z, C, fma(v.x, C, v.y*C)) (true). Only floor() sees the difference.
template <bool SKEW_Y = false>
__device__ MMG_NOISE_INLINE float simplex3_raw(float vx, float vy, float vz)
{
    const NoiseTab* T = noise_tab();
    const float C = 1.0f / 3.0f, D = 1.0f / 6.0f;
    float s = SKEW_Y ? fmaf(vz, C, fmaf(vx, C, vy * C)) : fmaf(vz, C, fmaf(vy, C, vx * C));
    float ix = floorf(vx + s), iy = floorf(vy + s), iz = floorf(vz + s);
    float t = fmaf(iz, D, fmaf(ix, D, iy * D));
    float x0x = (vx - ix) + t, x0y = (vy - iy) + t, x0z = (vz - iz) + t;
    // g = step(x0.yzx, x0) ; l = 1 - g ; i1 = min(g, l.zxy) ; i2 = max(g, l.zxy)   (all 0 / 1)
    const bool gx = !(x0x < x0y), gy = !(x0y < x0z), gz = !(x0z < x0x);
    const bool a1x = gx && !gz, a1y = gy && !gx, a1z = gz && !gy;
    const bool a2x = gx || !gz, a2y = gy || !gx, a2z = gz || !gy;
    float x1x = (x0x - (a1x ? 1.0f : 0.0f)) + D, x1y = (x0y - (a1y ? 1.0f : 0.0f)) + D, x1z = (x0z - (a1z ? 1.0f : 0.0f)) + D;
    float x2x = (x0x - (a2x ? 1.0f : 0.0f)) + C, x2y = (x0y - (a2y ? 1.0f : 0.0f)) + C, x2z = (x0z - (a2z ? 1.0f : 0.0f)) + C;
    float x3x = x0x - 0.5f, x3y = x0y - 0.5f, x3z = x0z - 0.5f;
    const int jx = (int)sx_mod289(ix), jy = (int)sx_mod289(iy), jz = (int)sx_mod289(iz);
    // p = permute(permute(permute(i.z + e.z) + i.y + e.y) + i.x + e.x) for the four corners
    const char* PB = reinterpret_cast<const char*>(T->perm);
    const char* GB = reinterpret_cast<const char*>(T->grad3);
    const int jx2 = jx + jx, jy2 = jy + jy, jz2 = jz + jz;
    // the two middle corners' first permute is one of the two the outer corners load anyway
    const int pz0 = MMG_P2(jz2), pz1 = MMG_P2(jz2 + 2);
    const int p0 = MMG_P2(MMG_P2(pz0 + jy2) + jx2);
    const int p1 = MMG_P2(MMG_P2((a1z ? pz1 : pz0) + jy2 + (a1y ? 2 : 0)) + jx2 + (a1x ? 2 : 0));
    const int p2 = MMG_P2(MMG_P2((a2z ? pz1 : pz0) + jy2 + (a2y ? 2 : 0)) + jx2 + (a2x ? 2 : 0));
    const int p3 = MMG_P2(MMG_P2(pz1 + jy2 + 2) + jx2 + 2);
    const float4 G0 = *reinterpret_cast<const float4*>(GB + 8 * p0), G1 = *reinterpret_cast<const float4*>(GB + 8 * p1);
    const float4 G2 = *reinterpret_cast<const float4*>(GB + 8 * p2), G3 = *reinterpret_cast<const float4*>(GB + 8 * p3);
    float m0 = fmaxf(0.6f - fmaf(x0z, x0z, fmaf(x0x, x0x, x0y * x0y)), 0.0f);
    float m1 = fmaxf(0.6f - fmaf(x1z, x1z, fmaf(x1x, x1x, x1y * x1y)), 0.0f);
    float m2 = fmaxf(0.6f - fmaf(x2z, x2z, fmaf(x2x, x2x, x2y * x2y)), 0.0f);
    float m3 = fmaxf(0.6f - fmaf(x3z, x3z, fmaf(x3x, x3x, x3y * x3y)), 0.0f);
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2; m3 = m3 * m3;
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2; m3 = m3 * m3;
    const float d0 = fmaf(x0z, G0.z, fmaf(x0x, G0.x, x0y * G0.y));
    const float d1 = fmaf(x1z, G1.z, fmaf(x1x, G1.x, x1y * G1.y));
    const float d2 = fmaf(x2z, G2.z, fmaf(x2x, G2.x, x2y * G2.y));
    const float d3 = fmaf(x3z, G3.z, fmaf(x3x, G3.x, x3y * G3.y));
    // 42 * dot(m*m, dots): (t.x + t.y) + (t.z + t.w) with fma on .y and .w
    float a = fmaf(m1, d1, m0 * d0);
    float b = fmaf(m3, d3, m2 * d2);
    return b + a;
}
template <bool SKEW_Y = false>
__device__ __forceinline__ float simplex3(float vx, float vy, float vz) { return simplex3_raw<SKEW_Y>(vx, vy, vz) * 42.0f; }

// ---------------------------------------------------------------- fbm (rng.hpp:166-191)
template <int OCT, bool SKEW_X = false>
__device__ __forceinline__ float fbm2(float x, float y)
{
    float f = 0.0f, amp = 1.0f;
#pragma unroll
    for (int i = 0; i < OCT; ++i)
    {
        amp *= 0.5f;
        f = fmaf(simplex2<SKEW_X>(x, y), amp, f);
        x = x + x; y = y + y;
    }
    return f;
}

template <int OCT, bool SKEW_Y = false>
__device__ __forceinline__ float fbm3(float x, float y, float z)
{
    float f = 0.0f, amp = 1.0f;
#pragma unroll
    for (int i = 0; i < OCT; ++i)
    {
        amp *= 0.5f;
        f = fmaf(simplex3<SKEW_Y>(x, y, z), amp, f);
        x = x + x; y = y + y; z = z + z;
    }
    return f;
}

}  // namespace mmg

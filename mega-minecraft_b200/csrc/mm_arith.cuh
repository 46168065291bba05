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
// and calls this once (all threads of the CTA, at a CTA-uniform point) before the first evaluation.
// The 10 KB arrive as ONE bulk copy (cp.async.bulk global -> shared, completion counted on an mbarrier) issued by thread 0:
// the copy engine moves the bytes, the CTA spends one instruction instead of 653 LDG + STS pairs and their loop - staging was
// 7 % of k_caves' executed instructions and 14 % of its stall samples (profiles/r02_k_caves_v2_src.txt) with a thread copy.
__device__ __forceinline__ void noise_tab_stage()
{
    __shared__ __align__(8) unsigned long long shTabBar;
    static_assert(kNoiseSmemBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&shTabBar);
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(mmg_dyn_smem);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kNoiseSmemBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(&g_noiseTab), "r"(kNoiseSmemBytes), "r"(bar) : "memory");
    }
    __syncthreads();      // the barrier is initialised (and the copy in flight) before anyone waits on it
    unsigned done;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    } while (!done);
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

// ---------------------------------------------------------------- packed fp32 (sm_100: FFMA2 / FADD2 / FMUL2)
// Two IEEE fp32 operations per issued instruction. Each half is rounded exactly like the scalar instruction of the same
// name (fma.rn / add.rn / mul.rn), so a pair of independent evaluations written with these gives the two scalar results
// bit for bit. The noise kernels are bound by issue slots, not by the FMA pipe (tools/ffma2_probe.cu: 16 FFMA + 16 integer
// instructions take 1.61 ms where 8 FFMA2 + 16 integer instructions take 1.08), which is what makes this worth having.
struct f32x2
{
    unsigned long long v;
};
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 f2_dup(float a) { return f2_make(a, a); }
__device__ __forceinline__ float f2_lo(f32x2 a) { return __uint_as_float((unsigned)a.v); }
__device__ __forceinline__ float f2_hi(f32x2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// Two independent simplex3_raw evaluations (lo and hi halves), operation for operation the scalar routine above: the vector
// arithmetic up to the falloffs m^4 runs packed, floor / compare / lattice hash / gradient fetch and the gradient dots run
// per half (the gradients come out of two different LDS.128, their components cannot be register pairs).
template <bool SKEW_Y = false>
__device__ MMG_NOISE_INLINE f32x2 simplex3x2_raw(f32x2 vx, f32x2 vy, f32x2 vz)
{
    const NoiseTab* T = noise_tab();
    const f32x2 C = f2_dup(1.0f / 3.0f), D = f2_dup(1.0f / 6.0f);
    const f32x2 s = SKEW_Y ? f2_fma(vz, C, f2_fma(vx, C, f2_mul(vy, C))) : f2_fma(vz, C, f2_fma(vy, C, f2_mul(vx, C)));
    const f32x2 qx = f2_add(vx, s), qy = f2_add(vy, s), qz = f2_add(vz, s);
    const f32x2 ix = f2_make(floorf(f2_lo(qx)), floorf(f2_hi(qx))), iy = f2_make(floorf(f2_lo(qy)), floorf(f2_hi(qy))),
                iz = f2_make(floorf(f2_lo(qz)), floorf(f2_hi(qz)));
    const f32x2 t = f2_fma(iz, D, f2_fma(ix, D, f2_mul(iy, D)));
    const f32x2 x0x = f2_add(f2_sub(vx, ix), t), x0y = f2_add(f2_sub(vy, iy), t), x0z = f2_add(f2_sub(vz, iz), t);
    float dots[2][4];
    float e1[2][3], e2[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        const float ax = h ? f2_hi(x0x) : f2_lo(x0x), ay = h ? f2_hi(x0y) : f2_lo(x0y), az = h ? f2_hi(x0z) : f2_lo(x0z);
        const bool gx = !(ax < ay), gy = !(ay < az), gz = !(az < ax);
        const bool a1x = gx && !gz, a1y = gy && !gx, a1z = gz && !gy;
        const bool a2x = gx || !gz, a2y = gy || !gx, a2z = gz || !gy;
        e1[h][0] = a1x ? 1.0f : 0.0f; e1[h][1] = a1y ? 1.0f : 0.0f; e1[h][2] = a1z ? 1.0f : 0.0f;
        e2[h][0] = a2x ? 1.0f : 0.0f; e2[h][1] = a2y ? 1.0f : 0.0f; e2[h][2] = a2z ? 1.0f : 0.0f;
    }
    const f32x2 x1x = f2_add(f2_sub(x0x, f2_make(e1[0][0], e1[1][0])), D), x1y = f2_add(f2_sub(x0y, f2_make(e1[0][1], e1[1][1])), D),
                x1z = f2_add(f2_sub(x0z, f2_make(e1[0][2], e1[1][2])), D);
    const f32x2 x2x = f2_add(f2_sub(x0x, f2_make(e2[0][0], e2[1][0])), C), x2y = f2_add(f2_sub(x0y, f2_make(e2[0][1], e2[1][1])), C),
                x2z = f2_add(f2_sub(x0z, f2_make(e2[0][2], e2[1][2])), C);
    const f32x2 H = f2_dup(0.5f);
    const f32x2 x3x = f2_sub(x0x, H), x3y = f2_sub(x0y, H), x3z = f2_sub(x0z, H);
    // mod289 of the cell: x - floor(x * (1/289)) * 289   (sx_mod289; the product and the final subtraction are separate roundings)
    const f32x2 R289 = f2_dup(1.0f / 289.0f), K289 = f2_dup(289.0f);
    const f32x2 mx = f2_mul(ix, R289), my = f2_mul(iy, R289), mz = f2_mul(iz, R289);
    const f32x2 jxf = f2_sub(ix, f2_mul(f2_make(floorf(f2_lo(mx)), floorf(f2_hi(mx))), K289));
    const f32x2 jyf = f2_sub(iy, f2_mul(f2_make(floorf(f2_lo(my)), floorf(f2_hi(my))), K289));
    const f32x2 jzf = f2_sub(iz, f2_mul(f2_make(floorf(f2_lo(mz)), floorf(f2_hi(mz))), K289));
    const f32x2 P6 = f2_dup(0.6f);
    // m = max(0.6 - |x|^2, 0)^4 for the four corners
    f32x2 m0 = f2_sub(P6, f2_fma(x0z, x0z, f2_fma(x0x, x0x, f2_mul(x0y, x0y))));
    f32x2 m1 = f2_sub(P6, f2_fma(x1z, x1z, f2_fma(x1x, x1x, f2_mul(x1y, x1y))));
    f32x2 m2 = f2_sub(P6, f2_fma(x2z, x2z, f2_fma(x2x, x2x, f2_mul(x2y, x2y))));
    f32x2 m3 = f2_sub(P6, f2_fma(x3z, x3z, f2_fma(x3x, x3x, f2_mul(x3y, x3y))));
    m0 = f2_make(fmaxf(f2_lo(m0), 0.0f), fmaxf(f2_hi(m0), 0.0f)); m1 = f2_make(fmaxf(f2_lo(m1), 0.0f), fmaxf(f2_hi(m1), 0.0f));
    m2 = f2_make(fmaxf(f2_lo(m2), 0.0f), fmaxf(f2_hi(m2), 0.0f)); m3 = f2_make(fmaxf(f2_lo(m3), 0.0f), fmaxf(f2_hi(m3), 0.0f));
    m0 = f2_mul(m0, m0); m1 = f2_mul(m1, m1); m2 = f2_mul(m2, m2); m3 = f2_mul(m3, m3);
    m0 = f2_mul(m0, m0); m1 = f2_mul(m1, m1); m2 = f2_mul(m2, m2); m3 = f2_mul(m3, m3);
    const char* PB = reinterpret_cast<const char*>(T->perm);
    const char* GB = reinterpret_cast<const char*>(T->grad3);
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        const int jx = (int)(h ? f2_hi(jxf) : f2_lo(jxf)), jy = (int)(h ? f2_hi(jyf) : f2_lo(jyf)), jz = (int)(h ? f2_hi(jzf) : f2_lo(jzf));
        const int jx2 = jx + jx, jy2 = jy + jy, jz2 = jz + jz;
        const bool a1x = e1[h][0] != 0.0f, a1y = e1[h][1] != 0.0f, a1z = e1[h][2] != 0.0f;
        const bool a2x = e2[h][0] != 0.0f, a2y = e2[h][1] != 0.0f, a2z = e2[h][2] != 0.0f;
        const int pz0 = MMG_P2(jz2), pz1 = MMG_P2(jz2 + 2);
        const int p0 = MMG_P2(MMG_P2(pz0 + jy2) + jx2);
        const int p1 = MMG_P2(MMG_P2((a1z ? pz1 : pz0) + jy2 + (a1y ? 2 : 0)) + jx2 + (a1x ? 2 : 0));
        const int p2 = MMG_P2(MMG_P2((a2z ? pz1 : pz0) + jy2 + (a2y ? 2 : 0)) + jx2 + (a2x ? 2 : 0));
        const int p3 = MMG_P2(MMG_P2(pz1 + jy2 + 2) + jx2 + 2);
        const float4 G0 = *reinterpret_cast<const float4*>(GB + 8 * p0), G1 = *reinterpret_cast<const float4*>(GB + 8 * p1);
        const float4 G2 = *reinterpret_cast<const float4*>(GB + 8 * p2), G3 = *reinterpret_cast<const float4*>(GB + 8 * p3);
#define MMG_H(a) (h ? f2_hi(a) : f2_lo(a))
        dots[h][0] = fmaf(MMG_H(x0z), G0.z, fmaf(MMG_H(x0x), G0.x, MMG_H(x0y) * G0.y));
        dots[h][1] = fmaf(MMG_H(x1z), G1.z, fmaf(MMG_H(x1x), G1.x, MMG_H(x1y) * G1.y));
        dots[h][2] = fmaf(MMG_H(x2z), G2.z, fmaf(MMG_H(x2x), G2.x, MMG_H(x2y) * G2.y));
        dots[h][3] = fmaf(MMG_H(x3z), G3.z, fmaf(MMG_H(x3x), G3.x, MMG_H(x3y) * G3.y));
#undef MMG_H
    }
    const f32x2 d0 = f2_make(dots[0][0], dots[1][0]), d1 = f2_make(dots[0][1], dots[1][1]);
    const f32x2 d2 = f2_make(dots[0][2], dots[1][2]), d3 = f2_make(dots[0][3], dots[1][3]);
    const f32x2 a = f2_fma(m1, d1, f2_mul(m0, d0));
    const f32x2 b = f2_fma(m3, d3, f2_mul(m2, d2));
    return f2_add(b, a);
}

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

// Two independent simplex2_raw evaluations, operation for operation the scalar routine.
template <bool SKEW_X = false>
__device__ MMG_NOISE_INLINE f32x2 simplex2x2_raw(f32x2 vx, f32x2 vy)
{
    const NoiseTab* T = noise_tab();
    const f32x2 C0 = f2_dup(0.211324865405187f), C1 = f2_dup(0.366025403784439f), C2 = f2_dup(-0.577350269189626f);
    const f32x2 s = SKEW_X ? f2_fma(vx, C1, f2_mul(vy, C1)) : f2_fma(vy, C1, f2_mul(vx, C1));
    const f32x2 qx = f2_add(vx, s), qy = f2_add(vy, s);
    const f32x2 ix = f2_make(floorf(f2_lo(qx)), floorf(f2_hi(qx))), iy = f2_make(floorf(f2_lo(qy)), floorf(f2_hi(qy)));
    const f32x2 t = f2_fma(ix, C0, f2_mul(iy, C0));
    const f32x2 x0x = f2_add(f2_sub(vx, ix), t), x0y = f2_add(f2_sub(vy, iy), t);
    const bool gtA = f2_lo(x0x) > f2_lo(x0y), gtB = f2_hi(x0x) > f2_hi(x0y);
    const f32x2 i1x = f2_make(gtA ? 1.0f : 0.0f, gtB ? 1.0f : 0.0f), i1y = f2_make(gtA ? 0.0f : 1.0f, gtB ? 0.0f : 1.0f);
    const f32x2 x1x = f2_sub(f2_add(x0x, C0), i1x), x1y = f2_sub(f2_add(x0y, C0), i1y);
    const f32x2 x2x = f2_add(x0x, C2), x2y = f2_add(x0y, C2);
    const f32x2 H = f2_dup(0.5f);
    f32x2 m0 = f2_sub(H, f2_fma(x0x, x0x, f2_mul(x0y, x0y)));
    f32x2 m1 = f2_sub(H, f2_fma(x1x, x1x, f2_mul(x1y, x1y)));
    f32x2 m2 = f2_sub(H, f2_fma(x2x, x2x, f2_mul(x2y, x2y)));
    m0 = f2_make(fmaxf(f2_lo(m0), 0.0f), fmaxf(f2_hi(m0), 0.0f)); m1 = f2_make(fmaxf(f2_lo(m1), 0.0f), fmaxf(f2_hi(m1), 0.0f));
    m2 = f2_make(fmaxf(f2_lo(m2), 0.0f), fmaxf(f2_hi(m2), 0.0f));
    m0 = f2_mul(m0, m0); m1 = f2_mul(m1, m1); m2 = f2_mul(m2, m2);
    m0 = f2_mul(m0, m0); m1 = f2_mul(m1, m1); m2 = f2_mul(m2, m2);
    const char* PB = reinterpret_cast<const char*>(T->perm);
    const char* GB = reinterpret_cast<const char*>(T->grad2);
    float res[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
#define MMG_H(a) (h ? f2_hi(a) : f2_lo(a))
        const float fix = MMG_H(ix), fiy = MMG_H(iy);
        const bool gt = h ? gtB : gtA;
        int jx, jy;
        if (fmaxf(fabsf(fix), fabsf(fiy)) < 4194304.0f)      // see simplex2_raw
        {
            jx = (int)fix % 289; jx += jx < 0 ? 289 : 0;
            jy = (int)fiy % 289; jy += jy < 0 ? 289 : 0;
        }
        else
        {
            jx = (int)(fix - 289.0f * floorf(fix / 289.0f));
            jy = (int)(fiy - 289.0f * floorf(fiy / 289.0f));
        }
        const int jx2 = jx + jx, jy2 = jy + jy;
        const int e1x = gt ? 2 : 0;
        const int py0 = MMG_P2(jy2), py1 = MMG_P2(jy2 + 2);
        const float4 G0 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2(py0 + jx2));
        const float4 G1 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2((gt ? py0 : py1) + jx2 + e1x));
        const float4 G2 = *reinterpret_cast<const float4*>(GB + 8 * MMG_P2(py1 + jx2 + 2));
        const float n0 = MMG_H(m0) * G0.z, n1 = MMG_H(m1) * G1.z, n2 = MMG_H(m2) * G2.z;
        const float g0 = fmaf(MMG_H(x0y), G0.y, MMG_H(x0x) * G0.x);
        const float g1 = fmaf(MMG_H(x1y), G1.y, MMG_H(x1x) * G1.x);
        const float g2 = fmaf(MMG_H(x2y), G2.y, MMG_H(x2x) * G2.x);
        res[h] = fmaf(g2, n2, fmaf(g0, n0, g1 * n1));
#undef MMG_H
    }
    return f2_make(res[0], res[1]);
}

// two fbm2 streams; simplex2() = 130 * raw rounded, then fma(that, amp, f) as in fbm2
template <int OCT, bool SKEW_X = false>
__device__ __forceinline__ f32x2 fbm2x2(f32x2 x, f32x2 y)
{
    f32x2 f = f2_dup(0.0f);
    float amp = 1.0f;
    const f32x2 K130 = f2_dup(130.0f);
#pragma unroll
    for (int i = 0; i < OCT; ++i)
    {
        amp *= 0.5f;
        f = f2_fma(f2_mul(K130, simplex2x2_raw<SKEW_X>(x, y)), f2_dup(amp), f);
        x = f2_add(x, x); y = f2_add(y, y);
    }
    return f;
}

// ---- the same fbm sums with the simplex evaluations taken two at a time (simplex3x2_raw): identical values, ~25 % fewer
// issued instructions. fbm3x2: two streams (lo / hi); fbm3_paired: one stream whose octaves 2k and 2k+1 share a call
// (the octave positions are exact doublings, so they can be formed ahead of the sum, which keeps its order).
template <int OCT, bool SKEW_Y = false>
__device__ __forceinline__ f32x2 fbm3x2(f32x2 x, f32x2 y, f32x2 z)
{
    f32x2 f = f2_dup(0.0f);
    float amp = 1.0f;
    const f32x2 K42 = f2_dup(42.0f);
#pragma unroll
    for (int i = 0; i < OCT; ++i)
    {
        amp *= 0.5f;
        f = f2_fma(f2_mul(simplex3x2_raw<SKEW_Y>(x, y, z), K42), f2_dup(amp), f);
        x = f2_add(x, x); y = f2_add(y, y); z = f2_add(z, z);
    }
    return f;
}

template <int OCT, bool SKEW_Y = false>
__device__ __forceinline__ float fbm3_paired(float x, float y, float z)
{
    float f = 0.0f, amp = 1.0f;
#pragma unroll
    for (int i = 0; i < OCT; i += 2)
    {
        if (i + 1 < OCT)
        {
            const float x2 = x + x, y2 = y + y, z2 = z + z;
            const f32x2 r = simplex3x2_raw<SKEW_Y>(f2_make(x, x2), f2_make(y, y2), f2_make(z, z2));
            amp *= 0.5f;
            f = fmaf(f2_lo(r) * 42.0f, amp, f);
            amp *= 0.5f;
            f = fmaf(f2_hi(r) * 42.0f, amp, f);
            x = x2 + x2; y = y2 + y2; z = z2 + z2;
        }
        else
        {
            amp *= 0.5f;
            f = fmaf(simplex3<SKEW_Y>(x, y, z), amp, f);
        }
    }
    return f;
}

// fbm3From3 (rng.hpp:188-191) with one skew variant for all three components: (o1, o2) as two streams, o3 by octave pairs
template <int OCT, bool SKEW_Y = false>
__device__ __forceinline__ void fbm3_from3(float ax, float ay, float az, float* o1, float* o2, float* o3)
{
    const f32x2 r = fbm3x2<OCT, SKEW_Y>(f2_make(ax, ax + 5923.45f), f2_make(ay, ay + 4129.42f), f2_make(az, az + 5790.48f));
    *o1 = f2_lo(r); *o2 = f2_hi(r);
    *o3 = fbm3_paired<OCT, SKEW_Y>(ax + 1765.68f, ay + 4704.36f, az + 5692.12f);
}

}  // namespace mmg

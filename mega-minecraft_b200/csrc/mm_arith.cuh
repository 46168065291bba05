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
    __device__ __forceinline__ explicit Minstd(uint32_t seed)
    {
        x = seed % 2147483647u;
        if (x == 0) x = 1;
    }
    __device__ __forceinline__ uint32_t next()
    {
        x = (uint32_t)(((uint64_t)x * 48271u) % 2147483647u);
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

// ---------------------------------------------------------------- simplex 2-D
// permute(x) = mod289((34x+1)x) on small non-negative integers: every step is exact in fp32,
// so contraction cannot change it (34*290+1 and its product with 290 are < 2^24).
__device__ __forceinline__ float sx_mod289(float x) { return x - floorf(x * (1.0f / 289.0f)) * 289.0f; }
__device__ __forceinline__ float sx_permute(float x) { return sx_mod289(fmaf(x, 34.0f, 1.0f) * x); }

// returns dot(m, g) BEFORE the final *130: callers that add to the result fuse that multiply
// (fma(raw, 130, c)); everything else uses 130*raw rounded.
// SKEW_X selects which product of the skew dot(v, C.yy) the reference build fused: most inlined
// copies compute fma(v.y, C1, v.x*C1) (false); the copies listed in DESIGN.md (direct simplex()
// calls hoisted in front of the biome loop) compute fma(v.x, C1, v.y*C1) (true). The value only
// feeds floor(), so the two differ only when v + s lands within an ulp of an integer.
template <bool SKEW_X = false>
__device__ MMG_NOISE_INLINE float simplex2_raw(float vx, float vy)
{
    const float C0 = 0.211324865405187f, C1 = 0.366025403784439f, C2 = -0.577350269189626f, C3 = 0.024390243902439f;
    // i = floor(v + dot(v, C.yy)); dot = fma(v.y, C1, v.x*C1)
    float s = SKEW_X ? fmaf(vx, C1, vy * C1) : fmaf(vy, C1, vx * C1);
    float ix = floorf(vx + s), iy = floorf(vy + s);
    // x0 = v - i + dot(i, C.xx); dot = fma(i.x, C0, i.y*C0)
    float t = fmaf(ix, C0, iy * C0);
    float x0x = (vx - ix) + t, x0y = (vy - iy) + t;
    float i1x = (x0x > x0y) ? 1.0f : 0.0f, i1y = (x0x > x0y) ? 0.0f : 1.0f;
    float x1x = (x0x + C0) - i1x, x1y = (x0y + C0) - i1y;
    float x2x = x0x + C2, x2y = x0y + C2;
    // mod(i, 289) = i - 289*floor(i/289) (true division; exact for these integers)
    float mx = ix - 289.0f * floorf(ix / 289.0f), my = iy - 289.0f * floorf(iy / 289.0f);
    float p0 = sx_permute(sx_permute(my + 0.0f) + mx + 0.0f);
    float p1 = sx_permute(sx_permute(my + i1y) + mx + i1x);
    float p2 = sx_permute(sx_permute(my + 1.0f) + mx + 1.0f);
    float m0 = fmaxf(0.5f - fmaf(x0x, x0x, x0y * x0y), 0.0f);
    float m1 = fmaxf(0.5f - fmaf(x1x, x1x, x1y * x1y), 0.0f);
    float m2 = fmaxf(0.5f - fmaf(x2x, x2x, x2y * x2y), 0.0f);
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    m0 = m0 * m0; m1 = m1 * m1; m2 = m2 * m2;
    // x = 2*fract(p*C.w) - 1 ; h = |x| - 0.5 ; ox = floor(x + 0.5) ; a0 = x - ox
    float q0 = p0 * C3, q1 = p1 * C3, q2 = p2 * C3;
    float gx0 = fmaf(q0 - floorf(q0), 2.0f, -1.0f);
    float gx1 = fmaf(q1 - floorf(q1), 2.0f, -1.0f);
    float gx2 = fmaf(q2 - floorf(q2), 2.0f, -1.0f);
    float h0 = fabsf(gx0) - 0.5f, h1 = fabsf(gx1) - 0.5f, h2 = fabsf(gx2) - 0.5f;
    float a0 = gx0 - floorf(gx0 + 0.5f), a1 = gx1 - floorf(gx1 + 0.5f), a2 = gx2 - floorf(gx2 + 0.5f);
    // m *= 1.79284291400159 - 0.85373472095314 * (a0*a0 + h*h)
    m0 = m0 * fmaf(fmaf(h0, h0, a0 * a0), -0.85373472095314f, 1.79284291400159f);
    m1 = m1 * fmaf(fmaf(h1, h1, a1 * a1), -0.85373472095314f, 1.79284291400159f);
    m2 = m2 * fmaf(fmaf(h2, h2, a2 * a2), -0.85373472095314f, 1.79284291400159f);
    // g = a0*x.x + h*x.y  -> fma(x.y, h, x.x*a0)
    float g0 = fmaf(x0y, h0, x0x * a0);
    float g1 = fmaf(x1y, h1, x1x * a1);
    float g2 = fmaf(x2y, h2, x2x * a2);
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
    const float C = 1.0f / 3.0f, D = 1.0f / 6.0f;
    const float NZ = 0.142857142857f;           // n_
    const float NX = NZ * 2.0f;                  // ns.x = n_*D.w - D.x
    const float NY = NZ * 0.5f - 1.0f;           // ns.y = n_*D.y - D.z
    float s = SKEW_Y ? fmaf(vz, C, fmaf(vx, C, vy * C)) : fmaf(vz, C, fmaf(vy, C, vx * C));
    float ix = floorf(vx + s), iy = floorf(vy + s), iz = floorf(vz + s);
    float t = fmaf(iz, D, fmaf(ix, D, iy * D));
    float x0x = (vx - ix) + t, x0y = (vy - iy) + t, x0z = (vz - iz) + t;
    // g = step(x0.yzx, x0) ; l = 1 - g ; i1 = min(g, l.zxy) ; i2 = max(g, l.zxy)
    float gx = (x0x < x0y) ? 0.0f : 1.0f, gy = (x0y < x0z) ? 0.0f : 1.0f, gz = (x0z < x0x) ? 0.0f : 1.0f;
    float lx = 1.0f - gx, ly = 1.0f - gy, lz = 1.0f - gz;
    float i1x = fminf(gx, lz), i1y = fminf(gy, lx), i1z = fminf(gz, ly);
    float i2x = fmaxf(gx, lz), i2y = fmaxf(gy, lx), i2z = fmaxf(gz, ly);
    float x1x = (x0x - i1x) + D, x1y = (x0y - i1y) + D, x1z = (x0z - i1z) + D;
    float x2x = (x0x - i2x) + C, x2y = (x0y - i2y) + C, x2z = (x0z - i2z) + C;
    float x3x = x0x - 0.5f, x3y = x0y - 0.5f, x3z = x0z - 0.5f;
    float mx = sx_mod289(ix), my = sx_mod289(iy), mz = sx_mod289(iz);
    float p[4];
    {
        const float ez[4] = {0.0f, i1z, i2z, 1.0f}, ey[4] = {0.0f, i1y, i2y, 1.0f}, ex[4] = {0.0f, i1x, i2x, 1.0f};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            p[k] = sx_permute(sx_permute(sx_permute(mz + ez[k]) + my + ey[k]) + mx + ex[k]);
    }
    float gxk[4], gyk[4], hk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        float j = p[k] - 49.0f * floorf((p[k] * NZ) * NZ);   // exact (integers)
        float x_ = floorf(j * NZ);
        float y_ = floorf(j - 7.0f * x_);
        gxk[k] = fmaf(x_, NX, NY);
        gyk[k] = fmaf(y_, NX, NY);
        hk[k] = (1.0f - fabsf(gxk[k])) - fabsf(gyk[k]);
    }
    const float xsx[4] = {x0x, x1x, x2x, x3x}, xsy[4] = {x0y, x1y, x2y, x3y}, xsz[4] = {x0z, x1z, x2z, x3z};
    float md[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        // a = b + s*sh with s = floor(b)*2+1, sh = -step(h, 0)
        float sh = (hk[k] <= 0.0f) ? -1.0f : 0.0f;
        float ax = fmaf(fmaf(floorf(gxk[k]), 2.0f, 1.0f), sh, gxk[k]);
        float ay = fmaf(fmaf(floorf(gyk[k]), 2.0f, 1.0f), sh, gyk[k]);
        float az = hk[k];
        // norm = taylorInvSqrt(dot(p,p)); dot3 = fma(z,z, fma(x,x, y*y))
        float nrm = fmaf(fmaf(az, az, fmaf(ax, ax, ay * ay)), -0.85373472095314f, 1.79284291400159f);
        ax = ax * nrm; ay = ay * nrm; az = az * nrm;
        float X = xsx[k], Y = xsy[k], Z = xsz[k];
        float m = fmaxf(0.6f - fmaf(Z, Z, fmaf(X, X, Y * Y)), 0.0f);
        m = m * m;
        m = m * m;
        float dpx = fmaf(Z, az, fmaf(X, ax, Y * ay));
        md[k] = m;
        gxk[k] = dpx;   // reuse: dot(p_k, x_k)
    }
    // 42 * dot(m*m, dots): (t.x + t.y) + (t.z + t.w) with fma on .y and .w
    float a = fmaf(md[1], gxk[1], md[0] * gxk[0]);
    float b = fmaf(md[3], gxk[3], md[2] * gxk[2]);
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

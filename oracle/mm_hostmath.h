// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product (the product carries its own
// copy for device code in mega-minecraft_b200/csrc/mm_hostmath.cuh).
//
// The reference evaluates isFeaturePos() (/root/reference/src/terrain/chunk.cu:999-1008, via
// rand2From3, rng.hpp:131-137) on the HOST, so its sin() is the C library's sinf, not libdevice's.
// glibc is not under /root/reference; this image ships glibc 2.39 (Ubuntu 2.39-0ubuntu8.5), whose
// x86-64 sinf resolves to the FMA build of sysdeps/ieee754/flt-32/s_sinf.c (ARM optimized-routines
// sincosf). The routine below restates that algorithm, with the fused operations exactly where the
// shipped binary has them (read from its disassembly) and the coefficient tables read from it, so
// the same bits can be produced on the GPU. tests/test_hostmath.py checks it against the C library
// over a dense sweep of all exponent ranges.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace mmo {

static const uint32_t hm_inv_pio4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
                                         0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
                                         0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
// {c0, c1, s1, c2, s2, c3, s3, c4} for quadrant sets 0 and 1
static const double hm_poly[2][8] = {
    {0x1.0000000000000p+0, -0x1.ffffffd0c621cp-2, -0x1.555545995a603p-3, 0x1.55553e1068f19p-5, 0x1.1107605230bc4p-7,
     -0x1.6c087e89a359dp-10, -0x1.994eb3774cf24p-13, 0x1.99343027bf8c3p-16},
    {-0x1.0000000000000p+0, 0x1.ffffffd0c621cp-2, -0x1.555545995a603p-3, -0x1.55553e1068f19p-5, 0x1.1107605230bc4p-7,
     0x1.6c087e89a359dp-10, -0x1.994eb3774cf24p-13, -0x1.99343027bf8c3p-16}};
static const double hm_sign[4] = {1.0, -1.0, -1.0, 1.0};

static inline float hm_sin_poly(double xs, double x2, const double* p)
{
    const double s1 = fma(x2, p[6], p[4]);
    const double x3 = x2 * xs;
    const double x7 = x2 * x3;
    const double s = fma(x3, p[2], xs);
    return (float)fma(s1, x7, s);
}
static inline float hm_cos_poly(double x2, const double* p)
{
    const double x4 = x2 * x2;
    const double c1 = fma(x2, p[1], p[0]);
    const double c2 = fma(x2, p[7], p[5]);
    const double x6 = x4 * x2;
    const double c = fma(x4, p[3], c1);
    return (float)fma(c2, x6, c);
}

static inline float hm_sinf(float y)
{
    uint32_t xi;
    std::memcpy(&xi, &y, 4);
    const double x = (double)y;
    const uint32_t top12 = (xi >> 20) & 0x7ff;
    if (top12 <= 0x3f3)            // |y| < pi/4
    {
        if (top12 <= 0x397) return y;   // |y| < 2^-12
        return hm_sin_poly(x, x * x, hm_poly[0]);
    }
    if (top12 <= 0x42e)            // |y| < 120
    {
        const double r = x * 0x1.45f306dc9c883p+23;
        const int32_t n = ((int32_t)r + 0x800000) >> 24;
        const double xr = fma(-(double)n, 0x1.921fb54442d18p+0, x);
        const double* p = hm_poly[(n & 2) ? 1 : 0];
        const double x2 = xr * xr;
        if (n & 1) return hm_cos_poly(x2, p);
        return hm_sin_poly(xr * hm_sign[n & 3], x2, p);
    }
    if (top12 <= 0x7f7)            // finite, large: 96-bit multiply by 4/pi
    {
        const uint32_t* arr = &hm_inv_pio4[(xi >> 26) & 15];
        const int shift = (xi >> 23) & 7;
        const uint32_t sign = xi >> 31;
        uint32_t m = ((xi & 0x7fffff) | 0x800000) << shift;
        uint64_t res0 = (uint32_t)(m * arr[0]);
        const uint64_t res1 = (uint64_t)m * arr[4];
        const uint64_t res2 = (uint64_t)m * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const uint64_t n = (res0 + (1ULL << 61)) >> 62;
        res0 -= n << 62;
        const double xr = (double)(int64_t)res0 * 0x1.921fb54442d18p-62;
        const uint32_t ns = (uint32_t)n + sign;
        const double* p = hm_poly[(ns & 2) ? 1 : 0];
        const double x2 = xr * xr;
        if (n & 1) return hm_cos_poly(x2, p);
        return hm_sin_poly(xr * hm_sign[ns & 3], x2, p);
    }
    return y - y;                  // inf/nan -> nan
}

}  // namespace mmo

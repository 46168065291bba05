// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// Feature rasterisers of the reference, restated:
//   SDFs / splines / helpers     /root/reference/src/terrain/featurePlacement.hpp:15-142
//   placeFeature (20 features)   featurePlacement.hpp:147-1107
//   placeCaveFeature (9)         featurePlacement.hpp:1110-1380
//   geometry helpers             /root/reference/src/util/rng.hpp:9-63
// RNG draw order: the reference's device build evaluates constructor arguments left to right
// (SURVEY.md B-3); every draw below is a separate statement in that order.
#pragma once
#include "mm_surface.h"

namespace mmo {

// Every fused multiply-add of the rasterisers goes through pf_fma, placed where the reference's sm_100 SASS has an FFMA.
// Diagnostic builds (oracle/Makefile: libmmoracle_unfused.so) round the product first instead: when the product and the
// reference disagree on a voxel, tests/test_reference_tour.py asks whether the un-fused (or the fully contracted,
// -ffp-contract=fast) arithmetic reproduces the reference's block there - i.e. whether the voxel sits on an fp32
// threshold that the placement of one rounding decides.
#ifdef MMO_PF_UNFUSED
static inline float pf_fma(float a, float b, float c) { volatile float p = a * b; return p + c; }
#else
static inline float pf_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// glm::dot(vec3): mul on y, fma x, fma z
static inline float dot3(V3 a, V3 b) { return pf_fma(a.z, b.z, pf_fma(a.x, b.x, a.y * b.y)); }
static inline float len3(V3 a) { return sqrtf(dot3(a, a)); }
static inline float len2(float x, float z) { return sqrtf(pf_fma(x, x, z * z)); }
// glm::mix(x, y, a) = x*(1-a) + y*a, compiled as fma(y, a, x*(1-a))
static inline float mixf(float x, float y, float a) { return pf_fma(y, a, x * (1.f - a)); }
static inline V3 mix3(V3 a, V3 b, float t) { return v3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
static inline V3 normalize3(V3 a) { const float inv = 1.f / sqrtf(dot3(a, a)); return a * inv; }
static inline V3 floor3(V3 a) { return v3(floorf(a.x), floorf(a.y), floorf(a.z)); }
static inline V3 ceil3(V3 a) { return v3(ceilf(a.x), ceilf(a.y), ceilf(a.z)); }
static inline bool in_range_f(float v, float lo, float hi) { return v >= lo && v <= hi; }
static inline bool in_range_i(int v, int lo, int hi) { return v >= lo && v <= hi; }
static inline bool saturated(float v) { return v >= 0.f && v <= 1.f; }
static inline float ratio_of(float v, float lo, float hi) { return (v - lo) / (hi - lo); }

constexpr float kPi = 3.14159265358979323846264338327f, kTwoPi = 6.28318530717958647692528676655f,
                kPiOverTwo = 1.57079632679489661923132169163f;

// rng.hpp:52-63. distFromLine = |vecLine*ratio - pointPos| (the product is fused into the subtraction)
static inline bool line_params(V3 pos, V3 l1, V3 l2, float* ratio, float* dist)
{
    const V3 vl = l2 - l1, pp = pos - l1;
    *ratio = dot3(pp, vl) / dot3(vl, vl);
    const V3 d = v3(pf_fma(vl.x, *ratio, -pp.x), pf_fma(vl.y, *ratio, -pp.y), pf_fma(vl.z, *ratio, -pp.z));
    *dist = len3(d);
    return saturated(*ratio);
}

// featurePlacement.hpp:68-74
static inline bool in_rasterized_line(int fx, int fy, int fz, V3 l1, V3 l2)
{
    float ratio, dist;
    const bool inLine = line_params(v3((float)fx + 0.5f, (float)fy + 0.5f, (float)fz + 0.5f), l1, l2, &ratio, &dist);
    if (!(inLine && dist < 2.f)) return false;
    const V3 m = floor3(mix3(l1, l2, ratio));
    return fx == (int)m.x && fy == (int)m.y && fz == (int)m.z;
}

// featurePlacement.hpp:40-66
template <int NC, int NS>
static inline void de_casteljau(const V3* ctrl, V3* spline)
{
    for (int i = 0; i < NS; ++i)
    {
        V3 c[NC];
        for (int j = 0; j < NC; ++j) c[j] = ctrl[j];
        const float t = (float)i / (float)(NS - 1);
        for (int points = NC; points > 1; --points)
            for (int j = 0; j < points - 1; ++j) c[j] = mix3(c[j], c[j + 1], t);
        spline[i] = c[0];
    }
}

// featurePlacement.hpp:80-90
static inline bool jungle_leaves(V3 pos, float maxHeight, float minRadius, float maxRadius, float rand)
{
    const float mult = pf_fma(rand, 0.4f, 0.8f);
    if (in_range_f(pos.y, 0.f, maxHeight))
    {
        const float r = mixf(maxRadius, minRadius, pos.y / maxHeight) * mult;
        return len2(pos.x, pos.z) < r;
    }
    return false;
}

// featurePlacement.hpp:92-125
static inline float crystal_radius(float ratio)
{
    const float coneStart = 0.8f, coneN = 1.f / (1.f - coneStart);
    return (ratio < coneStart) ? pf_fma(ratio, 0.25f, 0.8f) : coneN * (1.f - ratio);
}
static inline V3 cross3(V3 a, V3 b)
{
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline bool in_crystal(V3 pos, V3 p1, V3 p2, float radiusMul)
{
    float ratio, dist;
    if (!line_params(pos, p1, p2, &ratio, &dist)) return false;
    float radius = crystal_radius(ratio) * radiusMul;
    const float p = kPi / 6.f;
    const V3 line = p2 - p1;
    const V3 pp = pos - (p1 + line * ratio);
    float ang = 0.f;
    if (len3(pp) != 0.f)
    {
        const V3 a = normalize3(pp), b = normalize3(cross3(line, v3(1.f, 0.f, 0.f)));
        ang = dm_acosf(fminf(fmaxf(dot3(a, b), -1.f), 1.f)) + kTwoPi;
    }
    radius = radius * (dm_cosf(p) / dm_cosf(p - dm_fmodf(ang, 2.f * p)));
    return dist < radius;
}
static inline uint8_t random_crystal_block(float rand)
{
    const float r = rand * 3.f;
    return r < 1.f ? B_MAGENTA_CRYSTAL : (r < 2.f ? B_CYAN_CRYSTAL : B_GREEN_CRYSTAL);
}

// featurePlacement.hpp:147-1107. Returns true and sets *out when the voxel belongs to the feature.
static inline bool place_feature(const FeaturePlacement& fp, int wx, int wy, int wz, uint8_t* out)
{
    const int fx = wx - fp.x, fy = wy - fp.y, fz = wz - fp.z;
    V3 pos = v3((float)fx, (float)fy, (float)fz);
    Minstd frng = make_rng4(fp.x, fp.y, fp.z, 1293012);
    Minstd brng = make_rng4(wx, wy, wz, 57847812);
    switch (fp.feature)
    {
    case F_NONE: return false;
    case F_SPHERE:
        if (dot3(pos, pos) > 25.f) return false;
        *out = B_GRAVEL;
        return true;
    case F_CORAL:
    {
        if (fp.y > SEA_LEVEL - 6) return false;
        const float x2 = pos.x * pos.x, z2 = pos.z * pos.z;       // shared by the two length() calls: not fused
        if (sqrtf(x2 + z2) > 8.f) return false;
        const int coral = (int)(frng.u01() * 5.f);
        switch (coral)
        {
        case 0:
        case 1:
        {
            const float ys = pos.y * (coral == 0 ? 1.15f : 1.25f);
            float radius = coral == 0 ? pf_fma(frng.u01(), 1.4f, 2.8f) : pf_fma(frng.u01(), 1.7f, 2.2f);
            const float sc = coral == 0 ? 0.2f : 0.3f;
            radius = pf_fma(simplex3<true>((float)wx * sc, (float)wy * sc, (float)wz * sc), coral == 0 ? 0.4f : 1.2f, radius);
            if (sqrtf(z2 + pf_fma(ys, ys, x2)) < radius) { *out = coral == 0 ? B_BRAIN_CORAL_BLOCK : B_BUBBLE_CORAL_BLOCK; return true; }
            return false;
        }
        case 2:
        case 3:
        {
            const uint8_t block = coral == 2 ? B_FIRE_CORAL_BLOCK : B_HORN_CORAL_BLOCK;
            const float a = frng.u11(), b = frng.u01(), c = frng.u11();
            const V3 p1 = v3(a * 2.5f, b * 3.5f, c * 2.5f);
            if (in_rasterized_line(fx, fy, fz, v3(0, 0, 0), p1)) { *out = block; return true; }
            for (int i = 0; i < 5; ++i)
            {
                V3 p2 = p1;
                p2.x = pf_fma(frng.u11(), 4.f, p2.x);
                p2.y = p2.y + pf_fma(frng.u01(), 3.f, 2.f);
                p2.z = pf_fma(frng.u11(), 4.f, p2.z);
                if (in_rasterized_line(fx, fy, fz, p1, p2)) { *out = block; return true; }
            }
            return false;
        }
        case 4:
        {
            const Worley2 w = worley2((float)wx * 0.7f, (float)wz * 0.7f);
            float h = (1.f - w.d1) + (w.d2 - w.d1) * 0.5f;
            h = h * 3.5f;
            h = h * ss_t((sqrtf(x2 + z2) + -3.7f) / (2.5f - 3.7f));
            h = h - 2.f;
            if (in_range_f(pos.y, -1.f, h)) { *out = B_TUBE_CORAL_BLOCK; return true; }
            return false;
        }
        }
        return false;
    }
    case F_KELP:
    {
        if (fx != 0 || fz != 0) return false;
        int height = (int)pf_fma(frng.u01(), 15.f, 5.f);
        height = height < SEA_LEVEL - fp.y - 1 ? height : SEA_LEVEL - fp.y - 1;
        if (!in_range_i(fy, 0, height)) return false;
        *out = (fy == height) ? B_KELP_END : B_KELP_MAIN;
        return true;
    }
    case F_ICEBERG:
    {
        if (fp.y > SEA_LEVEL - 32) return false;
        pos.y = (float)(wy - SEA_LEVEL);
        const float hd = len2(pos.x, pos.z);
        const float radius = pf_fma(frng.u01(), 12.f, 20.f);
        const float center = 1.f - (hd / radius);
        if (center > 1.15f) return false;
        const float nx = (float)wx * 0.0450f, nz = (float)wz * 0.0450f;
        const float f = fbm2<3>(nx, nz);
        const float start = pf_fma(f, 14.f, pf_fma(center, -34.f, -6.f));
        const float end = pf_fma(f, 8.f, pf_fma(center, 20.f, -4.f));
        if (end < start || !in_range_f(pos.y, start, end)) return false;
        if (pos.y < -4.f) { *out = B_BLUE_ICE; return true; }
        const float packed = pf_fma(simplex2<true>(nx * 0.8000f, nz * 0.8000f), 1.2f, pf_fma(center, 5.6f, -2.2f));
        *out = (pos.y > end - packed) ? B_PACKED_ICE : B_BLUE_ICE;
        return true;
    }
    case F_ACACIA_TREE:
    {
        if ((abs(fx) > abs(fz) ? abs(fx) : abs(fz)) > 15) return false;
        const int trunk = (int)pf_fma(frng.u01(), 1.5f, 4.5f);
        if (fx == 0 && fz == 0 && in_range_i(fy, 0, trunk)) { *out = B_ACACIA_WOOD; return true; }
        float angle = frng.u01() * kTwoPi;
        V3 bs = v3(0.f, (float)trunk, 0.f), be = v3(0, 0, 0);
        dm_sincosf(angle, &be.z, &be.x);
        {
            const float s = pf_fma(frng.u01(), 1.5f, 2.f);
            be = v3(pf_fma(be.x, s, bs.x), pf_fma(be.y, s, bs.y), pf_fma(be.z, s, bs.z));
        }
        be.y = be.y + pf_fma(frng.u01(), 1.5f, 2.5f);
        if (in_rasterized_line(fx, fy, fz, floor3(bs), ceil3(be))) { *out = B_ACACIA_WOOD; return true; }
        V3 lp = v3((float)fx, (float)fy, (float)fz) - be;
        lp.y = lp.y + 0.5f;
        if (jungle_leaves(lp, 2.f, 2.f, 4.f, pf_fma(frng.u01(), 0.5f, 0.5f))) { *out = B_ACACIA_LEAVES; return true; }
        if (frng.u01() < 0.5f) return false;
        angle = angle + pf_fma(frng.u01(), kPi, kPiOverTwo);
        bs = v3(0.f, pf_fma(frng.u01(), -0.8f, (float)trunk - 0.8f), 0.f);
        be = v3(0, 0, 0);
        dm_sincosf(angle, &be.z, &be.x);
        {
            const float s = pf_fma(frng.u01(), 1.f, 1.5f);
            be = v3(pf_fma(be.x, s, bs.x), pf_fma(be.y, s, bs.y), pf_fma(be.z, s, bs.z));
        }
        be.y = be.y + pf_fma(frng.u01(), 1.f, 2.f);
        if (in_rasterized_line(fx, fy, fz, floor3(bs), ceil3(be))) { *out = B_ACACIA_WOOD; return true; }
        lp = v3((float)fx, (float)fy, (float)fz) - be;
        lp.y = lp.y + 0.5f;
        if (jungle_leaves(lp, 2.001f, 1.5f, 3.5f, pf_fma(frng.u01(), 0.5f, 0.5f))) { *out = B_ACACIA_LEAVES; return true; }
        return false;
    }
    case F_REDWOOD_TREE:
    {
        pos = pos * pf_fma(frng.u01(), 0.3f, 0.6f);
        const float height = pf_fma(frng.u01(), 13.f, 27.f);
        const float hd = len2(pos.x, pos.z);
        const float leavesStart = pf_fma(frng.u01(), 4.f, 10.f);
        if (pos.y > height + 8.f || hd > 12.f || (pos.y < leavesStart - 4.f && hd > 3.f)) return false;
        const float tr = ratio_of(pos.y, -4.f, height);
        if (saturated(tr))
        {
            float radius = 2.f / (tr + 2.f) + 0.08f / dm_powf(tr + 0.4f, 3.f);
            radius = pf_fma(simplex3<true>((float)wx * 0.1300f, (float)wy * 0.1300f, (float)wz * 0.1300f) * 0.3f,
                          ss_t((tr + -0.6f) / (0.2f - 0.6f)), radius);
            if (hd < radius) { *out = B_REDWOOD_WOOD; return true; }
        }
        const float leavesEnd = (height + 1.5f) + 1.f * frng.u01();
        if (!in_range_f(pos.y, leavesStart, leavesEnd)) return false;
        const int cellBase = (int)floorf(pos.y * 0.5f) * 2;
        const float branchSeed = 593.23f * hash_fract(pf_fma((float)fp.z, 640.88f, pf_fma((float)fp.x, 238.68f, (float)fp.y * 491.28f)));
        const float leavesSeed = 412.39f * hash_fract(branchSeed * 238.68f);
        const float leavesSimplex = 1.1f * simplex3<true>((float)wx * 0.2000f, (float)wy * 0.2000f, (float)wz * 0.2000f);
        bool inLeaves = false;
        for (int dy = -4; dy <= 4; dy += 2)
        {
            const int cell = cellBase + dy;
            const float fc = (float)cell;
            float hr = ratio_of(fc, leavesStart, leavesEnd);
            hr = pf_fma(hr, -0.5f, 1.1f);
            // rand3From2(vec2(cell, leavesSeed)) - 0.5
            V3 lc = v3(hash_fract(pf_fma(fc, 238.68f, leavesSeed * 491.28f)) - 0.5f, hash_fract(pf_fma(fc, 654.37f, leavesSeed * 560.45f)) - 0.5f,
                       hash_fract(pf_fma(fc, 640.88f, leavesSeed * 151.81f)) - 0.5f);
            lc = v3(lc.x * (7.5f * hr), lc.y * (1.3f * hr), lc.z * (7.5f * hr));
            lc.y = fminf(lc.y + fc, height + 0.8f);
            const V3 bs = v3(0.f, pf_fma(hash_fract((fc + branchSeed) * 238.68f), -1.5f, lc.y - 2.f), 0.f);
            float br, bd;
            if (line_params(pos, bs, lc, &br, &bd))
                if (saturated(br) && bd < 0.5f) { *out = B_REDWOOD_WOOD; return true; }
            if (inLeaves) continue;
            V3 lp = pos - lc;
            lp.y = lp.y * 1.7f;
            const float ld = len3(lp);
            if (ld > 5.0f) continue;
            float lr = pf_fma(hash_fract((fc + leavesSeed) * 238.68f), 0.5f, 2.5f) + leavesSimplex;
            lr = lr * hr;
            if (ld < lr) inLeaves = true;
        }
        if (inLeaves) { *out = B_REDWOOD_LEAVES; return true; }
        return false;
    }
    case F_CYPRESS_TREE:
    {
        const float trunkHeight = pf_fma(frng.u01(), 12.f, 25.f);
        const float td = len2(pos.x, pos.z);
        if (pos.y > trunkHeight + 4.f || td > 12.f) return false;
        const float tr = ratio_of(pos.y, -2.f, trunkHeight);
        if (saturated(tr))
        {
            float radius = pf_fma((1.3f + tr) / dm_powf(0.73f + tr, 4.f), 0.5f, 0.5f);
            radius = radius * pf_fma(simplex3<true>((float)wx * 0.1500f, (float)wy * 0.1500f, (float)wz * 0.1500f) * 0.3f,
                                   ss_t((tr + -0.55f) / (0.15f - 0.55f)), 1.f);
            if (td < radius) { *out = B_CYPRESS_WOOD; return true; }
        }
        if (jungle_leaves(pos - v3(0.f, trunkHeight, 0.f), 2.f, 3.f, 4.5f, frng.u01())) { *out = B_CYPRESS_LEAVES; return true; }
        const int numBranches = 6 + (int)(frng.u01() * 5.f);
        float branchHeight = trunkHeight - 1.f;
        float angle = frng.u01() * kTwoPi;
        for (int i = 0; i < numBranches; ++i)
        {
            branchHeight = branchHeight - pf_fma(frng.u01(), 3.6f, 1.f);
            angle = angle + pf_fma(frng.u01(), kPi, kPiOverTwo);
            const V3 bs = v3(0.f, branchHeight, 0.f);
            V3 be = v3(0, 0, 0);
            dm_sincosf(angle, &be.z, &be.x);
            be = be * pf_fma(frng.u01(), 1.5f, 4.f);
            be.y = pf_fma(frng.u01(), 1.2f, 2.2f);
            be = be * pf_fma(ratio_of(branchHeight, 0.f, trunkHeight), -0.3f, 1.f);
            be = be + bs;
            // isInRasterizedLine(ivec3(pos), ...): pos is truncated to ints by the implicit conversion
            if (in_rasterized_line((int)pos.x, (int)pos.y, (int)pos.z, bs, be)) { *out = B_CYPRESS_WOOD; return true; }
            V3 lp = v3((pos.x - be.x) + 0.3f, (pos.y - be.y) + 0.3f, (pos.z - be.z) + 0.3f);
            const float droop = hash_fract(pf_fma((float)wx, 238.68f, (float)wz * 491.28f));
            if (droop < 0.2f && in_range_f(lp.y, fmaxf(-2.f, droop * -10.f), 0.f)) lp.y = 0.f;
            if (jungle_leaves(lp, 2.f, 2.5f, 4.f, frng.u01())) { *out = B_CYPRESS_LEAVES; return true; }
        }
        return false;
    }
    case F_BIRCH_TREE:
    {
        int height = (int)pf_fma(frng.u01(), 4.f, 6.2f);
        const bool tall = frng.u01() < 0.08f;
        if (tall) height = (int)((float)height * 1.9f);
        if ((abs(fx) > abs(fz) ? abs(fx) : abs(fz)) > 8 || !in_range_i(fy, 0, height + 6)) return false;
        if (fx == 0 && fz == 0 && in_range_i(fy, 0, height)) { *out = B_BIRCH_WOOD; return true; }
        const float tm = tall ? 1.5f : 1.f;
        const float fh = (float)height;
        const float leavesStart = pf_fma(-pf_fma(frng.u01(), -2.2f, 3.0f), tm, fh);
        const float leavesEnd = pf_fma(pf_fma(frng.u01(), 1.2f, 4.2f), tm, fh);
        const float ratio = (pos.y - leavesStart) / (leavesEnd - leavesStart);
        if (!in_range_f(ratio, 0.f, 1.f)) return false;
        const float x = dm_powf(ratio, 0.8f);
        const float poly = (((0.5f * x) * x) * x - ((1.5f * x) * x)) + x;
        const float leavesRadius = (5.f * poly) * pf_fma(frng.u01(), 0.8f, 2.8f);
        if (len2(pos.x, pos.z) > leavesRadius) return false;
        const float lr = frng.u01();
        *out = lr < 0.1f ? B_YELLOW_BIRCH_LEAVES : (lr < 0.2f ? B_ORANGE_BIRCH_LEAVES : B_BIRCH_LEAVES);
        return true;
    }
    case F_PINE_TREE:
    {
        const int height = (int)pf_fma(frng.u01(), 4.f, 7.f);
        if (fy < 0 || fy > height + 4 || (abs(fx) > abs(fz) ? abs(fx) : abs(fz)) > 6) return false;
        if (fx == 0 && fz == 0 && fy <= height) { *out = B_PINE_WOOD; return true; }
        const float fh = (float)height;
        const float leavesStart = pf_fma(frng.u01(), -2.5f, fh - 4.f);
        const float leavesEnd = fh + 3.f;
        const float lr = (pos.y - leavesStart) / (leavesEnd - leavesStart);
        if (!in_range_f(lr, 0.f, 1.f)) return false;
        const float radius = mixf(3.f, 1.f, lr);
        if (len2(pos.x, pos.z) < radius) { *out = frng.u01() < 0.5f ? B_PINE_LEAVES_1 : B_PINE_LEAVES_2; return true; }
        return false;
    }
    case F_PINE_SHRUB:
    {
        const int height = (int)pf_fma(frng.u01(), 2.f, 2.f);
        if (fy < 0 || fy > height + 4 || (abs(fx) > abs(fz) ? abs(fx) : abs(fz)) > 6) return false;
        if (fx == 0 && fz == 0 && fy <= height) { *out = B_PINE_WOOD; return true; }
        const V3 lp = pos - v3(0.f, (float)height - 1.f, 0.f);
        if (jungle_leaves(lp, 2.5f, 1.5f, 2.5f, frng.u01())) { *out = frng.u01() < 0.5f ? B_PINE_LEAVES_1 : B_PINE_LEAVES_2; return true; }
        return false;
    }
    case F_MEDIUM_PURPLE_MUSHROOM:
    {
        if (abs(fx) + abs(fz) > 8) return false;
        const int height = (int)pf_fma(frng.u01(), 2.3f, 1.5f);
        if (fx == 0 && in_range_i(fy, 0, height) && fz == 0) { *out = B_MUSHROOM_STEM; return true; }
        const float radius = frng.u01() < 0.5f ? 1.8f : 2.5f;
        if (fy == height + 1 && len2(pos.x, pos.z) < radius) { *out = B_PURPLE_MUSHROOM_CAP; return true; }
        return false;
    }
    case F_PURPLE_MUSHROOM:
    {
        const float scale = pf_fma(frng.u01(), 1.2f, 1.f);
        pos = pos * scale;
        if (frng.u01() < 0.2f) pos = pos * 0.5f;
        const float height = pf_fma(frng.u01(), 30.f, 25.f);
        {
            const float x2 = pos.x * pos.x, z2 = pos.z * pos.z;     // shared squares: not fused
            const float dy = pos.y - height;
            if (pos.y < -1.f || pos.y > height + 12.f ||
                (sqrtf(x2 + z2) > 8.f && (pos.y < height + -12.f || sqrtf(z2 + pf_fma(dy, dy, x2)) > 35.f)))
                return false;
        }
        constexpr int NC = 5, NS = 7;
        V3 ctrl[NC];
        ctrl[0] = v3(0, 0, 0);
        for (int i = 1; i < NC; ++i)
        {
            const float a = frng.u11(), b = frng.u11(), c = frng.u11();
            V3 off = v3(a * 6.f, b * 2.f, c * 6.f);
            if (i == NC - 1) off = off * 0.6f;
            const float f = (float)i / 4.f;
            ctrl[i] = v3(pf_fma(0.f, f, off.x), pf_fma(height, f, off.y), pf_fma(0.f, f, off.z));
        }
        V3 spline[NS];
        de_casteljau<NC, NS>(ctrl, spline);
        for (int i = 0; i < NS; ++i)
        {
            const V3 p1 = spline[i];
            V3 p2;
            if (i < NS - 1)
            {
                p2 = spline[i + 1];
                if (pos.y < p1.y - 3.f || pos.y > p2.y + 3.f) continue;
            }
            else
            {
                const V3 n = normalize3(p1 - spline[i - 1]);
                const float l = pf_fma(frng.u01(), 1.5f, 3.f);
                p2 = v3(pf_fma(n.x, l, p1.x), pf_fma(n.y, l, p1.y), pf_fma(n.z, l, p1.z));
            }
            float ratio, dist;
            const bool inRatio = line_params(pos, p1, p2, &ratio, &dist);
            float radius;
            uint8_t block;
            if (i < NS - 1)
            {
                const float t = ((float)i + fminf(fmaxf(ratio, 0.f), 1.f)) / (float)(NS - 1);
                const float x = t - 0.5f;
                radius = pf_fma((4.f * x), x, 1.5f) * 1.2f;
                block = B_MUSHROOM_STEM;
            }
            else
            {
                radius = pf_fma(frng.u01(), 7.f, 12.f) * mixf(0.8f, 1.2f, (height - 33.f) / 40.f);
                block = (dist < radius - 1.8f && ratio < 0.5f && scale < 1.4f) ? B_MUSHROOM_UNDERSIDE : B_PURPLE_MUSHROOM_CAP;
            }
            if ((inRatio && dist <= radius) || (i < NS - 1 && ratio < 0.f && len3(p1 - pos) < radius) ||
                (i < NS - 2 && ratio > 1.f && len3(p2 - pos) < radius))
            {
                *out = block;
                return true;
            }
        }
        return false;
    }
    case F_RAFFLESIA:
    {
        if (pos.y > 10.f || len3(pos) > 15.f) return false;
        // the reference build fuses the scaling of y into what follows (its SASS: FFMA y, 0.8, -1 and FFMA y, 0.8, -3.2);
        // x and z are scaled by a plain multiply
        const float posY = pos.y;
        pos = pos * 0.8f;
        V3 c = pos;
        c.y = pf_fma(posY, 0.8f, -1.f);
        c.y = c.y * 1.4f;
        // the three sphere SDFs share x*x and z*z (computed once, rounded); only the y term is fused
        const float cx2 = c.x * c.x, cz2 = c.z * c.z;
        if (sqrtf(cz2 + pf_fma(c.y, c.y, cx2)) - 1.f < 0.f) { *out = B_RAFFLESIA_SPIKES; return true; }
        const float y1 = c.y - 1.f, y2 = c.y - 1.8f;
        float sdf = fabsf(sqrtf(cz2 + pf_fma(y1, y1, cx2)) - 2.0f) - 0.8f;
        const float hole = sqrtf(cz2 + pf_fma(y2, y2, cx2)) - 1.8f;
        sdf = fmaxf(sdf, -hole);
        if (sdf < 0.f) { *out = c.y > 1.f ? B_RAFFLESIA_CENTER : B_RAFFLESIA_STEM; return true; }
        // startAngle + i * (2 pi / 5): the reference build folds the constant and fuses the product u * 2 pi into the sum for
        // i >= 1 (FFMA u, 2pi, 1.2566371 ...), petal 0 uses the rounded product
        const float u = frng.u01();
        for (int i = 0; i < 5; ++i)
        {
            const float angle = i == 0 ? u * kTwoPi : pf_fma(u, kTwoPi, ((float)i * kTwoPi) * 0.2f);
            float s, co;
            dm_sincosf(-angle, &s, &co);
            V3 pp = v3(pf_fma(pos.x, co, pos.z * s), pf_fma(posY, 0.8f, -3.2f), pf_fma(pos.z, co, -(pos.x * s)));
            pp.y = pp.y - (float)(i % 2) * 0.53f;
            pp.y = pf_fma(fminf(fmaxf((fabsf(pp.x - 3.f) - 1.5f) / 1.5f, 0.f), 1.f), 1.3f, pp.y);
            pp.x = pp.x - 3.8f;
            pp.z = pp.z * 1.2f;
            // sdCappedCylinder(pp, 2.5, 0.5)
            const float dx = fabsf(len2(pp.x, pp.z)) - 2.5f, dy = fabsf(pp.y) - 0.5f;
            const float mx = fmaxf(dx, 0.f), my = fmaxf(dy, 0.f);
            const float sd = fminf(fmaxf(dx, dy), 0.0f) + sqrtf(pf_fma(mx, mx, my * my));
            if (sd < 0.f) { *out = B_RAFFLESIA_PETAL; return true; }
        }
        return false;
    }
    case F_LARGE_JUNGLE_TREE:
    {
        const float height = pf_fma(frng.u01(), 10.f, 18.f);
        if (pos.y > height + 6.f || len2(pos.x, pos.z) > 15.f) return false;
        const int tx = (int)floorf(pos.x), tz = (int)floorf(pos.z);
        if (in_range_f(pos.y, 0.f, height) && tx >= 0 && tx <= 1 && tz >= 0 && tz <= 1) { *out = B_JUNGLE_WOOD; return true; }
        pos = pos - v3(0.5f, 0.f, 0.5f);
        V3 lp = pos;
        lp.y = lp.y - (height - 2.f);
        if (jungle_leaves(lp, 4.f, 4.f, 7.f, frng.u01())) { *out = brng.u01() < 0.5f ? B_JUNGLE_LEAVES_FRUITS : B_JUNGLE_LEAVES_PLAIN; return true; }
        const float numBranches = pf_fma(frng.u01(), 2.5f, 0.5f);
        float branchHeight = height;
        for (int i = 0; (float)i < numBranches; ++i)
        {
            branchHeight = pf_fma(-pf_fma(frng.u01(), 3.f, 8.f), height / 30.f, branchHeight);
            const float angle = kTwoPi * frng.u01();
            const V3 bs = v3(0.f, branchHeight, 0.f);
            V3 be = v3(0, 0, 0);
            dm_sincosf(-angle, &be.z, &be.x);
            {
                const float s = pf_fma(frng.u01(), 1.5f, 3.f);
                be = v3(pf_fma(be.x, s, bs.x), pf_fma(be.y, s, bs.y), pf_fma(be.z, s, bs.z));
            }
            be.y = be.y + pf_fma(frng.u01(), 1.5f, 1.f);
            float ratio, dist;
            const bool inRatio = line_params(pos, bs, be, &ratio, &dist);
            const float radius = pf_fma(ratio, -0.4f, 1.2f);
            if (inRatio && dist < radius) { *out = B_JUNGLE_WOOD; return true; }
            lp = (pos - be) + v3(0.f, 0.2f, 0.f);
            if (jungle_leaves(lp, 2.f, 2.5f, 3.5f, frng.u01())) { *out = brng.u01() < 0.25f ? B_JUNGLE_LEAVES_FRUITS : B_JUNGLE_LEAVES_PLAIN; return true; }
        }
        return false;
    }
    case F_SMALL_JUNGLE_TREE:
    {
        const float height = pf_fma(frng.u01(), 4.f, 8.f);
        const float maxDist = pos.y < height - 2.f ? 2.f : 8.f;
        if (pos.y > height + 4.f || len2(pos.x, pos.z) > maxDist) return false;
        if (in_range_f(pos.y, 0.f, height) && (int)floorf(pos.x) == 0 && (int)floorf(pos.z) == 0) { *out = B_JUNGLE_WOOD; return true; }
        const V3 lp = pos - v3(0.f, height - 1.f, 0.f);
        if (jungle_leaves(lp, 3.f, 2.f, 4.f, frng.u01())) { *out = brng.u01() < 0.25f ? B_JUNGLE_LEAVES_FRUITS : B_JUNGLE_LEAVES_PLAIN; return true; }
        return false;
    }
    case F_TINY_JUNGLE_TREE:
    {
        if (fx + fy + fz > 8) return false;
        const int height = (int)pf_fma(frng.u01(), 2.5f, 0.5f);
        if (fx == 0 && in_range_i(fy, 0, height) && fz == 0) { *out = B_JUNGLE_WOOD; return true; }
        if (abs(fx) + abs(fy - height) + abs(fz) == 1) { *out = B_JUNGLE_LEAVES_PLAIN; return true; }
        return false;
    }
    case F_CACTUS:
    {
        if (abs(fx) > 5 || abs(fz) > 5) return false;
        const int height = (int)pf_fma(frng.u01(), 6.0f, 7.5f);
        if (pos.y > (float)height + 2.f) return false;
        if (fx == 0 && in_range_i(fy, 0, height) && fz == 0) { *out = B_CACTUS; return true; }
        for (int arm = 0; arm < 4; ++arm)
        {
            if (frng.u01() >= 0.35f) continue;
            const int armStart = (int)pf_fma(frng.u01(), (float)(height - 10), 4.f);
            const int armLength = (int)pf_fma(frng.u01(), 1.f, 2.f);
            int armHeight = (int)pf_fma(frng.u01(), 3.f, 3.f);
            armHeight = (height - armStart - 1) < armHeight ? (height - armStart - 1) : armHeight;
            const int dx = kDirVecs2d[arm * 2][0], dz = kDirVecs2d[arm * 2][1];
            const int p1[3] = {0, armStart, 0}, p2[3] = {dx * armLength, armStart, dz * armLength}, p3[3] = {p2[0], armStart + armHeight, p2[2]};
            auto inBox = [&](const int* a, const int* b) {
                const int lo[3] = {a[0] < b[0] ? a[0] : b[0], a[1] < b[1] ? a[1] : b[1], a[2] < b[2] ? a[2] : b[2]};
                const int hi[3] = {a[0] > b[0] ? a[0] : b[0], a[1] > b[1] ? a[1] : b[1], a[2] > b[2] ? a[2] : b[2]};
                return fx >= lo[0] && fx <= hi[0] && fy >= lo[1] && fy <= hi[1] && fz >= lo[2] && fz <= hi[2];
            };
            if (inBox(p1, p2) || inBox(p2, p3)) { *out = B_CACTUS; return true; }
        }
        return false;
    }
    case F_PALM_TREE:
    {
        if (fy < -2 || fy > 28 || abs(fx) + abs(fz) > 24) return false;
        constexpr int NC = 4, NS = 5;
        V3 minP = v3(0, 0, 0), maxP = v3(0, 0, 0), ctrl[NC], cur = v3(0, 0, 0);
        ctrl[0] = cur;
        for (int i = 1; i < NC; ++i)
        {
            const float walk = pf_fma((float)i / (float)NC, 5.f, 1.f);
            const float a = frng.u11(), b = frng.u01(), c = frng.u11();
            cur = v3(pf_fma(walk, a, cur.x), cur.y + pf_fma(b, 5.f, 3.f), pf_fma(walk, c, cur.z));
            ctrl[i] = cur;
            minP = v3(fminf(minP.x, cur.x), fminf(minP.y, cur.y), fminf(minP.z, cur.z));
            maxP = v3(fmaxf(maxP.x, cur.x), fmaxf(maxP.y, cur.y), fmaxf(maxP.z, cur.z));
        }
        {
            const V3 lo = minP - v3(7, 1, 7), hi = maxP + v3(7, 6, 7);
            const V3 a = v3(fminf(lo.x, hi.x), fminf(lo.y, hi.y), fminf(lo.z, hi.z)), b = v3(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y), fmaxf(lo.z, hi.z));
            if (!(pos.x >= a.x && pos.x <= b.x && pos.y >= a.y && pos.y <= b.y && pos.z >= a.z && pos.z <= b.z)) return false;
        }
        V3 spline[NS];
        de_casteljau<NC, NS>(ctrl, spline);
        const int ttx = (int)floorf(spline[NS - 1].x), tty = (int)floorf(spline[NS - 1].y), ttz = (int)floorf(spline[NS - 1].z);
        const int lx = fx - ttx, ly = fy - tty, lz = fz - ttz;
        float ld = len2((float)lx, (float)lz);
        {
            const float sat = fminf(fmaxf((float)(20 - tty) * 0.05f, 0.f), 1.f);
            ld = ld * pf_fma(frng.u01(), 0.3f, pf_fma(sat, 0.3f, 0.6f));
        }
        if (in_range_i(ly, -1, 0) && ld < 3.9f && (lx == 0 || lz == 0 || abs(lx) == abs(lz)))
        {
            const int lh = ld > 3.f ? -1 : 0;
            if (ly == lh) { *out = B_PALM_LEAVES; return true; }
        }
        for (int i = 0; i < NS - 1; ++i)
        {
            V3 p1 = spline[i], p2 = spline[i + 1];
            const V3 pad = normalize3(p2 - p1) * 0.5f;
            if (i > 0) p1 = p1 - pad;
            if (i + 1 < NS - 1) p2 = p2 + pad;
            if (in_rasterized_line(fx, fy, fz, p1, p2)) { *out = B_PALM_WOOD; return true; }
        }
        return false;
    }
    case F_MEDIUM_CRYSTAL:
    case F_CRYSTAL:
    {
        if (fp.y > 180) return false;
        pos = pos + v3(0.f, 2.f, 0.f);
        pos = pos * pf_fma(frng.u01(), 0.4f, 0.55f);
        if (fp.feature == F_MEDIUM_CRYSTAL) pos = pos * 2.f;
        if ((abs(fx) > abs(fz) ? abs(fx) : abs(fz)) > 25) return false;
        const float a = frng.u11(), b = frng.u01(), c = frng.u11();
        const V3 end = v3(12.f * a, pf_fma(b, 8.f, 18.f), 12.f * c);
        if (pos.y > end.y + 2.f) return false;
        const uint8_t block = random_crystal_block(frng.u01());
        if (in_crystal(pos, v3(0, 0, 0), end, pf_fma(frng.u01(), 1.2f, 4.f))) { *out = block; return true; }
        pos = pos * 0.8f;
        const int numSmall = (int)pf_fma(frng.u01(), 2.f, 4.f);
        float angle = frng.u01() * kTwoPi;
        for (int i = 0; i < numSmall; ++i)
        {
            angle = angle + pf_fma(frng.u01(), kPi, kPiOverTwo);
            V3 s = v3(0, 0, 0);
            dm_sincosf(angle, &s.z, &s.x);
            V3 e = s;
            s = s * 3.f;
            e = e * pf_fma(frng.u01(), 3.f, 6.f);
            e.y = pf_fma(frng.u01(), 5.f, 7.f);
            (void)s;
            if (in_crystal(pos, v3(0, 0, 0), e, pf_fma(frng.u01(), 1.5f, 1.5f))) { *out = block; return true; }
        }
        return false;
    }
    }
    return false;
}

// featurePlacement.hpp:1110-1380
static inline bool place_cave_feature(const CaveFeaturePlacement& cp, int wx, int wy, int wz, uint8_t* out)
{
    const int lh = cp.layerHeight;
    const int fx = wx - cp.x, fy = wy - cp.y, fz = wz - cp.z;
    const int tx = fx, ty = wy - (cp.y + lh), tz = fz;
    const V3 pos = v3((float)fx, (float)fy, (float)fz);
    V3 top = v3((float)tx, (float)ty, (float)tz);
    Minstd frng = make_rng4(cp.x, cp.y, cp.z, 398132);
    Minstd brng = make_rng4(wx, wy, wz, 9322743);
    switch (cp.feature)
    {
    case CF_NONE: return false;
    case CF_TEST_GLOWSTONE_PILLAR:
    case CF_TEST_SHROOMLIGHT_PILLAR:
        if (fx == 0 && fz == 0 && in_range_i(fy, 0, lh)) { *out = cp.feature == CF_TEST_GLOWSTONE_PILLAR ? B_GLOWSTONE : B_SHROOMLIGHT; return true; }
        return false;
    case CF_CAVE_VINE:
    {
        if (tx != 0 || tz != 0) return false;
        int height = (int)pf_fma(frng.u01(), 12.f, 3.f);
        height = height < lh ? height : lh;
        if (!in_range_i(ty, -height, 0)) return false;
        const bool glowing = brng.u01() < 0.2f;
        if (ty == -height) *out = glowing ? B_CAVE_VINES_GLOW_END : B_CAVE_VINES_END;
        else *out = glowing ? B_CAVE_VINES_GLOW_MAIN : B_CAVE_VINES_MAIN;
        return true;
    }
    case CF_GLOWSTONE_CLUSTER:
    {
        top.y = top.y * 1.35f;
        top = top * pf_fma(frng.u01(), 0.5f, 1.f);
        const float r = len3(top);
        if (r > 6.f) return false;
        const float angle = dm_atan2f(pos.z, pos.x);
        const float maxR = pf_fma(simplex2<true>(angle * 1.5f, (float)wy * 1.5f), 2.f, 3.5f);
        if (r < maxR) { *out = B_GLOWSTONE; return true; }
        return false;
    }
    case CF_STORMLIGHT_SPHERE:
    case CF_CEILING_STORMLIGHT_SPHERE:
    {
        const float radius = pf_fma(frng.u01(), 4.f, 3.5f);
        const float dist = cp.feature == CF_STORMLIGHT_SPHERE ? len3(pos) : len3(top);
        if (dist > radius) return false;
        const float rr = dist / radius;
        const float chance = ss_t((rr + -0.4f) / (0.2f - 0.4f));
        if (brng.u01() < chance) *out = B_GLOWSTONE;
        else *out = random_crystal_block(frng.u01());
        return true;
    }
    case CF_CRYSTAL_PILLAR:
    {
        if (pos.y < -8.f || top.y > 8.f) return false;
        float dist = len2(pos.x, pos.z);
        if (dist > 7.f) return false;
        float hr = pos.y / (float)lh;
        if (hr < 0.f) { hr = 0.f; dist = len3(pos); }
        else if (hr > 1.f) { hr = 1.f; dist = len3(top); }
        float radius = hr - 0.5f;
        radius = 4.f * pf_fma(2.f * radius, radius, 0.5f);
        if (dist > radius) return false;
        if (dist / radius < 0.4f) *out = B_GLOWSTONE;
        else *out = random_crystal_block(frng.u01());
        return true;
    }
    case CF_WARPED_FUNGUS:
    {
        if (abs(fx) + abs(fz) > 6) return false;
        const int height = (int)pf_fma(frng.u01(), 3.0f, 2.5f);
        if (fy < -2 || fy > height + 3) return false;
        if (fx == 0 && fz == 0 && in_range_i(fy, 0, height)) { *out = B_WARPED_STEM; return true; }
        const int sh = fy - (height - 1);
        if (in_range_i(sh, 0, 1) && abs(fx) + abs(fz) == 1)
        {
            const float chance = sh == 0 ? 0.2f : 0.5f;
            if (brng.u01() < chance) { *out = B_SHROOMLIGHT; return true; }
        }
        const float capRadius = len2(pos.x, pos.z);
        if (capRadius > 3.7f) return false;
        const int capEnd = height + 1 - (int)(capRadius / 2.5f);
        const float sx = ((float)wx + (float)cp.y) * 3.f, sz = ((float)wz + (float)cp.y) * 3.f;
        const int capStart = (int)((float)capEnd - (4.2f * simplex2<true>(sx, sz)) * fmaxf(capRadius - 2.3f, 0.f));
        if (in_range_i(fy, capStart, capEnd)) { *out = B_WARPED_WART; return true; }
        return false;
    }
    case CF_AMBER_FUNGUS:
    {
        const int m2 = abs(fx) + abs(fz);
        if (m2 > 4) return false;
        const int height = (int)pf_fma(frng.u01(), 4.5f, 4.5f);
        if (fy < -2 || fy > height + 3) return false;
        if (fx == 0 && fz == 0)
        {
            if (in_range_i(fy, 0, height)) { *out = B_AMBER_STEM; return true; }
            else if (fy == height + 1) { *out = B_AMBER_WART; return true; }
        }
        int capStart = height / 2;
        if (simplex2<true>((float)wx, (float)wz) < 0.f) capStart -= 1;
        if (in_range_i(fy, capStart, height))
        {
            const int capDist = (fy - capStart) < (height / 4 + 1) ? 2 : 1;
            if (m2 == capDist)
            {
                const int gx = (wx / 2) * 2, gy = (wy / 2) * 2, gz = (wz / 2) * 2;
                const float cx = (float)gx, cy = (float)gy, cz = (float)gz;
                // rand3From3(gridCorner) * 2
                const int rx = gx + (int)(hash_fract(pf_fma(cz, 402.98f, pf_fma(cx, 238.68f, cy * 491.28f))) * 2.f);
                const int ry = gy + (int)(hash_fract(pf_fma(cz, 747.42f, pf_fma(cx, 654.37f, cy * 560.45f))) * 2.f);
                const int rz = gz + (int)(hash_fract(pf_fma(cz, 674.81f, pf_fma(cx, 640.88f, cy * 151.81f))) * 2.f);
                if (wx == rx && wy == ry && wz == rz && brng.u01() < 0.65f) *out = B_SHROOMLIGHT;
                else *out = B_AMBER_WART;
                return true;
            }
        }
        return false;
    }
    }
    return false;
}

}  // namespace mmo

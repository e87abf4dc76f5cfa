// cp_traverse.cuh -- closest-hit / any-hit ray queries against the device BVH (sm_100a).
//
// Replaces the reference's two-level kd-tree query for this path:
//   ShapeKDTree::rayIntersect            src/librender/skdtree.cpp:112-142 (closest), :207-226 (shadow)
//   HairKDTree::rayIntersect             src/shapes/hair.cpp:200-237 (per-shape AABB clip of the interval)
//   rayIntersectHavran                   include/mitsuba/render/sahkdtree3.h:178-308 (maxt = t on every accepted hit)
// The kd-trees themselves are not reproduced: one 4-wide BVH spans the segments of all hair shapes.
// What IS reproduced is the interval each primitive test sees: [max(rayMinT, sceneNear, shapeNear),
// min(ray.maxt, sceneFar, shapeFar, best t so far)], including the adaptive epsilon that is applied
// only when ray.mint == Epsilon exactly (and without the inner clamp for shadow rays).
// Closest-hit ties (equal fp32 t) resolve by traversal order, as in the reference.
#pragma once
#include "cp_scene.cuh"

namespace cp {

struct RayHit { float t; uint32_t gv; V3 p; };

#define CP_STACK_SIZE 64

template <bool ANY, bool STATS>
CP_D bool traverse(const SceneDev &S, const V3 &o, const V3 &d, float rayMint, float rayMaxt, RayHit &hit,
                   uint32_t &nodesVisited, uint32_t &primsTested, int &overflow) {
    uint32_t fullTests = 0; (void) fullTests;
    hit.t = CP_INF; hit.gv = 0xffffffffu;
    const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float mint, maxt;
    if (!aabb_ray(S.sceneMin, S.sceneMax, o, d, dRcp, mint, maxt)) return false;
    float rayMinT = rayMint;
    if (rayMinT == kEpsilon) {
        float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
        if (!ANY) m = fmaxf(m, kEpsilon);                 // skdtree.cpp:126-129 vs :214-216
        rayMinT *= m;
    }
    if (rayMinT > mint) mint = rayMinT;
    if (rayMaxt < maxt) maxt = rayMaxt;
    if (!(maxt > mint)) return false;

    // per-shape clipped interval, cached for the last shape seen (hair.cpp:205-209)
    const bool multiShape = S.shapeCount > 1;
    uint32_t cachedShape = 0xffffffffu;
    float sNear = mint, sFar = CP_INF; bool sOk = true;
    float radius = S.shapes[0].radius;

    const BVH4Node *__restrict__ nodes = S.bvh.nodes;
    const uint32_t *__restrict__ prims = S.bvh.prims;
    const float4 *__restrict__ vtx = S.vtx;

    int stack[CP_STACK_SIZE];
    int sp = 0;
    int cur = 0;                                          // root (inner node 0)
    if (S.bvh.nodeCount == 0) return false;
    bool found = false;

    while (true) {
        if (cur >= 0) {
            if (STATS) nodesVisited++;
            const float4 *np = reinterpret_cast<const float4 *>(nodes + cur);
            const float4 lox = __ldg(np + 0), loy = __ldg(np + 1), loz = __ldg(np + 2);
            const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
            const int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 6));
            float tn[4]; int ci[4]; int nh = 0;
#define CP_SLAB(k, LX, LY, LZ, HX, HY, HZ, C) { \
                float x0 = (LX - o.x) * dRcp.x, x1 = (HX - o.x) * dRcp.x; \
                float y0 = (LY - o.y) * dRcp.y, y1 = (HY - o.y) * dRcp.y; \
                float z0 = (LZ - o.z) * dRcp.z, z1 = (HZ - o.z) * dRcp.z; \
                float tnear = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), mint)); \
                float tfar = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), maxt)); \
                if (C != (int) 0x80000000 && tnear <= tfar * 1.0000004f) { tn[nh] = tnear; ci[nh] = C; nh++; } }
            CP_SLAB(0, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, ch.x)
            CP_SLAB(1, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, ch.y)
            CP_SLAB(2, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, ch.z)
            CP_SLAB(3, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, ch.w)
#undef CP_SLAB
            if (nh == 0) {
                if (sp == 0) break;
                cur = stack[--sp];
                continue;
            }
            // sort hits by entry distance (nh <= 4): nearest is visited first, the rest are pushed far-to-near
            if (nh > 1) {
#define CP_CSWAP(a, b) if (tn[a] > tn[b]) { float tt = tn[a]; tn[a] = tn[b]; tn[b] = tt; int cc = ci[a]; ci[a] = ci[b]; ci[b] = cc; }
                if (nh == 2) { CP_CSWAP(0, 1) }
                else if (nh == 3) { CP_CSWAP(0, 1) CP_CSWAP(1, 2) CP_CSWAP(0, 1) }
                else { CP_CSWAP(0, 1) CP_CSWAP(2, 3) CP_CSWAP(0, 2) CP_CSWAP(1, 3) CP_CSWAP(1, 2) }
#undef CP_CSWAP
                for (int i = nh - 1; i >= 1; --i) {
                    if (sp < CP_STACK_SIZE) stack[sp++] = ci[i]; else overflow = 1;
                }
            }
            cur = ci[0];
            continue;
        }
        // ---- leaf
        {
            const uint32_t ref = ~(uint32_t) cur;
            const uint32_t first = ref >> 3, count = (ref & 7u) + 1u;
            for (uint32_t i = 0; i < count; ++i) {
                const uint32_t gv = __ldg(prims + first + i);
                const float4 v1 = __ldg(vtx + gv), v2 = __ldg(vtx + gv + 1);
                if (STATS) primsTested++;
                // fp32 early-out: the ray misses the infinite cylinder if its distance to the axis line exceeds the
                // radius by more than a margin that bounds the fp32 rounding of this estimate (never rejects a hit the
                // FP64 test would accept; skipped for nearly parallel ray/axis pairs where the estimate is ill-conditioned)
                {
                    const V3 a = vtx_pos(v2) - vtx_pos(v1), w = vtx_pos(v1) - o, n = cross(d, a);
                    const float nn = dot(n, n), aa = dot(a, a), wn = dot(w, n);
                    const float sin2 = nn / (aa * dot(d, d));
                    if (sin2 > 4e-4f) {
                        const float wmax = fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fabsf(w.z));
                        const float R = (multiShape ? S.shapes[vtx_shape(v1)].radius : radius) * 1.02f + wmax * (1e-6f + 1e-6f * rsqrtf(sin2));
                        if (wn * wn > R * R * nn) continue;
                    }
                }
                if (STATS) fullTests++;
                const float4 v0 = __ldg(vtx + (gv > 0 ? gv - 1 : 0)), v3 = __ldg(vtx + gv + 2);
                float tmin = mint, tmax = maxt;
                if (multiShape) {
                    const uint32_t sh = vtx_shape(v1);
                    if (sh != cachedShape) {
                        cachedShape = sh;
                        const ShapeDev &sd = S.shapes[sh];
                        radius = sd.radius;
                        sOk = aabb_ray(sd.bmin, sd.bmax, o, d, dRcp, sNear, sFar);
                    }
                    if (!sOk) continue;
                    if (sNear > tmin) tmin = sNear;
                    if (sFar < tmax) tmax = sFar;
                    if (!(tmax > tmin)) continue;
                }
                float t; V3 p;
                if (segment_intersect(v0, v1, v2, v3, radius, o, d, tmin, tmax, t, p)) {
                    hit.t = t; hit.gv = gv; hit.p = p;
                    if (ANY) return true;
                    maxt = t;
                    found = true;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return found;
}

} // namespace cp

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
//
// Execution model: persistent warps.  Incoherent rays inside a hair volume have wildly different traversal lengths
// (ncu on the first, one-thread-per-ray version: 4.3 of 32 lanes active per instruction).  Each warp therefore keeps
// its lanes busy by pulling new rays from a global counter whenever enough lanes have finished, and the loop is split
// into phases the whole warp enters together (node descent / fp32 segment pre-test / FP64 cylinder test), so lanes in
// the same phase execute together instead of serialising against each other.
#pragma once
#include "cp_scene.cuh"

namespace cp {

struct RayHit { float t; uint32_t gv; V3 p; };

#define CP_STACK_SIZE 56
#define CP_EMPTY_CHILD ((int) 0x80000000)
#ifndef CP_REFILL_THRESHOLD
#define CP_REFILL_THRESHOLD 8      // refill a warp once this many lanes are idle
#endif

#ifndef CP_MIN_BLOCKS
#define CP_MIN_BLOCKS 6        // resident CTAs per SM the traversal kernels are compiled for (80 registers; measured best of 4/5/6/8)
#endif

#ifndef CP_DESCENT_MIN_LANES
#define CP_DESCENT_MIN_LANES 8
#endif

struct TraceCounters { unsigned long long nodes, prims, fullTests; };

// IO concept:  bool load(uint32_t i, V3 &o, V3 &d, float &mint, float &maxt)   (false: slot carries no ray)
//              void store(uint32_t i, bool hit, const RayHit &h)
// MESH: the scene also holds triangles (cp_tri.cuh).  They live in the same BVH; a triangle reference skips the fp32 pre-test
// and is tested against the scene-level interval, as in the reference's top-level tree (skdtree.h:293-304).
template <bool ANY, bool STATS, bool MESH, class IO>
CP_D void trace_persistent(const SceneDev &S, IO &io, uint32_t n, uint32_t *__restrict__ rayCounter, TraceCounters &tc, int &overflow) {
    const BVH4Node *__restrict__ nodes = S.bvh.nodes;
    const float4 *__restrict__ leafSeg = S.bvh.leafSeg;
    const float4 *__restrict__ vtx = S.vtx;
    const bool multiShape = S.clipPerShape != 0;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanesBelow = (1u << lane) - 1u;

    uint2 stack[CP_STACK_SIZE];          // (node reference, entry distance bits): one 8-byte local store per push
    // per-lane ray state
    bool idle = true, exhausted = false;
    uint32_t rayIdx = 0;
    V3 o(0.0f), d(0.0f), dRcp(0.0f);
    float mint = 0, maxt = 0, radius = 0;
    int sp = 0, cur = CP_EMPTY_CHILD;
    RayHit hit; hit.t = CP_INF; hit.gv = 0xffffffffu; hit.p = V3(0.0f);
    bool found = false;
    uint32_t cachedShape = 0xffffffffu; float sNear = 0, sFar = 0; bool sOk = true;

    while (true) {
        // ------------------------------------------------------------------ refill idle lanes from the global ray counter
        {
            const unsigned idleMask = __ballot_sync(0xffffffffu, idle && !exhausted);
            const unsigned busyMask = __ballot_sync(0xffffffffu, !idle);
            if (idleMask && (__popc(idleMask) >= CP_REFILL_THRESHOLD || busyMask == 0u)) {
                const int leader = __ffs(idleMask) - 1;
                uint32_t base = 0;
                if ((int) lane == leader) base = atomicAdd(rayCounter, (uint32_t) __popc(idleMask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (idle && !exhausted) {
                    const uint32_t idx = base + __popc(idleMask & lanesBelow);
                    if (idx >= n) exhausted = true;
                    else {
                        rayIdx = idx;
                        float rmin, rmax;
                        found = false; hit.t = CP_INF; hit.gv = 0xffffffffu; hit.p = V3(0.0f);
                        bool alive = io.load(idx, o, d, rmin, rmax) && S.bvh.nodeCount > 0;
                        if (alive) {
                            dRcp = V3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            // scene-level interval (skdtree.cpp:112-142 / :207-226)
                            alive = aabb_ray(S.sceneMin, S.sceneMax, o, d, dRcp, mint, maxt);
                            float rayMinT = rmin;
                            if (rayMinT == kEpsilon) {
                                float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
                                if (!ANY) m = fmaxf(m, kEpsilon);
                                rayMinT *= m;
                            }
                            if (rayMinT > mint) mint = rayMinT;
                            if (rmax < maxt) maxt = rmax;
                            alive = alive && (maxt > mint);
                        }
                        if (alive) { idle = false; sp = 0; cur = 0; cachedShape = 0xffffffffu; radius = S.shapes[0].radius; sOk = true; }
                        else io.store(idx, false, hit);
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, !idle) == 0u && __ballot_sync(0xffffffffu, idle && !exhausted) == 0u) break;
        }

// pop the next node whose entry distance is still inside the (shrinking) interval; an empty stack finishes the ray
#define CP_POP() { \
            bool got_ = false; \
            while (sp > 0) { --sp; const uint2 e_ = stack[sp]; if (ANY || __uint_as_float(e_.y) <= maxt) { cur = (int) e_.x; got_ = true; break; } } \
            if (!got_) { io.store(rayIdx, found, hit); idle = true; cur = CP_EMPTY_CHILD; } }

        // ------------------------------------------------------------------ phase 1: descend inner nodes until this lane holds a leaf
        while (!idle && cur >= 0) {
            if (STATS) tc.nodes++;
            const float4 *np = reinterpret_cast<const float4 *>(nodes + cur);
            const float4 lox = __ldg(np + 0), loy = __ldg(np + 1), loz = __ldg(np + 2);
            const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
            const int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 6));
            // entry distances of the four children (+inf = not entered); everything stays in registers
#define CP_SLAB(T, LX, LY, LZ, HX, HY, HZ, C) { \
                float x0 = (LX - o.x) * dRcp.x, x1 = (HX - o.x) * dRcp.x; \
                float y0 = (LY - o.y) * dRcp.y, y1 = (HY - o.y) * dRcp.y; \
                float z0 = (LZ - o.z) * dRcp.z, z1 = (HZ - o.z) * dRcp.z; \
                float tnear = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), mint)); \
                float tfar = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), maxt)); \
                T = (C != CP_EMPTY_CHILD && tnear <= tfar * 1.0000004f) ? tnear : CP_INF; }
            float t0, t1, t2, t3; int c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
            CP_SLAB(t0, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, ch.x)
            CP_SLAB(t1, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, ch.y)
            CP_SLAB(t2, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, ch.z)
            CP_SLAB(t3, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, ch.w)
#undef CP_SLAB
            // 5-comparator sorting network, ascending entry distance (misses sink to the end)
#define CP_CSWAP(TA, CA, TB, CB) if (TA > TB) { float tt = TA; TA = TB; TB = tt; int cc = CA; CA = CB; CB = cc; }
            CP_CSWAP(t0, c0, t1, c1) CP_CSWAP(t2, c2, t3, c3) CP_CSWAP(t0, c0, t2, c2) CP_CSWAP(t1, c1, t3, c3) CP_CSWAP(t1, c1, t2, c2)
#undef CP_CSWAP
            if (t0 == CP_INF) { CP_POP() continue; }
            // nearest child first, the others are pushed far-to-near
            if (t3 != CP_INF) { if (sp < CP_STACK_SIZE) stack[sp++] = make_uint2((uint32_t) c3, __float_as_uint(t3)); else overflow = 1; }
            if (t2 != CP_INF) { if (sp < CP_STACK_SIZE) stack[sp++] = make_uint2((uint32_t) c2, __float_as_uint(t2)); else overflow = 1; }
            if (t1 != CP_INF) { if (sp < CP_STACK_SIZE) stack[sp++] = make_uint2((uint32_t) c1, __float_as_uint(t1)); else overflow = 1; }
            cur = c0;
            // Leave the descent once most of the warp is already waiting with a leaf: the stragglers resume next round,
            // together with the lanes that will have finished their leaves (keeps both phases reasonably full).
            if (__popc(__activemask()) <= CP_DESCENT_MIN_LANES) break;
        }

        // ------------------------------------------------------------------ phase 2: fp32 pre-test of the leaf's segments (all lanes holding a leaf)
        uint32_t candMask = 0, leafFirst = 0;
        const bool inLeaf = !idle && cur < 0;
        if (inLeaf) {
            const uint32_t ref = ~(uint32_t) cur;
            const uint32_t count = (ref & 7u) + 1u;
            leafFirst = ref >> 3;
            for (uint32_t i = 0; i < count; ++i) {
                const float4 v1 = __ldg(leafSeg + 2 * (size_t) (leafFirst + i)), v2 = __ldg(leafSeg + 2 * (size_t) (leafFirst + i) + 1);
                if (STATS) tc.prims++;
                if (MESH && (vtx_bits(v1) & 8u)) { candMask |= 1u << i; continue; }     // triangle reference
                // Conservative fp32 rejection (never rejects a hit the FP64 test would accept).  With n = d x a the ray and the
                // axis line are closest at ray parameter tc and axis parameter sc; every point of the infinite cylinder the
                // ray can touch lies within R/sin(theta) of tc and within R/(sin(theta)|a|) of sc.  R carries a margin that
                // bounds the fp32 rounding of these estimates; nearly parallel ray/axis pairs skip the test.
                const V3 a = vtx_pos(v2) - vtx_pos(v1), w = vtx_pos(v1) - o, nrm = cross(d, a);
                const float nn = dot(nrm, nrm), aa = dot(a, a), dd = dot(d, d), wn = dot(w, nrm);
                const float sin2 = nn / (aa * dd);
                if (sin2 > 4e-4f) {
                    const float wmax = fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fabsf(w.z));
                    const float rsin = rsqrtf(sin2);
                    const float rad = multiShape ? S.shapes[vtx_shape(v1)].radius : radius;
                    const float R = rad * 1.02f + wmax * (2e-6f + 2e-6f * rsin);
                    if (wn * wn > R * R * nn) continue;                                   // farther than R from the axis line
                    const float inn = 1.0f / nn;
                    const float tcl = dot(cross(w, a), nrm) * inn;                        // ray parameter of closest approach
                    const float slack = 1.01f * R * rsin * rsqrtf(dd) + 1e-5f * fabsf(tcl);
                    if (tcl + slack < mint || tcl - slack > maxt) continue;               // outside the ray interval / behind the best hit
                    if ((vtx_bits(v1) & 6u) == 6u) {                                       // both joints bend mildly: miter overshoot <= 2 r
                        const float scl = dot(cross(w, d), nrm) * inn;                    // axis parameter in [0,1] of closest approach
                        const float sslack = 1.01f * (R * rsin + 2.0f * rad) * rsqrtf(aa) + 1e-5f * (1.0f + fabsf(scl));
                        if (scl + sslack < 0.0f || scl - sslack > 1.0f) continue;         // beyond the segment's ends
                    }
                }
                candMask |= 1u << i;
            }
        }
        // ------------------------------------------------------------------ phase 3: FP64 mitred-cylinder test of the survivors
        while (candMask) {
            const int ci = __ffs(candMask) - 1;
            candMask &= candMask - 1;
            const uint32_t gv = __float_as_uint(__ldg(leafSeg + 2 * (size_t) (leafFirst + ci) + 1).w);
            if (MESH && (gv & CP_TRI_FLAG)) {
                const float4 *ta = S.mesh.triAccel + 3 * (size_t) (gv & ~CP_TRI_FLAG);
                const float4 A = __ldg(ta), B = __ldg(ta + 1), C = __ldg(ta + 2);
                if (STATS) tc.fullTests++;
                float t, u, v;
                if (tri_intersect(A, B, C, o, d, mint, maxt, u, v, t)) {
                    hit.t = t; hit.gv = gv; hit.p = V3(u, v, 0.0f); found = true;       // barycentrics ride in the point slot
                    if (ANY) break;
                    maxt = t;
                }
                continue;
            }
            const float4 v1 = __ldg(vtx + gv), v2 = __ldg(vtx + gv + 1);
            const float4 v0 = __ldg(vtx + (gv > 0 ? gv - 1 : 0)), v3 = __ldg(vtx + gv + 2);
            float tmin = mint, tmax = maxt;
            if (multiShape) {   // per-shape clipped interval, cached for the last shape seen (hair.cpp:205-209)
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
            if (STATS) tc.fullTests++;
            float t; V3 p;
            if (segment_intersect(v0, v1, v2, v3, radius, o, d, tmin, tmax, t, p)) {
                hit.t = t; hit.gv = gv; hit.p = p; found = true;
                if (ANY) break;
                maxt = t;
            }
        }
        if (inLeaf) {
            if (ANY && found) { io.store(rayIdx, true, hit); idle = true; cur = CP_EMPTY_CHILD; }
            else CP_POP()
        }
#undef CP_POP
    }
}

} // namespace cp

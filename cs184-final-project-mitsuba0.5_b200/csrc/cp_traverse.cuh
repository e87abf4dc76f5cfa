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
//
// One launch serves closest-hit and occlusion rays at once (MODE 2): the wavefront hands the closest-hit rays of bounce
// b+1 and the shadow rays of bounce b -- which start at the same hit points -- to the same persistent grid, in one
// coherence order, so a warp walks one part of the tree for both and the tail of one query kind is filled by the other.
//
// On-chip state: the top CP_SMEM_STACK entries of every lane's traversal stack live in shared memory (one 8-byte slot per
// lane and level, conflict-free), only deeper entries go to local memory; the hit point and the ray slot of a lane, which
// are written once per hit / ray and read once per ray, live there as well, and so does everything else a lane touches once per
// leaf / hit / ray rather than once per node -- the ray direction (the slab test only needs the origin and the reciprocal direction),
// the hit distance and primitive, the cached per-shape interval: 18 KB per CTA.  The kernel's rate is nearly proportional to its
// resident warps (a third of the persistent CTAs: 0.53 of the rate), and what caps them is registers: with the cold state out of the
// way the descent loop fits 72 registers (7 CTAs per SM) with the spills the 80-register version had.
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
#define CP_MIN_BLOCKS 7        // resident CTAs per SM the traversal kernels are compiled for (72 registers).  With the cold per-lane state in shared memory:
                               // 326 / 335 / 320 Mpaths/s at 6 / 7 / 8 CTAs against 312 at 6 with that state in registers (hair-curl, round 2, g31)
#endif

#ifndef CP_DESCENT_MIN_LANES
#define CP_DESCENT_MIN_LANES 8
#endif

#ifndef CP_SMEM_STACK
#define CP_SMEM_STACK 12       // stack levels per lane kept in shared memory (12 KB per 128-thread CTA)
#endif
#ifndef CP_HOLD_MIN
#define CP_HOLD_MIN 8          // lanes whose leaf produced FP64 candidates wait until this many lanes hold some (0: test at once)
#endif
#define CP_TRACE_THREADS 128

struct TraceCounters { unsigned long long nodes, prims, fullTests; };

enum : int { TRACE_CLOSEST = 0, TRACE_ANY = 1, TRACE_MIXED = 2 };

// IO concept:  bool load(uint32_t k, V3 &o, V3 &d, float &mint, float &maxt, bool &any, uint32_t &slot)   (false: slot carries no ray;
//                   `any` is only read in TRACE_MIXED; `slot` is handed back to store())
//              void store(uint32_t slot, bool any, bool hit, const RayHit &h)
// MESH: the scene also holds triangles (cp_tri.cuh).  They live in the same BVH; a triangle reference skips the fp32 pre-test
// and is tested against the scene-level interval, as in the reference's top-level tree (skdtree.h:293-304).
// tc[0] counts closest-hit rays, tc[1] occlusion rays (STATS only).
template <int MODE, bool STATS, bool MESH, class IO>
CP_D void trace_persistent(const SceneDev &S, IO &io, uint32_t n, uint32_t *rayCounter, TraceCounters *tc, int &overflow) {
    const BVH4Node *__restrict__ nodes = S.bvh.nodes;
    const float4 *__restrict__ leafSeg = S.bvh.leafSeg;
    const float4 *__restrict__ vtx = S.vtx;
    const bool multiShape = S.clipPerShape != 0;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanesBelow = (1u << lane) - 1u;
    const unsigned tid = threadIdx.x;

#if CP_SMEM_STACK > 0
    __shared__ uint2 s_stack[CP_SMEM_STACK * CP_TRACE_THREADS];
#endif
    __shared__ float s_hitP[3 * CP_TRACE_THREADS];
    __shared__ uint32_t s_slot[CP_TRACE_THREADS];
    // state a lane touches once per leaf / hit / ray, not per node: the ray direction (the slab test only needs the origin and the reciprocal
    // direction), the hit record and the cached per-shape interval -- in shared memory so that the descent loop fits a lower register cap
    __shared__ float s_d[3 * CP_TRACE_THREADS];
    __shared__ float s_hitT[CP_TRACE_THREADS];
    __shared__ uint32_t s_hitGv[CP_TRACE_THREADS];
    __shared__ float s_shapeIv[2 * CP_TRACE_THREADS];
    __shared__ uint32_t s_cachedShape[CP_TRACE_THREADS];      // bit 31: the ray misses the cached shape's box
    uint2 stack[CP_STACK_SIZE - CP_SMEM_STACK];          // (node reference, entry distance bits): one 8-byte store per push
    // per-lane ray state
    bool idle = true, exhausted = false, any = (MODE == TRACE_ANY);
    V3 o(0.0f), dRcp(0.0f);
#define CP_RAY_D() V3(s_d[tid], s_d[CP_TRACE_THREADS + tid], s_d[2 * CP_TRACE_THREADS + tid])
#define CP_HIT_T s_hitT[tid]
#define CP_HIT_GV s_hitGv[tid]
    float mint = 0, maxt = 0, radius = 0;
    int sp = 0, cur = CP_EMPTY_CHILD;
    bool found = false;
    uint32_t candMask = 0, leafFirst = 0;               // FP64 candidates of the leaf this lane holds (bit i: reference leafFirst + i)

#define CP_STORE_RAY() { RayHit h_; h_.t = CP_HIT_T; h_.gv = CP_HIT_GV; h_.p = V3(s_hitP[tid], s_hitP[CP_TRACE_THREADS + tid], s_hitP[2 * CP_TRACE_THREADS + tid]); \
                         io.store(s_slot[tid], any, found, h_); }

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
                        float rmin, rmax; uint32_t slot = idx; bool rayAny = (MODE == TRACE_ANY);
                        found = false; CP_HIT_T = CP_INF; CP_HIT_GV = 0xffffffffu;
                        V3 d(0.0f);
                        s_hitP[tid] = 0.0f; s_hitP[CP_TRACE_THREADS + tid] = 0.0f; s_hitP[2 * CP_TRACE_THREADS + tid] = 0.0f;
                        bool alive = io.load(idx, o, d, rmin, rmax, rayAny, slot) && S.bvh.nodeCount > 0;
                        if (MODE == TRACE_MIXED) any = rayAny;
                        s_slot[tid] = slot;
                        if (alive) {
                            dRcp = V3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            // scene-level interval (skdtree.cpp:112-142 / :207-226)
                            alive = aabb_ray(S.sceneMin, S.sceneMax, o, d, dRcp, mint, maxt);
                            float rayMinT = rmin;
                            if (rayMinT == kEpsilon) {
                                float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
                                if (!any) m = fmaxf(m, kEpsilon);
                                rayMinT *= m;
                            }
                            if (rayMinT > mint) mint = rayMinT;
                            if (rmax < maxt) maxt = rmax;
                            alive = alive && (maxt > mint);
                        }
                        s_d[tid] = d.x; s_d[CP_TRACE_THREADS + tid] = d.y; s_d[2 * CP_TRACE_THREADS + tid] = d.z;
                        if (alive) { idle = false; sp = 0; cur = 0; s_cachedShape[tid] = 0x7fffffffu; radius = S.shapes[0].radius; candMask = 0; }
                        else CP_STORE_RAY()
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, !idle) == 0u && __ballot_sync(0xffffffffu, idle && !exhausted) == 0u) break;
        }

#if CP_SMEM_STACK > 0
#define CP_STACK_AT(I) (*((I) < CP_SMEM_STACK ? &s_stack[(I) * CP_TRACE_THREADS + tid] : &stack[(I) - CP_SMEM_STACK]))
#else
#define CP_STACK_AT(I) (stack[(I)])
#endif
#define CP_PUSH(C, T) { if (sp < CP_STACK_SIZE) { CP_STACK_AT(sp) = make_uint2((uint32_t) (C), __float_as_uint(T)); ++sp; } else overflow = 1; }
// pop the next node whose entry distance is still inside the (shrinking) interval; an empty stack finishes the ray
#define CP_POP() { \
            bool got_ = false; \
            while (sp > 0) { --sp; const uint2 e_ = CP_STACK_AT(sp); if (any || __uint_as_float(e_.y) <= maxt) { cur = (int) e_.x; got_ = true; break; } } \
            if (!got_) { CP_STORE_RAY() idle = true; cur = CP_EMPTY_CHILD; } }

        // ------------------------------------------------------------------ phase 1: descend inner nodes until this lane holds a leaf
        while (!idle && cur >= 0) {
            if (STATS) tc[any ? 1 : 0].nodes++;
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
            if (t3 != CP_INF) CP_PUSH(c3, t3)
            if (t2 != CP_INF) CP_PUSH(c2, t2)
            if (t1 != CP_INF) CP_PUSH(c1, t1)
            cur = c0;
            // Leave the descent once most of the warp is already waiting with a leaf: the stragglers resume next round,
            // together with the lanes that will have finished their leaves (keeps both phases reasonably full).
            if (__popc(__activemask()) <= CP_DESCENT_MIN_LANES) break;
        }

        // ------------------------------------------------------------------ phase 2: fp32 pre-test of the leaf's segments (all lanes that just reached a leaf)
        const bool inLeaf = !idle && cur < 0;
        if (inLeaf && candMask == 0u) {
            const uint32_t ref = ~(uint32_t) cur;
            const uint32_t count = (ref & 7u) + 1u;
            leafFirst = ref >> 3;
            const V3 d = CP_RAY_D();
            for (uint32_t i = 0; i < count; ++i) {
                const float4 v1 = __ldg(leafSeg + 2 * (size_t) (leafFirst + i)), v2 = __ldg(leafSeg + 2 * (size_t) (leafFirst + i) + 1);
                if (STATS) tc[any ? 1 : 0].prims++;
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
            if (candMask == 0u) CP_POP()
        }
        // ------------------------------------------------------------------ phase 3: FP64 mitred-cylinder test of the survivors
        // The exact test is ~10x the instructions of a pre-test and only one leaf in twenty produces a candidate: run it when the
        // lanes holding candidates are many (they wait, the others keep descending) or when nobody else can make progress.
        {
            const unsigned holdMask = __ballot_sync(0xffffffffu, candMask != 0u);
            if (holdMask == 0u) continue;
#if CP_HOLD_MIN > 1
            const unsigned walkMask = __ballot_sync(0xffffffffu, !idle && candMask == 0u);
            if (__popc(holdMask) < CP_HOLD_MIN && walkMask != 0u) continue;
#endif
        }
        if (candMask) {
            bool done = false;
            const V3 d = CP_RAY_D();
            while (candMask) {
                const int ci = __ffs(candMask) - 1;
                candMask &= candMask - 1;
                const uint32_t gv = __float_as_uint(__ldg(leafSeg + 2 * (size_t) (leafFirst + ci) + 1).w);
                if (MESH && (gv & CP_TRI_FLAG)) {
                    const bool isRect = (gv & CP_RECT_FLAG) != 0u;
                    const float4 *ta = isRect ? S.mesh.rects + CP_RECT_STRIDE * (size_t) (gv & CP_PRIM_MASK) : S.mesh.triAccel + 3 * (size_t) (gv & CP_PRIM_MASK);
                    const float4 A = __ldg(ta), B = __ldg(ta + 1), C = __ldg(ta + 2);
                    if (STATS) tc[any ? 1 : 0].fullTests++;
                    float t, u, v;
                    if (isRect ? rect_intersect(A, B, C, o, d, mint, maxt, u, v, t) : tri_intersect(A, B, C, o, d, mint, maxt, u, v, t)) {
                        CP_HIT_T = t; CP_HIT_GV = gv; found = true;       // barycentrics (rectangle: local x, y) ride in the point slot
                        s_hitP[tid] = u; s_hitP[CP_TRACE_THREADS + tid] = v; s_hitP[2 * CP_TRACE_THREADS + tid] = 0.0f;
                        if (any) { done = true; break; }
                        maxt = t;
                    }
                    continue;
                }
                const float4 v1 = __ldg(vtx + gv), v2 = __ldg(vtx + gv + 1);
                const float4 v0 = __ldg(vtx + (gv > 0 ? gv - 1 : 0)), v3 = __ldg(vtx + gv + 2);
                float tmin = mint, tmax = maxt;
                if (multiShape) {   // per-shape clipped interval, cached for the last shape seen (hair.cpp:205-209)
                    const uint32_t sh = vtx_shape(v1);
                    float sNear, sFar; bool sOk;
                    if (sh != (s_cachedShape[tid] & 0x7fffffffu)) {
                        const ShapeDev &sd = S.shapes[sh];
                        radius = sd.radius;
                        sOk = aabb_ray(sd.bmin, sd.bmax, o, d, dRcp, sNear, sFar);
                        s_cachedShape[tid] = sh | (sOk ? 0u : 0x80000000u); s_shapeIv[tid] = sNear; s_shapeIv[CP_TRACE_THREADS + tid] = sFar;
                    } else { sOk = !(s_cachedShape[tid] & 0x80000000u); sNear = s_shapeIv[tid]; sFar = s_shapeIv[CP_TRACE_THREADS + tid]; }
                    if (!sOk) continue;
                    if (sNear > tmin) tmin = sNear;
                    if (sFar < tmax) tmax = sFar;
                    if (!(tmax > tmin)) continue;
                }
                if (STATS) tc[any ? 1 : 0].fullTests++;
                float t; V3 p;
                if (segment_intersect(v0, v1, v2, v3, radius, o, d, tmin, tmax, t, p)) {
                    CP_HIT_T = t; CP_HIT_GV = gv; found = true;
                    s_hitP[tid] = p.x; s_hitP[CP_TRACE_THREADS + tid] = p.y; s_hitP[2 * CP_TRACE_THREADS + tid] = p.z;
                    if (any) { done = true; break; }
                    maxt = t;
                }
            }
            candMask = 0u;
            if (done) { CP_STORE_RAY() idle = true; cur = CP_EMPTY_CHILD; }
            else CP_POP()
        }
#undef CP_POP
#undef CP_PUSH
#undef CP_STACK_AT
#undef CP_STORE_RAY
#undef CP_RAY_D
#undef CP_HIT_T
#undef CP_HIT_GV
    }
}

} // namespace cp

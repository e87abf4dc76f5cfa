// cp_tri.cuh -- triangle meshes next to the hair fibers (sm_100a).
//
// Replaces (reference file:line):
//   TriAccel::load / rayIntersect             include/mitsuba/render/triaccel.h:61-97, 99-158 (Wald's projection test)
//   ShapeKDTree::intersect, triangle branch   include/mitsuba/render/skdtree.h:293-304 (closest), :330-338 (shadow)
//   fillIntersectionRecord<true>, mesh branch include/mitsuba/render/skdtree.h:346-427 (barycentric hit point, interpolated normal)
//   Triangle::getAABB                         include/mitsuba/core/triangle.h:40-45
//
// HBM layout: one 48-byte record per triangle of the whole scene (3 x float4, the reference's TriAccel field for field):
//   a = (k, n_u, n_v, n_d)   b = (a_u, a_v, b_nu, b_nv)   c = (c_nu, c_nv, shapeIndex, primIndex)
// plus the flattened mesh arrays: float4 positions, float4 normals (optional per shape), uint32 indices (global vertex ids).
// A triangle primitive id is CP_TRI_FLAG | global triangle index; the barycentrics (u, v) travel in the hit record's point slot.
//
// The test is written with explicit round-to-nearest intrinsics: it is evaluated inside the traversal translation unit, which
// keeps FMA contraction for the slab and cylinder arithmetic, but the reference's x86 build has no FMA and t / u / v decide
// hits bit for bit.
#pragma once
#include "cp_common.cuh"
#include "cp_hair.cuh"

namespace cp {

#define CP_TRI_FLAG 0x80000000u

struct MeshDev {
    const float4 *triAccel;     // 3 x float4 per triangle
    const float4 *pos;          // xyz, -
    const float4 *nrm;          // xyz, - (zero where the owning shape has no vertex normals)
    const uint32_t *idx;        // 3 per triangle, indices into pos / nrm
    uint32_t triCount, vertCount;
};

// triaccel.h:99-158
CP_D bool tri_intersect(const float4 &A, const float4 &B, const float4 &C, const V3 &ro, const V3 &rd, float mint, float maxt,
                        float &u, float &v, float &t) {
    const uint32_t k = __float_as_uint(A.x);
    float o_u, o_v, o_k, d_u, d_v, d_k;
    if (k == 0u) { o_u = ro.y; o_v = ro.z; o_k = ro.x; d_u = rd.y; d_v = rd.z; d_k = rd.x; }
    else if (k == 1u) { o_u = ro.z; o_v = ro.x; o_k = ro.y; d_u = rd.z; d_v = rd.x; d_k = rd.y; }
    else if (k == 2u) { o_u = ro.x; o_v = ro.y; o_k = ro.z; d_u = rd.x; d_v = rd.y; d_k = rd.z; }
    else return false;
    const float num = __fsub_rn(__fsub_rn(__fsub_rn(A.w, __fmul_rn(o_u, A.y)), __fmul_rn(o_v, A.z)), o_k);
    const float den = __fadd_rn(__fadd_rn(__fmul_rn(d_u, A.y), __fmul_rn(d_v, A.z)), d_k);
    t = __fdiv_rn(num, den);
    if (t < mint || t > maxt) return false;
    const float hu = __fsub_rn(__fadd_rn(o_u, __fmul_rn(t, d_u)), B.x);
    const float hv = __fsub_rn(__fadd_rn(o_v, __fmul_rn(t, d_v)), B.y);
    u = __fadd_rn(__fmul_rn(hv, B.z), __fmul_rn(hu, B.w));
    v = __fadd_rn(__fmul_rn(hu, C.x), __fmul_rn(hv, C.y));
    return u >= 0 && v >= 0 && __fadd_rn(u, v) <= 1.0f;
}

// skdtree.h:346-427 with BarycentricPos = true (skdtree.cpp:136), then computeShadingFrame + wi = toLocal(-ray.d).
// Only called from translation units built with -fmad=false.
CP_D uint32_t fill_intersection_mesh(const MeshDev &M, const ShapeDev *__restrict__ shapes, uint32_t tri, float u, float v, const V3 &rd, HitRecord &rec) {
    const uint32_t shape = __float_as_uint(__ldg(M.triAccel + 3 * (size_t) tri + 2).z);
    const V3 b(1 - u - v, u, v);
    const uint32_t idx0 = __ldg(M.idx + 3 * (size_t) tri), idx1 = __ldg(M.idx + 3 * (size_t) tri + 1), idx2 = __ldg(M.idx + 3 * (size_t) tri + 2);
    const float4 q0 = __ldg(M.pos + idx0), q1 = __ldg(M.pos + idx1), q2 = __ldg(M.pos + idx2);
    const V3 p0(q0.x, q0.y, q0.z), p1(q1.x, q1.y, q1.z), p2(q2.x, q2.y, q2.z);
    rec.p = p0 * b.x + p1 * b.y + p2 * b.z;
    const V3 side1 = p1 - p0, side2 = p2 - p0;
    V3 faceNormal = cross(side1, side2);
    const float len = length(faceNormal);
    if (!isZero(faceNormal)) faceNormal = faceNormal / len;
    if (shapes[shape].hasNormals) {
        const float4 m0 = __ldg(M.nrm + idx0), m1 = __ldg(M.nrm + idx1), m2 = __ldg(M.nrm + idx2);
        rec.sh.n = normalize(V3(m0.x, m0.y, m0.z) * b.x + V3(m1.x, m1.y, m1.z) * b.y + V3(m2.x, m2.y, m2.z) * b.z);
        if (dot(faceNormal, rec.sh.n) < 0) faceNormal = -faceNormal;   // geometric and shading normals face the same way
    } else rec.sh.n = faceNormal;
    rec.geoN = faceNormal;
    rec.sh.s = normalize(side1 - rec.sh.n * dot(rec.sh.n, side1));      // computeShadingFrame(n, dpdu = side1), util.cpp:603-608
    rec.sh.t = cross(rec.sh.n, rec.sh.s);
    rec.wi = rec.sh.toLocal(-rd);
    return shape;
}

} // namespace cp

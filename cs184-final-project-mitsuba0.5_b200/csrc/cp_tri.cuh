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
#define CP_RECT_FLAG 0x40000000u      // with CP_TRI_FLAG: a `rectangle` shape (one analytic primitive of the top-level tree), index into MeshDev::rects
#define CP_PRIM_MASK 0x0fffffffu

#define CP_RECT_STRIDE 8              // float4 per rectangle: rows 0..2 of worldToObject | (dpdu, shape index) | (normal, -) | - | box min | box max

struct MeshDev {
    const float4 *triAccel;     // 3 x float4 per triangle
    const float4 *pos;          // xyz, -
    const float4 *nrm;          // xyz, - (zero where the owning shape has no vertex normals)
    const float2 *uv;           // texture coordinates per vertex (zero where the owning shape has none)
    const uint32_t *idx;        // 3 per triangle, indices into pos / nrm / uv
    const float4 *rects;        // CP_RECT_STRIDE x float4 per rectangle (src/shapes/rectangle.cpp)
    uint32_t triCount, vertCount, rectCount;
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

// Rectangle::rayIntersect (src/shapes/rectangle.cpp:127-151) after Transform::transformAffine(Ray) (transform.h:139-146,292-307): the ray goes to
// object space through the three rows of worldToObject, meets z = 0 and must land in [-1, 1]^2.  One IEEE operation per source operation.
CP_D bool rect_intersect(const float4 &R0, const float4 &R1, const float4 &R2, const V3 &ro, const V3 &rd, float mint, float maxt,
                         float &lx, float &ly, float &t) {
#define CP_ROW_P(R) __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R.x, ro.x), __fmul_rn(R.y, ro.y)), __fmul_rn(R.z, ro.z)), R.w)
#define CP_ROW_V(R) __fadd_rn(__fadd_rn(__fmul_rn(R.x, rd.x), __fmul_rn(R.y, rd.y)), __fmul_rn(R.z, rd.z))
    const float oz = CP_ROW_P(R2), dz = CP_ROW_V(R2);
    const float hit = __fdiv_rn(-oz, dz);
    if (!(hit >= mint && hit <= maxt)) return false;
    const float ox = CP_ROW_P(R0), oy = CP_ROW_P(R1), dx = CP_ROW_V(R0), dy = CP_ROW_V(R1);
#undef CP_ROW_P
#undef CP_ROW_V
    const float x = __fadd_rn(ox, __fmul_rn(hit, dx)), y = __fadd_rn(oy, __fmul_rn(hit, dy));      // Ray::operator()(t) = o + t * d
    if (fabsf(x) <= 1 && fabsf(y) <= 1) { t = hit; lx = x; ly = y; return true; }
    return false;
}

// Rectangle::fillIntersectionRecord (rectangle.cpp:158-171), then computeShadingFrame + wi (skdtree.h:426-427).
// Only called from translation units built with -fmad=false.
CP_D uint32_t fill_intersection_rect(const MeshDev &M, uint32_t rect, float lx, float ly, float t, const V3 &ro, const V3 &rd, HitRecord &rec) {
    const float4 du = __ldg(M.rects + CP_RECT_STRIDE * (size_t) rect + 3), nn = __ldg(M.rects + CP_RECT_STRIDE * (size_t) rect + 4);
    const V3 dpdu(du.x, du.y, du.z), n(nn.x, nn.y, nn.z);
    rec.geoN = n; rec.sh.n = n;
    rec.u = 0.5f * (lx + 1); rec.v = 0.5f * (ly + 1);
    rec.p = ro + rd * t;
    rec.sh.s = normalize(dpdu - n * dot(n, dpdu));
    rec.sh.t = cross(n, rec.sh.s);
    rec.wi = rec.sh.toLocal(-rd);
    return __float_as_uint(du.w);
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
    if (shapes[shape].hasUV) {                                           // skdtree.h:399-406
        const float2 t0 = __ldg(M.uv + idx0), t1 = __ldg(M.uv + idx1), t2 = __ldg(M.uv + idx2);
        rec.u = t0.x * b.x + t1.x * b.y + t2.x * b.z; rec.v = t0.y * b.x + t1.y * b.y + t2.y * b.z;
    } else { rec.u = b.y; rec.v = b.z; }
    rec.sh.s = normalize(side1 - rec.sh.n * dot(rec.sh.n, side1));      // computeShadingFrame(n, dpdu = side1), util.cpp:603-608
    rec.sh.t = cross(rec.sh.n, rec.sh.s);
    rec.wi = rec.sh.toLocal(-rd);
    return shape;
}

} // namespace cp

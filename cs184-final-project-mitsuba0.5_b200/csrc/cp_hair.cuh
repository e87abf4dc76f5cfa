// cp_hair.cuh -- HairShape segment geometry on the device (sm_100a).
//
// Replaces (reference file:line):
//   miter helpers                     src/shapes/hair.cpp:551-596
//   HairKDTree::intersect (FP64)      src/shapes/hair.cpp:485-542  (+ solveQuadraticDouble src/libcore/util.cpp:487-525)
//   HairShape::fillIntersectionRecord src/shapes/hair.cpp:825-862 (+ include/mitsuba/render/skdtree.h:426-427)
//   segment bounds getAABB(index)     src/shapes/hair.cpp:246-286, 368-397
//
// HBM layout: all shapes share one vertex array `float4 vtx[]` = (x, y, z, bits) with
//   bits & 1      = vertexStartsFiber[i]           (hair.cpp:649-650, sentinel included)
//   bits >> 8     = shape index
// so one segment test reads vtx[gv-1 .. gv+2] = 64 contiguous bytes.  A primitive is identified by
// `gv`, the global index of the segment's first vertex; the reference's primitive id is the
// shape-local vertex index iv = gv - shape.vertexOffset (hair.cpp:151-155).
#pragma once
#include "cp_common.cuh"

namespace cp {

struct ShapeDev {
    float bmin[3], bmax[3];   // union of segment bounds = HairKDTree::m_aabb (gkdtree.h:997-1002); meshes: union of the triangle boxes
    float radius;
    uint32_t vertexOffset, vertexCount;   // hair: range in SceneDev::vtx; mesh: range in MeshDev::pos / nrm
    int bsdf;
    int kind;                 // 0 = hair, 1 = triangle mesh (cp_tri.cuh)
    uint32_t triOffset, triCount;
    int hasNormals;
    int hasUV;                // mesh: per-vertex texture coordinates present (MeshDev::uv)
};

CP_D uint32_t vtx_bits(const float4 &v) { return __float_as_uint(v.w); }
CP_D bool vtx_starts(const float4 &v) { return vtx_bits(v) & 1u; }
CP_D uint32_t vtx_shape(const float4 &v) { return vtx_bits(v) >> 8; }
CP_D V3 vtx_pos(const float4 &v) { return V3(v.x, v.y, v.z); }

// aabb.h:308-338 -- slab test returning the signed near/far distances
CP_D bool aabb_ray(const float *bmin, const float *bmax, const V3 &o, const V3 &d, const V3 &dRcp, float &nearT, float &farT) {
    nearT = -CP_INF; farT = CP_INF;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float origin = comp(o, i), dir = comp(d, i), minVal = bmin[i], maxVal = bmax[i];
        if (dir == 0) {
            if (origin < minVal || origin > maxVal) return false;
        } else {
            float t1 = (minVal - origin) * comp(dRcp, i), t2 = (maxVal - origin) * comp(dRcp, i);
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
            nearT = fmaxf(t1, nearT);
            farT = fminf(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}

// FP64 arithmetic of the cylinder test, one IEEE operation per source operation: the reference is built for x86-64 without FMA
// (-march=nocona), and a fused multiply-add in the dot products or the discriminant moves the fp64 result by an ulp -- enough to move
// the fp32 hit distance by one ulp for about 5 hits per million (found by the full-size furball ray batch, 1 of 200 278 hits).
// -DCP_FP64_FMA restores the contracted arithmetic (tuning experiments only).
#if defined(__CUDA_ARCH__) && !defined(CP_FP64_FMA)
CP_D double xmul(double a, double b) { return __dmul_rn(a, b); }
CP_D double xadd(double a, double b) { return __dadd_rn(a, b); }
CP_D double xsub(double a, double b) { return __dsub_rn(a, b); }
#else
CP_HD double xmul(double a, double b) { return a * b; }
CP_HD double xadd(double a, double b) { return a + b; }
CP_HD double xsub(double a, double b) { return a - b; }
#endif
CP_HD D3 xadd(D3 a, D3 b) { return D3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
CP_HD D3 xsub(D3 a, D3 b) { return D3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
CP_HD D3 xscale(D3 a, double s) { return D3(xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)); }
CP_HD double xdot(D3 a, D3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
CP_HD D3 xnormalize(D3 a) { return xscale(a, 1.0 / sqrt(xdot(a, a))); }

// hair.cpp:485-542.  v0..v3 = vtx[gv-1..gv+2] (v0/v3 are only read when the neighbour segment exists).
// mint/maxt are the shape-clipped global interval with maxt already shrunk to the current best hit.
CP_D bool segment_intersect(const float4 &v0, const float4 &v1, const float4 &v2, const float4 &v3, float radius,
                            const V3 &ro, const V3 &rd, float mint, float maxt, float &tOut, V3 &pOut) {
    const D3 p1(vtx_pos(v1)), p2(vtx_pos(v2));
    const D3 axis = xnormalize(xsub(p2, p1));
    const D3 rayO(ro), rayD(rd);
    const D3 relOrigin = xsub(rayO, p1);
    const D3 projOrigin = xsub(relOrigin, xscale(axis, xdot(axis, relOrigin)));
    const D3 projDirection = xsub(rayD, xscale(axis, xdot(axis, rayD)));
    const double A = xdot(projDirection, projDirection);
    const double B = xmul(2.0, xdot(projOrigin, projDirection));
    const double C = xsub(xdot(projOrigin, projOrigin), (double) (radius * radius));
    double nearT, farT;
    // solveQuadraticDouble
    if (A == 0) {
        if (B != 0) nearT = farT = -C / B; else return false;
    } else {
        double discrim = xsub(xmul(B, B), xmul(xmul(4.0, A), C));
        if (discrim < 0) return false;
        double sqrtDiscrim = sqrt(discrim), temp;
        if (B < 0) temp = xmul(-0.5, xsub(B, sqrtDiscrim)); else temp = xmul(-0.5, xadd(B, sqrtDiscrim));
        nearT = temp / A; farT = C / temp;
        if (nearT > farT) { double t = nearT; nearT = farT; farT = t; }
    }
    if (!(nearT <= (double) maxt && farT >= (double) mint)) return false;   // NaN-aware
    // miter planes (hair.cpp:584-596): previous segment exists iff !startsFiber[iv], next iff !startsFiber[iv+2]
    D3 n1 = axis, n2 = axis;
    if (!vtx_starts(v1)) n1 = xnormalize(xadd(xnormalize(xsub(p1, D3(vtx_pos(v0)))), axis));
    if (!vtx_starts(v3)) n2 = xnormalize(xadd(axis, xnormalize(xsub(D3(vtx_pos(v3)), p2))));
    const D3 pointNear = xadd(rayO, xscale(rayD, nearT)), pointFar = xadd(rayO, xscale(rayD, farT));
    if (xdot(xsub(pointNear, p1), n1) >= 0 && xdot(xsub(pointNear, p2), n2) <= 0 && nearT >= (double) mint) {
        tOut = (float) nearT;
        pOut = V3((float) pointNear.x, (float) pointNear.y, (float) pointNear.z);   // Point(rayO + rayD * nearT), hair.cpp:524
    } else if (xdot(xsub(pointFar, p1), n1) >= 0 && xdot(xsub(pointFar, p2), n2) <= 0) {
        if (farT > (double) maxt) return false;
        tOut = (float) farT;                                                // rays starting inside a fiber exit through its far wall
        pOut = V3((float) pointFar.x, (float) pointFar.y, (float) pointFar.z);
    } else return false;
    return true;
}

struct HitRecord { V3 p; Frame sh; V3 geoN; V3 wi; float u, v; };      // u, v = its.uv (meshes and rectangles; the hair BSDFs carry no textures)

// hair.cpp:825-862 then computeShadingFrame (util.cpp:603-608) and wi = toLocal(-ray.d)
CP_D void fill_intersection(const float4 &v1, const float4 &v2, float radius, const V3 &pHit, const V3 &rd, HitRecord &rec) {
    const V3 axis = normalize(vtx_pos(v2) - vtx_pos(v1));
    const V3 rel = pHit - vtx_pos(v1);
    V3 n = normalize(rel - dot(axis, rel) * axis);
    V3 gt = cross(n, axis);
    // its.geoFrame.toLocal(rel) = (dot(rel,s), dot(rel,t), dot(rel,n))
    float ly = dot(rel, gt), lz = dot(rel, n);
    rec.p = pHit + n * (radius - sqrtf(ly * ly + lz * lz));
    rec.geoN = n;
    rec.sh.n = n;
    rec.sh.s = normalize(axis - n * dot(n, axis));
    rec.sh.t = cross(n, rec.sh.s);
    rec.wi = rec.sh.toLocal(-rd);
    rec.u = 0.0f; rec.v = 0.0f;
}

// ----- segment bounds (build only): tight box of the two miter-cut end ellipses, radius*(1-Epsilon)
CP_D bool cyl_plane_ellipse(V3 planePt, V3 planeNrml, V3 cylPt, V3 cylD, float radius, V3 &center, V3 &ax0, V3 &ax1) {
    if (fabsf(dot(planeNrml, cylD)) < kEpsilon) return false;
    V3 B, A = cylD - dot(cylD, planeNrml) * planeNrml;
    float len = length(A);
    bool same = planeNrml.x == cylD.x && planeNrml.y == cylD.y && planeNrml.z == cylD.z;
    if (len > kEpsilon && !same) { A = A / len; B = cross(planeNrml, A); }
    else coordinateSystem(planeNrml, A, B);
    V3 delta = planePt - cylPt, deltaProj = delta - cylD * dot(delta, cylD);
    float aDotD = dot(A, cylD), bDotD = dot(B, cylD);
    float c0 = 1 - aDotD * aDotD, c1 = 1 - bDotD * bDotD;
    float c2 = 2 * dot(A, deltaProj), c3 = 2 * dot(B, deltaProj);
    float c4 = dot(delta, deltaProj) - radius * radius;
    float lambda = (c2 * c2 / (4 * c0) + c3 * c3 / (4 * c1) - c4) / (c0 * c1);
    float alpha0 = -c2 / (2 * c0), beta0 = -c3 / (2 * c1);
    center = planePt + alpha0 * A + beta0 * B;
    ax0 = A * sqrtf(c1 * lambda);
    ax1 = B * sqrtf(c0 * lambda);
    return true;
}
// Boxes of the two miter-cut end ellipses (radius*(1-Epsilon)); their union is the reference's getAABB(index).
CP_D void segment_end_boxes(const float4 &v0, const float4 &v1, const float4 &v2, const float4 &v3, float radius,
                            float *minA, float *maxA, float *minB, float *maxB) {
    const V3 p1 = vtx_pos(v1), p2 = vtx_pos(v2);
    const V3 tangent = normalize(p2 - p1);
    V3 n1 = tangent, n2 = tangent;
    if (!vtx_starts(v1)) n1 = normalize(normalize(p1 - vtx_pos(v0)) + tangent);
    if (!vtx_starts(v3)) n2 = normalize(tangent + normalize(vtx_pos(v3) - p2));
    const float r = radius * (1 - kEpsilon);
#pragma unroll
    for (int end = 0; end < 2; ++end) {
        V3 c, a0, a1;
        const V3 pt = end ? p2 : p1, nn = end ? n2 : n1;
        if (!cyl_plane_ellipse(pt, nn, pt, tangent, r, c, a0, a1)) {
            // degenerate miter (never for loader output): fall back to a conservative bound
            c = pt; a0 = V3(radius * 2, 0, 0); a1 = V3(0, radius * 2, radius * 2);
        }
        const float rx = sqrtf(a0.x * a0.x + a1.x * a1.x), ry = sqrtf(a0.y * a0.y + a1.y * a1.y), rz = sqrtf(a0.z * a0.z + a1.z * a1.z);
        float *mn = end ? minB : minA, *mx = end ? maxB : maxA;
        mn[0] = c.x - rx; mx[0] = c.x + rx; mn[1] = c.y - ry; mx[1] = c.y + ry; mn[2] = c.z - rz; mx[2] = c.z + rz;
    }
}
CP_D void segment_bounds(const float4 &v0, const float4 &v1, const float4 &v2, const float4 &v3, float radius, float *bmin, float *bmax) {
    float minA[3], maxA[3], minB[3], maxB[3];
    segment_end_boxes(v0, v1, v2, v3, radius, minA, maxA, minB, maxB);
#pragma unroll
    for (int k = 0; k < 3; ++k) { bmin[k] = fminf(minA[k], minB[k]); bmax[k] = fmaxf(maxA[k], maxB[k]); }
}

} // namespace cp

// oracle/ref_shim/ref_geom.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The hair geometry code of the reference, executed as written: oracle/Makefile cuts these pieces of text out of /root/reference at build
// time (oracle/_ref/*.inc, never committed) and this file pastes them into a class that only supplies the members they touch:
//   ref_hair_intersect.inc   src/shapes/hair.cpp:480-596   IntersectionStorage, HairKDTree::intersect (FP64 mitred cylinder), vertex / tangent / miter helpers
//   ref_hair_cylplane.inc    src/shapes/hair.cpp:246-286   intersectCylPlane
//   ref_hair_aabb.inc        src/shapes/hair.cpp:368-397   getAABB(index): the segment bounds
//   ref_hair_record.inc      src/shapes/hair.cpp:825-862   HairShape::fillIntersectionRecord
//   ref_quadratic.inc        src/libcore/util.cpp:487-525  solveQuadraticDouble
//   ref_triaccel.inc         include/mitsuba/render/triaccel.h:37-158  struct TriAccel, load, rayIntersect (Wald's projection test)
//   ref_aabb_ray.inc         include/mitsuba/core/aabb.h:308-338       TAABB::rayIntersect (the slab test that clips rays to the kd-tree boxes)
//   ref_hair_rayintersect.inc  src/shapes/hair.cpp:199-237     HairKDTree::rayIntersect (closest / visibility): the per-shape clip of the interval
//   ref_skd_rayintersect.inc   src/librender/skdtree.cpp:112-142,207-226  ShapeKDTree::rayIntersect (closest / shadow): scene clip + adaptive epsilon
//   ref_kd_enlarge.inc         include/mitsuba/render/gkdtree.h:1219-1220  the two lines that enlarge a finished kd-tree's box
// In these, the Havran traversal itself (rayIntersectHavran) is replaced by a scan over all primitives that updates maxt on every accepted
// hit exactly like the leaf loop of sahkdtree3.h:262-290 -- same answers up to the visiting order of equal-t hits.
// (HairKDTree itself derives from the generic kd-tree templates of gkdtree.h / sahkdtree3.h, which need the scheduler and boost and cannot be
// compiled here; the kd-tree is not reproduced by the product anyway.)  Output: oracle/_ref/libref_geom.so.
#include "mitsuba_shim.h"

namespace mitsuba {
// include/mitsuba/core/vector.h / point.h with T = double: same operators as the float types in mitsuba_shim.h
struct Vector3d {
    double x, y, z;
    Vector3d() : x(0), y(0), z(0) {}
    Vector3d(double x, double y, double z) : x(x), y(y), z(z) {}
    explicit Vector3d(const Vector &v) : x((double) v.x), y((double) v.y), z((double) v.z) {}
    Vector3d operator+(const Vector3d &v) const { return Vector3d(x + v.x, y + v.y, z + v.z); }
    Vector3d operator-(const Vector3d &v) const { return Vector3d(x - v.x, y - v.y, z - v.z); }
    Vector3d operator*(double f) const { return Vector3d(x * f, y * f, z * f); }
    Vector3d operator/(double f) const { double recip = (double) 1 / f; return Vector3d(x * recip, y * recip, z * recip); }
    double lengthSquared() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(lengthSquared()); }
};
typedef Vector3d Point3d;
inline Vector3d operator*(double f, const Vector3d &v) { return v * f; }
inline double dot(const Vector3d &a, const Vector3d &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3d normalize(const Vector3d &v) { return v / v.length(); }
inline Vector toFloat(const Vector3d &v) { return Vector((Float) v.x, (Float) v.y, (Float) v.z); }
// `Point(rayO + rayD * nearT)` (hair.cpp:524,530): the explicit double -> float conversion of TPoint3
struct PointFromDouble : public Vector { explicit PointFromDouble(const Vector3d &v) : Vector((Float) v.x, (Float) v.y, (Float) v.z) {} };

struct Ray { Point o; Vector d, dRcp; Float mint = 0, maxt = 0, time = 0; };
#define FINLINE inline
#define MM_ALIGN16
struct PointDim { static const int dim = 3; };
struct AABB { Point min, max; AABB() : min(std::numeric_limits<Float>::infinity()), max(-std::numeric_limits<Float>::infinity()) {}
    typedef Ray RayType; typedef PointDim PointType;
#include "ref_aabb_ray.inc"
};
inline bool operator!=(const Vector &a, const Vector &b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

struct GeoFrame { Vector s, t, n; Vector toLocal(const Vector &v) const { return Vector(dot(v, s), dot(v, t), dot(v, n)); } };   // frame.h:71-77
struct HairShape;
struct GeoIntersection { Point p; GeoFrame geoFrame, shFrame; Point2 uv; Vector dpdu, dpdv; const HairShape *shape = nullptr, *instance = nullptr; bool hasUVPartials = false; Float time = 0; };

#include "ref_quadratic.inc"
#include "ref_triaccel.inc"
#include "ref_fresnel.inc"           // coordinateSystem (util.cpp:592-601), used by intersectCylPlane; fresnelDielectricExt comes along unused

class HairKDTree {
public:
    typedef uint32_t IndexType;
    // The pasted text spells the conversion `Point(...)`; inside this class that name means the double -> float constructor above.
    typedef Vector PointF;
    std::vector<Vector> m_vertices; std::vector<bool> m_vertexStartsFiber; std::vector<IndexType> m_segIndex; Float m_radius = 0;
    Float getRadius() const { return m_radius; }
#define Point PointFromDoubleOrFloat
    struct PointFromDoubleOrFloat : public Vector {
        PointFromDoubleOrFloat() {}
        PointFromDoubleOrFloat(const Vector &v) : Vector(v) {}
        explicit PointFromDoubleOrFloat(const Vector3d &v) : Vector((Float) v.x, (Float) v.y, (Float) v.z) {}
    };
#include "ref_hair_cylplane.inc"
#include "ref_hair_aabb.inc"
#include "ref_hair_intersect.inc"
#undef Point
    AABB m_aabb;                                                    // enlarged box of the finished tree (getAABB() of gkdtree.h)
    const AABB &getAABB() const { return m_aabb; }
    // stand-in for the kd traversal: all segments in index order, maxt shrinks on every accepted hit (sahkdtree3.h:262-290)
    template <bool shadowRay> bool rayIntersectHavran(const Ray &ray, Float mint, Float maxt, Float &t, void *temp) const {
        bool found = false;
        for (IndexType iv : m_segIndex) {
            Float primT;
            if (intersect(ray, iv, mint, maxt, primT, temp)) {
                if (shadowRay) return true;
                maxt = t = primT; found = true;
            }
        }
        return found;
    }
#include "ref_hair_rayintersect.inc"
};

#define Intersection GeoIntersection
struct HairShape {
    HairKDTree *m_kdtree = nullptr;
    void fillIntersectionRecord(const Ray &ray, const void *temp, Intersection &its) const;
};
#include "ref_hair_record.inc"
#undef Intersection

// gkdtree.h:50 and :1219-1220 (KDTreeBase::buildInternal): what getAABB() of a finished tree returns
#define MTS_KD_AABB_EPSILON 1e-3f
inline void enlargeKDBox(AABB &aabb) {
    typedef Vector VectorType;
    const Float eps = MTS_KD_AABB_EPSILON;
#include "ref_kd_enlarge.inc"
}

#define MTS_KD_INTERSECTION_TEMP 64
struct SceneIntersection { Float t = 0; int shapeIndex = -1; uint32_t iv = 0; };
static long raysTraced = 0, shadowRaysTraced = 0;
#define Intersection SceneIntersection
class ShapeKDTree {
public:
    AABB m_aabb; std::vector<HairKDTree *> m_shapes;
    bool rayIntersect(const Ray &ray, Intersection &its) const;
    bool rayIntersect(const Ray &ray) const;
    // stand-in for the kd traversal of the scene-level tree: one primitive per hair shape (skdtree.cpp:60-62), tested through
    // ShapeKDTree::intersect -> Shape::rayIntersect (skdtree.h:271-278, hair.cpp:816-823)
    template <bool shadowRay> bool rayIntersectHavran(const Ray &ray, Float mint, Float maxt, Float &t, void *temp) const {
        bool found = false;
        for (size_t i = 0; i < m_shapes.size(); ++i) {
            Float primT;
            HairKDTree::IntersectionStorage st;
            if (shadowRay ? m_shapes[i]->rayIntersect(ray, mint, maxt) : m_shapes[i]->rayIntersect(ray, mint, maxt, primT, &st)) {
                if (shadowRay) return true;
                maxt = t = primT; found = true;
                SceneIntersection *rec = (SceneIntersection *) temp; rec->shapeIndex = (int) i; rec->iv = st.iv;
            }
        }
        return found;
    }
    template <bool bary> void fillIntersectionRecord(const Ray &, const void *temp, Intersection &its) const {
        const SceneIntersection *rec = (const SceneIntersection *) temp; its.shapeIndex = rec->shapeIndex; its.iv = rec->iv;
    }
};
#include "ref_skd_rayintersect.inc"
#undef Intersection
} // namespace mitsuba

using namespace mitsuba;

extern "C" {
void *ref_hair_create(const float *xyz, const unsigned char *startsFiber, int n, float radius) {
    HairKDTree *k = new HairKDTree();
    for (int i = 0; i < n; ++i) { k->m_vertices.push_back(Vector(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])); k->m_vertexStartsFiber.push_back(startsFiber[i] != 0); }
    k->m_vertexStartsFiber.push_back(true);                                                     // the sentinel, hair.cpp:782
    k->m_radius = radius;
    for (int iv = 0; iv + 1 < n; ++iv) if (!k->m_vertexStartsFiber[iv + 1]) k->m_segIndex.push_back((uint32_t) iv);     // hair.cpp:117-124
    return k;
}
int ref_hair_segment_count(void *h) { return (int) ((HairKDTree *) h)->m_segIndex.size(); }
void ref_hair_segments(void *h, uint32_t *out) { const HairKDTree *k = (const HairKDTree *) h; for (size_t i = 0; i < k->m_segIndex.size(); ++i) out[i] = k->m_segIndex[i]; }
// HairKDTree::intersect for n (ray, segment, interval) tuples; outP = the fp32 hit point it stores
void ref_hair_intersect(void *h, int n, const float *o, const float *d, const uint32_t *iv, const float *mint, const float *maxt, int *outHit, float *outT, float *outP) {
    const HairKDTree *k = (const HairKDTree *) h;
    for (int i = 0; i < n; ++i) {
        Ray r; r.o = Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        HairKDTree::IntersectionStorage st; st.iv = 0; Float t = 0;
        const bool hit = k->intersect(r, iv[i], mint[i], maxt[i], t, &st);
        outHit[i] = hit ? 1 : 0; outT[i] = hit ? t : 0.0f;
        outP[3 * i] = hit ? st.p.x : 0; outP[3 * i + 1] = hit ? st.p.y : 0; outP[3 * i + 2] = hit ? st.p.z : 0;
    }
}
// getAABB(index) for every segment: 6 floats each
void ref_hair_segment_bounds(void *h, float *out) {
    const HairKDTree *k = (const HairKDTree *) h;
    for (size_t i = 0; i < k->m_segIndex.size(); ++i) {
        const AABB b = k->getAABB((HairKDTree::IndexType) i);
        out[6 * i] = b.min.x; out[6 * i + 1] = b.min.y; out[6 * i + 2] = b.min.z; out[6 * i + 3] = b.max.x; out[6 * i + 4] = b.max.y; out[6 * i + 5] = b.max.z;
    }
}
// Scene of several hair shapes: per-shape boxes = union of getAABB(index), enlarged; scene box = their union, enlarged again
void *ref_scene_create(int nShapes, void **shapes) {
    ShapeKDTree *sc = new ShapeKDTree();
    for (int i = 0; i < nShapes; ++i) {
        HairKDTree *k = (HairKDTree *) shapes[i];
        AABB box;
        for (size_t j = 0; j < k->m_segIndex.size(); ++j) {
            const AABB b = k->getAABB((HairKDTree::IndexType) j);
            for (int a = 0; a < 3; ++a) { box.min[a] = std::min(box.min[a], b.min[a]); box.max[a] = std::max(box.max[a], b.max[a]); }
        }
        enlargeKDBox(box);
        k->m_aabb = box;
        for (int a = 0; a < 3; ++a) { sc->m_aabb.min[a] = std::min(sc->m_aabb.min[a], box.min[a]); sc->m_aabb.max[a] = std::max(sc->m_aabb.max[a], box.max[a]); }
        sc->m_shapes.push_back(k);
    }
    enlargeKDBox(sc->m_aabb);
    return sc;
}
void ref_scene_bounds(void *h, float *out6) { const ShapeKDTree *sc = (const ShapeKDTree *) h; for (int a = 0; a < 3; ++a) { out6[a] = sc->m_aabb.min[a]; out6[3 + a] = sc->m_aabb.max[a]; } }
// ShapeKDTree::rayIntersect (closest: shape, segment, t) and the shadow-ray overload for n rays
void ref_scene_intersect(void *h, int n, const float *o, const float *d, const float *mint, const float *maxt, int *outShape, uint32_t *outIv, float *outT, int *outOccluded) {
    const ShapeKDTree *sc = (const ShapeKDTree *) h;
    for (int i = 0; i < n; ++i) {
        Ray r; r.o = Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        r.dRcp = Vector(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z); r.mint = mint[i]; r.maxt = maxt[i];
        SceneIntersection its;
        const bool hit = sc->rayIntersect(r, its);
        outShape[i] = hit ? its.shapeIndex : -1; outIv[i] = hit ? its.iv : 0xffffffffu; outT[i] = hit ? its.t : std::numeric_limits<float>::infinity();
        outOccluded[i] = sc->rayIntersect(r) ? 1 : 0;
    }
}
// TriAccel::load + rayIntersect for n (triangle, ray, interval) tuples; outAccel = k, n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv (k as float)
void ref_triaccel(int n, const float *A, const float *B, const float *C, const float *o, const float *d, const float *mint, const float *maxt,
                  float *outAccel, int *outHit, float *outTUV) {
    for (int i = 0; i < n; ++i) {
        TriAccel acc;
        acc.load(Vector(A[3 * i], A[3 * i + 1], A[3 * i + 2]), Vector(B[3 * i], B[3 * i + 1], B[3 * i + 2]), Vector(C[3 * i], C[3 * i + 1], C[3 * i + 2]));
        const float vals[10] = {(float) acc.k, acc.n_u, acc.n_v, acc.n_d, acc.a_u, acc.a_v, acc.b_nu, acc.b_nv, acc.c_nu, acc.c_nv};
        for (int k = 0; k < 10; ++k) outAccel[10 * i + k] = acc.k == 3 && k > 0 ? 0.0f : vals[k];
        Ray r; r.o = Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        Float u = 0, v = 0, t = 0;
        const bool hit = acc.rayIntersect(r, mint[i], maxt[i], u, v, t);
        outHit[i] = hit ? 1 : 0; outTUV[3 * i] = hit ? t : 0; outTUV[3 * i + 1] = hit ? u : 0; outTUV[3 * i + 2] = hit ? v : 0;
    }
}
// TAABB::rayIntersect for n (box, ray) pairs
void ref_aabb_ray(int n, const float *bmin, const float *bmax, const float *o, const float *d, int *outHit, float *outNearFar) {
    for (int i = 0; i < n; ++i) {
        AABB b; b.min = Vector(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2]); b.max = Vector(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2]);
        Ray r; r.o = Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        r.dRcp = Vector(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);                                       // ray.h:72-93 (setDirection)
        Float nearT = 0, farT = 0;
        const bool hit = b.rayIntersect(r, nearT, farT);
        outHit[i] = hit ? 1 : 0; outNearFar[2 * i] = hit ? nearT : 0; outNearFar[2 * i + 1] = hit ? farT : 0;
    }
}
// HairShape::fillIntersectionRecord for n (segment, stored hit point) pairs: out = p(3) n(3) s(3) t(3)
void ref_hair_records(void *h, int n, const uint32_t *iv, const float *p, float *out) {
    HairShape shape; shape.m_kdtree = (HairKDTree *) h;
    for (int i = 0; i < n; ++i) {
        HairKDTree::IntersectionStorage st; st.iv = iv[i]; st.p = Vector(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        Ray r; GeoIntersection its;
        shape.fillIntersectionRecord(r, &st, its);
        const Vector vs[4] = {its.p, its.geoFrame.n, its.geoFrame.s, its.geoFrame.t};
        for (int k = 0; k < 4; ++k) { out[12 * i + 3 * k] = vs[k].x; out[12 * i + 3 * k + 1] = vs[k].y; out[12 * i + 3 * k + 2] = vs[k].z; }
    }
}
}

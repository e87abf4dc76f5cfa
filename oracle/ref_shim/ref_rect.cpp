// oracle/ref_shim/ref_rect.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The `rectangle` shape and the `checkerboard` texture of the reference executed as written (models/teapot/scene.xml:41-61): oracle/Makefile
// cuts these pieces of text out of /root/reference at build time (oracle/_ref/ref_rect_*.inc, ref_tex_*.inc) and this file pastes them into
// classes that only supply the members they touch:
//   src/shapes/rectangle.cpp :100-125 configure / getAABB / getSurfaceArea, :127-156 rayIntersect (both), :158-171 fillIntersectionRecord
//   include/mitsuba/core/transform.h :108-125 operator()(Point), :139-146 transformAffine(Point, Point &), :175-196 operator()(Vector) (both forms),
//                                    :203-211 operator()(Normal), :292-307 transformAffine(Ray, Ray &)
//   src/libcore/transform.cpp :28-31 Transform::operator*, :49-63 Transform::scale;  matrix.h:743-757, matrix.inl:138-193 (product, Gauss-Jordan inverse)
//   src/libcore/util.cpp :592-601 coordinateSystem
//   src/textures/checkerboard.cpp :65-72 Checkerboard::eval(uv);  src/librender/texture.cpp :112-121 Texture2D::eval(its, filter);  math.h:67-70 modulo
// The constructor of Rectangle (:81-86, three statements on a Properties object) is restated in ref_rect_create.  Output: part of libref_geom.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>

namespace refrect {
typedef float Float;
#undef M_PI
#define M_PI 3.14159265358979323846f
#define BOOST_STATIC_ASSERT(x) static_assert(x, "")
#define MTS_EXPORT_CORE
#define FINLINE inline
static const Float Epsilon = 1e-4f;
enum ELogLevel { EDebug, EInfo, EWarn, EError };
#define SLog(level, ...) do { if (level >= EError) throw std::runtime_error("matrix is singular"); } while (0)
#define Log(level, ...) do { if (level >= EError) { char b_[256]; snprintf(b_, sizeof(b_), __VA_ARGS__); throw std::runtime_error(b_); } } while (0)
inline Float degToRad(Float value) { return value * (M_PI / 180.0f); }

struct Vector { Float x, y, z; Vector() : x(0), y(0), z(0) {} Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Vector operator/(Float f) const { Float recip = (Float) 1 / f; return Vector(x * recip, y * recip, z * recip); }      // vector.h: multiplies by the reciprocal
    Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
    Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
    Vector operator*(Float f) const { return Vector(x * f, y * f, z * f); }
    Float length() const { return (Float) std::sqrt((Float) (x * x + y * y + z * z)); } };
inline Vector operator*(Float f, const Vector &v) { return v * f; }
struct Normal : public Vector { Normal() {} Normal(Float x, Float y, Float z) : Vector(x, y, z) {} Normal(const Vector &v) : Vector(v) {} };
struct Point { Float x, y, z; Point() : x(0), y(0), z(0) {} Point(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Point operator+(const Vector &v) const { return Point(x + v.x, y + v.y, z + v.z); }
    Point operator/(Float f) const { Float recip = (Float) 1 / f; return Point(x * recip, y * recip, z * recip); } };
typedef Point Point3;
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {}
    Point2 operator+(const Point2 &p) const { return Point2(x + p.x, y + p.y); } };
struct Vector2 { Float x, y; Vector2() : x(0), y(0) {} Vector2(Float x, Float y) : x(x), y(y) {} };
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector cross(const Vector &a, const Vector &b) { return Vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline Vector normalize(const Vector &v) { return v / v.length(); }
#include "ref_sun_coordsys.inc"
struct Frame { Vector s, t; Normal n; Frame() {}
    Frame(const Vector &s, const Vector &t, const Normal &n) : s(s), t(t), n(n) {}                                     // frame.h:45-47
    Frame(const Vector &n) : n(n) { coordinateSystem(n, s, t); } };                                                    // frame.h:55-57
struct Ray { Point o; Float mint; Vector d; Float maxt; Vector dRcp; Float time = 0;
    inline Point operator()(Float t) const { return o + t * d; } };                                                    // ray.h:124
struct AABB { Point min, max; AABB() { reset(); }
    void reset() { const Float inf = std::numeric_limits<Float>::infinity(); min = Point(inf, inf, inf); max = Point(-inf, -inf, -inf); }
    void expandBy(const Point &p) { min.x = std::min(min.x, p.x); min.y = std::min(min.y, p.y); min.z = std::min(min.z, p.z);
                                    max.x = std::max(max.x, p.x); max.y = std::max(max.y, p.y); max.z = std::max(max.z, p.z); } };   // aabb.h:150-155

template <int M, int N, typename T> struct Matrix { T m[M][N]; bool invert(Matrix &target) const; };
#include "ref_cam_matmul.inc"
#include "ref_cam_invert.inc"
struct Matrix4x4 : public Matrix<4, 4, Float> {
    Matrix4x4() {}
    Matrix4x4(const Matrix<4, 4, Float> &o) { std::memcpy(m, o.m, sizeof(m)); }
    Matrix4x4(Float a00, Float a01, Float a02, Float a03, Float a10, Float a11, Float a12, Float a13, Float a20, Float a21, Float a22, Float a23, Float a30, Float a31, Float a32, Float a33) {
        const Float v[16] = {a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33}; std::memcpy(m, v, sizeof(m)); }
    std::string toString() const { return "matrix"; }
};
struct Transform {
    Matrix4x4 m_transform, m_invTransform;
    Transform() { *this = Transform(Matrix4x4(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1), Matrix4x4(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)); }
    Transform(const Matrix4x4 &trafo) : m_transform(trafo) {                                                           // transform.h:50-55
        bool success = m_transform.invert(m_invTransform);
        if (!success) SLog(EError, "Unable to invert singular matrix %s", trafo.toString().c_str()); }
    Transform(const Matrix4x4 &trafo, const Matrix4x4 &invTrafo) : m_transform(trafo), m_invTransform(invTrafo) {}
    Transform inverse() const { return Transform(m_invTransform, m_transform); }                                        // transform.h:62-64
    Transform operator*(const Transform &t) const;
    static Transform translate(const Vector &v); static Transform scale(const Vector &v); static Transform perspective(Float fov, Float clipNear, Float clipFar);
#include "ref_rect_apply.inc"
};
#include "ref_cam_transform.inc"

struct Shape;
struct Intersection { Float t; Point p; Frame geoFrame, shFrame; Point2 uv; Vector dpdu, dpdv; Float dudx, dudy, dvdx, dvdy, time; Vector wi; bool hasUVPartials; const Shape *shape; const Shape *instance; };
struct Shape { void configure() {} };
struct Rectangle : public Shape {
    Transform m_objectToWorld, m_worldToObject; Frame m_frame; Vector m_dpdu, m_dpdv; Float m_invSurfaceArea;
#include "ref_rect_members.inc"
};

namespace math {
#include "ref_tex_modulo.inc"
}
struct Spectrum { Float s[3]; };
struct Texture2D { Point2 m_uvOffset; Vector2 m_uvScale;
    virtual Spectrum eval(const Point2 &uv) const = 0;
    virtual Spectrum eval(const Point2 &uv, const Vector2 &d0, const Vector2 &d1) const = 0;
    Spectrum eval(const Intersection &its, bool filter) const; };
#include "ref_tex_texture2d.inc"
struct Checkerboard : public Texture2D { Spectrum m_color0, m_color1;
    Spectrum eval(const Point2 &uv, const Vector2 &, const Vector2 &) const override { return Checkerboard::eval(uv); }
#include "ref_tex_checker.inc"
};
} // namespace refrect

using namespace refrect;
extern "C" {
// Rectangle(const Properties &) rectangle.cpp:81-86, then configure()
void *ref_rect_create(const float *toWorld16, int flipNormals, char *err, int errSize) {
    try {
        Rectangle *r = new Rectangle();
        Matrix4x4 m; std::memcpy(m.m, toWorld16, sizeof(m.m));
        r->m_objectToWorld = Transform(m);
        if (flipNormals) r->m_objectToWorld = r->m_objectToWorld * Transform::scale(Vector(1, 1, -1));
        r->m_worldToObject = r->m_objectToWorld.inverse();
        r->configure();
        return r;
    } catch (const std::exception &e) { if (err) snprintf(err, errSize, "%s", e.what()); return nullptr; }
}
void ref_rect_bounds(void *h, float *out6) { const AABB a = ((Rectangle *) h)->getAABB(); out6[0] = a.min.x; out6[1] = a.min.y; out6[2] = a.min.z; out6[3] = a.max.x; out6[4] = a.max.y; out6[5] = a.max.z; }
// per ray: hit flag, t, then the record (p, geoFrame.n, shFrame.n, dpdu, uv: 14 floats) as fillIntersectionRecord leaves it
void ref_rect_intersect(void *h, int n, const float *o, const float *d, const float *mint, const float *maxt, int *outHit, float *outT, float *outRec14) {
    const Rectangle *r = (const Rectangle *) h;
    for (int i = 0; i < n; ++i) {
        Ray ray; ray.o = Point(o[3 * i], o[3 * i + 1], o[3 * i + 2]); ray.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]); ray.mint = mint[i]; ray.maxt = maxt[i];
        Float t = 0, temp[2] = {0, 0};
        const bool hit = r->rayIntersect(ray, mint[i], maxt[i], t, temp);
        outHit[i] = hit ? 1 : 0; outT[i] = hit ? t : 0;
        if (r->rayIntersect(ray, mint[i], maxt[i]) != hit) outHit[i] = -1;
        float *q = outRec14 + 14 * i; std::memset(q, 0, 14 * sizeof(float));
        if (!hit) continue;
        Intersection its; its.t = t;
        r->fillIntersectionRecord(ray, temp, its);
        q[0] = its.p.x; q[1] = its.p.y; q[2] = its.p.z; q[3] = its.geoFrame.n.x; q[4] = its.geoFrame.n.y; q[5] = its.geoFrame.n.z;
        q[6] = its.shFrame.n.x; q[7] = its.shFrame.n.y; q[8] = its.shFrame.n.z; q[9] = its.dpdu.x; q[10] = its.dpdu.y; q[11] = its.dpdu.z; q[12] = its.uv.x; q[13] = its.uv.y;
    }
}
void ref_checkerboard_eval(const float *color0, const float *color1, float uoffset, float voffset, float uscale, float vscale, int n, const float *uv, float *outRGB) {
    Checkerboard c; for (int k = 0; k < 3; ++k) { c.m_color0.s[k] = color0[k]; c.m_color1.s[k] = color1[k]; }
    c.m_uvOffset = Point2(uoffset, voffset); c.m_uvScale = Vector2(uscale, vscale);
    for (int i = 0; i < n; ++i) {
        Intersection its; its.uv = Point2(uv[2 * i], uv[2 * i + 1]); its.hasUVPartials = false;
        const Spectrum s = static_cast<const Texture2D &>(c).eval(its, true);
        outRGB[3 * i] = s.s[0]; outRGB[3 * i + 1] = s.s[1]; outRGB[3 * i + 2] = s.s[2];
    }
}
}

// oracle/ref_shim/ref_camera.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The perspective sensor of the reference executed as written: oracle/Makefile cuts these pieces of text out of /root/reference at build
// time (oracle/_ref/ref_cam_*.inc) and this file pastes them into classes that only supply the members they touch:
//   src/sensors/perspective.cpp :126-180 PerspectiveCameraImpl::configure(), :271-298 sampleRayDifferential()
//   src/libcore/transform.cpp   :28-31 Transform::operator*, :33-47 translate, :49-63 scale, :99-123 perspective
//   include/mitsuba/core/transform.h :108-125 operator()(Point), :128-136 transformAffine, :175-183 operator()(Vector)
//   include/mitsuba/core/matrix.h :743-757 matrix product;  include/mitsuba/core/matrix.inl :138-193 Matrix::invert (fp32 Gauss-Jordan)
// Point and Vector are distinct types here.  Output: part of oracle/_ref/libref_geom.so.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>

namespace refcam {
typedef float Float;
#undef M_PI
#define M_PI 3.14159265358979323846f
#define BOOST_STATIC_ASSERT(x) static_assert(x, "")
#define MTS_EXPORT_CORE
enum ELogLevel { EDebug, EInfo, EWarn, EError };
#define SLog(level, ...) do { if (level >= EError) throw std::runtime_error("matrix is singular"); } while (0)
inline Float degToRad(Float value) { return value * (M_PI / 180.0f); }              // util.h:297

struct Vector { Float x, y, z; Vector() : x(0), y(0), z(0) {} Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Vector operator/(Float f) const { Float recip = (Float) 1 / f; return Vector(x * recip, y * recip, z * recip); }
    Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
    Float length() const { return (Float) std::sqrt((Float) (x * x + y * y + z * z)); } };
struct Point { Float x, y, z; Point() : x(0), y(0), z(0) {} Point(Float x, Float y, Float z) : x(x), y(y), z(z) {} explicit Point(Float v) : x(v), y(v), z(v) {}
    Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
    Point operator/(Float f) const { Float recip = (Float) 1 / f; return Point(x * recip, y * recip, z * recip); } };
struct VectorFromPoint : public Vector { explicit VectorFromPoint(const Point &p) : Vector(p.x, p.y, p.z) {} };
inline Vector normalize(const Vector &v) { return v / v.length(); }
struct Vector2 { Float x, y; Vector2() : x(0), y(0) {} Vector2(Float x, Float y) : x(x), y(y) {} };
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} Point2 operator/(Float f) const { Float recip = (Float) 1 / f; return Point2(x * recip, y * recip); } };
struct Vector2i { int x, y; }; struct Point2i { int x, y; };
struct Spectrum { explicit Spectrum(Float) {} };

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
    Transform(const Matrix4x4 &trafo) : m_transform(trafo) {                                                     // transform.h:50-55
        bool success = m_transform.invert(m_invTransform);
        if (!success) SLog(EError, "Unable to invert singular matrix %s", trafo.toString().c_str()); }
    Transform(const Matrix4x4 &trafo, const Matrix4x4 &invTrafo) : m_transform(trafo), m_invTransform(invTrafo) {}
    Transform inverse() const { return Transform(m_invTransform, m_transform); }                                  // transform.h:62-64
    Transform operator*(const Transform &t) const;
    static Transform translate(const Vector &v); static Transform scale(const Vector &v); static Transform perspective(Float fov, Float clipNear, Float clipFar);
#include "ref_cam_apply.inc"
};
#include "ref_cam_transform.inc"

struct AABB2 { void reset() {} void expandBy(const Point2 &) {} Float getVolume() const { return 1.0f; } };
struct Film { Vector2i size; Point2i off; const Vector2i &getSize() const { return size; } const Vector2i &getCropSize() const { return size; } const Point2i &getCropOffset() const { return off; } };
struct AnimatedTransform { Transform t; const Transform &eval(Float) const { return t; } };
struct RayDifferential { Point o, rxOrigin, ryOrigin; Vector d, rxDirection, ryDirection; Float mint, maxt, time; bool hasDifferentials;
    void setOrigin(const Point &p) { o = p; } void setDirection(const Vector &v) { d = v; } };
struct PerspectiveCamera { Film *m_film = nullptr; Float m_aspect = 1; Vector2 m_resolution, m_invResolution;
    void configure() {                                                                                           // src/librender/sensor.cpp:101-107
        m_aspect = m_film->getSize().x / (Float) m_film->getSize().y;
        m_resolution = Vector2((Float) m_film->getCropSize().x, (Float) m_film->getCropSize().y);
        m_invResolution = Vector2((Float) 1 / m_resolution.x, (Float) 1 / m_resolution.y); } };
#define Vector(...) VectorCtor(__VA_ARGS__)
// `Vector(nearP)` in the pasted text converts a Point; `Vector(a, b, c)` builds one
inline Vector VectorCtor(const Point &p) { return VectorFromPoint(p); }
inline Vector VectorCtor(Float x, Float y, Float z) { Vector v; v.x = x; v.y = y; v.z = z; return v; }
class PerspectiveCameraImpl : public PerspectiveCamera {
public:
    Float m_xfov = 35, m_nearClip = 1e-2f, m_farClip = 1e4f, m_normalization = 0; Transform m_cameraToSample, m_sampleToCamera, m_clipTransform; Vector m_dx, m_dy; AABB2 m_imageRect;
    AnimatedTransform *m_worldTransform = nullptr;
    Float sampleTime(Float) const { return 0; }
#include "ref_cam_members.inc"
};
#undef Vector
} // namespace refcam

using namespace refcam;
extern "C" {
void *ref_camera_create(const float *toWorld16, float xfov, float nearClip, float farClip, int w, int h) {
    PerspectiveCameraImpl *c = new PerspectiveCameraImpl();
    c->m_film = new Film(); c->m_film->size.x = w; c->m_film->size.y = h; c->m_film->off.x = c->m_film->off.y = 0;
    c->m_xfov = xfov; c->m_nearClip = nearClip; c->m_farClip = farClip;
    c->m_worldTransform = new AnimatedTransform();
    Matrix4x4 m; std::memcpy(m.m, toWorld16, sizeof(m.m));
    c->m_worldTransform->t = Transform(m);
    c->configure();
    return c;
}
void ref_camera_matrices(void *h, float *sampleToCamera16, float *dxdy6) {
    const PerspectiveCameraImpl *c = (const PerspectiveCameraImpl *) h;
    std::memcpy(sampleToCamera16, c->m_sampleToCamera.m_transform.m, 64);
    dxdy6[0] = c->m_dx.x; dxdy6[1] = c->m_dx.y; dxdy6[2] = c->m_dx.z; dxdy6[3] = c->m_dy.x; dxdy6[4] = c->m_dy.y; dxdy6[5] = c->m_dy.z;
}
void ref_camera_rays(void *h, int n, const float *pxy, float *outO, float *outD, float *outMinMax, float *outRxRy) {
    const PerspectiveCameraImpl *c = (const PerspectiveCameraImpl *) h;
    for (int i = 0; i < n; ++i) {
        RayDifferential r;
        c->sampleRayDifferential(r, Point2(pxy[2 * i], pxy[2 * i + 1]), Point2(0.5f, 0.5f), 0.5f);
        outO[3 * i] = r.o.x; outO[3 * i + 1] = r.o.y; outO[3 * i + 2] = r.o.z; outD[3 * i] = r.d.x; outD[3 * i + 1] = r.d.y; outD[3 * i + 2] = r.d.z;
        outMinMax[2 * i] = r.mint; outMinMax[2 * i + 1] = r.maxt;
        outRxRy[6 * i] = r.rxDirection.x; outRxRy[6 * i + 1] = r.rxDirection.y; outRxRy[6 * i + 2] = r.rxDirection.z;
        outRxRy[6 * i + 3] = r.ryDirection.x; outRxRy[6 * i + 4] = r.ryDirection.y; outRxRy[6 * i + 5] = r.ryDirection.z;
    }
}
// Matrix4x4::invert alone (the envmap's toWorld goes through it as well)
int ref_matrix_invert(const float *m16, float *out16) {
    Matrix4x4 m, inv; std::memcpy(m.m, m16, 64);
    const bool ok = m.invert(inv);
    std::memcpy(out16, inv.m, 64);
    return ok ? 1 : 0;
}
}

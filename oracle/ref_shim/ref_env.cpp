// oracle/ref_shim/ref_env.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The environment-map emitter of the reference executed as written: oracle/Makefile cuts these member functions out of
// /root/reference at build time (oracle/_ref/ref_env_*.inc) and this file pastes them into classes that only supply the members they touch:
//   src/emitters/envmap.cpp  :260-329 configure() (row / column CDFs, row weights, normalisation), :358-374 fillDirectSamplingRecord,
//                            :380-410 evalEnvironment, :516-543 sampleDirect, :545-556 pdfDirect, :567-602 internalSampleDirection,
//                            :603-635 internalPdfDirection, :657-662 sampleReuse
//   include/mitsuba/render/mipmap.h :503-566 evalTexel (repeat / clamp boundary handling), :577-596 evalBilinear (+ evalBox)
//   include/mitsuba/core/bsphere.h :88-95 BSphere::rayIntersect;  src/libcore/util.cpp :448-482 solveQuadratic;
//   src/libcore/warp.cpp :143-162 intervalToTent / squareToTent
// Texels are stored as IEEE binary16 (`_Float16`, round to nearest even), the compiler's own conversion -- independent of the oracle's
// half routines; the MIP pyramid above level 0 and the EWA lookup (mipmap.h:629-700) are not part of this path.
// Output: part of oracle/_ref/libref_geom.so.
#include "mitsuba_shim.h"

namespace mitsuba {
namespace envpin {
typedef _Float16 half;
struct Vector2i { int x, y; Vector2i() : x(0), y(0) {} Vector2i(int x, int y) : x(x), y(y) {} };
struct SpectrumHalf { typedef void IsSpectrumType; half s[3]; };

namespace math {
    using namespace mitsuba::math;
    template <typename Scalar> inline int floorToInt(Scalar value) { return (int) std::floor(value); }      // math.h:100
    inline int32_t modulo(int32_t a, int32_t b) { int32_t r = a % b; return (r < 0) ? r + b : r; }          // math.h:67-70
}
struct ReconstructionFilter { enum EBoundaryCondition { EClamp = 0, ERepeat, EMirror, EZero, EOne }; };
struct Array2D { Vector2i size; std::vector<SpectrumHalf> data;
    const Vector2i &getSize() const { return size; }
    const SpectrumHalf &operator()(int x, int y) const { return data[(size_t) y * size.x + x]; } };
struct MIPMap {
    typedef Spectrum Value; typedef Array2D Array2DType;
    Array2D m_pyramid[1]; int m_levels = 1;
    ReconstructionFilter::EBoundaryCondition m_bcu = ReconstructionFilter::ERepeat, m_bcv = ReconstructionFilter::EClamp;    // envmap.cpp:178-179
    const Array2D &getArray() const { return m_pyramid[0]; }
#include "ref_env_mipmap.inc"
    Value eval(const Point2 &uv, const Vector2 &, const Vector2 &) const { return evalBilinear(0, uv); }      // EWA: outside this path
};

// include/mitsuba/core/transform.h:175-183 (vectors only are transformed here) with both matrices stored like Transform does
struct Transform { Float m[4][4], inv[4][4];
    Transform inverse() const { Transform t; std::memcpy(t.m, inv, sizeof(m)); std::memcpy(t.inv, m, sizeof(m)); return t; }
    Vector operator()(const Vector &v) const {
        Float x = m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z;
        Float y = m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z;
        Float z = m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z;
        return Vector(x, y, z); } };
struct AnimatedTransform { Transform t; const Transform &eval(Float) const { return t; } };
struct Ray { Point o; Vector d; Float mint, maxt, time;
    Ray(const Point &o, const Vector &d, Float time) : o(o), d(d), mint(Epsilon), maxt(std::numeric_limits<Float>::infinity()), time(time) {}
    Point operator()(Float t) const { return o + t * d; } };                                                  // ray.h:96
struct RayDifferential : public Ray { bool hasDifferentials = false; Vector rxDirection, ryDirection;
    RayDifferential(const Point &o, const Vector &d, Float time) : Ray(o, d, time) {} };
#include "ref_env_quadratic.inc"
struct BSphere { Point center; Float radius = 0;
#include "ref_env_bsphere.inc"
};
struct DirectSamplingRecord { Point ref, p; Normal n; Vector d; Float dist = 0, pdf = 0, time = 0; EMeasure measure = ESolidAngle; const void *object = nullptr; };
namespace stats { struct Counter { Counter &operator++() { return *this; } void incrementBase() {} }; static Counter filteredLookups; }
struct Timer { int getMilliseconds() const { return 0; } };
namespace warp {
#include "ref_env_tent.inc"
}
struct Emitter { void configure() {} };

class EnvironmentMap : public Emitter {
public:
    MIPMap *m_mipmap = nullptr; float *m_cdfRows = nullptr, *m_cdfCols = nullptr; Float *m_rowWeights = nullptr; Vector2i m_size;
    Float m_normalization = 0, m_power = 0, m_invSurfaceArea = 0, m_scale = 1; Vector2 m_pixelSize; BSphere m_sceneBSphere;
    AnimatedTransform *m_worldTransform = nullptr;
#include "ref_env_members.inc"
};
} // namespace envpin
} // namespace mitsuba

using namespace mitsuba;
using namespace mitsuba::envpin;
extern "C" {
void *ref_env_create(const float *rgb, int w, int h, const float *toWorld16, const float *toLocal16, float scale, const float *bsCenter, float bsRadius) {
    EnvironmentMap *e = new EnvironmentMap();
    e->m_mipmap = new MIPMap();
    e->m_mipmap->m_pyramid[0].size = Vector2i(w, h);
    e->m_mipmap->m_pyramid[0].data.resize((size_t) w * h);
    for (size_t i = 0; i < (size_t) w * h; ++i) for (int k = 0; k < 3; ++k) e->m_mipmap->m_pyramid[0].data[i].s[k] = (half) std::max(rgb[3 * i + k], 0.0f);   // mipmap.h:232-240 clamps negatives
    e->m_worldTransform = new AnimatedTransform();
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { e->m_worldTransform->t.m[i][j] = toWorld16[4 * i + j]; e->m_worldTransform->t.inv[i][j] = toLocal16[4 * i + j]; }
    e->m_scale = scale;
    e->m_sceneBSphere.center = Vector(bsCenter[0], bsCenter[1], bsCenter[2]); e->m_sceneBSphere.radius = bsRadius;
    e->configure();
    return e;
}
void ref_env_tables(void *h, float *cdfRows, float *cdfCols, float *rowWeights, float *normalization) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    std::memcpy(cdfRows, e->m_cdfRows, 4 * (size_t) (e->m_size.y + 1)); std::memcpy(cdfCols, e->m_cdfCols, 4 * (size_t) (e->m_size.x + 1) * e->m_size.y);
    std::memcpy(rowWeights, e->m_rowWeights, 4 * (size_t) e->m_size.y); *normalization = e->m_normalization;
}
void ref_env_eval(void *h, int n, const float *d, float *outRGB, float *outPdf) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    for (int i = 0; i < n; ++i) {
        RayDifferential r(Vector(0, 0, 0), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]), 0);
        const mitsuba::Spectrum v = e->evalEnvironment(r);
        outRGB[3 * i] = v[0]; outRGB[3 * i + 1] = v[1]; outRGB[3 * i + 2] = v[2];
        DirectSamplingRecord dRec; dRec.d = r.d; dRec.measure = ESolidAngle;
        outPdf[i] = e->pdfDirect(dRec);
    }
}
void ref_env_sample(void *h, int n, const float *ref, const float *sample, float *outD, float *outValue, float *outPdfDist) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    for (int i = 0; i < n; ++i) {
        DirectSamplingRecord dRec; dRec.ref = Vector(ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]);
        const mitsuba::Spectrum v = e->sampleDirect(dRec, Point2(sample[2 * i], sample[2 * i + 1]));
        outD[3 * i] = dRec.d.x; outD[3 * i + 1] = dRec.d.y; outD[3 * i + 2] = dRec.d.z;
        outValue[3 * i] = v[0]; outValue[3 * i + 1] = v[1]; outValue[3 * i + 2] = v[2];
        outPdfDist[2 * i] = dRec.pdf; outPdfDist[2 * i + 1] = dRec.dist;
    }
}
void ref_env_fill(void *h, int n, const float *o, const float *d, int *outOk) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    for (int i = 0; i < n; ++i) {
        DirectSamplingRecord dRec; envpin::Ray r(Vector(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]), 0);
        outOk[i] = e->fillDirectSamplingRecord(dRec, r) ? 1 : 0;
    }
}
}

// oracle/ref_shim/ref_env.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The environment-map emitter of the reference executed as written: oracle/Makefile cuts these member functions out of
// /root/reference at build time (oracle/_ref/ref_env_*.inc) and this file pastes them into classes that only supply the members they touch:
//   src/emitters/envmap.cpp  :260-329 configure() (row / column CDFs, row weights, normalisation), :358-374 fillDirectSamplingRecord,
//                            :380-410 evalEnvironment, :516-543 sampleDirect, :545-556 pdfDirect, :567-602 internalSampleDirection,
//                            :603-635 internalPdfDirection, :657-662 sampleReuse
//   include/mitsuba/render/mipmap.h :503-566 evalTexel (repeat / clamp boundary handling), :577-596 evalBilinear (+ evalBox),
//                            :629-725 eval (EWA filter type: ellipse from the uv Jacobian, anisotropy clamp, level choice), :760-836 evalEWA
//   include/mitsuba/core/rfilter.h :107-460 Resampler<Scalar> (whole struct); src/rfilters/lanczos.cpp :42-55 LanczosSincFilter::eval;
//   src/libcore/math.cpp :74-86 hypot2, :103-106 log2 -- the MIP pyramid is built by the driver loop below exactly like
//   TMIPMap's constructor (mipmap.h:245-271) + Bitmap::resample (bitmap.cpp:2230-2328) drive these pieces
//   include/mitsuba/core/bsphere.h :88-95 BSphere::rayIntersect;  src/libcore/util.cpp :448-482 solveQuadratic;
//   src/libcore/warp.cpp :143-162 intervalToTent / squareToTent
// Texels are stored as IEEE binary16 (`_Float16`, round to nearest even), the compiler's own conversion -- independent of the oracle's
// half routines.
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
    template <typename Scalar> inline int ceilToInt(Scalar value) { return (int) std::ceil(value); }        // math.h:105
    inline int32_t modulo(int32_t a, int32_t b) { int32_t r = a % b; return (r < 0) ? r + b : r; }          // math.h:67-70
    inline float fastlog(float value) { return (float) ::log((double) value); }                             // math.h:193-195 (Linux / x86-64)
    inline float fastexp(float value) { return (float) ::exp((double) value); }                             // math.h:185-187
#include "ref_env_mathfn.inc"
}
#define SAssert(x) ((void) 0)
#define EXPECT_NOT_TAKEN(x) (x)
#define FINLINE inline
#define MTS_EXPORT_CORE
#define MTS_MIPMAP_LUT_SIZE 64
struct ReconstructionFilter {
    enum EBoundaryCondition { EClamp = 0, ERepeat, EMirror, EZero, EOne };
    Float m_radius = 2;                                     // LanczosSincFilter(props): lobes = 2 (envmap.cpp:160-163)
    Float getRadius() const { return m_radius; }
#include "ref_env_lanczos.inc"
};
#include "ref_env_resampler.inc"
enum EMIPFilterType { ENearest = 0, EBilinear = 1, ETrilinear = 2, EEWA = 3 };
struct Array2D { Vector2i size; std::vector<SpectrumHalf> data;
    const Vector2i &getSize() const { return size; }
    const SpectrumHalf &operator()(int x, int y) const { return data[(size_t) y * size.x + x]; } };
namespace stats { struct Counter { Counter &operator++() { return *this; } Counter &operator+=(int) { return *this; } void incrementBase() {} };
                  static Counter filteredLookups, clampedAnisotropy, avgEWASamples; }
struct MIPMap {
    typedef Spectrum Value; typedef Array2D Array2DType;
    Array2D m_pyramid[20]; Vector2 m_sizeRatio[20]; int m_levels = 1;
    ReconstructionFilter::EBoundaryCondition m_bcu = ReconstructionFilter::ERepeat, m_bcv = ReconstructionFilter::EClamp;    // envmap.cpp:178-179
    EMIPFilterType m_filterType = EEWA; Float m_maxAnisotropy = 10.0f; Float m_weightLut[MTS_MIPMAP_LUT_SIZE];                 // envmap.cpp:150-152
    const Array2D &getArray() const { return m_pyramid[0]; }
#include "ref_env_mipmap.inc"
    // TMIPMap constructor, steps 1-3 (mipmap.h:180-192, 232-271, 296-302) with bitmap->resample(rfilter, bcu, bcv, size, 0.0f, inf) spelled
    // out as Bitmap::resample does it (bitmap.cpp:2258-2326): X pass into a temporary, then Y pass, both through resampleAndClamp
    void build(const float *rgb, int w, int h) {
        std::vector<float> img(rgb, rgb + (size_t) 3 * w * h);
        for (float &v : img) v = std::max(v, 0.0f);
        auto store = [&](int level, const std::vector<float> &src, int lw, int lh) {
            m_pyramid[level].size = Vector2i(lw, lh); m_pyramid[level].data.resize((size_t) lw * lh);
            for (size_t i = 0; i < (size_t) lw * lh; ++i) for (int k = 0; k < 3; ++k) m_pyramid[level].data[i].s[k] = (half) src[3 * i + k];
        };
        store(0, img, w, h);
        m_sizeRatio[0] = Vector2(1, 1);
        ReconstructionFilter rfilter;
        const Float minValue = 0.0f, maxValue = std::numeric_limits<Float>::infinity();
        Vector2i size(w, h);
        m_levels = 1;
        while (size.x > 1 || size.y > 1) {
            const Vector2i srcSize = size;
            size.x = std::max(1, (size.x + 1) / 2);
            size.y = std::max(1, (size.y + 1) / 2);
            std::vector<float> target((size_t) 3 * size.x * size.y), temp;
            const float *source = img.data();
            if (srcSize.x != size.x) {
                Resampler<float> r(&rfilter, m_bcu, srcSize.x, size.x);
                float *out = target.data();
                if (srcSize.y != size.y) { temp.resize((size_t) 3 * size.x * srcSize.y); out = temp.data(); }
                for (int y = 0; y < srcSize.y; ++y) r.resampleAndClamp(source + (size_t) y * srcSize.x * 3, 1, out + (size_t) y * size.x * 3, 1, 3, minValue, maxValue);
                source = out;
            }
            if (srcSize.y != size.y) {
                Resampler<float> r(&rfilter, m_bcv, srcSize.y, size.y);
                for (int x = 0; x < size.x; ++x) r.resampleAndClamp(source + (size_t) x * 3, (size_t) size.x, target.data() + (size_t) x * 3, (size_t) size.x, 3, minValue, maxValue);
            } else if (srcSize.x == size.x) target = img;
            store(m_levels, target, size.x, size.y);
            m_sizeRatio[m_levels] = Vector2((Float) size.x / (Float) m_pyramid[0].getSize().x, (Float) size.y / (Float) m_pyramid[0].getSize().y);
            img.swap(target);
            ++m_levels;
        }
        for (int i = 0; i < MTS_MIPMAP_LUT_SIZE; ++i) {
            Float r2 = (Float) i / (Float) (MTS_MIPMAP_LUT_SIZE - 1);
            m_weightLut[i] = math::fastexp(-2.0f * r2) - math::fastexp(-2.0f);
        }
    }
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
    e->m_mipmap->build(rgb, w, h);          // level 0: (half) max(rgb, 0) (mipmap.h:232-240 clamps negatives), then the Lanczos pyramid
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
// evalEnvironment with ray differentials (the EWA branch of MIPMap::eval)
void ref_env_eval_filtered(void *h, int n, const float *d, const float *rx, const float *ry, float *outRGB) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    for (int i = 0; i < n; ++i) {
        RayDifferential r(Vector(0, 0, 0), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]), 0);
        r.hasDifferentials = true;
        r.rxDirection = Vector(rx[3 * i], rx[3 * i + 1], rx[3 * i + 2]); r.ryDirection = Vector(ry[3 * i], ry[3 * i + 1], ry[3 * i + 2]);
        const mitsuba::Spectrum v = e->evalEnvironment(r);
        outRGB[3 * i] = v[0]; outRGB[3 * i + 1] = v[1]; outRGB[3 * i + 2] = v[2];
    }
}
int ref_env_mip_level(void *h, int level, int *outW, int *outH, float *outRGB) {
    const EnvironmentMap *e = (const EnvironmentMap *) h;
    const MIPMap *m = e->m_mipmap;
    if (level < 0 || level >= m->m_levels) return -1;
    const Array2D &a = m->m_pyramid[level];
    if (outW) *outW = a.size.x;
    if (outH) *outH = a.size.y;
    if (outRGB) for (size_t i = 0; i < a.data.size(); ++i) for (int k = 0; k < 3; ++k) outRGB[3 * i + k] = (float) a.data[i].s[k];
    return m->m_levels;
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

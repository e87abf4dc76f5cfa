// oracle/ref_shim/ref_sunrad.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The spectral side of the sunsky bake executed as the reference wrote it.  oracle/Makefile cuts these pieces out of /root/reference at
// build time (oracle/_ref/ref_sun_*.inc) and this file supplies only the declarations they need:
//   src/emitters/sunsky/sunmodel.h :252-371   the Iqbal absorption tables, the solar spectrum and computeSunRadiance (Preetham et al.)
//   src/libcore/spectrum.cpp       :172-185 Spectrum::fromContinuousSpectrum (RGB build), :222-227 fromXYZ, :503-505 ProductSpectrum::eval,
//                                  :546-568 ContinuousSpectrum::average (adaptive Gauss-Lobatto over 50 nm steps),
//                                  :604-615 InterpolatedSpectrum(wavelengths, values, n), :688-714 InterpolatedSpectrum::eval, the CIE 1931 tables
//   src/libcore/quad.cpp           :287-420   GaussLobattoIntegrator (constants, constructor, integrate, calculateAbsTolerance, adaptive step)
// boost::function / boost::bind become std::function / std::bind.  Output: oracle/_ref/libref_sun.so.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <functional>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_EXPORT_CORE
#define SPECTRUM_SAMPLES 3
#define SLog(level, ...) do { if (level >= 2) { char b_[512]; snprintf(b_, sizeof(b_), __VA_ARGS__); throw std::runtime_error(b_); } } while (0)
namespace boost { using std::function; using namespace std::placeholders; template <class... A> auto bind(A &&...a) { return std::bind(std::forward<A>(a)...); } }
using namespace std::placeholders;
using std::endl;

namespace mitsuba {
typedef float Float;                                             // -DSINGLE_PRECISION
#undef M_PI
#define M_PI 3.14159265358979323846f                             // include/mitsuba/core/constants.h:63,80
static const Float Epsilon = 1e-4f;
enum { EDebug = 0, EWarn = 1, EError = 2 };
namespace math {
    inline float fastexp(float value) { return (float) ::exp((double) value); }                                // math.h:185-187 (Linux / x86-64)
    template <typename Scalar> inline Scalar lerp(Scalar t, Scalar v1, Scalar v2) { return ((Scalar) 1 - t) * v1 + t * v2; }   // math.h:56-58
}
inline std::string indent(const std::string &s) { return s; }

class ContinuousSpectrum {
public:
    virtual Float eval(Float lambda) const = 0;
    virtual Float average(Float lambdaMin, Float lambdaMax) const;
    virtual std::string toString() const { return ""; }
    virtual ~ContinuousSpectrum() { }
};
class ProductSpectrum : public ContinuousSpectrum {
public:
    ProductSpectrum(const ContinuousSpectrum &s1, const ContinuousSpectrum &s2) : m_spec1(s1), m_spec2(s2) { }
    virtual Float eval(Float lambda) const;
private:
    const ContinuousSpectrum &m_spec1;
    const ContinuousSpectrum &m_spec2;
};
class InterpolatedSpectrum : public ContinuousSpectrum {
public:
    InterpolatedSpectrum(const Float *wavelengths, const Float *values, size_t nEntries);
    Float eval(Float lambda) const;
protected:
    std::vector<Float> m_wavelengths, m_values;
};
struct Spectrum {
    enum EConversionIntent { EReflectance, EIlluminant };
    Float s[3];
    Spectrum() { s[0] = s[1] = s[2] = 0; }
    void fromContinuousSpectrum(const ContinuousSpectrum &smooth);
    void fromXYZ(Float x, Float y, Float z, EConversionIntent intent = EReflectance);
    inline void clampNegative() { for (int i = 0; i < 3; i++) s[i] = std::max((Float) 0.0f, s[i]); }       // spectrum.h:537-540
};
class GaussLobattoIntegrator {                                    // include/mitsuba/core/quad.h:132-210 (declarations only)
public:
    typedef boost::function<Float (Float)> Integrand;
    GaussLobattoIntegrator(size_t maxEvals, Float absError = 0, Float relError = 0, bool useConvergenceEstimate = true, bool warn = true);
    Float integrate(const Integrand &f, Float a, Float b, size_t *evals = NULL) const;
protected:
    Float adaptiveGaussLobattoStep(const Integrand &f, Float a, Float b, Float fa, Float fb, Float is, size_t &evals) const;
    Float calculateAbsTolerance(const Integrand &f, Float a, Float b, size_t &evals) const;
    Float m_absError, m_relError;
    size_t m_maxEvals;
    bool m_useConvergenceEstimate;
    bool m_warn;
    static const Float m_alpha, m_beta, m_x1, m_x2, m_x3;
};

static const int CIE_samples = 471;
#include "cie_tables.inc"
static InterpolatedSpectrum CIE_X_interp(CIE_wavelengths, CIE_X_entries, CIE_samples);                     // spectrum.cpp:60-62
static InterpolatedSpectrum CIE_Y_interp(CIE_wavelengths, CIE_Y_entries, CIE_samples);
static InterpolatedSpectrum CIE_Z_interp(CIE_wavelengths, CIE_Z_entries, CIE_samples);

#include "ref_sun_quad.inc"
#include "ref_sun_spectrum.inc"
#include "ref_sun_model.inc"
} // namespace mitsuba

extern "C" void ref_sun_radiance(float theta, float turbidity, float *outRGB) {
    const mitsuba::Spectrum s = mitsuba::computeSunRadiance(theta, turbidity);
    outRGB[0] = s.s[0]; outRGB[1] = s.s[1]; outRGB[2] = s.s[2];
}

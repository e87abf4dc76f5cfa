// oracle/ref_shim/ref_bsdf.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's own BSDF plugin sources, which oracle/Makefile compiles unmodified from /root/reference
// against the scaffolding in fake_bsdf/mitsuba_shim.h:  kajiyakay.cpp, thindielectric.cpp, marschnerdielectric.cpp, marschner_diffuse.cpp
// (with its own microfacet.h, rtrans.h, ior.h, gausssexylingerie.hpp, InterpolatedDistribution1D.hpp) and src/libcore/spline.cpp.
// Output: oracle/_ref/libref_bsdf.so.  The oracle restatements (oracle/o_bsdf.h) are pinned against it in tests/test_oracle_cpu.py.
#include "mitsuba_shim.h"
#include <functional>

namespace boost { using std::function; using namespace std::placeholders; template <class... A> auto bind(A &&...a) { return std::bind(std::forward<A>(a)...); } }
using namespace std::placeholders;

namespace mitsuba {
#include "ref_fresnel.inc"          // src/libcore/util.cpp:592-601 (coordinateSystem) and :651-681 (fresnelDielectricExt), cut out at build time
class GaussLobattoIntegrator {      // include/mitsuba/core/quad.h:132-210 (declarations only); bodies: src/libcore/quad.cpp:287-420, cut out at build time
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
#include "ref_bsdf_quad.inc"
namespace {
#include "ref_fdr_integrand.inc"    // src/libcore/util.cpp:809-811 (the integrand; its anonymous namespace is supplied here)
}
#include "ref_fdr.inc"              // src/libcore/util.cpp:814-862 (fresnelDiffuseReflectance)
namespace math {
#include "ref_math.inc"             // src/libcore/math.cpp:25-86 (erfinv, erf, hypot2: what Beckmann visible-normal sampling runs on), cut out at build time
}
namespace warp {
#include "ref_warp.inc"             // src/libcore/warp.cpp:43-52 and :81-102, cut out at build time
}
}

using namespace mitsuba;

extern "C" {
void *ref_create_KajiyaKay(const Properties *);
void *ref_create_ThinDielectric(const Properties *);
void *ref_create_MarschnerDielectric(const Properties *);
void *ref_create_MarschnerDiffuse(const Properties *);
void *ref_create_Marschner(const Properties *);
void *ref_create_RoughPlastic(const Properties *);
void *ref_create_SmoothDiffuse(const Properties *);
void *ref_create_TwoSidedBRDF(const Properties *);
void *ref_create_SmoothPlastic(const Properties *);
void *ref_create_Mirror(const Properties *);
void *ref_create_MarschnerFull(const Properties *);           // oracle/_ref/marschner_full.cpp: marschner.cpp with the five edits listed in oracle/Makefile
float ref_fresnel_diffuse_reflectance(float eta) { return fresnelDiffuseReflectance(eta, false); }

// params: nFloat (name, value) pairs and nSpec (name, r, g, b) entries
void *ref_bsdf_create(const char *plugin, int nFloat, const char **floatNames, const float *floatValues, int nSpec, const char **specNames, const float *specValues) {
    try {
        Properties props(plugin);
        for (int i = 0; i < nFloat; ++i) {
            if (std::string(floatNames[i]) == "distribution") props.setString("distribution", floatValues[i] == 0 ? "beckmann" : floatValues[i] == 1 ? "ggx" : "phong");
            else props.setFloat(floatNames[i], floatValues[i]);
        }
        for (int i = 0; i < nSpec; ++i) { Spectrum s; for (int k = 0; k < 3; ++k) s[k] = specValues[3 * i + k]; props.setSpectrum(specNames[i], s); }
        const std::string p = plugin;
        BSDF *b = nullptr;
        if (p == "kajiyakay") b = (BSDF *) ref_create_KajiyaKay(&props);
        else if (p == "thindielectric") b = (BSDF *) ref_create_ThinDielectric(&props);
        else if (p == "marschnerdielectric") b = (BSDF *) ref_create_MarschnerDielectric(&props);
        else if (p == "marschner") b = (BSDF *) ref_create_MarschnerDiffuse(&props);        // the plugin named `marschner` is built from marschner_diffuse.cpp (src/bsdfs/SConscript:30-31)
        else if (p == "marschner_fixed") b = (BSDF *) ref_create_Marschner(&props);           // src/bsdfs/marschner.cpp, the class the fork's build leaves out
        else if (p == "roughplastic") b = (BSDF *) ref_create_RoughPlastic(&props);
        else if (p == "diffuse") b = (BSDF *) ref_create_SmoothDiffuse(&props);
        else if (p == "plastic") b = (BSDF *) ref_create_SmoothPlastic(&props);
        else if (p == "mirror") b = (BSDF *) ref_create_Mirror(&props);
        else if (p == "marschner_full") b = (BSDF *) ref_create_MarschnerFull(&props);
        else if (p == "twosided" || p == "twosided:plastic" || p == "twosided:roughplastic" || p == "twosided:mirror") {  // <bsdf type="twosided"><bsdf type="diffuse | plastic | roughplastic"/></bsdf>
            b = (BSDF *) ref_create_TwoSidedBRDF(&props);
            BSDF *nested = (BSDF *) (p == "twosided" ? ref_create_SmoothDiffuse(&props) : p == "twosided:plastic" ? ref_create_SmoothPlastic(&props) : p == "twosided:mirror" ? ref_create_Mirror(&props) : ref_create_RoughPlastic(&props));
            nested->configure();
            b->addChild("", nested);
        }
        else return nullptr;
        b->configure();
        return b;
    } catch (const std::exception &e) { fprintf(stderr, "ref_bsdf_create: %s\n", e.what()); return nullptr; } catch (...) { return nullptr; }
}
unsigned ref_bsdf_type(void *h) { return ((BSDF *) h)->getType(); }

// measure: 1 = ESolidAngle, 4 = EDiscrete (common.h:56-67); typeMask = EAll, component = -1 as BSDFSamplingRecord's ctor sets them
void ref_bsdf_eval(void *h, int n, const float *wi, const float *wo, int measure, float *outEval, float *outPdf) {
    const BSDF *b = (const BSDF *) h;
    for (int i = 0; i < n; ++i) {
        BSDFSamplingRecord r; r.typeMask = BSDF::EAll; r.component = -1;
        r.wi = Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]); r.wo = Vector(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        const Spectrum e = b->eval(r, (EMeasure) measure);
        for (int k = 0; k < 3; ++k) outEval[3 * i + k] = e[k];
        outPdf[i] = b->pdf(r, (EMeasure) measure);
    }
}
// extra (optional, 4 per tuple): what bRec.sampler hands out inside sample() (marschner.cpp:473-474)
void ref_bsdf_sample_ex(void *h, int n, const float *wi, const float *sample, const float *extra, float *outWo, float *outWeight, float *outPdf, int *outType);
void ref_bsdf_sample(void *h, int n, const float *wi, const float *sample, float *outWo, float *outWeight, float *outPdf, int *outType) {
    ref_bsdf_sample_ex(h, n, wi, sample, nullptr, outWo, outWeight, outPdf, outType);
}
void ref_bsdf_sample_ex(void *h, int n, const float *wi, const float *sample, const float *extra, float *outWo, float *outWeight, float *outPdf, int *outType) {
    const BSDF *b = (const BSDF *) h;
    Sampler sampler;
    for (int i = 0; i < n; ++i) {
        BSDFSamplingRecord r; r.typeMask = BSDF::EAll; r.component = -1; r.sampler = &sampler;
        sampler.pos = 0; sampler.queue.clear();
        if (extra) sampler.queue.assign(extra + 4 * i, extra + 4 * i + 4);
        r.wi = Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        Float pdf = 0;
        const Spectrum w = b->sample(r, pdf, Point2(sample[2 * i], sample[2 * i + 1]));
        outWo[3 * i] = r.wo.x; outWo[3 * i + 1] = r.wo.y; outWo[3 * i + 2] = r.wo.z;
        for (int k = 0; k < 3; ++k) outWeight[3 * i + k] = w[k];
        outPdf[i] = pdf; outType[i] = (int) r.sampledType | (r.sampledComponent << 8);
    }
}
}

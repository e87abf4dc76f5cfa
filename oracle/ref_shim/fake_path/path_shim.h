// oracle/ref_shim/fake_path/path_shim.h -- TEST INFRASTRUCTURE ONLY.
//
// Stand-ins for the librender declarations that src/integrators/path/path.cpp uses, so that the file compiles UNMODIFIED from
// /root/reference into oracle/_ref/libref_path.so.  Scene, Intersection, BSDF and Emitter forward to a callback table
// (ref_path_callbacks.h); nothing of the integrator's own logic lives here.  Mirrored semantics, with their reference lines:
//   RadianceQueryRecord::rayIntersect            include/mitsuba/render/records.inl:117-144 (no media on this path)
//   query type bits                              include/mitsuba/render/integrator.h:139-188
//   Ray / RayDifferential constructors           include/mitsuba/core/ray.h:42-93,130-168
//   MonteCarloIntegrator parameters              src/librender/integrator.cpp:190-225
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include "../ref_path_callbacks.h"

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_DECLARE_CLASS()
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" void *ref_create_##name(const mitsuba::Properties *props) { return new mitsuba::name(*props); }
#define EXPECT_TAKEN(x) (x)
#define EXPECT_NOT_TAKEN(x) (x)

namespace mitsuba {
using std::endl;
typedef float Float;
static const Float Epsilon = 1e-4f;
enum EMeasure { EInvalidMeasure = 0, ESolidAngle = 1, ELength = 2, EArea = 3, EDiscrete = 4 };
enum ETransportMode { ERadiance = 0, EImportance = 1 };
enum EStatsType { ENumberValue = 0, EByteCount, EPercentage, EMinimumValue, EMaximumValue, EAverage };
struct StatsCounter { StatsCounter(const char *, const char *, EStatsType) {} void incrementBase() {} StatsCounter &operator+=(int) { return *this; } };

struct Vector { Float x, y, z; Vector() : x(0), y(0), z(0) {} Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Vector operator-() const { return Vector(-x, -y, -z); } };
typedef Vector Point; typedef Vector Normal;
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} };

// include/mitsuba/core/spectrum.h (TSpectrum<Float, 3>): the operators path.cpp uses
struct Spectrum {
    Float s[3];
    Spectrum() { s[0] = s[1] = s[2] = 0; }
    explicit Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
    Spectrum operator*(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * o.s[i]; return r; }
    Spectrum operator*(Float f) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * f; return r; }
    Spectrum &operator+=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] += o.s[i]; return *this; }
    Spectrum &operator*=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] *= o.s[i]; return *this; }
    Spectrum &operator/=(Float f) { Float recip = 1.0f / f; for (int i = 0; i < 3; ++i) s[i] *= recip; return *this; }
    bool isZero() const { return s[0] == 0 && s[1] == 0 && s[2] == 0; }
    Float max() const { return std::max(s[0], std::max(s[1], s[2])); }
};

struct Ray { Point o; Vector d; Float mint, maxt, time;
    Ray() : mint(Epsilon), maxt(std::numeric_limits<Float>::infinity()), time(0) {}
    Ray(const Point &o, const Vector &d, Float time) : o(o), d(d), mint(Epsilon), maxt(std::numeric_limits<Float>::infinity()), time(time) {} };
struct RayDifferential : public Ray { Vector rxDirection, ryDirection; bool hasDifferentials;
    RayDifferential() : hasDifferentials(false) {}
    RayDifferential(const RayDifferential &r) = default;
    RayDifferential &operator=(const RayDifferential &r) = default;
    void operator=(const Ray &r) { o = r.o; d = r.d; mint = r.mint; maxt = r.maxt; time = r.time; hasDifferentials = false; } };   // ray.h:160-168

struct Frame { Vector s, t, n;
    Vector toLocal(const Vector &v) const { return Vector(dot(v, s), dot(v, t), dot(v, n)); }
    Vector toWorld(const Vector &v) const { return Vector(s.x * v.x + t.x * v.y + n.x * v.z, s.y * v.x + t.y * v.y + n.y * v.z, s.z * v.x + t.z * v.y + n.z * v.z); }   // s*v.x + t*v.y + n*v.z
    static Float cosTheta(const Vector &v) { return v.z; } };

class Scene; class Sampler; class Medium; class Emitter; class BSDF;
struct Intersection;
struct BSDFSamplingRecord {
    const Intersection &its; Sampler *sampler; Vector wi, wo; Float eta; ETransportMode mode; unsigned int typeMask; int component; unsigned int sampledType; int sampledComponent;
    inline BSDFSamplingRecord(const Intersection &its, Sampler *sampler, ETransportMode mode = ERadiance);                       // records.inl:24-30
    inline BSDFSamplingRecord(const Intersection &its, const Vector &wo, ETransportMode mode = ERadiance);                         // records.inl:32-37
};
class BSDF { public:
    enum { ENull = 0x1, EDiffuseReflection = 0x2, EDiffuseTransmission = 0x4, EGlossyReflection = 0x8, EGlossyTransmission = 0x10, EDeltaReflection = 0x20, EDeltaTransmission = 0x40,
           EDiffuse = EDiffuseReflection | EDiffuseTransmission, EGlossy = EGlossyReflection | EGlossyTransmission, ESmooth = EDiffuse | EGlossy,
           EDelta = ENull | EDeltaReflection | EDeltaTransmission };
    const RefPathCallbacks *cb; int id; mutable int depth;
    unsigned int getType() const { return cb->bsdfType(cb->user, id); }
    Spectrum eval(const BSDFSamplingRecord &b, EMeasure = ESolidAngle) const { Spectrum r; cb->bsdfEval(cb->user, id, &b.wi.x, &b.wo.x, r.s); return r; }
    Float pdf(const BSDFSamplingRecord &b, EMeasure = ESolidAngle) const { return cb->bsdfPdf(cb->user, id, &b.wi.x, &b.wo.x); }
    Spectrum sample(BSDFSamplingRecord &b, Float &pdf, const Point2 &sample) const {
        Spectrum w; cb->bsdfSample(cb->user, id, depth, &b.wi.x, sample.x, sample.y, &b.wo.x, w.s, &pdf, &b.sampledType, &b.eta); return w; }
};
struct Intersection {
    Point p; Float t; Frame geoFrame, shFrame; Vector wi; BSDF bsdf;
    Intersection() : t(std::numeric_limits<Float>::infinity()) {}
    bool isValid() const { return t != std::numeric_limits<Float>::infinity(); }
    const BSDF *getBSDF(const RayDifferential &) const { return &bsdf; }
    bool isEmitter() const { return false; } Spectrum Le(const Vector &) const { return Spectrum(0.0f); }
    bool hasSubsurface() const { return false; } Spectrum LoSub(const Scene *, Sampler *, const Vector &, int) const { return Spectrum(0.0f); }
    bool isMediumTransition() const { return false; }
    Vector toLocal(const Vector &v) const { return shFrame.toLocal(v); } Vector toWorld(const Vector &v) const { return shFrame.toWorld(v); }
};
inline BSDFSamplingRecord::BSDFSamplingRecord(const Intersection &its, Sampler *sampler, ETransportMode mode)
    : its(its), sampler(sampler), wi(its.wi), eta(1.0f), mode(mode), typeMask(0xffffffffu), component(-1), sampledType(0), sampledComponent(-1) {}
inline BSDFSamplingRecord::BSDFSamplingRecord(const Intersection &its, const Vector &wo, ETransportMode mode)
    : its(its), sampler(nullptr), wi(its.wi), wo(wo), eta(1.0f), mode(mode), typeMask(0xffffffffu), component(-1), sampledType(0), sampledComponent(-1) {}

struct DirectSamplingRecord { Point ref; Vector d; Float pdf; EMeasure measure; const void *object;
    DirectSamplingRecord(const Intersection &its) : ref(its.p), pdf(0), measure(ESolidAngle), object(nullptr) {}
    void setQuery(const Ray &, const Intersection &) {} };
class Emitter { public: const RefPathCallbacks *cb;
    bool isOnSurface() const { return true; }                                                                                   // envmap.cpp:107: EOnSurface
    Spectrum evalEnvironment(const RayDifferential &ray) const { Spectrum r; cb->evalEnvironment(cb->user, &ray.d.x, ray.hasDifferentials, &ray.rxDirection.x, &ray.ryDirection.x, r.s); return r; }
    bool fillDirectSamplingRecord(DirectSamplingRecord &dRec, const Ray &ray) const { dRec.d = ray.d; dRec.measure = ESolidAngle; dRec.object = this; return cb->fillDirectSamplingRecord(cb->user, &ray.o.x, &ray.d.x) != 0; } };
class Scene { public: const RefPathCallbacks *cb; Emitter env;
    bool rayIntersect(const Ray &ray, Intersection &its) const {
        RefPathIts r; const int hit = cb->rayIntersect(cb->user, &ray.o.x, &ray.d.x, ray.mint, ray.maxt, &r);
        its.t = hit ? r.t : std::numeric_limits<Float>::infinity();
        if (hit) { its.p = Point(r.p[0], r.p[1], r.p[2]); its.geoFrame.n = Vector(r.geoN[0], r.geoN[1], r.geoN[2]);
            its.shFrame.s = Vector(r.shS[0], r.shS[1], r.shS[2]); its.shFrame.t = Vector(r.shT[0], r.shT[1], r.shT[2]); its.shFrame.n = Vector(r.shN[0], r.shN[1], r.shN[2]);
            its.wi = Vector(r.wi[0], r.wi[1], r.wi[2]); its.bsdf.cb = cb; its.bsdf.id = r.bsdf; }
        return hit != 0; }
    Spectrum sampleEmitterDirect(DirectSamplingRecord &dRec, const Point2 &sample) const {
        Spectrum v; cb->sampleEmitterDirect(cb->user, &dRec.ref.x, sample.x, sample.y, v.s, &dRec.d.x, &dRec.pdf); dRec.measure = ESolidAngle; dRec.object = &env; return v; }
    Spectrum evalEnvironment(const RayDifferential &ray) const { return cb->hasEnvironment(cb->user) ? env.evalEnvironment(ray) : Spectrum(0.0f); }          // scene.h:871-876
    const Emitter *getEnvironmentEmitter() const { return cb->hasEnvironment(cb->user) ? &env : nullptr; }
    Float pdfEmitterDirect(const DirectSamplingRecord &dRec) const { return cb->pdfEmitterDirect(cb->user, &dRec.d.x); } };

struct RadianceQueryRecord {
    enum ERadianceQuery { EEmittedRadiance = 0x0001, ESubsurfaceRadiance = 0x0002, EDirectSurfaceRadiance = 0x0004, EIndirectSurfaceRadiance = 0x0008, ECausticRadiance = 0x0010,
        EDirectMediumRadiance = 0x0020, EIndirectMediumRadiance = 0x0040, EDistance = 0x0080, EOpacity = 0x0100, EIntersection = 0x0200,
        EVolumeRadiance = EDirectMediumRadiance | EIndirectMediumRadiance,
        ERadianceNoEmission = ESubsurfaceRadiance | EDirectSurfaceRadiance | EIndirectSurfaceRadiance | ECausticRadiance | EDirectMediumRadiance | EIndirectMediumRadiance | EIntersection,
        ERadiance = ERadianceNoEmission | EEmittedRadiance, ESensorRay = ERadiance | EOpacity };
    int type, depth; Intersection its; Float alpha, dist; const Scene *scene; Sampler *sampler; const Medium *medium; int extra;
    RadianceQueryRecord(const Scene *scene, Sampler *sampler) : type(0), depth(0), alpha(0), dist(-1), scene(scene), sampler(sampler), medium(nullptr), extra(0) {}
    void newQuery(int _type, const Medium *_medium) { type = _type; medium = _medium; depth = 1; extra = 0; alpha = 1; dist = -1; }                    // integrator.h:218-225
    // records.inl:117-144 without the medium branches
    bool rayIntersect(const RayDifferential &ray) {
        if (type & EIntersection) {
            scene->rayIntersect(ray, its);
            if (type & EOpacity) { if (its.isValid()) alpha = 1.0f; else alpha = 0.0f; }
            if (type & EDistance) dist = its.t;
            type ^= EIntersection;
        }
        return its.isValid();
    }
    // sampler draws: forwarded with the vertex they belong to.  path.cpp draws, per vertex and in this order: the emitter sample (only when the
    // BSDF has a smooth component), the BSDF sample, and -- from depth rrDepth on -- one number for the roulette, after depth was incremented.
    int lastDepth = -1, drawsAtDepth = 0;
    Point2 nextSample2D() {
        if (depth != lastDepth) { lastDepth = depth; drawsAtDepth = 0; }
        const bool smooth = (its.bsdf.getType() & BSDF::ESmooth) != 0;
        const int which = (smooth && drawsAtDepth == 0) ? 0 : 1;
        ++drawsAtDepth; its.bsdf.depth = depth;
        float o[2]; scene->cb->next2D(scene->cb->user, depth, which, o); return Point2(o[0], o[1]);
    }
    Float nextSample1D() { return scene->cb->next1D(scene->cb->user, depth - 1); }
};

class Stream; class InstanceManager;
class Properties { public: int maxDepth = -1, rrDepth = 5; bool strictNormals = false, hideEmitters = false;
    int getInteger(const std::string &n, int d) const { return n == "maxDepth" ? maxDepth : n == "rrDepth" ? rrDepth : d; }
    bool getBoolean(const std::string &n, bool d) const { return n == "strictNormals" ? strictNormals : n == "hideEmitters" ? hideEmitters : d; } };
// src/librender/integrator.cpp:190-225
class MonteCarloIntegrator { public:
    MonteCarloIntegrator(const Properties &props) { m_rrDepth = props.getInteger("rrDepth", 5); m_maxDepth = props.getInteger("maxDepth", -1);
        m_strictNormals = props.getBoolean("strictNormals", false); m_hideEmitters = props.getBoolean("hideEmitters", false);
        if (m_rrDepth <= 0) throw std::runtime_error("'rrDepth' must be set to a value greater than zero!");
        if (m_maxDepth <= 0 && m_maxDepth != -1) throw std::runtime_error("'maxDepth' must be set to -1 (infinite) or a value greater than zero!"); }
    MonteCarloIntegrator(Stream *, InstanceManager *) {}
    virtual ~MonteCarloIntegrator() {}
    virtual Spectrum Li(const RayDifferential &ray, RadianceQueryRecord &rRec) const = 0;
    void serialize(Stream *, InstanceManager *) const {}
protected:
    int m_maxDepth, m_rrDepth; bool m_strictNormals, m_hideEmitters;
};
} // namespace mitsuba

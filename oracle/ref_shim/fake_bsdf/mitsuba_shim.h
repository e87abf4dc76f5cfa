// oracle/ref_shim/fake_bsdf/mitsuba_shim.h -- TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the handful of libmitsuba-core / librender declarations that the reference's BSDF plugin sources use, so that
//   src/bsdfs/kajiyakay.cpp, src/bsdfs/thindielectric.cpp, src/bsdfs/marschnerdielectric.cpp
// compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/libref_bsdf.so (see oracle/Makefile).
// The real headers cannot be used: they pull in boost, OpenEXR and the rest of the library.
// Everything here is interface scaffolding; the arithmetic under test is the plugins' own.  Three helper functions the plugins call are
// not restated either: their bodies are cut out of src/libcore/util.cpp / warp.cpp at build time (oracle/_ref/*.inc).
// Semantics mirrored here, with the reference lines they follow:
//   BSDF::ensureEnergyConservation + the `scale` texture     src/librender/bsdf.cpp:88-146, src/textures/scale.cpp
//   ConstantSpectrumTexture / ConstantFloatTexture            include/mitsuba/hw/basicshader.h:33-120
//   BSDFSamplingRecord, EBSDFType, EMeasure                   include/mitsuba/render/bsdf.h:40-285, include/mitsuba/render/common.h:56-67
//   Spectrum::getLuminance (linear RGB, SPECTRUM_SAMPLES=3)   src/libcore/spectrum.cpp (0.212671, 0.715160, 0.072169)
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <array>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// Every plugin source is compiled with its own -DMTS_PLUGIN_NS=<name>: several of them define classes of the same name (three different
// `Azimuthal`s, two `MicrofacetDistribution` users ...), which must not be merged by the linker inside one shared object.
#ifdef MTS_PLUGIN_NS
#define MTS_NAMESPACE_BEGIN namespace mitsuba { namespace MTS_PLUGIN_NS {
#define MTS_NAMESPACE_END } }
#define MTS_PLUGIN_SCOPE mitsuba::MTS_PLUGIN_NS
#else
#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_PLUGIN_SCOPE mitsuba
#endif
#define MTS_EXPORT_CORE
#define MTS_DECLARE_CLASS()
#define MTS_IMPLEMENT_CLASS(name, abstract, super)
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#ifdef MTS_PLUGIN_EXPORT_NAME     /* a second build of one plugin source (oracle/Makefile: marschner_full) exports under its own name */
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" void *MTS_PLUGIN_EXPORT_NAME(const mitsuba::Properties *props) { return new MTS_PLUGIN_SCOPE::name(*props); }
#else
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" void *ref_create_##name(const mitsuba::Properties *props) { return new MTS_PLUGIN_SCOPE::name(*props); }
#endif
#define MTS_CLASS(x) (#x)
#define EXPECT_NOT_TAKEN(x) (x)
#define EXPECT_TAKEN(x) (x)
#define Assert(x) assert(x)
#define SLog(level, ...) mitsuba::shim_log(level, __VA_ARGS__)
#define Log(level, ...) mitsuba::shim_log(level, __VA_ARGS__)
#undef M_PI
#define M_PI        3.14159265358979323846f
#define M_PI_FLT    3.14159265358979323846f
#define INV_PI      0.31830988618379067154f
#define INV_TWOPI   0.15915494309189533577f
#define INV_FOURPI  0.07957747154594766788f

namespace boost { inline std::string to_lower_copy(std::string s) { for (auto &c : s) c = (char) std::tolower((unsigned char) c); return s; } }

namespace mitsuba {
using std::endl;
typedef float Float;
enum ELogLevel { ETrace, EDebug, EInfo, EWarn, EError };
inline void shim_log(ELogLevel level, const char *fmt, ...) { if (level >= EError) throw std::runtime_error(fmt); }
static const Float DeltaEpsilon = 1e-3f;     // include/mitsuba/core/constants.h:28-31
static const Float Epsilon = 1e-4f;

// include/mitsuba/core/vector.h (TVector3<float>, with the fork's component-wise product and max())
struct Vector {
    Float x, y, z;
    Vector() : x(0), y(0), z(0) {}
    Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    explicit Vector(Float v) : x(v), y(v), z(v) {}
    Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
    Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
    Vector &operator+=(const Vector &v) { x += v.x; y += v.y; z += v.z; return *this; }
    Vector &operator-=(const Vector &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    Vector operator*(Float f) const { return Vector(x * f, y * f, z * f); }
    Vector operator*(Vector v) const { return Vector(x * v.x, y * v.y, z * v.z); }
    Vector &operator*=(Float f) { x *= f; y *= f; z *= f; return *this; }
    Vector &operator*=(Vector v) { x *= v.x; y *= v.y; z *= v.z; return *this; }
    Vector operator-() const { return Vector(-x, -y, -z); }
    Vector operator/(Float f) const { Float recip = (Float) 1 / f; return Vector(x * recip, y * recip, z * recip); }
    Vector &operator/=(Float f) { Float recip = (Float) 1 / f; x *= recip; y *= recip; z *= recip; return *this; }
    Float &operator[](int i) { return (&x)[i]; }
    Float operator[](int i) const { return (&x)[i]; }
    Float lengthSquared() const { return x * x + y * y + z * z; }
    Float max() const { return std::max(x, std::max(y, z)); }
    Float length() const { return (Float) std::sqrt((Float) lengthSquared()); }
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
    std::string toString() const { std::ostringstream o; o << "[" << x << ", " << y << ", " << z << "]"; return o.str(); }
};
inline Vector operator*(Float f, const Vector &v) { return v * f; }                 // vector.h: `return v*f;`
inline Vector normalize(const Vector &v) { return v / v.length(); }                  // vector.h
typedef Vector Normal;
typedef Vector Point3; typedef Vector Point;
typedef Vector Vector3f;
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float absDot(const Vector &a, const Vector &b) { return std::abs(dot(a, b)); }
inline Vector cross(const Vector &a, const Vector &b) { return Vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }   // vector.h
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} explicit Point2(Float v) : x(v), y(v) {}
    Float &operator[](int i) { return (&x)[i]; } Float operator[](int i) const { return (&x)[i]; }
    Point2 operator+(const Point2 &p) const { return Point2(x + p.x, y + p.y); } };
struct Vector2 { Float x, y; Vector2() : x(0), y(0) {} Vector2(Float x, Float y) : x(x), y(y) {} explicit Vector2(Float v) : x(v), y(v) {} };
struct Size2 { size_t x, y; Size2() : x(0), y(0) {} Size2(size_t x, size_t y) : x(x), y(y) {} size_t &operator[](int i) { return (&x)[i]; } size_t operator[](int i) const { return (&x)[i]; } };
struct Size3 { size_t x, y, z; Size3() : x(0), y(0), z(0) {} Size3(size_t x, size_t y, size_t z) : x(x), y(y), z(z) {} size_t &operator[](int i) { return (&x)[i]; } size_t operator[](int i) const { return (&x)[i]; } };
void coordinateSystem(const Vector &a, Vector &b, Vector &c);      // body cut out of src/libcore/util.cpp:592-601 at build time
// include/mitsuba/core/frame.h:55-57,84-99
struct Frame {
    Vector s, t, n;
    Frame() {}
    Frame(const Vector &n) : n(n) { coordinateSystem(n, s, t); }
    Vector toWorld(const Vector &v) const { return s * v.x + t * v.y + n * v.z; }
    static Float cosTheta(const Vector &v) { return v.z; }
    static Float cosTheta2(const Vector &v) { return v.z * v.z; }
    static Float sinTheta2(const Vector &v) { return 1.0f - v.z * v.z; }
    static Float sinTheta(const Vector &v) { Float t = sinTheta2(v); if (t <= 0.0f) return 0.0f; return std::sqrt(t); }
    static Float tanTheta(const Vector &v) { Float t = 1.0f - v.z * v.z; if (t <= 0.0f) return 0.0f; return std::sqrt(t) / v.z; }
};

struct Spectrum {
    Float s[3];
    Spectrum() { s[0] = s[1] = s[2] = 0; }
    explicit Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
    explicit Spectrum(const Float v[3]) { s[0] = v[0]; s[1] = v[1]; s[2] = v[2]; }
    // TSpectrum(const TSpectrum<T2, N> &): component-wise conversion from another sample type, e.g. half (spectrum.h)
    template <typename S2> explicit Spectrum(const S2 &o, typename S2::IsSpectrumType * = nullptr) { s[0] = (Float) o.s[0]; s[1] = (Float) o.s[1]; s[2] = (Float) o.s[2]; }
    void fromLinearRGB(Float r, Float g, Float b) { s[0] = r; s[1] = g; s[2] = b; }        // spectrum.h, SPECTRUM_SAMPLES == 3
    Spectrum operator-(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] - o.s[i]; return r; }
    Spectrum operator/(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] / o.s[i]; return r; }
    Spectrum &operator*=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] *= o.s[i]; return *this; }
    Spectrum &operator/=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] /= o.s[i]; return *this; }
    Spectrum &operator/=(Float f) { Float recip = 1.0f / f; for (int i = 0; i < 3; ++i) s[i] *= recip; return *this; }
    Float min() const { return std::min(s[0], std::min(s[1], s[2])); }
    Float &operator[](int i) { return s[i]; } const Float &operator[](int i) const { return s[i]; }
    Spectrum operator+(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] + o.s[i]; return r; }
    Spectrum &operator+=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] += o.s[i]; return *this; }
    Spectrum operator*(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * o.s[i]; return r; }
    Spectrum operator*(Float f) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * f; return r; }
    Spectrum &operator*=(Float f) { for (int i = 0; i < 3; ++i) s[i] *= f; return *this; }
    Spectrum operator/(Float f) const { Spectrum r; Float recip = 1.0f / f; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * recip; return r; }   // spectrum.h: multiplies by the reciprocal
    bool isZero() const { return s[0] == 0 && s[1] == 0 && s[2] == 0; }
    Float max() const { return std::max(s[0], std::max(s[1], s[2])); }
    Float average() const { return (s[0] + s[1] + s[2]) * (1.0f / 3.0f); }
    Float getLuminance() const { return s[0] * 0.212671f + s[1] * 0.715160f + s[2] * 0.072169f; }
    std::string toString() const { std::ostringstream o; o << "[" << s[0] << ", " << s[1] << ", " << s[2] << "]"; return o.str(); }
};
inline Spectrum operator*(Float f, const Spectrum &s) { return s * f; }
inline std::string indent(const std::string &s) { return s; }

namespace math {
    inline Float safe_sqrt(Float v) { return std::sqrt(std::max((Float) 0, v)); }
    inline void sincos(Float t, Float *s, Float *c) { ::sincosf(t, s, c); }
    template <typename T> inline T clamp(T v, T lo, T hi) { return std::min(hi, std::max(lo, v)); }
    inline float fastexp(float value) { return (float) ::exp((double) value); }      // math.h:185-195: the Linux / x86-64 branch
    inline float fastlog(float value) { return (float) ::log((double) value); }
    inline Float signum(Float v) { return v < 0 ? (Float) -1 : (v > 0 ? (Float) 1 : (Float) 0); }
    float hypot2(float a, float b); Float erf(Float x); Float erfinv(Float x);   // bodies cut out of src/libcore/math.cpp:25-86 at build time (ref_bsdf.cpp)
    inline Float safe_acos(Float v) { return std::acos(std::min((Float) 1, std::max((Float) -1, v))); }
    inline Float safe_asin(Float v) { return std::asin(std::min((Float) 1, std::max((Float) -1, v))); }
}
static const Float RCPOVERFLOW = 2.93873587705571876e-39f;       // constants.h
#define SAssert(x) assert(x)
#define SIZE_T_FMT "%zu"
inline std::string formatString(const char *fmt, ...) { char buf[1024]; va_list a; va_start(a, fmt); vsnprintf(buf, sizeof(buf), fmt, a); va_end(a); return buf; }
inline std::string memString(size_t n) { return std::to_string(n) + " B"; }
using std::cout;

template <typename T> class ref {
public:
    ref() : p(nullptr) {} ref(T *q) : p(q) {}
    template <typename U> ref(const ref<U> &o) : p(o.get()) {}
    ref &operator=(T *q) { p = q; return *this; }
    T *operator->() const { return p; } T *get() const { return p; } operator T *() const { return p; }
    std::string toString() const { return p ? p->toString() : std::string("null"); }
private:
    T *p;                       // never freed: the shim lives for the duration of a test process
};

struct Intersection { Vector wi; };
class Stream { public:
    enum EByteOrder { EBigEndian = 0, ELittleEndian = 1 };
    virtual ~Stream() {}
    Float readFloat() { return 0; } void writeFloat(Float) {} bool readBool() { return false; } void writeBool(bool) {}
    unsigned int readUInt() { return 0; } void writeUInt(unsigned int) {} };
namespace fs { struct path { std::string p; path() {} path(const std::string &s) : p(s) {} path(const char *s) : p(s) {} std::string string() const { return p; }
    path filename() const { size_t k = p.rfind('/'); return path(k == std::string::npos ? p : p.substr(k + 1)); } }; }
// read-only little-endian binary file (fstream.h); the data files are little endian and so is the host
class FileStream : public Stream { public:
    enum EFileMode { EReadOnly = 0 };
    FileStream(const fs::path &path, EFileMode) : m_in(path.p, std::ios::binary) {
        if (!m_in) throw std::runtime_error("FileStream: cannot open " + path.p);
        m_in.seekg(0, std::ios::end); m_size = (size_t) m_in.tellg(); m_in.seekg(0); }
    void setByteOrder(EByteOrder) {}
    void read(void *dst, size_t n) { m_in.read((char *) dst, (std::streamsize) n); if (!m_in) throw std::runtime_error("FileStream: read past the end"); }
    char readChar() { char c; read(&c, 1); return c; }
    unsigned char readUChar() { unsigned char c; read(&c, 1); return c; }
    size_t readSize() { uint64_t v; read(&v, 8); return (size_t) v; }          // stream.h:367: readULong()
    Float readSingle() { float v; read(&v, 4); return v; }
    Float readFloat() { return readSingle(); }
    void readSingleArray(float *dst, size_t n) { read(dst, 4 * n); }
    void readFloatArray(float *dst, size_t n) { read(dst, 4 * n); }
    size_t getPos() { return (size_t) m_in.tellg(); } size_t getSize() const { return m_size; }
private:
    std::ifstream m_in; size_t m_size = 0; };
// FileResolver: "data/microfacet/x.dat" -> $REF_DATA_DIR/microfacet/x.dat (the mirrored run-time data, or /root/reference/data)
class FileResolver { public: fs::path resolve(const fs::path &p) const {
    const char *d = getenv("REF_DATA_DIR"); std::string rel = p.p; if (rel.compare(0, 5, "data/") == 0) rel = rel.substr(5);
    return fs::path(std::string(d ? d : "/root/reference/data") + "/" + rel); } };
class Thread { public: static Thread *getThread() { static Thread t; return &t; } FileResolver *getFileResolver() { static FileResolver r; return &r; } };
class ConfigurableObject;
class InstanceManager { public: ConfigurableObject *getInstance(Stream *) { return nullptr; } void serialize(Stream *, const ConfigurableObject *) {} };
struct ShimClass { std::string name; bool derivesFrom(const char *other) const { return name == other; } };

class Properties {
public:
    enum EPropertyType { EBoolean, EInteger, EFloat, EString, ESpectrum };
    Properties(const std::string &plugin = "") : m_plugin(plugin) {}
    void setFloat(const std::string &n, Float v) { m_float[n] = v; }
    void setString(const std::string &n, const std::string &v) { m_string[n] = v; }
    void setSpectrum(const std::string &n, const Spectrum &v) { m_spec[n] = v; }
    bool hasProperty(const std::string &n) const { return m_float.count(n) || m_string.count(n) || m_spec.count(n); }
    EPropertyType getType(const std::string &n) const { return m_float.count(n) ? EFloat : m_string.count(n) ? EString : ESpectrum; }
    Float getFloat(const std::string &n) const { return m_float.at(n); }
    Float getFloat(const std::string &n, Float d) const { auto it = m_float.find(n); return it == m_float.end() ? d : it->second; }
    std::string getString(const std::string &n) const { return m_string.at(n); }
    std::string getString(const std::string &n, const std::string &d) const { auto it = m_string.find(n); return it == m_string.end() ? d : it->second; }
    Spectrum getSpectrum(const std::string &n) const { return m_spec.at(n); }
    Spectrum getSpectrum(const std::string &n, const Spectrum &d) const { auto it = m_spec.find(n); return it == m_spec.end() ? d : it->second; }
    bool getBoolean(const std::string &n, bool d) const { auto it = m_float.find(n); return it == m_float.end() ? d : it->second != 0; }     // booleans ride as floats in this scaffolding
private:
    std::string m_plugin; std::map<std::string, Float> m_float; std::map<std::string, std::string> m_string; std::map<std::string, Spectrum> m_spec;
};

class Object { public: virtual ~Object() {} };
class ConfigurableObject {
public:
    ConfigurableObject() {}
    explicit ConfigurableObject(const Properties &) {}
    ConfigurableObject(Stream *, InstanceManager *) {}
    virtual ~ConfigurableObject() {}
    virtual void configure() {}
    virtual void addChild(const std::string &, ConfigurableObject *) {}
    virtual const ShimClass *getClass() const { static ShimClass c{"ConfigurableObject"}; return &c; }
    virtual std::string toString() const { return "object"; }
    std::string getID() const { return "unnamed"; }
};

class Texture : public ConfigurableObject {
public:
    virtual Spectrum eval(const Intersection &its, bool filter = true) const = 0;
    virtual Spectrum getAverage() const = 0;
    virtual Spectrum getMaximum() const = 0;
    virtual Spectrum getMinimum() const { return getMaximum(); }      // constant textures only
    virtual bool isConstant() const { return true; }
    virtual bool usesRayDifferentials() const { return false; }
    const ShimClass *getClass() const override { static ShimClass c{"Texture"}; return &c; }
};
class ConstantSpectrumTexture : public Texture {
public:
    explicit ConstantSpectrumTexture(const Spectrum &v) : m_value(v) {}
    Spectrum eval(const Intersection &, bool) const override { return m_value; }
    Spectrum getAverage() const override { return m_value; }
    Spectrum getMaximum() const override { return m_value; }
    std::string toString() const override { return m_value.toString(); }
private:
    Spectrum m_value;
};
class ConstantFloatTexture : public Texture {
public:
    explicit ConstantFloatTexture(Float v) : m_value(v) {}
    Spectrum eval(const Intersection &, bool) const override { return Spectrum(m_value); }
    Spectrum getAverage() const override { return Spectrum(m_value); }
    Spectrum getMaximum() const override { return Spectrum(m_value); }
private:
    Float m_value;
};
// src/textures/scale.cpp: nested texture times a constant
class ScaleTexture : public Texture {
public:
    ScaleTexture(Texture *nested, Float scale) : m_nested(nested), m_scale(scale) {}
    Spectrum eval(const Intersection &its, bool f) const override { return m_nested->eval(its, f) * m_scale; }
    Spectrum getAverage() const override { return m_nested->getAverage() * m_scale; }
    Spectrum getMaximum() const override { return m_nested->getMaximum() * m_scale; }
private:
    Texture *m_nested; Float m_scale;
};

class PhaseFunction : public ConfigurableObject { public: const ShimClass *getClass() const override { static ShimClass c{"PhaseFunction"}; return &c; } };
class PluginManager { public: static PluginManager *getInstance() { static PluginManager m; return &m; }
    ConfigurableObject *createObject(const char *, const Properties &) { return new PhaseFunction(); } };        // only ever asked for the `kkay` phase function, which is never evaluated
// src/bsdfs/marschner_diffuse.cpp:131 assigns to `m_exponent`, which the class never declares -- the plugin does not compile as committed
// (SURVEY.md fact 2).  A namespace-scope object of that name lets the unmodified file compile; nothing ever reads it.
static ref<Texture> m_exponent;
enum EMeasure { EInvalidMeasure = 0, ESolidAngle = 1, ELength = 2, EArea = 3, EDiscrete = 4 };
enum ETransportMode { ERadiance = 0, EImportance = 1 };
// Sampler: hands out the numbers the test harness queued (marschner.cpp:473-474 draws two more 2-D samples inside sample())
class Sampler { public:
    std::vector<Float> queue; size_t pos = 0;
    Float next1D() { return pos < queue.size() ? queue[pos++] : 0.5f; }
    Point2 next2D() { Float a = next1D(), b = next1D(); return Point2(a, b); }       // sampler.h: two consecutive 1-D draws (independent.cpp:98-103)
};
class GPUProgram { public:
    int getParameterID(const std::string &, bool = true) const { return 0; }
    template <typename T> void setParameter(int, const T &) {}
};
class Renderer;
class Shader;

struct BSDFSamplingRecord {
    Intersection its; Sampler *sampler = nullptr;
    Vector wi, wo; Float eta = 1; ETransportMode mode = ERadiance;
    unsigned int typeMask = 0xffffffffu; int component = -1;
    unsigned int sampledType = 0; int sampledComponent = -1;
};

class BSDF : public ConfigurableObject {
public:
    enum EBSDFType {
        ENull = 0x00001, EDiffuseReflection = 0x00002, EDiffuseTransmission = 0x00004, EGlossyReflection = 0x00008, EGlossyTransmission = 0x00010,
        EDeltaReflection = 0x00020, EDeltaTransmission = 0x00040, EDelta1DReflection = 0x00080, EDelta1DTransmission = 0x00100,
        EAnisotropic = 0x01000, ESpatiallyVarying = 0x02000, ENonSymmetric = 0x04000, EFrontSide = 0x08000, EBackSide = 0x10000, EUsesSampler = 0x20000
    };
    enum ETypeCombinations {
        EReflection = EDiffuseReflection | EDeltaReflection | EDelta1DReflection | EGlossyReflection,
        ETransmission = EDiffuseTransmission | EDeltaTransmission | EDelta1DTransmission | EGlossyTransmission | ENull,
        EDiffuse = EDiffuseReflection | EDiffuseTransmission, EGlossy = EGlossyReflection | EGlossyTransmission, ESmooth = EDiffuse | EGlossy,
        EDelta = ENull | EDeltaReflection | EDeltaTransmission, EDelta1D = EDelta1DReflection | EDelta1DTransmission,
        EAll = EDiffuse | EGlossy | EDelta | EDelta1D
    };
    static constexpr const char *m_theClass = "BSDF";
    const ShimClass *getClass() const override { static ShimClass c{"BSDF"}; return &c; }
    explicit BSDF(const Properties &props) : m_ensureEnergyConservation(props.getBoolean("ensureEnergyConservation", true)) {}
    BSDF(Stream *, InstanceManager *) {}
    void configure() override { m_combinedType = 0; for (unsigned c : m_components) m_combinedType |= c; }      // bsdf.cpp:56-70
    void addChild(const std::string &, ConfigurableObject *) override {}
    virtual void serialize(Stream *, InstanceManager *) const {}
    unsigned int getType() const { return m_combinedType; }
    unsigned int getType(int component) const { return m_components[component]; }
    int getComponentCount() const { return (int) m_components.size(); }
    bool usesRayDifferentials() const { return m_usesRayDifferentials; }
    virtual Spectrum sample(BSDFSamplingRecord &bRec, const Point2 &sample) const = 0;
    virtual Float getEta() const { return 1.0f; }
    virtual Float getRoughness(const Intersection &, int) const { return 0.0f; }
    virtual Spectrum getDiffuseReflectance(const Intersection &) const { return Spectrum(0.0f); }
    virtual Shader *createShader(Renderer *) const { return nullptr; }
    virtual Spectrum eval(const BSDFSamplingRecord &bRec, EMeasure measure) const = 0;
    virtual Float pdf(const BSDFSamplingRecord &bRec, EMeasure measure) const = 0;
    virtual Spectrum sample(BSDFSamplingRecord &bRec, Float &pdf, const Point2 &sample) const = 0;
    // src/librender/bsdf.cpp:88-113
    Texture *ensureEnergyConservation(Texture *texture, const std::string &, Float max) const {
        if (!m_ensureEnergyConservation) return texture;
        Float actualMax = texture->getMaximum().max();
        if (actualMax > max) { Float scale = 0.99f * (max / actualMax); return new ScaleTexture(texture, scale); }
        return texture;
    }
    // src/librender/bsdf.cpp:115-146
    std::pair<Texture *, Texture *> ensureEnergyConservation(Texture *tex1, Texture *tex2, const std::string &, const std::string &, Float max) const {
        if (!m_ensureEnergyConservation) return std::make_pair(tex1, tex2);
        Float actualMax = (tex1->getMaximum() + tex2->getMaximum()).max();
        if (actualMax > max) { Float scale = 0.99f * (max / actualMax); return std::make_pair((Texture *) new ScaleTexture(tex1, scale), (Texture *) new ScaleTexture(tex2, scale)); }
        return std::make_pair(tex1, tex2);
    }
protected:
    std::vector<unsigned int> m_components; unsigned int m_combinedType = 0;
    bool m_usesRayDifferentials = false, m_ensureEnergyConservation = true;
};

// hardware shaders (OpenGL preview): declared so that the plugins' shader classes compile; never instantiated
class Shader : public ConfigurableObject {
public:
    enum EShaderType { EBSDFShader = 0 }; enum EFlags { ETransparent = 1 };
    Shader(Renderer *, EShaderType) {}
    virtual Float getAlpha() const { return 1; }
    virtual bool isComplete() const { return true; }
    virtual void putDependencies(std::vector<Shader *> &) {}
    virtual void cleanup(Renderer *) {}
    virtual void generateCode(std::ostringstream &, const std::string &, const std::vector<std::string> &) const {}
protected:
    int m_flags = 0;
};
class Renderer { public: Shader *registerShaderForResource(const ConfigurableObject *) { return nullptr; } void unregisterShaderForResource(const ConfigurableObject *) {} };

// src/libcore/spline.cpp is compiled as it is; these are its declarations (include/mitsuba/core/spline.h:58-306)
Float evalCubicInterp1D(Float x, const Float *values, size_t size, Float min, Float max, bool extrapolate = false);
Float evalCubicInterp1DN(Float x, const Float *nodes, const Float *values, size_t size, bool extrapolate = false);
Float integrateCubicInterp1D(size_t idx, const Float *values, size_t size, Float min, Float max);
Float integrateCubicInterp1DN(size_t idx, const Float *nodes, const Float *values, size_t size);
Float sampleCubicInterp1D(size_t idx, const Float *values, size_t size, Float min, Float max, Float sample, Float *fval = NULL);
Float sampleCubicInterp1DN(size_t idx, const Float *nodes, const Float *values, size_t size, Float sample, Float *fval = NULL);
Float evalCubicInterp2D(const Point2 &p, const Float *values, const Size2 &size, const Point2 &min, const Point2 &max, bool extrapolate = false);
Float evalCubicInterp2DN(const Point2 &p, const Float **nodes, const Float *values, const Size2 &size, bool extrapolate = false);
Float evalCubicInterp3D(const Point3 &p, const Float *values, const Size3 &size, const Point3 &min, const Point3 &max, bool extrapolate = false);
Float evalCubicInterp3DN(const Point3 &p, const Float **nodes, const Float *values, const Size3 &size, bool extrapolate = false);

// bodies cut out of the reference at build time (oracle/Makefile): src/libcore/util.cpp:651-681, src/libcore/warp.cpp:43-52,81-102
Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta);
inline Float fresnelDielectricExt(Float cosThetaI, Float eta) { Float cosThetaT; return fresnelDielectricExt(cosThetaI, cosThetaT, eta); }   // util.h:479-480
Float fresnelDiffuseReflectance(Float eta, bool fast = false);                                                                              // util.h:593-594; body cut out of util.cpp:814-862 (ref_bsdf.cpp)
namespace warp {
    Point2 squareToUniformDiskConcentric(const Point2 &sample);
    Vector squareToCosineHemisphere(const Point2 &sample);
    inline Float squareToCosineHemispherePdf(const Vector &d) { return INV_PI * Frame::cosTheta(d); }    // warp.h:55-56
}
} // namespace mitsuba

// oracle/ref_shim/ref_loader.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's hair file loader executed as written: oracle/Makefile cuts the constructor HairShape::HairShape(const Properties &)
// (src/shapes/hair.cpp:609-785: binary and ASCII formats, toWorld, radius scaling, the angle-threshold vertex merge, the sentinel) out
// of /root/reference at build time (oracle/_ref/ref_hair_loader.inc) and this file supplies the few types its text touches.
// Point and Vector are distinct types here (the loader transforms points affinely and one vector linearly), unlike in mitsuba_shim.h.
// Output: part of oracle/_ref/libref_geom.so.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace reflsdr {              // own namespace: nothing here may clash with the scaffolding of the other translation units
typedef float Float;
#undef M_PI
#define M_PI 3.14159265358979323846f
#define SIZE_T_FMT "%zu"
enum ELogLevel { ETrace, EDebug, EInfo, EWarn, EError };
inline void logShim(ELogLevel level, const char *fmt, ...) { if (level >= EError) throw std::runtime_error(fmt); }
#define Log(level, ...) logShim(level, __VA_ARGS__)

// include/mitsuba/core/vector.h / point.h (TVector3<float>, TPoint3<float>): the members the loader uses
struct Vector {
    Float x, y, z;
    Vector() : x(0), y(0), z(0) {}
    Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    explicit Vector(Float v) : x(v), y(v), z(v) {}
    Vector operator/(Float f) const { Float recip = (Float) 1 / f; return Vector(x * recip, y * recip, z * recip); }
    Float lengthSquared() const { return x * x + y * y + z * z; }
    Float length() const { return (Float) std::sqrt((Float) lengthSquared()); }
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
};
struct Point {
    Float x, y, z;
    Point() : x(0), y(0), z(0) {}
    Point(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    explicit Point(Float v) : x(v), y(v), z(v) {}
    Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
    Point operator/(Float f) const { Float recip = (Float) 1 / f; return Point(x * recip, y * recip, z * recip); }
    bool operator!=(const Point &p) const { return p.x != x || p.y != y || p.z != z; }
};
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector normalize(const Vector &v) { return v / v.length(); }
inline Float degToRad(Float value) { return value * (M_PI / 180.0f); }          // util.h:297

// include/mitsuba/core/transform.h:108-125,175-183
struct Transform {
    Float m[4][4];
    Transform() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = i == j ? 1.0f : 0.0f; }
    Point operator()(const Point &p) const {
        Float x = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
        Float y = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
        Float z = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
        Float w = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
        if (w == 1.0f) return Point(x, y, z); else return Point(x, y, z) / w;
    }
    Vector operator()(const Vector &v) const {
        Float x = m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z;
        Float y = m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z;
        Float z = m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z;
        return Vector(x, y, z);
    }
};

template <typename T> struct ref { T *p; ref(T *q = nullptr) : p(q) {} ref &operator=(T *q) { p = q; return *this; } T *operator->() const { return p; } };
namespace fs {
    struct path { std::string s; path() {} path(const std::string &v) : s(v) {} std::string string() const { return s; } path filename() const { return *this; } };
    struct ifstream : public std::ifstream { explicit ifstream(const path &p) : std::ifstream(p.s) {} };
}
struct FileResolver { fs::path resolve(const std::string &s) const { return fs::path(s); } };
struct Thread { static Thread *getThread() { static Thread t; return &t; } FileResolver *getFileResolver() { static FileResolver r; return &r; } };
// Random: the SFMT-19937 state, its recursion and seeding cut out of src/libcore/random.cpp at build time (ref_random_*.inc: parameters :72-96,
// helpers :117-220 with the non-SSE branch, struct Random::State :225-392, init_gen_rand :396-405); the four members the loader reaches are declared here
#define MTS_SFMT_SSE 0
#define FINLINE inline
#define SAssert(x) ((void) 0)
struct RStream { void readULongArray(uint64_t *, size_t) {} int readInt() { return 0; } void writeULongArray(const uint64_t *, size_t) const {} void writeInt(int) const {} };
#define Stream RStream
#include "ref_random_params.inc"
#include "ref_random_helpers.inc"
struct Random {
    struct State;
    State *mt;
    Random();
    uint64_t nextULong();
    Float nextFloat();
};
#include "ref_random_state.inc"
const uint32_t Random::State::parity[4] = {PARITY1, PARITY2, PARITY3, PARITY4};
#include "ref_random_init.inc"
#undef Stream
#undef N
Random::Random() : mt(new State()) { mt->init_gen_rand(5489ULL); }              // Random() -> seed() with the default of include/mitsuba/core/random.h:113
uint64_t Random::nextULong() { return mt->gen_rand64(); }                        // random.cpp:551-553
Float Random::nextFloat() { union { uint32_t u; float f; } x; x.u = ((nextULong() & 0xFFFFFFFF) >> 9) | 0x3f800000UL; return x.f - 1.0f; }   // :630-639
struct Timer { int getMilliseconds() const { return 0; } };
struct Stream { enum EByteOrder { EBigEndian, ELittleEndian }; };
struct FileStream : public Stream {
    enum EFileMode { EReadOnly };
    std::ifstream in;
    FileStream(const fs::path &p, EFileMode) : in(p.s, std::ios::binary) { if (!in) throw std::runtime_error("cannot open " + p.s); }
    void setByteOrder(EByteOrder) {}
    void read(void *dst, size_t n) { in.read((char *) dst, (std::streamsize) n); if (in.gcount() != (std::streamsize) n) throw std::runtime_error("read past the end"); }     // fstream.cpp: EOFException
    uint32_t readUInt() { uint32_t v; read(&v, 4); return v; }
    float readSingle() { float v; read(&v, 4); return v; }
};
struct Properties {
    std::map<std::string, Float> f; std::map<std::string, std::string> s; Transform t;
    std::string getString(const std::string &n) const { return s.at(n); }
    Float getFloat(const std::string &n, Float d) const { auto it = f.find(n); return it == f.end() ? d : it->second; }
    Transform getTransform(const std::string &, const Transform &) const { return t; }
};
struct Shape { explicit Shape(const Properties &) {} };
struct HairKDTree {
    std::vector<Point> vertices; std::vector<bool> starts; Float radius;
    HairKDTree(const std::vector<Point> &v, const std::vector<bool> &s, Float r) : vertices(v), starts(s), radius(r) {}
};
struct HairShape : public Shape {
    ref<HairKDTree> m_kdtree;
    HairShape(const Properties &props);
};
#include "ref_hair_loader.inc"
} // namespace reflsdr

using namespace reflsdr;
extern "C" {
// returns a handle or null (message in *err, 256 bytes); sizes / arrays through the accessors (sentinel excluded like the product's loader)
void ref_random_floats(int n, float *out) { Random r; for (int i = 0; i < n; ++i) out[i] = r.nextFloat(); }
void *ref_hair_load_reduced(const char *path, float radius, float angleThresholdDeg, const float *toWorld16, float reduction, char *err);
void *ref_hair_load(const char *path, float radius, float angleThresholdDeg, const float *toWorld16, char *err) { return ref_hair_load_reduced(path, radius, angleThresholdDeg, toWorld16, 0.0f, err); }
void *ref_hair_load_reduced(const char *path, float radius, float angleThresholdDeg, const float *toWorld16, float reduction, char *err) {
    try {
        Properties props; props.s["filename"] = path; props.f["radius"] = radius; props.f["angleThreshold"] = angleThresholdDeg; props.f["reduction"] = reduction;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) props.t.m[i][j] = toWorld16[4 * i + j];
        HairShape *h = new HairShape(props);
        return h->m_kdtree.p;
    } catch (const std::exception &e) { if (err) std::snprintf(err, 256, "%s", e.what()); return nullptr; }
}
int ref_hair_load_count(void *h) { return (int) ((HairKDTree *) h)->vertices.size(); }
float ref_hair_load_radius(void *h) { return ((HairKDTree *) h)->radius; }
void ref_hair_load_copy(void *h, float *xyz, unsigned char *starts) {
    const HairKDTree *k = (const HairKDTree *) h;
    for (size_t i = 0; i < k->vertices.size(); ++i) { xyz[3 * i] = k->vertices[i].x; xyz[3 * i + 1] = k->vertices[i].y; xyz[3 * i + 2] = k->vertices[i].z; starts[i] = k->starts[i] ? 1 : 0; }
}
}

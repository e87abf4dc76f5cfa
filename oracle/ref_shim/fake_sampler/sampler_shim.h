// oracle/ref_shim/fake_sampler/sampler_shim.h -- TEST INFRASTRUCTURE ONLY.
//
// Stand-ins for the libmitsuba interfaces that src/samplers/sobol.cpp and src/samplers/sobolseq.{h,cpp} touch, so that those three files compile
// UNMODIFIED from where they lie under /root/reference (oracle/Makefile: libref_sobol.so).  Declarations only -- none of the sampler's arithmetic:
//   Sampler base class                include/mitsuba/render/sampler.h:66-181 (members the plugin reads and writes; the sample-array bookkeeping of
//                                     src/librender/sampler.cpp:44-118 reduced to what generate() needs)
//   sampleTEA                         include/mitsuba/core/qmc.h:146-156 (cut out of the reference at build time: ref_sobol_tea.inc)
//   math::roundToPowerOfTwo / log2i   src/libcore/math.cpp:108-134 (cut out at build time: ref_sobol_math.inc)
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_EXPORT_RENDER
#define MTS_EXPORT_CORE
#define MTS_DECLARE_CLASS()
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" void *ref_create_##name(const mitsuba::Properties *props) { return new mitsuba::name(*props); }
#define SINGLE_PRECISION 1
#define ONE_MINUS_EPS_FLT 0.999999940395355225f          /* include/mitsuba/core/constants.h:51 */
#define ONE_MINUS_EPS_DBL 0.999999999999999888
using std::endl;

namespace mitsuba {
typedef float Float;
enum ELogLevel { ETrace, EDebug, EInfo, EWarn, EError };
#define Log(level, ...) do { if ((level) >= mitsuba::EError) { char b_[512]; snprintf(b_, sizeof(b_), __VA_ARGS__); throw std::runtime_error(b_); } } while (0)
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} };
struct Point2i { int x, y; Point2i() : x(0), y(0) {} explicit Point2i(int v) : x(v), y(v) {} Point2i(int x, int y) : x(x), y(y) {} };
struct Vector2i { int x, y; Vector2i() : x(0), y(0) {} Vector2i(int x, int y) : x(x), y(y) {} };
template <typename T> struct ref { T *p; ref(T *q = nullptr) : p(q) {} T *get() const { return p; } T *operator->() const { return p; } };
struct Properties { std::map<std::string, size_t> sizes; size_t getSize(const std::string &n, size_t d) const { auto it = sizes.find(n); return it == sizes.end() ? d : it->second; } };
struct Stream { uint64_t readULong() { return 0; } Float readFloat() { return 0; } unsigned readUInt() { return 0; } size_t readSize() { return 0; }
                void writeULong(uint64_t) {} void writeFloat(Float) {} void writeUInt(unsigned) {} void writeSize(size_t) {} };
struct InstanceManager {};
struct ConfigurableObject { virtual ~ConfigurableObject() {} };
namespace math { namespace {            /* internal linkage: the bodies are definitions and this header is seen by three translation units */
#include "ref_sobol_math.inc"
} }
#include "ref_sobol_tea.inc"

class Sampler : public ConfigurableObject {
public:
    virtual ref<Sampler> clone() { return nullptr; }
    virtual void setFilmResolution(const Vector2i &, bool) {}
    virtual void generate(const Point2i &) {}
    virtual void advance() {}
    virtual void setSampleIndex(size_t) {}
    virtual Float next1D() = 0;
    virtual Point2 next2D() = 0;
    virtual void request2DArray(size_t size) { m_req2D.push_back(size); m_sampleArrays2D.push_back(new Point2[m_sampleCount * size]); }    // sampler.cpp:95-100
    virtual void request1DArray(size_t size) { m_req1D.push_back(size); m_sampleArrays1D.push_back(new Float[m_sampleCount * size]); }     // sampler.cpp:88-93
    size_t getSampleCount() const { return m_sampleCount; }
    virtual void serialize(Stream *, InstanceManager *) const {}
protected:
    Sampler(const Properties &) : m_sampleCount(0), m_sampleIndex(0), m_dimension1DArray(0), m_dimension2DArray(0) {}
    Sampler(Stream *, InstanceManager *) : m_sampleCount(0), m_sampleIndex(0), m_dimension1DArray(0), m_dimension2DArray(0) {}
    size_t m_sampleCount, m_sampleIndex;
    std::vector<size_t> m_req1D, m_req2D;
    std::vector<Float *> m_sampleArrays1D;
    std::vector<Point2 *> m_sampleArrays2D;
    size_t m_dimension1DArray, m_dimension2DArray;
};
} // namespace mitsuba

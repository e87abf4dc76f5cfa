// oracle/ref_shim/ref_sunsplat.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The sun-disc rasterisation of SunSkyEmitter executed as written: oracle/Makefile cuts these pieces out of /root/reference at build time
// (oracle/_ref/ref_sun_*.inc) and this file supplies the types and locals they touch:
//   src/emitters/sunsky.cpp :180-211 the QMC splat of the `sunRadiusScale != 0` branch (sample count, per-sample value, texel, 1/sin weight)
//   src/emitters/sunsky/sunmodel.h :90-105 toSphere / fromSphere
//   src/libcore/warp.cpp :54-63 squareToUniformCone;  include/mitsuba/core/qmc.h :43-60,82-87,115-120 radicalInverse2Single, sobol2Single, sample02
//   src/libcore/util.cpp :592-601 coordinateSystem (Frame(n), frame.h)
// The sun radiance (computeSunRadiance -> RGB) and the sun position are inputs.  Part of oracle/_ref/libref_geom.so.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace refsun {
typedef float Float;
#undef M_PI
#define M_PI 3.14159265358979323846f
#define SINGLE_PRECISION_SUN 1
#define MTS_EXPORT_CORE
#define SUN_APP_RADIUS 0.5358                                                               // sunmodel.h:28
struct Vector { Float x, y, z; Vector() : x(0), y(0), z(0) {} Vector(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Vector operator*(Float f) const { return Vector(x * f, y * f, z * f); }
    Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); } };
inline Vector cross(const Vector &v1, const Vector &v2) { return Vector(v1.y * v2.z - v1.z * v2.y, v1.z * v2.x - v1.x * v2.z, v1.x * v2.y - v1.y * v2.x); }   // vector.h
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} };
struct Point2i { int x, y; Point2i(int x, int y) : x(x), y(y) {} };
struct Spectrum { Float s[3];                                                               // spectrum.h, SPECTRUM_SAMPLES = 3
    Spectrum operator*(Float f) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * f; return r; }
    Spectrum operator/(Float f) const { Spectrum r; Float recip = 1.0f / f; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * recip; return r; }
    Spectrum &operator+=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] += o.s[i]; return *this; } };
namespace math {
    inline float safe_acos(float value) { return std::acos(std::min(1.0f, std::max(-1.0f, value))); }   // math.h:250-252
    inline float safe_sqrt(float value) { return std::sqrt(std::max(0.0f, value)); }                   // math.h:260-262
    inline void sincos(float theta, float *s, float *c) { ::sincosf(theta, s, c); }                    // math.h (glibc build)
}
inline Float degToRad(Float value) { return value * (M_PI / 180.0f); }                      // util.h:297
struct SphericalCoordinates { Float elevation, azimuth; SphericalCoordinates() {} SphericalCoordinates(Float e, Float a) : elevation(e), azimuth(a) {} };   // sunmodel.h:63-70
#include "ref_sun_sphere.inc"
#include "ref_sun_coordsys.inc"
struct Frame { Vector s, t, n;
    Frame(const Vector &n) : n(n) { coordinateSystem(n, s, t); }                            // frame.h:59-61
    Vector toWorld(const Vector &v) const { return s * v.x + t * v.y + n * v.z; } };        // frame.h:95-97
namespace warp {
#include "ref_sun_cone.inc"
}
#define SINGLE_PRECISION 1
#include "ref_sun_qmc.inc"
struct BitmapDims { int w, h; int getWidth() const { return w; } int getHeight() const { return h; } };
}

// data: resolution x resolution/2 RGB texels, accumulated in place.  The sun position follows computeSunCoordinates(sunDir, identity)
// (sunmodel.h:206-208: fromSphere(normalize(sunDir))), then `sun.elevation *= stretch` (sunsky.cpp:172).
extern "C" void ref_sun_splat(float *rgb, int resolution, const float sunRadianceRGB[3], const float sunDir[3], float stretch, float sunRadiusScale) {
    using namespace refsun;
    BitmapDims dims{resolution, resolution / 2}, *bitmap = &dims;
    Spectrum *data = reinterpret_cast<Spectrum *>(rgb);
    Spectrum sunRadiance; std::memcpy(sunRadiance.s, sunRadianceRGB, 12);
    Vector sd(sunDir[0], sunDir[1], sunDir[2]);
    sd = sd * (1.0f / std::sqrt(sd.x * sd.x + sd.y * sd.y + sd.z * sd.z));                     // normalize: vector.h, v / v.length() multiplies by the reciprocal
    SphericalCoordinates sun = fromSphere(sd);
    sun.elevation *= stretch;
    Point2 factor;
    // sunsky.cpp:173-175
    Frame sunFrame = Frame(toSphere(sun));
    Float theta = degToRad(SUN_APP_RADIUS * 0.5f);
    {
#include "ref_sun_splat.inc"
    }
}

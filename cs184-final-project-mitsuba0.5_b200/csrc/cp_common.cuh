// cp_common.cuh -- shared device/host helpers of the cudapath hot path (sm_100a).
//
// Conventions follow the reference's single-precision RGB build:
// Float = fp32, Spectrum = 3 x fp32, Epsilon = 1e-4f, ShadowEpsilon = 1e-3f
// (include/mitsuba/core/constants.h:28-31), M_PI = fp32 literal (:63,80).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#define CP_HD __host__ __device__ __forceinline__
#define CP_D __device__ __forceinline__

namespace cp {

constexpr float kEpsilon = 1e-4f;
constexpr float kShadowEpsilon = 1e-3f;
constexpr float kDeltaEpsilon = 1e-3f;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInvFourPi = 0.07957747154594766788f;
#define CP_INF (__int_as_float(0x7f800000))

// BSDF::EBSDFType bits that occur on this path (include/mitsuba/render/bsdf.h:230-270)
enum : int { ENull = 0x1, EDiffuseReflection = 0x2, EGlossyReflection = 0x8, EDeltaReflection = 0x20, EDeltaTransmission = 0x40,
              EDelta = ENull | EDeltaReflection | EDeltaTransmission };   // bsdf.h:230-285

struct V3 {
    float x, y, z;
    CP_HD V3() {}
    CP_HD explicit V3(float a) : x(a), y(a), z(a) {}
    CP_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
CP_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
CP_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
CP_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
CP_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
CP_HD V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
CP_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
// The reference divides vectors/spectra by a scalar as multiplication with the reciprocal
// (include/mitsuba/core/vector.h:551, spectrum.h:421).
CP_HD V3 operator/(V3 a, float s) { float r = 1.0f / s; return V3(a.x * r, a.y * r, a.z * r); }
CP_HD V3 &operator+=(V3 &a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
CP_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CP_HD V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CP_HD float length(V3 a) { return sqrtf(dot(a, a)); }
CP_HD V3 normalize(V3 a) { return a / length(a); }
CP_HD bool isZero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
CP_HD float maxc(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
CP_HD float luminance(V3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; } // spectrum.h:724-727
CP_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct D3 {
    double x, y, z;
    CP_HD D3() {}
    CP_HD D3(double a, double b, double c) : x(a), y(b), z(c) {}
    CP_HD explicit D3(V3 v) : x(v.x), y(v.y), z(v.z) {}
};
CP_HD D3 operator+(D3 a, D3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
CP_HD D3 operator-(D3 a, D3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
CP_HD D3 operator*(D3 a, double s) { return D3(a.x * s, a.y * s, a.z * s); }
CP_HD double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CP_HD D3 normalize(D3 a) { double r = 1.0 / sqrt(dot(a, a)); return a * r; }

// Elementary functions, two families.
//   cr_*  correctly rounded fp32 (fp64 evaluation rounded once).  The reference calls the platform's fp32 libm, whose last bit is
//         build-dependent, and M() / sampleM() of the Marschner model amplify that bit by O(1/v) = O(400); pinning a call to the
//         correctly rounded value keeps the device within a few ulp of any faithful CPU evaluation.
//   nc_*  "not critical": calls whose last bits are NOT amplified (table coordinates, sampled directions, the exp / log inside M(),
//         whose argument is O(1) after the cancellation).  Strict translation units (default) map them to cr_*, so device and oracle
//         agree bit for bit; translation units compiled with -DCP_FAST_MATH map them to the 1-2 ulp fp32 CUDA functions, which
//         leaves BSDF values within ~1e-6 of the strict mode (north_star asks for 1e-4) at a fraction of the FP64 work.
CP_HD float cr_sin(float x) { return (float) sin((double) x); }
CP_HD float cr_cos(float x) { return (float) cos((double) x); }
CP_HD void cr_sincos(float x, float *s, float *c) { double ds, dc; sincos((double) x, &ds, &dc); *s = (float) ds; *c = (float) dc; }
CP_HD float cr_asin(float x) { return (float) asin((double) x); }
CP_HD float cr_acos(float x) { return (float) acos((double) x); }
CP_HD float cr_atan2(float y, float x) { return (float) atan2((double) y, (double) x); }
CP_HD float cr_exp(float x) { return (float) exp((double) x); }
CP_HD float cr_log(float x) { return (float) log((double) x); }
CP_HD float cr_log2(float x) { return (float) log2((double) x); }
CP_HD float cr_pow(float x, float y) { return (float) pow((double) x, (double) y); }
CP_HD float cr_sinh(float x) { return (float) sinh((double) x); }
CP_HD float cr_tan(float x) { return (float) tan((double) x); }
#if defined(__CUDA_ARCH__) && defined(CP_FAST_MATH)
#define CP_MATH_IS_FAST 1
CP_D float nc_sin(float x) { return sinf(x); }
CP_D float nc_cos(float x) { return cosf(x); }
CP_D void nc_sincos(float x, float *s, float *c) { sincosf(x, s, c); }
CP_D float nc_asin(float x) { return asinf(x); }
CP_D float nc_acos(float x) { return acosf(x); }
CP_D float nc_atan2(float y, float x) { return atan2f(y, x); }
CP_D float nc_exp(float x) { return expf(x); }
CP_D float nc_log(float x) { return logf(x); }
CP_D float nc_log2(float x) { return log2f(x); }
CP_D float nc_pow(float x, float y) { return powf(x, y); }
CP_D float nc_sinh(float x) { return sinhf(x); }
CP_D float nc_tan(float x) { return tanf(x); }
#else
#define CP_MATH_IS_FAST 0
CP_HD float nc_sin(float x) { return cr_sin(x); }
CP_HD float nc_cos(float x) { return cr_cos(x); }
CP_HD void nc_sincos(float x, float *s, float *c) { cr_sincos(x, s, c); }
CP_HD float nc_asin(float x) { return cr_asin(x); }
CP_HD float nc_acos(float x) { return cr_acos(x); }
CP_HD float nc_atan2(float y, float x) { return cr_atan2(y, x); }
CP_HD float nc_exp(float x) { return cr_exp(x); }
CP_HD float nc_log(float x) { return cr_log(x); }
CP_HD float nc_log2(float x) { return cr_log2(x); }
CP_HD float nc_pow(float x, float y) { return cr_pow(x, y); }
CP_HD float nc_sinh(float x) { return cr_sinh(x); }
CP_HD float nc_tan(float x) { return cr_tan(x); }
#endif
CP_HD float cr_hypot(float x, float y) { return (float) sqrt((double) x * x + (double) y * y); }

CP_HD float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(lo, v)); }
CP_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
CP_HD float safe_sqrt(float v) { return sqrtf(fmaxf(0.0f, v)); }
CP_HD float safe_acos(float v) { return nc_acos(fminf(1.0f, fmaxf(-1.0f, v))); }

// src/libcore/util.cpp:592-601
CP_HD void coordinateSystem(const V3 &a, V3 &b, V3 &c) {
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// include/mitsuba/core/frame.h:55-85
struct Frame {
    V3 s, t, n;
    CP_HD Frame() {}
    CP_HD explicit Frame(const V3 &n_) : n(n_) { coordinateSystem(n, s, t); }
    CP_HD V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    CP_HD V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
};

// src/libcore/util.cpp:448-482 (fp32 quadratic used for the bounding-sphere exit distance)
CP_HD bool solveQuadratic(float a, float b, float c, float &x0, float &x1) {
    if (a == 0) {
        if (b != 0) { x0 = x1 = -c / b; return true; }
        return false;
    }
    float discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    float temp, sqrtDiscrim = sqrtf(discrim);
    if (b < 0) temp = -0.5f * (b - sqrtDiscrim);
    else       temp = -0.5f * (b + sqrtDiscrim);
    x0 = temp / a;
    x1 = c / temp;
    if (x0 > x1) { float t = x0; x0 = x1; x1 = t; }
    return true;
}

// src/libcore/util.cpp:651-681 through the 2-argument wrapper (util.h:479)
CP_HD float fresnelDielectricExt(float cosThetaI_, float eta) {
    if (eta == 1.0f) return 0.0f;
    float scale = (cosThetaI_ > 0) ? 1.0f / eta : eta,
          cosThetaTSqr = 1.0f - (1.0f - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) return 1.0f;
    float cosThetaI = fabsf(cosThetaI_);
    float cosThetaT = sqrtf(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return 0.5f * (Rs * Rs + Rp * Rp);
}

// src/libcore/warp.cpp:81-102
CP_HD void squareToUniformDiskConcentric(float sx, float sy, float &ox, float &oy) {
    float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (kPi / 4.0f) * (r2 / r1); }
    else { r = r2; phi = (kPi / 2.0f) - (r1 / r2) * (kPi / 4.0f); }
    float sinPhi, cosPhi;
    nc_sincos(phi, &sinPhi, &cosPhi);
    ox = r * cosPhi; oy = r * sinPhi;
}
// src/libcore/warp.cpp:43-52
CP_HD V3 squareToCosineHemisphere(float sx, float sy) {
    float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    float z = safe_sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}
// src/libcore/warp.cpp:143-156
CP_HD float intervalToTent(float sample) {
    float sign;
    if (sample < 0.5f) { sign = 1; sample *= 2; }
    else { sign = -1; sample = 2 * (sample - 0.5f); }
    return sign * (1 - sqrtf(sample));
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG replacing the sampler plugins (north_star item 4):
// Philox4x32-10, key = 64-bit seed, counter = (pixel index, sample index, path vertex, block).
// vertex 0 / block 0 -> {jitter.x, jitter.y}; vertex k>=1 / block 0 -> {nee.x, nee.y, bsdf.x, bsdf.y};
// vertex k / block 1 -> {russian roulette}.  Matches the dimension order of
// src/librender/integrator.cpp:171 and src/integrators/path/path.cpp:176,209,276-285.
// ---------------------------------------------------------------------------------------------
struct Philox4 { uint32_t v[4]; };
CP_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t) M0 * c0, p1 = (uint64_t) M1 * c2;
        uint32_t hi0 = (uint32_t) (p0 >> 32), lo0 = (uint32_t) p0, hi1 = (uint32_t) (p1 >> 32), lo1 = (uint32_t) p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3; return o;
}
CP_HD float u32_to_unit(uint32_t x) { return (float) (x >> 8) * (1.0f / 16777216.0f); }

} // namespace cp

// oracle/o_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Small numeric helpers restated from the reference (Mitsuba 0.5 fork, SINGLE_PRECISION,
// SPECTRUM_SAMPLES=3).  Each function cites the reference file:line it follows
// (paths relative to /root/reference).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <vector>
#include <string>
#include <stdexcept>

namespace orc {

// include/mitsuba/core/constants.h:28-31,63-69 (single-precision build)
static const float kEpsilon = 1e-4f;
static const float kShadowEpsilon = 1e-3f;
static const float kDeltaEpsilon = 1e-3f;   // constants.h:28-31
static const float kPi = 3.14159265358979323846f;
static const float kInvPi = 0.31830988618379067154f;
static const float kInvTwoPi = 0.15915494309189533577f;
static const float kInvFourPi = 0.07957747154594766788f;
static const float kInf = std::numeric_limits<float>::infinity();

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float &operator[](int i) { return (&x)[i]; }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator/(V3 a, float s) { float r = 1.0f / s; return V3(a.x * r, a.y * r, a.z * r); } // vector.h: multiplies by reciprocal
static inline V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
static inline V3 &operator*=(V3 &a, V3 b) { a = a * b; return a; }
static inline V3 &operator*=(V3 &a, float s) { a = a * s; return a; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a / length(a); }
static inline bool isZero(V3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
static inline float maxc(V3 a) { return std::max(a.x, std::max(a.y, a.z)); }
// include/mitsuba/core/spectrum.h:724-727 (RGB luminance)
static inline float luminance(V3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }

struct D3 {
    double x, y, z;
    D3() : x(0), y(0), z(0) {}
    D3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit D3(V3 v) : x(v.x), y(v.y), z(v.z) {}
};
static inline D3 operator+(D3 a, D3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline D3 operator-(D3 a, D3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline D3 operator*(D3 a, double s) { return D3(a.x * s, a.y * s, a.z * s); }
static inline D3 operator*(double s, D3 a) { return D3(a.x * s, a.y * s, a.z * s); }
static inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline D3 normalize(D3 a) { double r = 1.0 / std::sqrt(dot(a, a)); return a * r; }

// Elementary functions.  The reference calls the platform's fp32 libm (sinf, expf, ...), whose last bit differs between
// libm builds; several formulas on this path amplify that bit by O(1/v) = O(400).  The oracle therefore pins every
// elementary function to its correctly rounded fp32 value (fp64 evaluation rounded once) -- any faithful libm agrees with
// that to 1 ulp -- and the CUDA path does the same, so oracle and device can be compared almost bit for bit.
// The timing build (liboracle_fast.so, -DORC_FAST) calls the platform's fp32 libm instead, as the reference's binaries do.
namespace cr {
#ifdef ORC_FAST
static inline float sin(float x) { return ::sinf(x); }
static inline float cos(float x) { return ::cosf(x); }
static inline float tan(float x) { return ::tanf(x); }
static inline float asin(float x) { return ::asinf(x); }
static inline float acos(float x) { return ::acosf(x); }
static inline float atan2(float y, float x) { return ::atan2f(y, x); }
static inline float exp(float x) { return (float) std::exp((double) x); }        // math::fastexp / fastlog on Linux x86-64 (math.h:175-199)
static inline float log(float x) { return (float) std::log((double) x); }
static inline float log2(float x) { return ::log2f(x); }
static inline float pow(float x, float y) { return ::powf(x, y); }
static inline float sinh(float x) { return ::sinhf(x); }
static inline float hypot(float x, float y) { return ::hypotf(x, y); }
static inline void sincos(float x, float *s, float *c) { ::sincosf(x, s, c); }
#else
static inline float sin(float x) { return (float) std::sin((double) x); }
static inline float cos(float x) { return (float) std::cos((double) x); }
static inline float tan(float x) { return (float) std::tan((double) x); }
static inline float asin(float x) { return (float) std::asin((double) x); }
static inline float acos(float x) { return (float) std::acos((double) x); }
static inline float atan2(float y, float x) { return (float) std::atan2((double) y, (double) x); }
static inline float exp(float x) { return (float) std::exp((double) x); }
static inline float log(float x) { return (float) std::log((double) x); }
static inline float log2(float x) { return (float) std::log2((double) x); }
static inline float pow(float x, float y) { return (float) std::pow((double) x, (double) y); }
static inline float sinh(float x) { return (float) std::sinh((double) x); }
static inline float hypot(float x, float y) { return (float) std::sqrt((double) x * x + (double) y * y); }
static inline void sincos(float x, float *s, float *c) { *s = (float) std::sin((double) x); *c = (float) std::cos((double) x); }
#endif
}

static inline float clampf(float v, float lo, float hi) { return std::min(hi, std::max(lo, v)); }
static inline int clampi(int v, int lo, int hi) { return std::min(hi, std::max(lo, v)); }
static inline float safe_sqrt(float v) { return std::sqrt(std::max(0.0f, v)); }  // math.h:260
static inline float safe_acos(float v) { return cr::acos(std::min(1.0f, std::max(-1.0f, v))); } // math.h:250
static inline int floorToInt(float v) { return (int) std::floor(v); } // math.h:100
static inline int ceilToInt(float v) { return (int) std::ceil(v); }   // math.h:105
static inline int modulo(int a, int b) { int r = a % b; return (r < 0) ? r + b : r; } // math.h:67

// src/libcore/util.cpp:592-601
static inline void coordinateSystem(const V3 &a, V3 &b, V3 &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
        float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// include/mitsuba/core/frame.h:55-85
struct Frame {
    V3 s, t, n;
    Frame() {}
    explicit Frame(const V3 &n_) : n(n_) { coordinateSystem(n, s, t); }
    V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
};

// src/libcore/util.cpp:603-608
static inline void computeShadingFrame(const V3 &n, const V3 &dpdu, Frame &frame) {
    frame.n = n;
    frame.s = normalize(dpdu - frame.n * dot(frame.n, dpdu));
    frame.t = cross(frame.n, frame.s);
}

// src/libcore/util.cpp:487-525
static inline bool solveQuadraticDouble(double a, double b, double c, double &x0, double &x1) {
    if (a == 0) {
        if (b != 0) { x0 = x1 = -c / b; return true; }
        return false;
    }
    double discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    double temp, sqrtDiscrim = std::sqrt(discrim);
    if (b < 0) temp = -0.5f * (b - sqrtDiscrim);
    else       temp = -0.5f * (b + sqrtDiscrim);
    x0 = temp / a;
    x1 = c / temp;
    if (x0 > x1) std::swap(x0, x1);
    return true;
}

// src/libcore/util.cpp:448-482 (single-precision variant, used by BSphere::rayIntersect)
static inline bool solveQuadratic(float a, float b, float c, float &x0, float &x1) {
    if (a == 0) {
        if (b != 0) { x0 = x1 = -c / b; return true; }
        return false;
    }
    float discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    float temp, sqrtDiscrim = std::sqrt(discrim);
    if (b < 0) temp = -0.5f * (b - sqrtDiscrim);
    else       temp = -0.5f * (b + sqrtDiscrim);
    x0 = temp / a;
    x1 = c / temp;
    if (x0 > x1) std::swap(x0, x1);
    return true;
}

// src/libcore/util.cpp:651-681 (+ the 2-argument wrapper include/mitsuba/core/util.h:479)
static inline float fresnelDielectricExt(float cosThetaI_, float eta) {
    if (eta == 1) return 0.0f;
    float scale = (cosThetaI_ > 0) ? 1 / eta : eta,
          cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) return 1.0f;
    float cosThetaI = std::abs(cosThetaI_);
    float cosThetaT = std::sqrt(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return 0.5f * (Rs * Rs + Rp * Rp);
}

// src/libcore/warp.cpp:81-102
// ---------------------------------------------------------------------------------------------
// GaussLobattoIntegrator -- src/libcore/quad.cpp:287-420 (adaptive Gauss-Lobatto with Kronrod extension, fp32), restated with
// the reference's defaults useConvergenceEstimate = true (include/mitsuba/core/quad.h:155-159).  The six recursive calls of one
// step are summed left to right.
// ---------------------------------------------------------------------------------------------
struct GaussLobatto {
    float absError, relError; size_t maxEvals; bool useConvergenceEstimate = true;
    GaussLobatto(size_t maxEvals_, float absError_, float relError_) : absError(absError_), relError(relError_), maxEvals(maxEvals_) {}
    static float alpha() { return (float) std::sqrt(2.0 / 3.0); }
    static float beta() { return (float) (1.0 / std::sqrt(5.0)); }
    template <class F> float integrate(const F &f, float a, float b) const {
        float factor = 1; size_t evals = 0;
        if (a == b) return 0;
        if (b < a) { std::swap(a, b); factor = -1; }
        const float absTolerance = calculateAbsTolerance(f, a, b, evals);
        evals += 2;
        const float fa = f(a), fb = f(b);
        return factor * step(f, a, b, fa, fb, absTolerance, evals);
    }
    template <class F> float calculateAbsTolerance(const F &f, float a, float b, size_t &evals) const {
        const float x1 = (float) 0.94288241569547971906, x2 = (float) 0.64185334234578130578, x3 = (float) 0.23638319966214988028;
        const float m = (a + b) / 2, h = (b - a) / 2;
        const float y1 = f(a), y3 = f(m - alpha() * h), y5 = f(m - beta() * h), y7 = f(m), y9 = f(m + beta() * h), y11 = f(m + alpha() * h), y13 = f(b);
        const float f1a = f(m - x1 * h), f1b = f(m + x1 * h), f2a = f(m - x2 * h), f2b = f(m + x2 * h), f3a = f(m - x3 * h), f3b = f(m + x3 * h);
        float acc = h * ((float) 0.0158271919734801831 * (y1 + y13) + (float) 0.0942738402188500455 * (f1a + f1b)
                       + (float) 0.1550719873365853963 * (y3 + y11) + (float) 0.1888215739601824544 * (f2a + f2b)
                       + (float) 0.1997734052268585268 * (y5 + y9) + (float) 0.2249264653333395270 * (f3a + f3b)
                       + (float) 0.2426110719014077338 * y7);
        evals += 13;
        float r = 1.0f;
        if (useConvergenceEstimate) {
            const float integral2 = (h / 6) * (y1 + y13 + 5 * (y5 + y9));
            const float integral1 = (h / 1470) * (77 * (y1 + y13) + 432 * (y3 + y11) + 625 * (y5 + y9) + 672 * y7);
            if (std::abs(integral2 - acc) != 0.0f) r = std::abs(integral1 - acc) / std::abs(integral2 - acc);
            if (r == 0.0f || r > 1.0f) r = 1.0f;
        }
        float result = std::numeric_limits<float>::infinity();
        const float eps = std::numeric_limits<float>::epsilon();
        if (relError != 0 && acc != 0) result = acc * std::max(relError, eps) / (r * eps);
        if (absError != 0) result = std::min(result, absError / (r * eps));
        return result;
    }
    template <class F> float step(const F &f, float a, float b, float fa, float fb, float acc, size_t &evals) const {
        const float h = (b - a) / 2, m = (a + b) / 2;
        const float mll = m - alpha() * h, ml = m - beta() * h, mr = m + beta() * h, mrr = m + alpha() * h;
        const float fmll = f(mll), fml = f(ml), fm = f(m), fmr = f(mr), fmrr = f(mrr);
        const float integral2 = (h / 6) * (fa + fb + 5 * (fml + fmr));
        const float integral1 = (h / 1470) * (77 * (fa + fb) + 432 * (fmll + fmrr) + 625 * (fml + fmr) + 672 * fm);
        evals += 5;
        if (evals >= maxEvals) return integral1;
        const float dist = acc + (integral1 - integral2);
        if (dist == acc || mll <= a || b <= mrr) return integral1;
        const float s0 = step(f, a, mll, fa, fmll, acc, evals);
        const float s1 = step(f, mll, ml, fmll, fml, acc, evals);
        const float s2 = step(f, ml, m, fml, fm, acc, evals);
        const float s3 = step(f, m, mr, fm, fmr, acc, evals);
        const float s4 = step(f, mr, mrr, fmr, fmrr, acc, evals);
        const float s5 = step(f, mrr, b, fmrr, fb, acc, evals);
        return s0 + s1 + s2 + s3 + s4 + s5;
    }
};

// src/libcore/util.cpp:807-862 with fast = false: the diffuse Fresnel reflectance by adaptive quadrature of F(sqrt(xi), eta) over [0, 1]
static inline float fresnelDiffuseReflectance(float eta) {
    GaussLobatto quad(1024, 0, 1e-5f);
    return quad.integrate([eta](float xi) { return fresnelDielectricExt(std::sqrt(xi), eta); }, 0.0f, 1.0f);
}

static inline void squareToUniformDiskConcentric(float sx, float sy, float &ox, float &oy) {
    float r1 = 2.0f * sx - 1.0f;
    float r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (kPi / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (kPi / 2.0f) - (r1 / r2) * (kPi / 4.0f);
    }
    float cosPhi = cr::cos(phi), sinPhi = cr::sin(phi);
    ox = r * cosPhi; oy = r * sinPhi;
}

// src/libcore/warp.cpp:43-52
static inline V3 squareToCosineHemisphere(float sx, float sy) {
    float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    float z = safe_sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}

// src/libcore/warp.cpp:54-63
static inline V3 squareToUniformCone(float cosCutoff, float sx, float sy) {
    float cosTheta = (1 - sx) + sx * cosCutoff;
    float sinTheta = safe_sqrt(1.0f - cosTheta * cosTheta);
    float phi = 2.0f * kPi * sy;
    return V3(cr::cos(phi) * sinTheta, cr::sin(phi) * sinTheta, cosTheta);
}

// src/libcore/warp.cpp:143-162
static inline float intervalToTent(float sample) {
    float sign;
    if (sample < 0.5f) { sign = 1; sample *= 2; }
    else { sign = -1; sample = 2 * (sample - 0.5f); }
    return sign * (1 - std::sqrt(sample));
}

// ---------------------------------------------------------------------------
// Counter-based RNG shared (by specification, not by code) with the CUDA path:
// Philox4x32-10 (Salmon et al. 2011), key = 64-bit seed, counter =
// (pixel index, sample index, path vertex, block).  Replaces the reference's
// sampler plugins (include/mitsuba/render/sampler.h:105-117) as north_star asks.
// Dimension layout: vertex 0 block 0 -> {jitter.x, jitter.y};
// vertex k>=1 block 0 -> {nee.x, nee.y, bsdf.x, bsdf.y}; vertex k block 1 -> {rr}.
// ---------------------------------------------------------------------------
struct Philox4 { uint32_t v[4]; };
static inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t) M0 * c0, p1 = (uint64_t) M1 * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3; return o;
}
static inline float u32_to_unit(uint32_t x) { return (float) (x >> 8) * (1.0f / 16777216.0f); } // [0,1)

// IEEE binary16 round-to-nearest-even quantisation (OpenEXR `half`, used for envmap texels:
// src/emitters/envmap.cpp:102-103)
static inline uint16_t float_to_half(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t) ((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t) (sign | 0x7c00u | (man ? 0x200u : 0)); // inf/nan
    if (exp >= 31) return (uint16_t) (sign | 0x7c00u); // overflow -> inf
    if (exp <= 0) {
        if (exp < -10) return (uint16_t) sign; // underflow to zero
        man |= 0x800000u;
        int shift = 14 - exp; // 14..24
        uint32_t h = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1))) h++;
        return (uint16_t) (sign | h);
    }
    uint32_t h = ((uint32_t) exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++; // may carry into exponent (correct)
    return (uint16_t) (sign | h);
}
static inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400u));
            x = sign | ((uint32_t) (127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}

// 4x4 row-major matrix helpers (include/mitsuba/core/transform.h: point transform divides by w)
// ---------------------------------------------------------------------------------------------
// Mitsuba's Random (SFMT-19937), src/libcore/random.cpp: parameters :72-96, 128-bit shifts :129-173, do_recursion :204-219, period_certification
// :325-347, gen_rand_all :373-389, init_gen_rand :396-405, nextULong :551-553, single-precision nextFloat :630-639.  Default seed 5489
// (include/mitsuba/core/random.h:113).  The hair loader's `reduction` draws from it (hair.cpp:629,672-673,769-770).
// ---------------------------------------------------------------------------------------------
struct MitsubaRandom {
    enum { MEXP = 19937, N = MEXP / 128 + 1, N32 = N * 4, N64 = N * 2, POS1 = 122, SL1 = 18, SL2 = 1, SR1 = 11, SR2 = 1 };
    union W128 { uint64_t u64[2]; uint32_t u[4]; };
    union { W128 sfmt[N]; uint32_t psfmt32[N32]; uint64_t psfmt64[N64]; };
    int idx = -1;
    explicit MitsubaRandom(uint64_t seed = 5489ULL) {
        psfmt64[0] = seed;
        for (int i = 1; i < N64; ++i) psfmt64[i] = 6364136223846793005ULL * (psfmt64[i - 1] ^ (psfmt64[i - 1] >> 62)) + (uint64_t) i;
        idx = N32;
        const uint32_t parity[4] = {0x00000001U, 0x00000000U, 0x00000000U, 0x13c9e684U};
        int inner = 0;
        for (int i = 0; i < 4; ++i) inner ^= psfmt32[i] & parity[i];
        for (int i = 16; i > 0; i >>= 1) inner ^= inner >> i;
        inner &= 1;
        if (inner == 1) return;
        for (int i = 0; i < 4; ++i) {
            uint32_t work = 1;
            for (int j = 0; j < 32; ++j) { if ((work & parity[i]) != 0) { psfmt32[i] ^= work; return; } work = work << 1; }
        }
    }
    static void rshift128(W128 &out, const W128 &in, int shift) {
        const uint64_t th = in.u64[1], tl = in.u64[0];
        out.u64[1] = th >> (shift * 8); out.u64[0] = (tl >> (shift * 8)) | (th << (64 - shift * 8));
    }
    static void lshift128(W128 &out, const W128 &in, int shift) {
        const uint64_t th = in.u64[1], tl = in.u64[0];
        out.u64[0] = tl << (shift * 8); out.u64[1] = (th << (shift * 8)) | (tl >> (64 - shift * 8));
    }
    static void doRecursion(W128 &r, const W128 &a, const W128 &b, const W128 &c, const W128 &d) {
        const uint32_t msk[4] = {0xdfffffefU, 0xddfecb7fU, 0xbffaffffU, 0xbffffff6U};
        W128 x, y;
        lshift128(x, a, SL2); rshift128(y, c, SR2);
        for (int k = 0; k < 4; ++k) r.u[k] = a.u[k] ^ x.u[k] ^ ((b.u[k] >> SR1) & msk[k]) ^ y.u[k] ^ (d.u[k] << SL1);
    }
    void genRandAll() {
        W128 *r1 = &sfmt[N - 2], *r2 = &sfmt[N - 1];
        int i = 0;
        for (; i < N - POS1; ++i) { doRecursion(sfmt[i], sfmt[i], sfmt[i + POS1], *r1, *r2); r1 = r2; r2 = &sfmt[i]; }
        for (; i < N; ++i) { doRecursion(sfmt[i], sfmt[i], sfmt[i + POS1 - N], *r1, *r2); r1 = r2; r2 = &sfmt[i]; }
    }
    uint64_t nextULong() { if (idx >= N32) { genRandAll(); idx = 0; } uint64_t r = psfmt64[idx / 2]; idx += 2; return r; }
    float nextFloat() { union { uint32_t u; float f; } x; x.u = (uint32_t) ((nextULong() & 0xFFFFFFFF) >> 9) | 0x3f800000UL; return x.f - 1.0f; }
};

// ---------------------------------------------------------------------------------------------
// The `sobol` sampler (src/samplers/sobol.cpp) over Gruenschloss' Sobol code (src/samplers/sobolseq.h:59-135): sampleSingle, look_up
// (the (0,2)-sequence enumerated per pixel), sampleTEA for the scramble (include/mitsuba/core/qmc.h:146-156).  The direction numbers
// (src/samplers/sobolseq.cpp, data) are read from refdata/sobol.bin (tools/mirror_refdata.py).
// ---------------------------------------------------------------------------------------------
struct SobolTables {
    enum { NumDimensions = 1024, Size = 52 };
    std::vector<uint32_t> m32; std::vector<uint64_t> vdc, inv; uint32_t rowsVdc = 0, rowsInv = 0;
    void load(const std::string &path) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path + " (run tools/mirror_refdata.py)");
        uint32_t h[5];
        bool ok = fread(h, 4, 5, f) == 5 && h[0] == 0x4c424f53u && h[1] == NumDimensions && h[2] == Size;
        if (ok) { rowsVdc = h[3]; rowsInv = h[4]; m32.resize((size_t) NumDimensions * Size); vdc.resize((size_t) rowsVdc * Size); inv.resize((size_t) rowsInv * Size);
                  ok = fread(m32.data(), 4, m32.size(), f) == m32.size() && fread(vdc.data(), 8, vdc.size(), f) == vdc.size() && fread(inv.data(), 8, inv.size(), f) == inv.size(); }
        fclose(f);
        if (!ok) throw std::runtime_error("malformed " + path);
    }
    float sample(uint64_t index, uint32_t dimension, uint32_t scramble) const {          // sobolseq.h:59-74
        uint32_t result = scramble;
        for (uint32_t i = dimension * Size; index; index >>= 1, ++i) if (index & 1) result ^= m32[i];
        return std::min(result * (1.0f / (1ULL << 32)), 0.999999940395355225f);
    }
    uint64_t lookUp(uint32_t m, uint32_t frame, uint32_t px, uint32_t py, uint64_t scramble) const {   // sobolseq.h:104-133 (SINGLE_PRECISION)
        if (m < 1 || m > rowsVdc || m > rowsInv) throw std::runtime_error("sobol: film resolution outside the enumeration tables");
        const uint32_t m2 = m << 1;
        uint64_t index = uint64_t(frame) << m2, delta = 0;
        for (uint32_t c = 0; frame; frame >>= 1, ++c) if (frame & 1) delta ^= vdc[(size_t) (m - 1) * Size + c];
        scramble = (scramble & 0xFFFFFFFF) >> (32 - m);
        uint64_t b = (((uint64_t) (px ^ scramble) << m) | (py ^ scramble)) ^ delta;
        for (uint32_t c = 0; b; b >>= 1, ++c) if (b & 1) index ^= inv[(size_t) (m - 1) * Size + c];
        return index;
    }
};
static inline uint64_t sampleTEA(uint32_t v0, uint32_t v1, int rounds = 4) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9;
        v0 += ((v1 << 4) + 0xA341316C) ^ (v1 + sum) ^ ((v1 >> 5) + 0xC8013EA4);
        v1 += ((v0 << 4) + 0xAD90777D) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7E95761E);
    }
    return ((uint64_t) v1 << 32) + v0;
}
// SobolSampler as a block-based render drives it (sobol.cpp:90-105 ctor, :146-156 setFilmResolution(res, bucketed = true), :170-200 generate,
// :202-217 advance / setSampleIndex, :219-245 next1D / next2D; no sample arrays are requested by the path tracer, so arrayStartDim = arrayEndDim = 5)
struct SobolSampler {
    const SobolTables *T = nullptr;
    uint64_t scramble = 0, sobolSampleIndex = 0; size_t sampleIndex = 0;
    float resolution = 1; uint32_t logResolution = 0, dimension = 0; int px = 0, py = 0;
    void configure(const SobolTables *t, uint64_t scrambleProp, int filmW, int filmH) {
        T = t; scramble = scrambleProp;
        if (scramble) scramble = sampleTEA((uint32_t) scramble, (uint32_t) (scramble >> 32));
        uint32_t r = (uint32_t) std::max(filmW, filmH);                                  // math::roundToPowerOfTwo, math.cpp:128-134
        r--; r |= r >> 1; r |= r >> 2; r |= r >> 4; r |= r >> 8; r |= r >> 16; r++;
        resolution = (float) r;
        logResolution = 0; while ((1u << (logResolution + 1)) <= r) ++logResolution;     // math::log2i
    }
    void generate(int x, int y) { px = x; py = y; setSampleIndex(0); }
    void advance() { setSampleIndex(sampleIndex + 1); }
    void setSampleIndex(size_t i) {
        dimension = 0; sampleIndex = i;
        if (logResolution > 1 && px >= 0) sobolSampleIndex = T->lookUp(logResolution, (uint32_t) sampleIndex, (uint32_t) px, (uint32_t) py, scramble);
        else sobolSampleIndex = (uint64_t) sampleIndex;
    }
    void check(uint32_t d) const { if (d >= SobolTables::NumDimensions) throw std::runtime_error("Lookup dimension exceeds the direction number table size! You may have to reduce the 'maxDepth' parameter of your integrator."); }
    // The dimensions reserved for sample arrays start and end at 5 when none are requested (:176-178); the range tests still fire: a 2-D request
    // that would begin at dimension 4 moves to 5 (4 + 1 >= arrayStartDim && 4 < arrayEndDim), so dimension 4 is never used by a path tracer
    // (pixel 0-1, emitter 2-3, BSDF 5-6, ...).
    enum { ArrayStartDim = 5, ArrayEndDim = 5 };
    float next1D() {
        if (dimension >= ArrayStartDim && dimension < ArrayEndDim) dimension = ArrayEndDim;
        check(dimension); return T->sample(sobolSampleIndex, dimension++, (uint32_t) scramble);
    }
    void next2D(float &a, float &b) {
        if (dimension + 1 >= ArrayStartDim && dimension < ArrayEndDim) dimension = ArrayEndDim;
        check(dimension + 1);
        if (dimension == 0 && sobolSampleIndex != (uint64_t) sampleIndex) {
            a = T->sample(sobolSampleIndex, dimension++, (uint32_t) scramble) * resolution - px;
            b = T->sample(sobolSampleIndex, dimension++, (uint32_t) scramble) * resolution - py;
        } else {
            a = T->sample(sobolSampleIndex, dimension++, (uint32_t) scramble);
            b = T->sample(sobolSampleIndex, dimension++, (uint32_t) scramble);
        }
    }
};

struct M44 {
    float m[4][4];
    static M44 identity() { M44 r; std::memset(r.m, 0, sizeof(r.m)); for (int i = 0; i < 4; ++i) r.m[i][i] = 1; return r; }
    static M44 fromRowMajor(const float *p) { M44 r; std::memcpy(r.m, p, 64); return r; }
};
static inline M44 mul(const M44 &a, const M44 &b) {
    M44 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        float s = 0;
        for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
        r.m[i][j] = s;
    }
    return r;
}
static inline V3 xfmPoint(const M44 &t, const V3 &p) { // transform.h: Point operator()
    float x = t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3];
    float y = t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3];
    float z = t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3];
    float w = t.m[3][0] * p.x + t.m[3][1] * p.y + t.m[3][2] * p.z + t.m[3][3];
    if (w == 1.0f) return V3(x, y, z);
    return V3(x, y, z) / w;
}
static inline V3 xfmVector(const M44 &t, const V3 &v) {
    return V3(t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z,
              t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
              t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z);
}
// Matrix<4,4,float>::invert (include/mitsuba/core/matrix.inl:138-193): Gauss-Jordan with full pivoting, in place, in fp32 -- what
// Transform(const Matrix4x4 &) runs on every matrix read from a scene file and on Transform::perspective (transform.h:50-55)
static inline bool invert(const M44 &a, M44 &out) {
    const int N = 4;
    int indxc[N], indxr[N], ipiv[N] = {0, 0, 0, 0};
    std::memcpy(out.m, a.m, sizeof(out.m));
    for (int i = 0; i < N; i++) {
        int irow = -1, icol = -1;
        float big = 0;
        for (int j = 0; j < N; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < N; k++) {
                    if (ipiv[k] == 0) {
                        if (std::abs(out.m[j][k]) >= big) { big = std::abs(out.m[j][k]); irow = j; icol = k; }
                    } else if (ipiv[k] > 1) return false;
                }
            }
        }
        ++ipiv[icol];
        if (irow != icol) for (int k = 0; k < N; ++k) std::swap(out.m[irow][k], out.m[icol][k]);
        indxr[i] = irow; indxc[i] = icol;
        if (out.m[icol][icol] == 0) return false;
        float pivinv = 1.f / out.m[icol][icol];
        out.m[icol][icol] = 1.f;
        for (int j = 0; j < N; j++) out.m[icol][j] *= pivinv;
        for (int j = 0; j < N; j++) {
            if (j != icol) {
                float save = out.m[j][icol];
                out.m[j][icol] = 0;
                for (int k = 0; k < N; k++) out.m[j][k] -= out.m[icol][k] * save;
            }
        }
    }
    for (int j = N - 1; j >= 0; j--)
        if (indxr[j] != indxc[j]) for (int k = 0; k < N; k++) std::swap(out.m[k][indxr[j]], out.m[k][indxc[j]]);
    return true;
}

} // namespace orc

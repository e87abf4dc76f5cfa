// cp_host_sunsky.cpp -- host-side bake of the `sunsky` emitter into a lat-long RGB map.
//
// Replaces (reference file:line):
//   SunSkyEmitter::SunSkyEmitter          src/emitters/sunsky.cpp:100-236   (sky raster + QMC sun-disc splat)
//   SkyEmitter ctor / getSkyRadiance      src/emitters/sky.cpp:219-256, 413-447
//   Hosek-Wilkie RGB model evaluation     src/emitters/sunsky/skymodel.cpp:76-283,340-394 (algorithm of Hosek & Wilkie 2012;
//                                         the coefficient tables are read from <data>/sunsky/hosek_rgb.f64)
//   computeSunCoordinates / computeSunRadiance   src/emitters/sunsky/sunmodel.h:90-105, 206-222, 260-371
//   Spectrum::fromContinuousSpectrum (RGB)        src/libcore/spectrum.cpp:172-185, 222-227, 546-568
//   sample02, squareToUniformCone                 include/mitsuba/core/qmc.h:43-60,82-87,115-120; src/libcore/warp.cpp:54-63
// This is load-time set-up (about 130 k sky evaluations and a few thousand sun samples); it stays on the CPU.
#include "cp_host.h"
#include <cmath>
#include <cstdio>
#include <algorithm>

namespace cp {
namespace {

const float PI_F = 3.14159265358979323846f;

struct SkyState { double config[9]; double radiance; };

// quintic Bezier in the cube-root-warped solar elevation, bilinear in (albedo, turbidity)
double bezier5(const double *ctrl, int stride, double x) {
    const double y = 1.0 - x;
    return std::pow(y, 5.0) * ctrl[0] + 5.0 * std::pow(y, 4.0) * x * ctrl[stride] + 10.0 * std::pow(y, 3.0) * std::pow(x, 2.0) * ctrl[2 * stride] +
           10.0 * std::pow(y, 2.0) * std::pow(x, 3.0) * ctrl[3 * stride] + 5.0 * y * std::pow(x, 4.0) * ctrl[4 * stride] + std::pow(x, 5.0) * ctrl[5 * stride];
}
bool cookState(const double *cfg /*1080*/, const double *rad /*120*/, double turbidity, double albedo, double elevation, SkyState &st) {
    const int it = (int) turbidity;
    if (it < 1 || it > 10) return false;
    const double tr = turbidity - (double) it;
    // the reference's M_PI is the fp32 literal in single-precision builds (constants.h:63,80)
    const double x = std::pow(elevation / ((double) PI_F / 2.0), 1.0 / 3.0);
    for (int i = 0; i < 9; ++i) {
        double v = (1.0 - albedo) * (1.0 - tr) * bezier5(cfg + 9 * 6 * (it - 1) + i, 9, x);
        v += albedo * (1.0 - tr) * bezier5(cfg + 9 * 6 * 10 + 9 * 6 * (it - 1) + i, 9, x);
        if (it != 10) {
            v += (1.0 - albedo) * tr * bezier5(cfg + 9 * 6 * it + i, 9, x);
            v += albedo * tr * bezier5(cfg + 9 * 6 * 10 + 9 * 6 * it + i, 9, x);
        }
        st.config[i] = v;
    }
    double r = (1.0 - albedo) * (1.0 - tr) * bezier5(rad + 6 * (it - 1), 1, x);
    r += albedo * (1.0 - tr) * bezier5(rad + 6 * 10 + 6 * (it - 1), 1, x);
    if (it != 10) {
        r += (1.0 - albedo) * tr * bezier5(rad + 6 * it, 1, x);
        r += albedo * tr * bezier5(rad + 6 * 10 + 6 * it, 1, x);
    }
    st.radiance = r;
    return true;
}
double skyRadiance(const SkyState &st, double theta, double gamma) {
    const double *c = st.config;
    const double cg = std::cos(gamma), ct = std::cos(theta);
    const double expM = std::exp(c[4] * gamma);
    const double rayM = cg * cg;
    const double mieM = (1.0 + cg * cg) / std::pow(1.0 + c[8] * c[8] - 2.0 * c[8] * cg, 1.5);
    const double zenith = std::sqrt(ct);
    return (1.0 + c[0] * std::exp(c[1] / (ct + 0.01))) * (c[2] + c[3] * expM + c[5] * rayM + c[6] * mieM + c[7] * zenith) * st.radiance;
}

// piecewise-linear spectrum with zero outside its support
struct Curve {
    std::vector<float> x, y;
    Curve(const float *xs, const float *ys, size_t n) : x(xs, xs + n), y(ys, ys + n) {}
    float operator()(float l) const {
        if (x.size() < 2 || l < x.front() || l > x.back()) return 0.0f;
        auto r = std::equal_range(x.begin(), x.end(), l);
        size_t i1 = r.first - x.begin(), i2 = r.second - x.begin();
        if (i1 != i2) return y[i1];
        const float a = x[i1 - 1], b = x[i1], t = (l - a) / (b - a);
        return (1.0f - t) * y[i1 - 1] + t * y[i1];
    }
};
// mean of f*g over [a,b]: both are piecewise linear, so Simpson's rule per linear piece is exact.  (The reference
// uses an adaptive Gauss-Lobatto rule with 1e-4 tolerance, spectrum.cpp:546-568; the results agree to that tolerance.)
double meanProduct(const Curve &f, const Curve &g, float a, float b) {
    std::vector<float> k;
    for (float v : f.x) if (v > a && v < b) k.push_back(v);
    for (float v : g.x) if (v > a && v < b) k.push_back(v);
    k.push_back(a); k.push_back(b);
    std::sort(k.begin(), k.end());
    k.erase(std::unique(k.begin(), k.end()), k.end());
    double sum = 0;
    for (size_t i = 0; i + 1 < k.size(); ++i) {
        const double lo = k[i], hi = k[i + 1];
        const double f0 = f((float) lo), f1 = f((float) hi), g0 = g((float) lo), g1 = g((float) hi);
        sum += (hi - lo) / 6.0 * (f0 * g0 + 4.0 * (0.5 * (f0 + f1)) * (0.5 * (g0 + g1)) + f1 * g1);
    }
    return sum / ((double) b - (double) a);
}

// Preetham et al. sun: direct solar spectrum attenuated by Rayleigh, aerosol, ozone, mixed-gas and water-vapour terms
// (physical tables from Iqbal, "An Introduction to Solar Radiation")
bool sunRadianceRGB(float theta, float turbidity, const std::vector<float> &cie, float out[3]) {
    static const float koW[64] = {300,305,310,315,320,325,330,335,340,345,350,355,445,450,455,460,465,470,475,480,485,490,495,500,505,510,515,520,525,530,535,540,545,550,555,560,565,570,575,580,585,590,595,600,605,610,620,630,640,650,660,670,680,690,700,710,720,730,740,750,760,770,780,790};
    static const float koA[64] = {10.0f,4.8f,2.7f,1.35f,.8f,.380f,.160f,.075f,.04f,.019f,.007f,.0f,.003f,.003f,.004f,.006f,.008f,.009f,.012f,.014f,.017f,.021f,.025f,.03f,.035f,.04f,.045f,.048f,.057f,.063f,.07f,.075f,.08f,.085f,.095f,.103f,.110f,.12f,.122f,.12f,.118f,.115f,.12f,.125f,.130f,.12f,.105f,.09f,.079f,.067f,.057f,.048f,.036f,.028f,.023f,.018f,.014f,.011f,.010f,.009f,.007f,.004f,.0f,.0f};
    static const float kgW[4] = {759,760,770,771}, kgA[4] = {0,3.0f,0.210f,0};
    static const float kwW[13] = {689,690,700,710,720,730,740,750,760,770,780,790,800};
    static const float kwA[13] = {0,0.160e-1f,0.240e-1f,0.125e-1f,0.100e+1f,0.870f,0.610e-1f,0.100e-2f,0.100e-4f,0.100e-4f,0.600e-3f,0.175e-1f,0.360e-1f};
    static const float solW[38] = {380,390,400,410,420,430,440,450,460,470,480,490,500,510,520,530,540,550,560,570,580,590,600,610,620,630,640,650,660,670,680,690,700,710,720,730,740,750};
    static const float solA[38] = {16559.0f,16233.7f,21127.5f,25888.2f,25829.1f,24232.3f,26760.5f,29658.3f,30545.4f,30057.5f,30663.7f,28830.4f,28712.1f,27825.0f,27100.6f,27233.6f,26361.3f,25503.8f,25060.2f,25311.6f,25355.9f,25134.2f,24631.5f,24173.2f,23685.3f,23212.1f,22827.7f,22339.8f,21970.2f,21526.7f,21097.9f,20728.3f,20240.4f,19870.8f,19427.2f,19072.4f,18628.9f,18259.2f};
    const Curve ko(koW, koA, 64), kg(kgW, kgA, 4), kwa(kwW, kwA, 13), sol(solW, solA, 38);
    float data[91], wl[91];
    const float beta = 0.04608365822050f * turbidity - 0.04586025928522f;
    const float m = 1.0f / (std::cos(theta) + 0.15f * std::pow(93.885f - theta / PI_F * 180.0f, -1.253f));   // relative optical mass
    float lambda = 350;
    for (int i = 0; i < 91; ++i, lambda += 5) {
        const float um = lambda / 1000.0f;
        const float tauR = std::exp(-m * 0.008735f * std::pow(um, -4.08f));
        const float tauA = std::exp(-m * beta * std::pow(um, -1.3f));
        const float tauO = std::exp(-m * ko(lambda) * .35f);
        const float tauG = std::exp(-1.41f * kg(lambda) * m / std::pow(1 + 118.93f * kg(lambda) * m, 0.45f));
        const float tauWA = std::exp(-0.2385f * kwa(lambda) * 2.0f * m / std::pow(1 + 20.07f * kwa(lambda) * 2.0f * m, 0.45f));
        data[i] = sol(lambda) * tauR * tauA * tauO * tauG * tauWA;
        wl[i] = lambda;
    }
    const Curve spectrum(wl, data, 91);
    const size_t n = cie.size() / 4;
    const Curve cx(cie.data(), cie.data() + n, n), cy(cie.data(), cie.data() + 2 * n, n), cz(cie.data(), cie.data() + 3 * n, n);
    std::vector<float> ones(n, 1.0f);
    const Curve unit(cie.data(), ones.data(), n);
    const float a = cie[0], b = cie[n - 1];
    float X = (float) meanProduct(spectrum, cx, a, b), Y = (float) meanProduct(spectrum, cy, a, b), Z = (float) meanProduct(spectrum, cz, a, b);
    const float norm = 1.0f / (float) meanProduct(unit, cy, a, b);
    X *= norm; Y *= norm; Z *= norm;
    // XYZ -> linear Rec.709 (spectrum.cpp:222-227), negative lobes clamped (sunmodel.h:368)
    out[0] = std::max(0.0f, 3.240479f * X + -1.537150f * Y + -0.498535f * Z);
    out[1] = std::max(0.0f, -0.969256f * X + 1.875991f * Y + 0.041556f * Z);
    out[2] = std::max(0.0f, 0.055648f * X + -0.204043f * Y + 1.057311f * Z);
    return true;
}

bool readAll(const std::string &path, void *dst, size_t bytes) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    const bool ok = std::fread(dst, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok;
}
inline float safeAcos(float v) { return std::acos(std::min(1.0f, std::max(-1.0f, v))); }

} // namespace

bool bake_sunsky(const std::string &dataDir, const SunSkyParams &P, std::vector<float> &rgb, int &W, int &H, std::string &err, float *sunRadianceOut) {
    std::vector<double> hosek(3 * 1200);
    std::vector<float> cie(4 * 471);
    if (!readAll(dataDir + "/sunsky/hosek_rgb.f64", hosek.data(), hosek.size() * 8)) { err = "cannot read " + dataDir + "/sunsky/hosek_rgb.f64"; return false; }
    if (!readAll(dataDir + "/cie1931.f32", cie.data(), cie.size() * 4)) { err = "cannot read " + dataDir + "/cie1931.f32"; return false; }
    if (P.turbidity < 1 || P.turbidity > 10) { err = "The turbidity parameter must be in the range [1,10]!"; return false; }
    if (P.stretch < 1 || P.stretch > 2) { err = "The stretch parameter must be in the range [1,2]!"; return false; }
    if (P.sunRadiusScale == 0) { err = "sunRadiusScale = 0 (directional sun) is not supported on this path"; return false; }
    W = P.resolution; H = P.resolution / 2;
    rgb.assign((size_t) 3 * W * H, 0.0f);
    // sun position from the direction vector (sunmodel.h:98-105,206-208)
    float sx = P.sunDirection[0], sy = P.sunDirection[1], sz = P.sunDirection[2];
    { const float r = 1.0f / std::sqrt(sx * sx + sy * sy + sz * sz); sx *= r; sy *= r; sz *= r; }
    float sunAz = std::atan2(sx, -sz); const float sunEl = safeAcos(sy);
    if (sunAz < 0) sunAz += 2 * PI_F;
    const float sunElevationAboveHorizon = 0.5f * PI_F - sunEl;
    if (sunElevationAboveHorizon < 0) { err = "The sun is below the horizon -- this is not supported by the sky model."; return false; }
    SkyState st[3];
    for (int c = 0; c < 3; ++c)
        if (!cookState(hosek.data() + 1200 * c, hosek.data() + 1200 * c + 1080, P.turbidity, P.albedo[c], sunElevationAboveHorizon, st[c])) { err = "sky model: turbidity out of range"; return false; }
    const float fx = (2 * PI_F) / W, fy = PI_F / H;
    const float cosSunEl = std::cos(sunEl), sinSunEl = std::sin(sunEl);
    for (int y = 0; y < H; ++y) {
        const float th0 = (y + .5f) * fy;
        for (int x = 0; x < W; ++x) {
            const float ph0 = (x + .5f) * fx;
            // toSphere then fromSphere, as the reference routes each texel through a ray direction (sunsky.cpp:140-145, sky.cpp:392)
            const float dx = std::sin(ph0) * std::sin(th0), dy = std::cos(th0), dz = -std::cos(ph0) * std::sin(th0);
            float az = std::atan2(dx, -dz); const float el = safeAcos(dy);
            if (az < 0) az += 2 * PI_F;
            const float theta = el / P.stretch;
            float *t = &rgb[3 * ((size_t) y * W + x)];
            if (std::cos(theta) <= 0) continue;                 // below the horizon: black (extend = false)
            const float cosGamma = std::cos(theta) * cosSunEl + std::sin(theta) * sinSunEl * std::cos(az - sunAz);
            const float gamma = safeAcos(cosGamma);
            for (int c = 0; c < 3; ++c)
                t[c] = std::max((float) (skyRadiance(st[c], theta, gamma) / 106.856980), 0.0f) * P.skyScale;
        }
    }
    // sun disc (sunsky.cpp:163-216)
    float sunRGB[3];
    if (!sunRadianceRGB(sunEl, P.turbidity, cie, sunRGB)) { err = "sun radiance failed"; return false; }
    for (int c = 0; c < 3; ++c) sunRGB[c] *= P.sunScale;
    if (sunRadianceOut) for (int c = 0; c < 3; ++c) sunRadianceOut[c] = sunRGB[c];
    const float sEl = sunEl * P.stretch;
    const V3 sunDir(std::sin(sunAz) * std::sin(sEl), std::cos(sEl), -std::cos(sunAz) * std::sin(sEl));
    const Frame sunFrame(sunDir);
    const float halfAngle = (0.5358f * 0.5f) * (PI_F / 180.0f);
    const size_t pixelCount = (size_t) P.resolution * P.resolution / 2;
    const float cosCut = std::cos(halfAngle * P.sunRadiusScale);
    const float covered = 0.5f * (1 - cosCut);
    const size_t nSamples = (size_t) std::max(100.0f, (pixelCount * covered * 1000));
    const float gx = W / (2 * PI_F), gy = H / PI_F;
    // sunRadiance * solidAngle * texels / (2 pi^2 n), one rounding per Spectrum operation; Spectrum / Float multiplies by the reciprocal (sunsky.cpp:195-198, spectrum.h:415-425)
    const float solid = 2 * PI_F * (1 - std::cos(halfAngle)), texels = (float) (W * H), recip = 1.0f / (2 * PI_F * PI_F * nSamples);
    const V3 value(sunRGB[0] * solid * texels * recip, sunRGB[1] * solid * texels * recip, sunRGB[2] * solid * texels * recip);
    for (size_t i = 0; i < nSamples; ++i) {
        // (0,2)-sequence point: van der Corput radical inverse and Sobol' dimension 2
        uint32_t n = (uint32_t) i, v = __builtin_bswap32(n);
        v = ((v & 0x0f0f0f0f) << 4) | ((v & 0xf0f0f0f0) >> 4);
        v = ((v & 0x33333333) << 2) | ((v & 0xcccccccc) >> 2);
        v = ((v & 0x55555555) << 1) | ((v & 0xaaaaaaaa) >> 1);
        const float u0 = (float) (v >> 8) / (float) (1U << 24);
        uint32_t sob = 0;
        for (uint32_t dirv = 1U << 31, nn = n; nn != 0; nn >>= 1, dirv ^= dirv >> 1) if (nn & 1) sob ^= dirv;
        const float u1 = (float) sob / (float) (1ULL << 32);
        // squareToUniformCone
        const float ct = (1 - u0) + u0 * cosCut, stn = std::sqrt(std::max(0.0f, 1.0f - ct * ct)), ph = 2.0f * PI_F * u1;
        const V3 dir = sunFrame.toWorld(V3(std::cos(ph) * stn, std::sin(ph) * stn, ct));
        const float sinTheta = std::sqrt(std::max(0.0f, 1 - dir.y * dir.y));
        float az = std::atan2(dir.x, -dir.z); const float el = safeAcos(dir.y);
        if (az < 0) az += 2 * PI_F;
        const int px = std::min(std::max(0, (int) (az * gx)), W - 1), py = std::min(std::max(0, (int) (el * gy)), H - 1);
        const V3 add = value / std::max(1e-3f, sinTheta);
        float *t = &rgb[3 * ((size_t) py * W + px)];
        t[0] += add.x; t[1] += add.y; t[2] += add.z;
    }
    return true;
}

} // namespace cp

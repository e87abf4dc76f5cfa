// oracle/o_bsdf.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Restates the two hair BSDF plugins of the reference, quirks included:
//   KajiyaKay           src/bsdfs/kajiyakay.cpp:60-273
//   MarschnerDiffuse    src/bsdfs/marschner_diffuse.cpp:39-109,113-160,193-247,279-374,377-847
//   (the plugin that is actually built as `marschner`, src/bsdfs/SConscript:30-31)
//   RoughTransmittance  src/bsdfs/rtrans.h:81-149,183-408; splines src/libcore/spline.cpp:23-60,236-304,~380-447
//   GaussLegendre<140>  src/bsdfs/gausssexylingerie.hpp;  InterpolatedDistribution1D.hpp
#pragma once
#include "o_math.h"
#include <cstdio>
#include <memory>

namespace orc {

// BSDF::EBSDFType bits used on this path (include/mitsuba/render/bsdf.h:230-270)
enum { ENull = 0x1, EDiffuseReflection = 0x2, EGlossyReflection = 0x8, EDeltaReflection = 0x20, EDeltaTransmission = 0x40,
       EDelta = ENull | EDeltaReflection | EDeltaTransmission };

struct BSDFSample {
    V3 wo; V3 weight; float pdf = 0; int sampledType = 0; int sampledComponent = -1; float eta = 1;
};

// ---------------------------------------------------------------------------------------------
// KajiyaKay
// ---------------------------------------------------------------------------------------------
struct KajiyaKay {
    V3 diffuse, specular; float exponent; float specularSamplingWeight;

    // kajiyakay.cpp:60-107 (+ ensureEnergyConservation src/librender/bsdf.cpp:115-146)
    void configure(V3 diff, V3 spec, float expo) {
        float actualMax = maxc(spec + diff);
        if (actualMax > 1.0f) {
            float scale = 0.99f * (1.0f / actualMax);
            spec = spec * scale; diff = diff * scale;
        }
        diffuse = diff; specular = spec; exponent = expo;
        float dAvg = luminance(diff), sAvg = luminance(spec);
        specularSamplingWeight = sAvg / (dAvg + sAvg);
    }
    static V3 reflect(const V3 &wi) { return V3(-wi.x, -wi.y, wi.z); }

    // kajiyakay.cpp:122-180
    V3 eval(const V3 &wi, const V3 &wo) const {
        if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
        V3 result(0.0f);
        float tl = std::abs(wi.x), te = std::abs(wo.x);
        float sin_tl = std::sqrt(1 - tl * tl), sin_te = std::sqrt(1 - te * te);
        float alpha = tl * te + sin_tl * sin_te;
        if (alpha > 0.0f && wi.x * wo.x < 0) {
            V3 res = 0.15f * specular * ((exponent + 2) * kInvFourPi * cr::pow(alpha, exponent));
            result += res;
        }
        result += diffuse * kInvPi;
        return result * wo.z;
    }
    // kajiyakay.cpp:182-214
    float pdf(const V3 &wi, const V3 &wo) const {
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        float diffuseProb = kInvPi * wo.z; // warp::squareToCosineHemispherePdf
        float specProb = 0.0f;
        float alpha = dot(wo, reflect(wi));
        if (alpha > 0) specProb = cr::pow(alpha, exponent) * (exponent + 1.0f) / (2.0f * kPi);
        return specularSamplingWeight * specProb + (1 - specularSamplingWeight) * diffuseProb;
    }
    // kajiyakay.cpp:216-273
    BSDFSample sample(const V3 &wi, float sx, float sy) const {
        BSDFSample r; r.weight = V3(0.0f);
        bool choseSpecular = true;
        if (sx <= specularSamplingWeight) {
            sx /= specularSamplingWeight;
        } else {
            sx = (sx - specularSamplingWeight) / (1 - specularSamplingWeight);
            choseSpecular = false;
        }
        if (choseSpecular) {
            V3 R = reflect(wi);
            float sinAlpha = std::sqrt(1 - cr::pow(sy, 2 / (exponent + 1)));
            float cosAlpha = cr::pow(sy, 1 / (exponent + 1));
            float phi = (2.0f * kPi) * sx;
            V3 localDir(sinAlpha * cr::cos(phi), sinAlpha * cr::sin(phi), cosAlpha);
            r.wo = Frame(R).toWorld(localDir);
            r.sampledComponent = 1; r.sampledType = EGlossyReflection; // labels swapped in the reference
            if (r.wo.z <= 0) return r;
        } else {
            r.wo = squareToCosineHemisphere(sx, sy);
            r.sampledComponent = 0; r.sampledType = EDiffuseReflection;
        }
        r.eta = 1.0f;
        r.pdf = pdf(wi, r.wo);
        if (r.pdf == 0) return r;
        r.weight = eval(wi, r.wo) / r.pdf;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// Cubic-spline helpers (src/libcore/spline.cpp)
// ---------------------------------------------------------------------------------------------
// spline.cpp:23-60
static inline float evalCubicInterp1D(float x, const float *values, size_t size, float mn, float mx) {
    if (!(x >= mn && x <= mx)) return 0.0f;
    float t = ((x - mn) * (size - 1)) / (mx - mn);
    size_t k = std::max((size_t) 0, std::min((size_t) t, size - 2));
    float f0 = values[k], f1 = values[k + 1], d0, d1;
    if (k > 0) d0 = 0.5f * (values[k + 1] - values[k - 1]); else d0 = values[k + 1] - values[k];
    if (k + 2 < size) d1 = 0.5f * (values[k + 2] - values[k]); else d1 = values[k + 1] - values[k];
    t = t - (float) k;
    float t2 = t * t, t3 = t2 * t;
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
}
// per-dimension knot weights shared by the 2-D/3-D variants (spline.cpp:242-287)
static inline bool splineWeights(float p, size_t size, size_t &knot, float *weights) {
    if (!(p >= 0.0f && p <= 1.0f)) return false;
    float t = ((p - 0.0f) * (size - 1)) / (1.0f - 0.0f);
    knot = std::min((size_t) t, size - 2);
    t = t - (float) knot;
    float t2 = t * t, t3 = t2 * t;
    weights[0] = 0.0f; weights[1] = 2 * t3 - 3 * t2 + 1; weights[2] = -2 * t3 + 3 * t2; weights[3] = 0.0f;
    float d0 = t3 - 2 * t2 + t, d1 = t3 - t2;
    if (knot > 0) { weights[2] += 0.5f * d0; weights[0] -= 0.5f * d0; }
    else { weights[2] += d0; weights[1] -= d0; }
    if (knot + 2 < size) { weights[3] += 0.5f * d1; weights[1] -= 0.5f * d1; }
    else { weights[2] += d1; weights[1] -= d1; }
    return true;
}
// spline.cpp:236-304 (min=0,max=1)
static inline float evalCubicInterp2D(float px, float py, const float *values, size_t sx, size_t sy) {
    float w[2][4]; size_t knot[2];
    if (!splineWeights(px, sx, knot[0], w[0])) return 0.0f;
    if (!splineWeights(py, sy, knot[1], w[1])) return 0.0f;
    float result = 0.0f;
    for (int y = -1; y <= 2; ++y) {
        float wy = w[1][y + 1];
        for (int x = -1; x <= 2; ++x) {
            float wxy = w[0][x + 1] * wy;
            if (wxy == 0) continue;
            size_t pos = (knot[1] + y) * sx + knot[0] + x;
            result += values[pos] * wxy;
        }
    }
    return result;
}
// spline.cpp (evalCubicInterp3D, ~:380-447)
static inline float evalCubicInterp3D(float px, float py, float pz, const float *values, size_t sx, size_t sy, size_t sz) {
    float w[3][4]; size_t knot[3];
    if (!splineWeights(px, sx, knot[0], w[0])) return 0.0f;
    if (!splineWeights(py, sy, knot[1], w[1])) return 0.0f;
    if (!splineWeights(pz, sz, knot[2], w[2])) return 0.0f;
    float result = 0.0f;
    for (int z = -1; z <= 2; ++z) {
        float wz = w[2][z + 1];
        for (int y = -1; y <= 2; ++y) {
            float wyz = w[1][y + 1] * wz;
            for (int x = -1; x <= 2; ++x) {
                float wxyz = w[0][x + 1] * wyz;
                if (wxyz == 0) continue;
                size_t pos = ((knot[2] + z) * sy + (knot[1] + y)) * sx + knot[0] + x;
                result += values[pos] * wxyz;
            }
        }
    }
    return result;
}

// ---------------------------------------------------------------------------------------------
// RoughTransmittance (src/bsdfs/rtrans.h)
// ---------------------------------------------------------------------------------------------
struct RoughTransmittance {
    size_t etaSamples = 0, alphaSamples = 0, thetaSamples = 0;
    bool etaFixed = false, alphaFixed = false;
    float etaMin = 0, etaMax = 0, alphaMin = 0, alphaMax = 0;
    std::vector<float> trans, diffTrans;

    // rtrans.h:81-149
    void load(const std::string &path) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("oracle: cannot open " + path);
        char hdr[17];
        if (std::fread(hdr, 1, 17, f) != 17 || std::memcmp(hdr, "MTS_TRANSMITTANCE", 17) != 0) { std::fclose(f); throw std::runtime_error("oracle: bad transmittance file"); }
        uint64_t sz[3];
        if (std::fread(sz, 8, 3, f) != 3) { std::fclose(f); throw std::runtime_error("oracle: bad transmittance file"); }
        etaSamples = sz[0]; alphaSamples = sz[1]; thetaSamples = sz[2];
        size_t transSize = 2 * etaSamples * alphaSamples * thetaSamples, diffSize = 2 * etaSamples * alphaSamples;
        float rng[4];
        if (std::fread(rng, 4, 4, f) != 4) { std::fclose(f); throw std::runtime_error("oracle: bad transmittance file"); }
        etaMin = rng[0]; etaMax = rng[1]; alphaMin = rng[2]; alphaMax = rng[3];
        std::vector<float> temp(transSize + diffSize);
        if (std::fread(temp.data(), 4, temp.size(), f) != temp.size()) { std::fclose(f); throw std::runtime_error("oracle: truncated transmittance file"); }
        std::fclose(f);
        trans.resize(transSize); diffTrans.resize(diffSize);
        const float *ptr = temp.data(); size_t fdr = 0, de = 0;
        for (size_t i = 0; i < 2 * etaSamples; ++i)
            for (size_t j = 0; j < alphaSamples; ++j) {
                for (size_t k = 0; k < thetaSamples; ++k) trans[de++] = *ptr++;
                diffTrans[fdr++] = *ptr++;
            }
        etaFixed = alphaFixed = false;
    }
    float warpAlpha(float alpha) const { return cr::pow((alpha - alphaMin) / (alphaMax - alphaMin), 0.25f); }
    // rtrans.h:292-343
    void setEta(float eta) {
        if (etaFixed) return;
        const float *tr = trans.data(), *dt = diffTrans.data();
        if (eta < 1) { tr += etaSamples * alphaSamples * thetaSamples; dt += etaSamples * alphaSamples; eta = 1.0f / eta; }
        if (eta < etaMin) eta = etaMin;
        float warpedEta = cr::pow((eta - etaMin) / (etaMax - etaMin), 0.25f);
        std::vector<float> nt(alphaSamples * thetaSamples), nd(alphaSamples);
        float dAlpha = 1.0f / (alphaSamples - 1), dTheta = 1.0f / (thetaSamples - 1);
        for (size_t i = 0; i < alphaSamples; ++i) {
            for (size_t j = 0; j < thetaSamples; ++j)
                nt[i * thetaSamples + j] = evalCubicInterp3D(j * dTheta, i * dAlpha, warpedEta, tr, thetaSamples, alphaSamples, etaSamples);
            nd[i] = evalCubicInterp2D(i * dAlpha, warpedEta, dt, alphaSamples, etaSamples);
        }
        trans.swap(nt); diffTrans.swap(nd); etaFixed = true;
    }
    // rtrans.h:351-384
    void setAlpha(float alpha) {
        if (!etaFixed) throw std::runtime_error("setAlpha(): needs a preceding call to setEta()!");
        if (alphaFixed) return;
        float wa = warpAlpha(alpha);
        std::vector<float> nt(thetaSamples), nd(1);
        float dTheta = 1.0f / (thetaSamples - 1);
        for (size_t i = 0; i < thetaSamples; ++i)
            nt[i] = evalCubicInterp2D(i * dTheta, wa, trans.data(), thetaSamples, alphaSamples);
        nd[0] = evalCubicInterp1D(wa, diffTrans.data(), alphaSamples, 0.0f, 1.0f);
        trans.swap(nt); diffTrans.swap(nd); alphaFixed = true;
    }
    // rtrans.h:183-234 (only the branches reachable on this path)
    float eval(float cosTheta, float alpha = 0) const {
        float warpedCosTheta = cr::pow(std::abs(cosTheta), 0.25f), result;
        if (alphaFixed && etaFixed) {
            if (!(cosTheta >= 0)) return 0.f;
            result = evalCubicInterp1D(warpedCosTheta, trans.data(), thetaSamples, 0.0f, 1.0f);
        } else if (etaFixed) {
            if (!(cosTheta >= 0)) return 0.f;
            result = evalCubicInterp2D(warpedCosTheta, warpAlpha(alpha), trans.data(), thetaSamples, alphaSamples);
        } else throw std::runtime_error("oracle: 3-D rough transmittance lookup not on this path");
        return std::min(1.0f, std::max(0.0f, result));
    }
    // rtrans.h:249-290
    float evalDiffuse(float alpha = 0) const {
        float result;
        if (alphaFixed && etaFixed) result = diffTrans[0];
        else if (etaFixed) result = evalCubicInterp1D(warpAlpha(alpha), diffTrans.data(), alphaSamples, 0.0f, 1.0f);
        else throw std::runtime_error("oracle: 2-D diffuse transmittance lookup not on this path");
        return std::min(1.0f, std::max(0.0f, result));
    }
    void checkRanges(float eta, float alpha) const { // rtrans.h:386-405
        if (alpha < alphaMin || alpha > alphaMax) throw std::runtime_error("roughness outside the supported range");
        if (eta < 1) eta = 1 / eta;
        if (eta < etaMin || eta > etaMax) throw std::runtime_error("IOR outside the supported range");
    }
};

// ---------------------------------------------------------------------------------------------
// Gauss-Legendre nodes (gausssexylingerie.hpp:14-68) -- roots/weights in double, rounded to float
// ---------------------------------------------------------------------------------------------
template <int N> struct GaussLegendre {
    float points[N], weights[N];
    static double legendre(double x, int n) {
        if (n == 0) return 1.0;
        if (n == 1) return x;
        double P0 = 1.0, P1 = x;
        for (int i = 2; i <= n; ++i) { double Pi = ((2.0 * i - 1.0) * x * P1 - (i - 1.0) * P0) / i; P0 = P1; P1 = Pi; }
        return P1;
    }
    static double legendreDeriv(double x, int n) { return n / (x * x - 1.0) * (x * legendre(x, n) - legendre(x, n - 1)); }
    static double kthRoot(int k) {
        double x = std::cos(kPi * (4.0 * k - 1.0) / (4.0 * N + 2.0)) * (1.0 - 1.0 / (8.0 * N * N) + 1.0 / (8.0 * N * N * N));
        for (int i = 0; i < 100; ++i) {
            double f = legendre(x, N);
            x -= f / legendreDeriv(x, N);
            if (std::abs(f) < 1e-6) break;
        }
        return x;
    }
    GaussLegendre() {
        for (int i = 0; i < N; ++i) {
            points[i] = float(kthRoot(i + 1));
            weights[i] = float(2.0 / ((1.0 - points[i] * points[i]) * legendreDeriv(points[i], N) * legendreDeriv(points[i], N)));
        }
    }
};

// InterpolatedDistribution1D.hpp:37-110
struct InterpolatedDistribution1D {
    int size = 0, num = 0;
    std::vector<float> pdfs, cdfs, sums;
    float &cdf(int x, int d) { return cdfs[x + d * (size + 1)]; }
    float cdf(int x, int d) const { return cdfs[x + d * (size + 1)]; }
    float &pdfv(int x, int d) { return pdfs[x + d * size]; }
    float pdfv(int x, int d) const { return pdfs[x + d * size]; }
    void init(std::vector<float> weights, int size_, int num_) {
        size = size_; num = num_; pdfs = std::move(weights);
        cdfs.assign((size + 1) * num, 0.0f); sums.assign(num, 0.0f);
        for (int dist = 0; dist < num; ++dist) {
            cdf(0, dist) = 0.0f;
            for (int x = 0; x < size; ++x) cdf(x + 1, dist) = pdfv(x, dist) + cdf(x, dist);
            sums[dist] = cdf(size, dist);
            if (sums[dist] < 1e-4f) {
                float ratio = 1.0f / size;
                for (int x = 0; x < size; ++x) { pdfv(x, dist) = ratio; cdf(x, dist) = x * ratio; }
            } else {
                float scale = 1.0f / sums[dist];
                for (int x = 0; x < size; ++x) { pdfv(x, dist) *= scale; cdf(x, dist) *= scale; }
            }
            cdf(size, dist) = 1.0f;
        }
    }
    void warp(float distribution, float &u, int &x) const {
        int d0 = clampi(int(distribution), 0, num - 1);
        int d1 = std::min(d0 + 1, num - 1);
        float v = clampf(distribution - d0, 0.0f, 1.0f);
        int lower = 0, upper = size;
        float lowerU = 0.0f, upperU = 1.0f;
        while (upper - lower != 1) {
            int midpoint = (upper + lower) / 2;
            float midpointU = cdf(midpoint, d0) * (1.0f - v) + cdf(midpoint, d1) * v;
            if (midpointU < u) { lower = midpoint; lowerU = midpointU; }
            else { upper = midpoint; upperU = midpointU; }
        }
        x = lower;
        u = clampf((u - lowerU) / (upperU - lowerU), 0.0f, 1.0f);
    }
    float pdf(float distribution, int x) const { // InterpolatedDistribution1D.hpp:94-101
        int d0 = clampi(int(distribution), 0, num - 1);
        int d1 = std::min(d0 + 1, num - 1);
        float v = clampf(distribution - d0, 0.0f, 1.0f);
        return pdfv(x, d0) * (1.0f - v) + pdfv(x, d1) * v;
    }
    float sum(float distribution) const {
        int d0 = clampi(int(distribution), 0, num - 1);
        int d1 = std::min(d0 + 1, num - 1);
        float v = clampf(distribution - d0, 0.0f, 1.0f);
        return sums[d0] * (1.0f - v) + sums[d1] * v;
    }
};

// marschner_diffuse.cpp:39-109
struct Azimuthal {
    static const int Res = 64;
    std::vector<V3> table;
    InterpolatedDistribution1D sampler;
    void init(std::vector<V3> t) {
        table = std::move(t);
        const int Size = Res;
        std::vector<float> weights(Size * Size);
        for (int i = 0; i < Size * Size; ++i) weights[i] = maxc(table[i]);
        for (int y = 0; y < Size; ++y) {
            for (int x = 0; x < Size - 1; ++x) weights[x + y * Size] = std::max(weights[x + y * Size], weights[x + 1 + y * Size]);
            for (int x = Size - 1; x > 0; --x) weights[x + y * Size] = std::max(weights[x + y * Size], weights[x - 1 + y * Size]);
        }
        for (int x = 0; x < Size; ++x) {
            for (int y = 0; y < Size - 1; ++y) weights[x + y * Size] = std::max(weights[x + y * Size], weights[x + (y + 1) * Size]);
            for (int y = Size - 1; y > 0; --y) weights[x + y * Size] = std::max(weights[x + y * Size], weights[x + (y - 1) * Size]);
        }
        sampler.init(std::move(weights), Size, Size);
    }
    void sample(float cosThetaD, float xi, float &phi) const {
        float v = (Res - 1) * cosThetaD;
        int x;
        sampler.warp(v, xi, x);
        phi = 2.0f * kPi * (x + xi) * (1.0f / Res);
    }
    V3 eval(float phi, float cosThetaD) const {
        float u = (Res - 1) * phi * (1.0f / (2.0f * kPi));
        float v = (Res - 1) * cosThetaD;
        int x0 = clampi(int(u), 0, Res - 2), y0 = clampi(int(v), 0, Res - 2);
        int x1 = x0 + 1, y1 = y0 + 1;
        u = clampf(u - x0, 0.0f, 1.0f);
        v = clampf(v - y0, 0.0f, 1.0f);
        return (table[x0 + y0 * Res] * (1.0f - u) + table[x1 + y0 * Res] * u) * (1.0f - v) +
               (table[x0 + y1 * Res] * (1.0f - u) + table[x1 + y1 * Res] * u) * v;
    }
    float weight(float cosThetaD) const {
        float v = (Res - 1) * cosThetaD;
        return sampler.sum(v) * (2.0f * kPi / Res);
    }
    // marschner.cpp:91-96 (the unbuilt `Marschner`): pdf of the azimuth in the cell that contains phi
    float pdf(float phi, float cosThetaD) const {
        float u = (Res - 1) * phi * (1.0f / (2.0f * kPi));
        float v = (Res - 1) * cosThetaD;
        return sampler.pdf(v, int(u)) * float(Res * (1.0f / (2.0f * kPi)));
    }
};

// ---------------------------------------------------------------------------------------------
// MarschnerDiffuse (the `marschner` plugin as built)
// ---------------------------------------------------------------------------------------------
struct Marschner {
    float eta = 1.5046f / 1.000277f;
    V3 sigmaA = V3(0.5f), diffuse = V3(0.5f), specularReflectance = V3(0.5f);
    bool nonlinear = false;
    float alpha = 0.1f;
    int distribution = 0; // 0 beckmann, 1 ggx, 2 phong
    float betaR = 0.1f, betaTT = 0.05f, betaTRT = 0.2f, scaleAngleRad = -0.1f;
    float vR, vTT, vTRT, invEta2, specularSamplingWeight;
    Azimuthal nR, nTT, nTRT;
    RoughTransmittance extRT, intRT;

    static float I0(float x) { // :279-290
        float result = 1.0f, xSq = x * x, xi = xSq, denom = 4.0f;
        for (int i = 1; i <= 10; ++i) { result += xi / denom; xi *= xSq; denom *= 4.0f * float((i + 1) * (i + 1)); }
        return result;
    }
    static float logI0(float x) { // :292-299
        if (x > 12.0f) return x + 0.5f * (cr::log(1.0f / (kPi * 2.0f * x)) + 1.0f / (8.0f * x));
        else return cr::log(I0(x));
    }
    static float g(float beta, float theta) { return cr::exp(-theta * theta / (2.0f * beta * beta)) / (std::sqrt(2.0f * kPi) * beta); } // :301-303
    static float D(float beta, float phi) { // :305-315
        float result = 0.0f, delta, shift = 0.0f;
        do {
            delta = g(beta, phi + shift) + g(beta, phi - shift - 2 * kPi);
            result += delta;
            shift += 2 * kPi;
        } while (delta > 1e-4f);
        return result;
    }
    static float Phi(float gammaI, float gammaT, int p) { return 2.0f * p * gammaT - 2.0f * gammaI + p * kPi; } // :317-319
    static float M(float v, float sinThetaI, float sinThetaO, float cosThetaI, float cosThetaO) { // :364-374
        float a = cosThetaI * cosThetaO / v, b = sinThetaI * sinThetaO / v;
        if (v < 0.1f) return cr::exp(-b + logI0(a) - 1.0f / v + 0.6931f + cr::log(1.0f / (2.0f * v)));
        else return cr::exp(-b) * I0(a) / (2.0f * v * cr::sinh(1.0f / v));
    }
    static float trigInverse(float x) { return std::min(std::sqrt(std::max(1.0f - x * x, 0.0f)), 1.0f); } // :484-486

    // :751-847
    void precomputeAzimuthalDistributions() {
        const int Resolution = Azimuthal::Res;
        std::vector<V3> valuesR(Resolution * Resolution), valuesTT(Resolution * Resolution), valuesTRT(Resolution * Resolution);
        const int NumPoints = 140;
        GaussLegendre<NumPoints> integrator;
        const float *points = integrator.points, *weights = integrator.weights;
        float gammaIs[NumPoints];
        for (int i = 0; i < NumPoints; ++i) gammaIs[i] = cr::asin(points[i]);
        const int NumGaussianSamples = 2048;
        std::vector<float> Ds(NumGaussianSamples); // identical for p=0,1,2: all use _betaR (:778)
        for (int i = 0; i < NumGaussianSamples; ++i) Ds[i] = D(betaR, i / (NumGaussianSamples - 1.0f) * 2 * kPi);
        auto approxD = [&](float phi) {
            float u = (float) std::abs(phi * (1.0 / (2 * kPi) * (NumGaussianSamples - 1))); // double arithmetic as in :783
            int x0 = int(u), x1 = x0 + 1;
            u -= x0;
            return Ds[x0 % NumGaussianSamples] * (1.0f - u) + Ds[x1 % NumGaussianSamples] * u;
        };
        for (int y = 0; y < Resolution; ++y) {
            float cosHalfAngle = y / (Resolution - 1.0f);
            float iorPrime = std::sqrt(eta * eta - (1.0f - cosHalfAngle * cosHalfAngle)) / cosHalfAngle;
            float cosThetaT = std::sqrt(1.0f - (1.0f - cosHalfAngle * cosHalfAngle) * (1.0f / eta) * (1.0f / eta));
            V3 sigmaAPrime = sigmaA / cosThetaT;
            float fresnelTerms[NumPoints], gammaTs[NumPoints]; V3 absorptions[NumPoints];
            for (int i = 0; i < NumPoints; ++i) {
                gammaTs[i] = cr::asin(clampf(points[i] / iorPrime, -1.0f, 1.0f));
                fresnelTerms[i] = fresnelDielectricExt(1.0f / eta, cosHalfAngle * cr::cos(gammaIs[i])); // swapped arguments, verbatim (:809)
                V3 e = -sigmaAPrime * 2.0f * cr::cos(gammaTs[i]);
                absorptions[i] = V3(cr::exp(e.x), cr::exp(e.y), cr::exp(e.z));
            }
            for (int phiI = 0; phiI < Resolution; ++phiI) {
                float phi = kPi * 2 * phiI / (Resolution - 1.0f);
                float integralR = 0.0f; V3 integralTT(0.0f), integralTRT(0.0f);
                for (int i = 0; i < NumPoints; ++i) {
                    float fR = fresnelTerms[i]; V3 T = absorptions[i];
                    float AR = fR;
                    V3 ATT = (1.0f - fR) * (1.0f - fR) * T;
                    V3 ATRT = ATT * fR * T;
                    integralR += weights[i] * approxD(phi - Phi(gammaIs[i], gammaTs[i], 0)) * AR;
                    integralTT += weights[i] * approxD(phi - Phi(gammaIs[i], gammaTs[i], 1)) * ATT;
                    integralTRT += weights[i] * approxD(phi - Phi(gammaIs[i], gammaTs[i], 2)) * ATRT;
                }
                valuesR[phiI + y * Resolution] = V3(0.5f * integralR);
                valuesTT[phiI + y * Resolution] = 0.5f * integralTT;
                valuesTRT[phiI + y * Resolution] = 0.5f * integralTRT;
            }
        }
        nR.init(std::move(valuesR)); nTT.init(std::move(valuesTT)); nTRT.init(std::move(valuesTRT));
    }

    // ctor :113-160 + configure :193-247.  `dataDir` must contain microfacet/<distr>.dat
    void configure(float intIOR, float extIOR, V3 diff, V3 specRefl, float alpha_, int distr, bool nonlin, const std::string &dataDir) {
        eta = intIOR / extIOR;
        sigmaA = V3(0.5f);
        nonlinear = nonlin; distribution = distr;
        alpha = std::max(alpha_, 1e-4f); // microfacet.h:131
        precomputeAzimuthalDistributions();
        vR = betaR * betaR; vTT = betaTT * betaTT; vTRT = betaTRT * betaTRT;
        float mx = maxc(specRefl); // ensureEnergyConservation(specularReflectance, 1.0) bsdf.cpp:88-113
        if (mx > 1.0f) specRefl = specRefl * (0.99f * (1.0f / mx));
        specularReflectance = specRefl; diffuse = diff;
        float dAvg = luminance(diff), sAvg = luminance(specRefl);
        specularSamplingWeight = sAvg / (dAvg + sAvg);
        invEta2 = 1.0f / (eta * eta);
        static const char *names[3] = {"beckmann", "ggx", "phong"};
        extRT.load(dataDir + "/microfacet/" + names[distr] + ".dat");
        extRT.checkRanges(eta, alpha);
        intRT = extRT;          // clone() before setEta (:230)
        extRT.setEta(eta);
        intRT.setEta(1 / eta);
        extRT.setAlpha(alpha);
    }

    // :377-482
    V3 eval(const V3 &wi, const V3 &wo) const {
        float sinThetaI = wi.y, sinThetaO = wo.y;
        float cosThetaO = trigInverse(sinThetaO);
        float thetaI = cr::asin(clampf(sinThetaI, -1.0f, 1.0f));
        float thetaO = cr::asin(clampf(sinThetaO, -1.0f, 1.0f));
        float thetaD = (thetaO - thetaI) * 0.5f;
        float cosThetaD = cr::cos(thetaD);
        float phi = cr::atan2(wo.x, wo.z);
        if (phi < 0.0f) phi += kPi * 2.0f;
        float thetaIR = thetaI - 2.0f * scaleAngleRad;
        float thetaITT = thetaI + scaleAngleRad;
        float thetaITRT = thetaI + 4.0f * scaleAngleRad;
        float MR = M(vR, cr::sin(thetaIR), sinThetaO, cr::cos(thetaIR), cosThetaO);
        float MTT = M(vTT, cr::sin(thetaITT), sinThetaO, cr::cos(thetaITT), cosThetaO);
        float MTRT = M(vTRT, cr::sin(thetaITRT), sinThetaO, cr::cos(thetaITRT), cosThetaO);
        V3 result = 0.15f * MR * nR.eval(phi, cosThetaD) + MTT * nTT.eval(phi, cosThetaD) + MTRT * nTRT.eval(phi, cosThetaD);
        // diffuse term (typeMask=EAll, component=-1 on this path)
        V3 diff = diffuse;
        float T12 = extRT.eval(wi.z, alpha);
        float T21 = extRT.eval(wo.z, alpha);
        float Fdr = 1 - intRT.evalDiffuse(alpha);
        if (nonlinear) diff = V3(diff.x / (1.0f - diff.x * Fdr), diff.y / (1.0f - diff.y * Fdr), diff.z / (1.0f - diff.z * Fdr));
        else diff = diff / (1 - Fdr);
        result += diff * (kInvPi * wo.z * T12 * T21 * invEta2);
        return result;
    }
    // :488-520 -- constant 1 whenever the diffuse component is requested (always here)
    float pdf(const V3 &, const V3 &) const { return 1.0f; }
    // :582-592
    float sampleM(float v, float sinThetaI, float cosThetaI, float xi1, float xi2) const {
        float cosTheta = 1.0f + v * cr::log(xi1 + (1.0f - xi1) * cr::exp(-2.0f / v));
        float sinTheta = trigInverse(cosTheta);
        float cosPhi = cr::cos(2 * kPi * xi2);
        return -cosTheta * sinThetaI + sinTheta * cosPhi * cosThetaI;
    }
    // :594-744
    BSDFSample sample(const V3 &wi, float sx, float sy) const {
        BSDFSample r;
        float sinThetaI = wi.y;
        float cosThetaI = trigInverse(sinThetaI);
        float thetaI = cr::asin(clampf(sinThetaI, -1.0f, 1.0f));
        float thetaIR = thetaI - 2.0f * scaleAngleRad;
        float thetaITT = thetaI + scaleAngleRad;
        float thetaITRT = thetaI + 4.0f * scaleAngleRad;
        float weightR = nR.weight(cosThetaI), weightTT = nTT.weight(cosThetaI), weightTRT = nTRT.weight(cosThetaI);
        const Azimuthal *lobe; float v, theta;
        float target = sx * (weightR + weightTT + weightTRT);
        if (target < weightR) { r.sampledComponent = 5; v = vR; theta = thetaIR; lobe = &nR; }
        else if (target < weightR + weightTT) { r.sampledComponent = 6; v = vTT; theta = thetaITT; lobe = &nTT; }
        else { r.sampledComponent = 7; v = vTRT; theta = thetaITRT; lobe = &nTRT; }
        float sinThetaO = sampleM(v, cr::sin(theta), cr::cos(theta), sx, sy);
        float cosThetaO = trigInverse(sinThetaO);
        float thetaO = cr::asin(clampf(sinThetaO, -1.0f, 1.0f));
        float thetaD = (thetaO - thetaI) * 0.5f;
        float cosThetaD = cr::cos(thetaD);
        float phi;
        lobe->sample(cosThetaD, sy, phi);
        float sinPhi = cr::sin(phi), cosPhi = cr::cos(phi);
        float probSpecular = 1 - extRT.eval(wi.z, alpha);
        probSpecular = (probSpecular * specularSamplingWeight) /
                       (probSpecular * specularSamplingWeight + (1 - probSpecular) * (1 - specularSamplingWeight));
        bool choseSpecular = sy < probSpecular;
        if (choseSpecular) {
            r.wo = V3(sinPhi * cosThetaO, sinThetaO, cosPhi * cosThetaO);
            r.sampledType = EDeltaReflection;
        } else {
            r.sampledComponent = 1;
            r.sampledType = EDiffuseReflection;
            r.wo = squareToCosineHemisphere(sx, sy);
        }
        r.eta = 1.0f;
        r.pdf = 1.0f;
        r.weight = eval(wi, r.wo) / r.pdf;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// Marschner ("fixed" mode): src/bsdfs/marschner.cpp, the file the fork leaves out of the build.  Same azimuthal tables
// with sigmaA = 0.22 and the amber / air default IORs (:110-138); eval keeps only the TRT lobe (MR = MTT = 0, :333-334), pdf is the
// lobe-weighted product of M and the azimuthal cell pdf (:347-407), sample draws two EXTRA 2-D numbers from the sampler (xiN, xiM,
// :473-474), ignores its `sample` argument and rejects pdf <= 0 or pdf > 1 (:530).
// ---------------------------------------------------------------------------------------------
struct MarschnerFixed {
    Marschner base;      // tables, variances, scale angle
    int lobeMask = 4;    // lobes eval() keeps: bit 0 R, bit 1 TT, bit 2 TRT.  4 = the file as committed (MR = MTT = 0, :333-334); 7 = all three
    void configure(float intIOR, float extIOR) { configure(intIOR, extIOR, V3(0.22f), 0.1f, -0.1f, 4); }
    // the scene-driven variant: what the constructor hard-codes (:122, :131-137) comes from the scene
    void configure(float intIOR, float extIOR, V3 sigmaA, float betaR, float scaleAngleRad, int mask) {
        base.eta = intIOR / extIOR;
        base.sigmaA = sigmaA;
        base.betaR = betaR; base.betaTT = betaR * 0.5f; base.betaTRT = betaR * 2.0f;
        base.scaleAngleRad = scaleAngleRad;
        lobeMask = mask;
        base.precomputeAzimuthalDistributions();
        base.vR = base.betaR * base.betaR; base.vTT = base.betaTT * base.betaTT; base.vTRT = base.betaTRT * base.betaTRT;
    }
    V3 eval(const V3 &wi, const V3 &wo) const { // :309-341
        float sinThetaI = wi.y, sinThetaO = wo.y;
        float cosThetaO = Marschner::trigInverse(sinThetaO);
        float thetaI = cr::asin(clampf(sinThetaI, -1.0f, 1.0f));
        float thetaO = cr::asin(clampf(sinThetaO, -1.0f, 1.0f));
        float thetaD = (thetaO - thetaI) * 0.5f;
        float cosThetaD = cr::cos(thetaD);
        float phi = cr::atan2(wo.x, wo.z);
        if (phi < 0.0f) phi += kPi * 2.0f;
        float thetaIR = thetaI - 2.0f * base.scaleAngleRad, thetaITT = thetaI + base.scaleAngleRad, thetaITRT = thetaI + 4.0f * base.scaleAngleRad;
        // MR and MTT are computed and then zeroed in the reference; 0 * eval() contributes exactly +0 for finite tables
        float MR = (lobeMask & 1) ? Marschner::M(base.vR, cr::sin(thetaIR), sinThetaO, cr::cos(thetaIR), cosThetaO) : 0.0f;
        float MTT = (lobeMask & 2) ? Marschner::M(base.vTT, cr::sin(thetaITT), sinThetaO, cr::cos(thetaITT), cosThetaO) : 0.0f;
        float MTRT = (lobeMask & 4) ? Marschner::M(base.vTRT, cr::sin(thetaITRT), sinThetaO, cr::cos(thetaITRT), cosThetaO) : 0.0f;
        return MR * base.nR.eval(phi, cosThetaD) + MTT * base.nTT.eval(phi, cosThetaD) + MTRT * base.nTRT.eval(phi, cosThetaD);
    }
    float pdf(const V3 &wi, const V3 &wo) const { // :347-407
        float sinThetaI = wi.y, sinThetaO = wo.y;
        float cosThetaI = Marschner::trigInverse(sinThetaI), cosThetaO = Marschner::trigInverse(sinThetaO);
        float thetaI = cr::asin(clampf(sinThetaI, -1.0f, 1.0f));
        float thetaO = cr::asin(clampf(sinThetaO, -1.0f, 1.0f));
        float thetaD = (thetaO - thetaI) * 0.5f;
        float cosThetaD = cr::cos(thetaD);
        float phi = cr::atan2(wo.x, wo.z);
        if (phi < 0.0f) phi += 2.0f * kPi;
        float thetaIR = thetaI - 2.0f * base.scaleAngleRad, thetaITT = thetaI + base.scaleAngleRad, thetaITRT = thetaI + 4.0f * base.scaleAngleRad;
        float weightR = base.nR.weight(cosThetaI), weightTT = base.nTT.weight(cosThetaI), weightTRT = base.nTRT.weight(cosThetaI);
        float weightSum = weightR + weightTT + weightTRT;
        float pdfR = weightR * Marschner::M(base.vR, cr::sin(thetaIR), sinThetaO, cr::cos(thetaIR), cosThetaO);
        float pdfTT = weightTT * Marschner::M(base.vTT, cr::sin(thetaITT), sinThetaO, cr::cos(thetaITT), cosThetaO);
        float pdfTRT = weightTRT * Marschner::M(base.vTRT, cr::sin(thetaITRT), sinThetaO, cr::cos(thetaITRT), cosThetaO);
        return (1.0f / weightSum) * (pdfR * base.nR.pdf(phi, cosThetaD) + pdfTT * base.nTT.pdf(phi, cosThetaD) + pdfTRT * base.nTRT.pdf(phi, cosThetaD));
    }
    // :421-535; xiN / xiM are the two extra sampler->next2D() draws
    BSDFSample sample(const V3 &wi, float xiNx, float xiNy, float xiMx, float xiMy) const {
        BSDFSample r; r.weight = V3(0.0f);
        float sinThetaI = wi.y;
        float cosThetaI = Marschner::trigInverse(sinThetaI);
        float thetaI = cr::asin(clampf(sinThetaI, -1.0f, 1.0f));
        float thetaIR = thetaI - 2.0f * base.scaleAngleRad, thetaITT = thetaI + base.scaleAngleRad, thetaITRT = thetaI + 4.0f * base.scaleAngleRad;
        float weightR = base.nR.weight(cosThetaI), weightTT = base.nTT.weight(cosThetaI), weightTRT = base.nTRT.weight(cosThetaI);
        const Azimuthal *lobe; float v, theta;
        float target = xiNx * (weightR + weightTT + weightTRT);
        if (target < weightR) { r.sampledComponent = 0; v = base.vR; theta = thetaIR; lobe = &base.nR; }
        else if (target < weightR + weightTT) { r.sampledComponent = 1; v = base.vTT; theta = thetaITT; lobe = &base.nTT; }
        else { r.sampledComponent = 2; v = base.vTRT; theta = thetaITRT; lobe = &base.nTRT; }
        float sinThetaO = base.sampleM(v, cr::sin(theta), cr::cos(theta), xiMx, xiMy);
        float cosThetaO = Marschner::trigInverse(sinThetaO);
        float thetaO = cr::asin(clampf(sinThetaO, -1.0f, 1.0f));
        float thetaD = (thetaO - thetaI) * 0.5f;
        float cosThetaD = cr::cos(thetaD);
        float phi;
        lobe->sample(cosThetaD, xiNy, phi);
        float sinPhi = cr::sin(phi), cosPhi = cr::cos(phi);
        r.wo = V3(sinPhi * cosThetaO, sinThetaO, cosPhi * cosThetaO);
        r.pdf = pdf(wi, r.wo);
        r.sampledType = EDeltaReflection; r.eta = 1.0f;
        if (r.pdf <= 0 || r.pdf > 1) return r;               // also lets NaN through, like the reference's comparison
        r.weight = eval(wi, r.wo) / r.pdf;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// RoughPlastic (`roughplastic`, the BSDF of the default models/*/scene.xml files)
//   MicrofacetDistribution  src/bsdfs/microfacet.h:184-232 (eval), :238-279 (sample/pdf), :284-386 (sampleAll), :389-447 (visible
//   normals), :470-510 (smithG1, G), :529-538 (projectRoughness), :555-673 (sampleVisible11), :677-680 (Phong exponent)
//   math::erf / erfinv / hypot2  src/libcore/math.cpp:25-86;  RoughPlastic  src/bsdfs/roughplastic.cpp:186-235,258-304,325-494
// Isotropic roughness only (the plugin rejects anisotropic distributions, roughplastic.cpp:212-214).
// ---------------------------------------------------------------------------------------------
namespace mf {
static inline float signum(float v) { return v < 0 ? -1.0f : (v > 0 ? 1.0f : 0.0f); }
static inline float erfinv(float x) { // math.cpp:25-53
    float w = -cr::log((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
    } else {
        w = std::sqrt(w) - 3.0f;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
    }
    return p * x;
}
static inline float erf(float x) { // math.cpp:55-72
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    float sign = signum(x);
    x = std::abs(x);
    float t = 1.0f / (1.0f + p * x);
    float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * cr::exp(-x * x);
    return sign * y;
}
static inline float hypot2(float a, float b) { // math.cpp:74-86
    float r;
    if (std::abs(a) > std::abs(b)) { r = b / a; r = std::abs(a) * std::sqrt(1.0f + r * r); }
    else if (b != 0.0f) { r = a / b; r = std::abs(b) * std::sqrt(1.0f + r * r); }
    else r = 0.0f;
    return r;
}
}

struct MicrofacetDistribution {
    int type = 0;            // 0 beckmann, 1 ggx, 2 phong
    float alpha = 0.1f, exponent = 0;
    bool sampleVis = true;
    void configure(int t, float a, bool visible) {
        type = t; alpha = std::max(a, 1e-4f); sampleVis = visible;
        if (type == 2) { sampleVis = false; exponent = std::max(2.0f / (alpha * alpha) - 2.0f, 0.0f); }
    }
    float eval(const V3 &m) const { // :184-232
        if (m.z <= 0) return 0.0f;
        float cosTheta2 = m.z * m.z;
        float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        float result;
        if (type == 0) result = cr::exp(-beckmannExponent) / (kPi * alpha * alpha * cosTheta2 * cosTheta2);
        else if (type == 1) { float root = (1.0f + beckmannExponent) * cosTheta2; result = 1.0f / (kPi * alpha * alpha * root * root); }
        else result = std::sqrt((exponent + 2) * (exponent + 2)) * kInvTwoPi * cr::pow(m.z, exponent);
        if (result * m.z < 1e-20f) result = 0;
        return result;
    }
    float smithG1(const V3 &v, const V3 &m) const { // :470-505
        if (dot(v, m) * v.z <= 0) return 0.0f;
        float temp = 1 - v.z * v.z;                      // Frame::tanTheta (frame.h): sqrt(1 - cos^2) / cos, 0 when temp <= 0
        float tanTheta = temp <= 0.0f ? 0.0f : std::abs(std::sqrt(temp) / v.z);
        if (tanTheta == 0.0f) return 1.0f;
        if (type != 1) {
            float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        }
        float root = alpha * tanTheta;
        return 2.0f / (1.0f + mf::hypot2(1.0f, root));
    }
    float G(const V3 &wi, const V3 &wo, const V3 &m) const { return smithG1(wi, m) * smithG1(wo, m); }
    float pdfVisible(const V3 &wi, const V3 &m) const { // :442-447
        if (wi.z == 0) return 0.0f;
        return smithG1(wi, m) * std::abs(dot(wi, m)) * eval(m) / std::abs(wi.z);
    }
    float pdf(const V3 &wi, const V3 &m) const { return sampleVis ? pdfVisible(wi, m) : eval(m) * m.z; }
    V3 sampleAll(float sx, float sy) const { // :284-380 (isotropic)
        float cosThetaM, sinPhiM, cosPhiM;
        if (type == 0) {
            sinPhiM = cr::sin((2.0f * kPi) * sy); cosPhiM = cr::cos((2.0f * kPi) * sy);
            float tanThetaMSqr = alpha * alpha * -cr::log(1.0f - sx);
            cosThetaM = 1.0f / std::sqrt(1.0f + tanThetaMSqr);
        } else if (type == 1) {
            sinPhiM = cr::sin((2.0f * kPi) * sy); cosPhiM = cr::cos((2.0f * kPi) * sy);
            float tanThetaMSqr = alpha * alpha * sx / (1.0f - sx);
            cosThetaM = 1.0f / std::sqrt(1.0f + tanThetaMSqr);
        } else {
            float phiM = (2.0f * kPi) * sy;
            sinPhiM = cr::sin(phiM); cosPhiM = cr::cos(phiM);
            cosThetaM = cr::pow(sx, 1.0f / (exponent + 2.0f));
        }
        float sinThetaM = std::sqrt(std::max(0.0f, 1 - cosThetaM * cosThetaM));
        return V3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
    }
    void sampleVisible11(float thetaI, float sx, float sy, float &slopeX, float &slopeY) const { // :555-673
        const float SQRT_PI_INV = 1 / std::sqrt(kPi);
        if (type == 0) {
            if (thetaI < 1e-4f) {
                float r = std::sqrt(-cr::log(1.0f - sx));
                float sinPhi = cr::sin(2 * kPi * sy), cosPhi = cr::cos(2 * kPi * sy);
                slopeX = r * cosPhi; slopeY = r * sinPhi; return;
            }
            float tanThetaI = cr::tan(thetaI), cotThetaI = 1 / tanThetaI;
            float a = -1, c = mf::erf(cotThetaI);
            float sample_x = std::max(sx, 1e-6f);
            float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * cr::pow(1 - sample_x, fit);
            float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * cr::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                float invErf = mf::erfinv(b);
                float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * cr::exp(-invErf * invErf)) - sample_x;
                float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::abs(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slopeX = mf::erfinv(b);
            slopeY = mf::erfinv(2.0f * std::max(sy, 1e-6f) - 1.0f);
        } else {
            if (thetaI < 1e-4f) {
                float r = safe_sqrt(sx / (1 - sx));
                float sinPhi = cr::sin(2 * kPi * sy), cosPhi = cr::cos(2 * kPi * sy);
                slopeX = r * cosPhi; slopeY = r * sinPhi; return;
            }
            float tanThetaI = cr::tan(thetaI);
            float a = 1 / tanThetaI;
            float G1 = 2.0f / (1.0f + safe_sqrt(1.0f + 1.0f / (a * a)));
            float A = 2.0f * sx / G1 - 1.0f;
            if (std::abs(A) == 1) A -= mf::signum(A) * kEpsilon;
            float tmp = 1.0f / (A * A - 1.0f);
            float B = tanThetaI;
            float D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
            slopeX = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
            float S;
            if (sy > 0.5f) { S = 1.0f; sy = 2.0f * (sy - 0.5f); }
            else { S = -1.0f; sy = 2.0f * (0.5f - sy); }
            float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                      (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
            slopeY = S * z * std::sqrt(1.0f + slopeX * slopeX);
        }
    }
    V3 sampleVisible(const V3 &_wi, float sx, float sy) const { // :389-439
        V3 wi = normalize(V3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) { theta = cr::acos(wi.z); phi = cr::atan2(wi.y, wi.x); }
        float sinPhi = cr::sin(phi), cosPhi = cr::cos(phi);
        float slx, sly;
        sampleVisible11(theta, sx, sy, slx, sly);
        float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= alpha; ry *= alpha;
        float normalization = 1.0f / std::sqrt(rx * rx + ry * ry + 1.0f);
        return V3(-rx * normalization, -ry * normalization, normalization);
    }
    V3 sample(const V3 &wi, float sx, float sy) const { return sampleVis ? sampleVisible(wi, sx, sy) : sampleAll(sx, sy); }
};

struct RoughPlastic {
    float eta = 1.5f, invEta2 = 0, alpha = 0.1f, specularSamplingWeight = 0;
    V3 diffuse = V3(0.5f), specular = V3(1.0f);
    bool nonlinear = false;
    MicrofacetDistribution distr;
    RoughTransmittance extRT, intRT;
    // ctor :186-235 + configure :258-304
    void configure(float intIOR, float extIOR, V3 diff, V3 spec, float alpha_, int type, bool sampleVisible, bool nonlin, const std::string &dataDir) {
        eta = intIOR / extIOR;
        nonlinear = nonlin;
        distr.configure(type, alpha_, sampleVisible);
        alpha = distr.alpha;
        float mx = maxc(spec); if (mx > 1.0f) spec = spec * (0.99f * (1.0f / mx));   // ensureEnergyConservation(.., 1.0f), bsdf.cpp:88-113
        mx = maxc(diff); if (mx > 1.0f) diff = diff * (0.99f * (1.0f / mx));
        specular = spec; diffuse = diff;
        float dAvg = luminance(diff), sAvg = luminance(spec);
        specularSamplingWeight = sAvg / (dAvg + sAvg);
        invEta2 = 1.0f / (eta * eta);
        static const char *names[3] = {"beckmann", "ggx", "phong"};
        extRT.load(dataDir + "/microfacet/" + names[type] + ".dat");
        extRT.checkRanges(eta, alpha);
        intRT = extRT;
        extRT.setEta(eta);
        intRT.setEta(1 / eta);
        extRT.setAlpha(alpha);
    }
    V3 eval(const V3 &wi, const V3 &wo) const { // :325-375
        if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
        V3 result(0.0f);
        const V3 H = normalize(wo + wi);
        const float D = distr.eval(H);
        const float F = fresnelDielectricExt(dot(wi, H), eta);
        const float G = distr.G(wi, wo, H);
        float value = F * D * G / (4.0f * wi.z);
        result += specular * value;
        V3 diff = diffuse;
        float T12 = extRT.eval(wi.z, alpha), T21 = extRT.eval(wo.z, alpha);
        float Fdr = 1 - intRT.evalDiffuse(alpha);
        if (nonlinear) diff = V3(diff.x / (1.0f - diff.x * Fdr), diff.y / (1.0f - diff.y * Fdr), diff.z / (1.0f - diff.z * Fdr));
        else diff = diff / (1 - Fdr);
        result += diff * (kInvPi * wo.z * T12 * T21 * invEta2);
        return result;
    }
    float probSpecularOf(float cosThetaI) const {
        float probSpecular = 1 - extRT.eval(cosThetaI, alpha);
        return (probSpecular * specularSamplingWeight) / (probSpecular * specularSamplingWeight + (1 - probSpecular) * (1 - specularSamplingWeight));
    }
    float pdf(const V3 &wi, const V3 &wo) const { // :377-436
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        const V3 H = normalize(wo + wi);
        float probSpecular = probSpecularOf(wi.z), probDiffuse = 1 - probSpecular;
        const float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
        const float prob = distr.pdf(wi, H);
        float result = prob * dwh_dwo * probSpecular;
        result += probDiffuse * (kInvPi * wo.z);
        return result;
    }
    BSDFSample sample(const V3 &wi, float sx, float sy) const { // :438-494
        BSDFSample r; r.weight = V3(0.0f); r.pdf = 0;
        if (wi.z <= 0) return r;
        bool choseSpecular = true;
        float probSpecular = probSpecularOf(wi.z);
        if (sy < probSpecular) sy /= probSpecular;
        else { sy = (sy - probSpecular) / (1 - probSpecular); choseSpecular = false; }
        if (choseSpecular) {
            V3 m = distr.sample(wi, sx, sy);
            r.wo = 2 * dot(wi, m) * m - wi;
            r.sampledComponent = 0; r.sampledType = EGlossyReflection;
            if (r.wo.z <= 0) return r;
        } else {
            r.sampledComponent = 1; r.sampledType = EDiffuseReflection;
            r.wo = squareToCosineHemisphere(sx, sy);
        }
        r.eta = 1.0f;
        r.pdf = pdf(wi, r.wo);
        if (r.pdf == 0) return r;
        r.weight = eval(wi, r.wo) / r.pdf;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// SmoothDiffuse (`diffuse` plugin, constant reflectance) -- src/bsdfs/diffuse.cpp:92-156, optionally wrapped in
// `twosided` (src/bsdfs/twosided.cpp:101-181) with the same nested BRDF on both sides.  Used for triangle meshes.
// ---------------------------------------------------------------------------------------------
// A 2-D texture as the BSDFs of this path see it: a constant Spectrum (ConstantSpectrumTexture, include/mitsuba/hw/basicshader.h:33-70) or
// the procedural `checkerboard` (src/textures/checkerboard.cpp:47-72) behind Texture2D's uv transform (src/librender/texture.cpp:81-121).
// A ScaleTexture wrapped around it by ensureEnergyConservation (bsdf.cpp:88-113) is folded into the colours (same product per lookup).
struct Texture2D {
    int kind = 0;                      // 0 constant, 1 checkerboard
    V3 color0 = V3(0.5f), color1 = V3(0.5f);     // constant: color0
    float uoffset = 0, voffset = 0, uscale = 1, vscale = 1;
    void setConstant(V3 v) { kind = 0; color0 = color1 = v; }
    void setCheckerboard(V3 c0, V3 c1, float uo, float vo, float us, float vs) { kind = 1; color0 = c0; color1 = c1; uoffset = uo; voffset = vo; uscale = us; vscale = vs; }
    bool isConstant() const { return kind == 0; }
    V3 getMaximum() const { return kind == 0 ? color0 : V3(std::max(color0.x, color1.x), std::max(color0.y, color1.y), std::max(color0.z, color1.z)); }
    V3 getAverage() const { return kind == 0 ? color0 : (color0 + color1) * 0.5f; }
    void scale(float f) { color0 = color0 * f; color1 = color1 * f; }
    // bsdf.cpp:88-113
    void ensureEnergyConservation(float mx = 1.0f) { const float actualMax = maxc(getMaximum()); if (actualMax > mx) scale(0.99f * (mx / actualMax)); }
    V3 eval(float u, float v) const {
        if (kind == 0) return color0;
        const float uu = u * uscale + uoffset, vv = v * vscale + voffset;        // texture.cpp:112-113
        const int x = 2 * modulo((int) (uu * 2), 2) - 1, y = 2 * modulo((int) (vv * 2), 2) - 1;    // checkerboard.cpp:65-72
        return x * y == 1 ? color0 : color1;
    }
};

struct SmoothDiffuse {
    Texture2D reflectance; bool twoSided = false;
    void configure(V3 r, bool two) { reflectance.setConstant(r); reflectance.ensureEnergyConservation(); twoSided = two; }
    void configure(const Texture2D &t, bool two) { reflectance = t; reflectance.ensureEnergyConservation(); twoSided = two; }
    V3 eval(V3 wi, V3 wo, float u = 0, float v = 0) const {
        if (twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }     // twosided.cpp:101-115 (wi.z > 0 ? front : flipped)
        if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
        return reflectance.eval(u, v) * (kInvPi * wo.z);
    }
    float pdf(V3 wi, V3 wo) const {
        if (twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }
        if (wi.z <= 0 || wo.z <= 0) return 0.0f;
        return kInvPi * wo.z;                                       // warp::squareToCosineHemispherePdf
    }
    BSDFSample sample(V3 wi, float sx, float sy, float u = 0, float v = 0) const {
        BSDFSample r; r.weight = V3(0.0f); r.pdf = 0;
        bool flipped = false;
        if (twoSided && wi.z < 0) { wi.z *= -1; flipped = true; }   // twosided.cpp:162-181
        if (wi.z <= 0) return r;
        r.wo = squareToCosineHemisphere(sx, sy);
        r.eta = 1.0f; r.sampledComponent = 0; r.sampledType = EDiffuseReflection;
        r.pdf = kInvPi * r.wo.z;
        r.weight = reflectance.eval(u, v);
        if (flipped && !isZero(r.weight) && r.pdf != 0) { r.wo.z *= -1; r.sampledComponent += 1; }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// SmoothPlastic (`plastic` plugin, models/teapot/scene.xml:31-38) -- src/bsdfs/plastic.cpp:140-167 (ctor), :186-217 (configure),
// :246-281 (eval), :283-313 (pdf), :381-445 (sample with pdf).  A delta reflection (component 0) over a diffuse base (component 1),
// front side only; typeMask = EAll, component = -1 as the path tracer asks.  specularReflectance is a constant here.
// ---------------------------------------------------------------------------------------------
struct SmoothPlastic {
    float eta, invEta2, fdrInt, fdrExt, specularSamplingWeight; bool nonlinear = false;
    V3 specR; Texture2D diffuse;
    void configure(float intIOR, float extIOR, const Texture2D &d, V3 s, bool nonlin) {
        eta = intIOR / extIOR; nonlinear = nonlin;
        Texture2D st; st.setConstant(s); st.ensureEnergyConservation(); specR = st.color0;
        diffuse = d; diffuse.ensureEnergyConservation();
        fdrInt = fresnelDiffuseReflectance(1 / eta); fdrExt = fresnelDiffuseReflectance(eta);
        const float dAvg = luminance(diffuse.getAverage()), sAvg = luminance(specR);
        specularSamplingWeight = sAvg / (dAvg + sAvg);
        invEta2 = 1 / (eta * eta);
    }
    static V3 reflect(const V3 &wi) { return V3(-wi.x, -wi.y, wi.z); }
    V3 diffTerm(float u, float v) const {
        V3 diff = diffuse.eval(u, v);
        if (nonlinear) diff = V3(diff.x / (1.0f - diff.x * fdrInt), diff.y / (1.0f - diff.y * fdrInt), diff.z / (1.0f - diff.z * fdrInt));
        else diff = diff / (1 - fdrInt);
        return diff;
    }
    float probSpecular(float Fi) const { return (Fi * specularSamplingWeight) / (Fi * specularSamplingWeight + (1 - Fi) * (1 - specularSamplingWeight)); }
    V3 eval(const V3 &wi, const V3 &wo, bool discrete, float u = 0, float v = 0) const {
        if (wo.z <= 0 || wi.z <= 0) return V3(0.0f);
        const float Fi = fresnelDielectricExt(wi.z, eta);
        if (discrete) {
            if (std::abs(dot(reflect(wi), wo) - 1) < kDeltaEpsilon) return specR * Fi;
        } else {
            const float Fo = fresnelDielectricExt(wo.z, eta);
            return diffTerm(u, v) * ((kInvPi * wo.z) * invEta2 * (1 - Fi) * (1 - Fo));
        }
        return V3(0.0f);
    }
    float pdf(const V3 &wi, const V3 &wo, bool discrete) const {
        if (wo.z <= 0 || wi.z <= 0) return 0.0f;
        const float ps = probSpecular(fresnelDielectricExt(wi.z, eta));
        if (discrete) {
            if (std::abs(dot(reflect(wi), wo) - 1) < kDeltaEpsilon) return ps;
        } else return (kInvPi * wo.z) * (1 - ps);
        return 0.0f;
    }
    BSDFSample sample(const V3 &wi, float sx, float sy, float u = 0, float v = 0) const {
        BSDFSample r; r.weight = V3(0.0f); r.pdf = 0;
        if (wi.z <= 0) return r;
        const float Fi = fresnelDielectricExt(wi.z, eta);
        r.eta = 1.0f;
        const float ps = probSpecular(Fi);
        if (sx < ps) {
            r.sampledComponent = 0; r.sampledType = EDeltaReflection; r.wo = reflect(wi); r.pdf = ps;
            r.weight = specR * Fi / ps;
        } else {
            r.sampledComponent = 1; r.sampledType = EDiffuseReflection;
            r.wo = squareToCosineHemisphere((sx - ps) / (1 - ps), sy);
            const float Fo = fresnelDielectricExt(r.wo.z, eta);
            r.pdf = (1 - ps) * (kInvPi * r.wo.z);
            r.weight = diffTerm(u, v) * (invEta2 * (1 - Fi) * (1 - Fo) / (1 - ps));
        }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// Mirror (`mirror`, the fork's own plugin: models/teapot/mirror_scene.xml:32-36) -- src/bsdfs/mirror.cpp:187-199 (ctor), :210-221 (configure),
// :233-247 (eval), :249-254 (pdf), :256-276 (sample).  One EDeltaReflection component.  As committed: eval() is identically zero (it asks for
// ESolidAngle AND for a flag that needs EDiscrete), pdf() is 1 in the discrete measure whatever the directions, sample() mirrors wi on the front side.
// ---------------------------------------------------------------------------------------------
struct Mirror {
    V3 specR;
    void configure(V3 r) { Texture2D t; t.setConstant(r); t.ensureEnergyConservation(); specR = t.color0; }
    V3 eval(const V3 &, const V3 &, bool) const { return V3(0.0f); }
    float pdf(const V3 &, const V3 &, bool discrete) const { return discrete ? 1.0f : 0.0f; }
    BSDFSample sample(const V3 &wi, float, float) const {
        BSDFSample r; r.weight = V3(0.0f); r.pdf = 0;
        if (wi.z <= 0) return r;
        r.sampledComponent = 0; r.sampledType = EDeltaReflection; r.wo = V3(-wi.x, -wi.y, wi.z); r.eta = 1.0f; r.pdf = 1.0f; r.weight = specR;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// ThinDielectric (`thindielectric`, models/straight-hair/scene_thindielectric.xml) -- src/bsdfs/thindielectric.cpp:73-300.
// Two discrete components (EDeltaReflection, ENull), both sides; nothing smooth, so the path tracer skips emitter sampling
// at such a vertex (path.cpp:174-175).  eval/pdf are non-zero only in the discrete measure.
// ---------------------------------------------------------------------------------------------
static inline V3 ensureEnergyConservationConst(V3 v, float mx = 1.0f) {        // bsdf.cpp:88-113 for a constant texture
    const float actualMax = maxc(v);
    if (actualMax > mx) v = v * (0.99f * (mx / actualMax));
    return v;
}
// R' = R + TRT + TR^3T + ...  (thindielectric.cpp:150-154 and five more copies; marschnerdielectric.cpp:268-272 etc.)
static inline float thinSlabReflectance(float cosThetaI, float eta) {
    float R = fresnelDielectricExt(cosThetaI, eta), T = 1 - R;
    if (R < 1) R += T * T * R / (1 - R * R);
    return R;
}

struct ThinDielectric {
    float eta; V3 specR, specT;
    void configure(float intIOR, float extIOR, V3 r, V3 t) {
        eta = intIOR / extIOR;
        specR = ensureEnergyConservationConst(r); specT = ensureEnergyConservationConst(t);
    }
    static V3 reflect(const V3 &wi) { return V3(-wi.x, -wi.y, wi.z); }
    // thindielectric.cpp:143-168
    V3 eval(const V3 &wi, const V3 &wo, bool discrete) const {
        const float R = thinSlabReflectance(std::abs(wi.z), eta);
        if (wi.z * wo.z >= 0) {
            if (!discrete || std::abs(dot(reflect(wi), wo) - 1) > kDeltaEpsilon) return V3(0.0f);
            return specR * R;
        } else {
            if (!discrete || std::abs(dot(-wi, wo) - 1) > kDeltaEpsilon) return V3(0.0f);
            return specT * (1 - R);
        }
    }
    // thindielectric.cpp:170-194
    float pdf(const V3 &wi, const V3 &wo, bool discrete) const {
        const float R = thinSlabReflectance(std::abs(wi.z), eta);
        if (wi.z * wo.z >= 0) {
            if (!discrete || std::abs(dot(reflect(wi), wo) - 1) > kDeltaEpsilon) return 0.0f;
            return R;
        } else {
            if (!discrete || std::abs(dot(-wi, wo) - 1) > kDeltaEpsilon) return 0.0f;
            return 1 - R;
        }
    }
    // thindielectric.cpp:196-245 (typeMask = EAll: both components)
    BSDFSample sample(const V3 &wi, float sx, float /*sy*/) const {
        BSDFSample r; r.eta = 1.0f;
        const float R = thinSlabReflectance(std::abs(wi.z), eta);
        if (sx <= R) { r.sampledComponent = 0; r.sampledType = EDeltaReflection; r.wo = reflect(wi); r.pdf = R; r.weight = specR; }
        else { r.sampledComponent = 1; r.sampledType = ENull; r.wo = -wi; r.pdf = 1 - R; r.weight = specT; }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// MarschnerDielectric (`marschnerdielectric`, models/straight-hair/scene_dielectric*.xml) -- the fork's third hair BSDF,
// src/bsdfs/marschnerdielectric.cpp:128-167 (ctor), :189-221 (configure), :245-307 (eval), :309-376 (pdf), :419-499 (sample).
// A thin dielectric with a Kajiya-Kay cone and a diffuse term bolted on.  As committed:
//   * eval() is identically zero: it returns 0 unless measure == ESolidAngle, and in that measure `sampleReflection` /
//     `sampleTransmission` (which require EDiscrete) are false, so both branches return before the cone / diffuse terms (:258-260,:274-300);
//   * pdf() in the solid-angle measure is the cosine-hemisphere density of the diffuse component (:322-347), 0 otherwise;
//   * sample(): the specular branch (probability m_specularSamplingWeight = (s+t)/(d+s+t), :211-214) behaves like the thin dielectric;
//     the diffuse branch draws a cosine direction and returns eval()/pdf() = 0, which ends the path (:486-498).
// It has a diffuse component, so the path tracer does sample the emitter at every vertex (and traces the shadow ray) although the
// product with eval() is always zero.
// ---------------------------------------------------------------------------------------------
struct MarschnerDielectric {
    float eta, exponent, specularSamplingWeight; V3 diffuse, specR, specT;
    void configure(float intIOR, float extIOR, V3 d, V3 r, V3 t, float expo) {
        eta = intIOR / extIOR; exponent = expo; diffuse = d;
        specR = ensureEnergyConservationConst(r); specT = ensureEnergyConservationConst(t);
        const float dAvg = luminance(diffuse), sAvg = luminance(specR), tAvg = luminance(specT);
        specularSamplingWeight = (sAvg + tAvg) / (dAvg + sAvg + tAvg);
    }
    V3 eval(const V3 &, const V3 &, bool) const { return V3(0.0f); }
    float pdf(const V3 &wi, const V3 &wo, bool discrete) const {
        if (discrete || wi.z <= 0 || wo.z <= 0) return 0.0f;
        return kInvPi * wo.z;
    }
    BSDFSample sample(const V3 &wi, float sx, float sy) const {
        BSDFSample r; r.weight = V3(0.0f); r.eta = 1.0f;
        bool choseSpecular = true;
        if (sx <= specularSamplingWeight) sx /= specularSamplingWeight;
        else { sx = (sx - specularSamplingWeight) / (1 - specularSamplingWeight); choseSpecular = false; }
        if (choseSpecular) {
            const float R = thinSlabReflectance(std::abs(wi.z), eta);
            if (sx <= R) { r.sampledComponent = 0; r.sampledType = EDeltaReflection; r.wo = ThinDielectric::reflect(wi); r.pdf = R; r.weight = specR; }
            else { r.sampledComponent = 1; r.sampledType = ENull; r.wo = -wi; r.pdf = 1 - R; r.weight = specT; }
            return r;
        }
        r.wo = squareToCosineHemisphere(sx, sy);
        r.sampledComponent = 2; r.sampledType = EDiffuseReflection;
        r.pdf = pdf(wi, r.wo, false);
        return r;                                   // eval()/pdf = 0 (or pdf == 0): zero weight
    }
};

} // namespace orc

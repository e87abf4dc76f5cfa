// cp_bsdf.cuh -- hair BSDF device code (sm_100a): KajiyaKay and the as-built `marschner` plugin.
//
// Replaces (reference file:line, quirks reproduced on purpose -- SURVEY.md Appendix A):
//   KajiyaKay::eval/pdf/sample          src/bsdfs/kajiyakay.cpp:122-180, 182-214, 216-273
//   MarschnerDiffuse::eval/pdf/sample   src/bsdfs/marschner_diffuse.cpp:377-482, 488-520, 594-744
//   Azimuthal::eval/sample/weight       src/bsdfs/marschner_diffuse.cpp:68-105
//   InterpolatedDistribution1D::warp    src/bsdfs/InterpolatedDistribution1D.hpp:69-92
//   RoughTransmittance::eval (1-D)      src/bsdfs/rtrans.h:183-194 + src/libcore/spline.cpp:23-60
//
// Numerics: everything is fp32 like the reference.  The longitudinal term M() exponentiates a sum of
// O(1/v) = O(400) terms, so a 1-ulp difference in any input (a libm call, a fused multiply-add) becomes
// ~3e-5 relative in the result, and sampleM() feeds such values back into M().  To stay inside the 1e-4
// parity tolerance every elementary function is correctly rounded (cr_* in cp_common.cuh) and the translation
// units that contain BSDF code are compiled with -fmad=false (the reference's x86 build has no FMA either).
#pragma once
#include "cp_common.cuh"

namespace cp {

struct BsdfDev {
    int kind;            // 0 = kajiyakay, 1 = marschner (as built), 2 = diffuse (constant reflectance in `diffuse`; meshes),
                         // 3 = the unbuilt `Marschner` of src/bsdfs/marschner.cpp ("fixed" mode: TRT-only eval, real pdf)
    int twoSided;        // kind 2 only: wrapped in `twosided` with the same nested BRDF on both sides
    // kajiyakay (kajiyakay.cpp:60-107) / marschner diffuse colour
    V3 diffuse, specular;
    float exponent;
    float specW;         // m_specularSamplingWeight
    // marschner (marschner_diffuse.cpp:113-160,193-247)
    float eta, invEta2, alpha, Fdr, vR, vTT, vTRT, scaleAngle;
    int nonlinear, rtSize;
    const float4 *tab;   // 3 lobes x 64x64 x (r,g,b,-)
    const float *cdf;    // 3 x 64 rows x 65
    const float *sums;   // 3 x 64
    const float *pdfs;   // 3 x 64 rows x 64 (normalised row pdfs; only the fixed mode reads them)
    const float *rt;     // external rough transmittance, 1-D slice (rtSize samples over |cos|^(1/4))
};

struct BsdfSampleOut { V3 wo, weight; float pdf; int type; int component; };


// ------------------------------------------------------------------------------------------ KajiyaKay
CP_D V3 kk_reflect(const V3 &wi) { return V3(-wi.x, -wi.y, wi.z); }

CP_D V3 kk_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
    V3 result(0.0f);
    float tl = fabsf(wi.x), te = fabsf(wo.x);
    float sin_tl = sqrtf(1 - tl * tl), sin_te = sqrtf(1 - te * te);
    float alpha = tl * te + sin_tl * sin_te;
    if (alpha > 0.0f && wi.x * wo.x < 0)  // no back-scatter lobe (kajiyakay.cpp:157)
        result += 0.15f * b.specular * ((b.exponent + 2) * kInvFourPi * cr_pow(alpha, b.exponent));
    result += b.diffuse * kInvPi;
    return result * wo.z;
}
CP_D float kk_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    float diffuseProb = kInvPi * wo.z;
    float specProb = 0.0f;
    float alpha = dot(wo, kk_reflect(wi));
    if (alpha > 0) specProb = cr_pow(alpha, b.exponent) * (b.exponent + 1.0f) / (2.0f * kPi);
    return b.specW * specProb + (1 - b.specW) * diffuseProb;
}
CP_D BsdfSampleOut kk_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    BsdfSampleOut r; r.weight = V3(0.0f); r.pdf = 0.0f; r.wo = V3(0.0f);
    bool choseSpecular = true;
    if (sx <= b.specW) sx /= b.specW;
    else { sx = (sx - b.specW) / (1 - b.specW); choseSpecular = false; }
    if (choseSpecular) {
        V3 R = kk_reflect(wi);
        float sinAlpha = sqrtf(1 - cr_pow(sy, 2 / (b.exponent + 1)));
        float cosAlpha = cr_pow(sy, 1 / (b.exponent + 1));
        float phi = (2.0f * kPi) * sx, sp, cp_;
        cr_sincos(phi, &sp, &cp_);
        V3 localDir(sinAlpha * cp_, sinAlpha * sp, cosAlpha);
        r.wo = Frame(R).toWorld(localDir);
        r.component = 1; r.type = EGlossyReflection;   // component labels swapped in the reference (:255-263)
        if (r.wo.z <= 0) return r;
    } else {
        r.wo = squareToCosineHemisphere(sx, sy);
        r.component = 0; r.type = EDiffuseReflection;
    }
    r.pdf = kk_pdf(b, wi, r.wo);
    if (r.pdf == 0) return r;
    r.weight = kk_eval(b, wi, r.wo) / r.pdf;
    return r;
}

// ------------------------------------------------------------------------------------------ Marschner
// :484-486.  1 - x*x is evaluated un-fused (the reference's x86 build has no FMA) because it feeds the ill-conditioned M()
CP_D float ma_trigInverse(float x) { return fminf(sqrtf(fmaxf(__fsub_rn(1.0f, __fmul_rn(x, x)), 0.0f)), 1.0f); }

CP_D float ma_I0(float x) { // :279-290
    float result = 1.0f, xSq = x * x, xi = xSq, denom = 4.0f;
#pragma unroll
    for (int i = 1; i <= 10; ++i) { result += xi / denom; xi *= xSq; denom *= 4.0f * float((i + 1) * (i + 1)); }
    return result;
}
CP_D float ma_logI0(float x) { // :292-299
    if (x > 12.0f) return x + 0.5f * (cr_log(1.0f / (kPi * 2.0f * x)) + 1.0f / (8.0f * x));
    return cr_log(ma_I0(x));
}
CP_D float ma_M(float v, float sinThetaI, float sinThetaO, float cosThetaI, float cosThetaO) { // :364-374
    float a = cosThetaI * cosThetaO / v, b = sinThetaI * sinThetaO / v;
    if (v < 0.1f) return cr_exp(-b + ma_logI0(a) - 1.0f / v + 0.6931f + cr_log(1.0f / (2.0f * v)));
    return cr_exp(-b) * ma_I0(a) / (2.0f * v * cr_sinh(1.0f / v));
}

// Azimuthal::eval :79-92 -- bilinear lookup in one 64x64 RGB table
CP_D V3 ma_azimuthal(const float4 *__restrict__ tab, float phi, float cosThetaD) {
    float u = 63 * phi * (1.0f / (2.0f * kPi));
    float v = 63 * cosThetaD;
    int x0 = clampi(int(u), 0, 62), y0 = clampi(int(v), 0, 62);
    u = clampf(u - x0, 0.0f, 1.0f);
    v = clampf(v - y0, 0.0f, 1.0f);
    float4 t00 = __ldg(tab + x0 + y0 * 64), t10 = __ldg(tab + x0 + 1 + y0 * 64);
    float4 t01 = __ldg(tab + x0 + (y0 + 1) * 64), t11 = __ldg(tab + x0 + 1 + (y0 + 1) * 64);
    V3 a = (V3(t00.x, t00.y, t00.z) * (1.0f - u) + V3(t10.x, t10.y, t10.z) * u) * (1.0f - v);
    V3 c = (V3(t01.x, t01.y, t01.z) * (1.0f - u) + V3(t11.x, t11.y, t11.z) * u) * v;
    return a + c;
}
// Azimuthal::weight :101-105 / InterpolatedDistribution1D::sum
CP_D float ma_weight(const float *__restrict__ sums, float cosThetaD) {
    float v = 63 * cosThetaD;
    int d0 = clampi(int(v), 0, 63), d1 = min(d0 + 1, 63);
    float f = clampf(v - d0, 0.0f, 1.0f);
    return (__ldg(sums + d0) * (1.0f - f) + __ldg(sums + d1) * f) * (2.0f * kPi / 64);
}
// Azimuthal::sample :68-77 + InterpolatedDistribution1D::warp
CP_D float ma_sample_phi(const float *__restrict__ cdf, float cosThetaD, float xi) {
    float dist = 63 * cosThetaD;
    int d0 = clampi(int(dist), 0, 63), d1 = min(d0 + 1, 63);
    float v = clampf(dist - d0, 0.0f, 1.0f);
    const float *c0 = cdf + d0 * 65, *c1 = cdf + d1 * 65;
    int lower = 0, upper = 64;
    float lowerU = 0.0f, upperU = 1.0f;
    while (upper - lower != 1) {
        int midpoint = (upper + lower) / 2;
        float midpointU = __ldg(c0 + midpoint) * (1.0f - v) + __ldg(c1 + midpoint) * v;
        if (midpointU < xi) { lower = midpoint; lowerU = midpointU; }
        else { upper = midpoint; upperU = midpointU; }
    }
    xi = clampf((xi - lowerU) / (upperU - lowerU), 0.0f, 1.0f);
    return 2.0f * kPi * (lower + xi) * (1.0f / 64);
}
// RoughTransmittance::eval, alpha and eta fixed (rtrans.h:183-194,233) -> Catmull-Rom over rtSize samples
CP_D float ma_T(const BsdfDev &b, float cosTheta) {
    float warped = cr_pow(fabsf(cosTheta), 0.25f);
    if (!(cosTheta >= 0)) return 0.0f;
    float x = warped;
    if (!(x >= 0.0f && x <= 1.0f)) return 0.0f;       // spline.cpp:25-26 (min(1,max(0,0)) = 0)
    const int size = b.rtSize;
    float t = ((x - 0.0f) * (size - 1)) / (1.0f - 0.0f);
    int k = max(0, min((int) t, size - 2));
    float f0 = __ldg(b.rt + k), f1 = __ldg(b.rt + k + 1), d0, d1;
    if (k > 0) d0 = 0.5f * (f1 - __ldg(b.rt + k - 1)); else d0 = f1 - f0;
    if (k + 2 < size) d1 = 0.5f * (__ldg(b.rt + k + 2) - f0); else d1 = f1 - f0;
    t = t - (float) k;
    float t2 = t * t, t3 = t2 * t;
    float result = (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
    return fminf(1.0f, fmaxf(0.0f, result));
}

CP_D V3 ma_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    float sinThetaI = wi.y, sinThetaO = wo.y;            // the `t` axis, not the tangent (quirk 2)
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = cr_atan2(wo.x, wo.z);                     // depends on wo only (quirk 2)
    if (phi < 0.0f) phi += kPi * 2.0f;
    float thetaIR = thetaI - 2.0f * b.scaleAngle;
    float thetaITT = thetaI + b.scaleAngle;
    float thetaITRT = thetaI + 4.0f * b.scaleAngle;
    float MR = ma_M(b.vR, cr_sin(thetaIR), sinThetaO, cr_cos(thetaIR), cosThetaO);
    float MTT = ma_M(b.vTT, cr_sin(thetaITT), sinThetaO, cr_cos(thetaITT), cosThetaO);
    float MTRT = ma_M(b.vTRT, cr_sin(thetaITRT), sinThetaO, cr_cos(thetaITRT), cosThetaO);
    V3 result = 0.15f * MR * ma_azimuthal(b.tab, phi, cosThetaD)
              + MTT * ma_azimuthal(b.tab + 4096, phi, cosThetaD)
              + MTRT * ma_azimuthal(b.tab + 8192, phi, cosThetaD);
    V3 diff = b.diffuse;
    float T12 = ma_T(b, wi.z), T21 = ma_T(b, wo.z);
    if (b.nonlinear) diff = V3(diff.x / (1.0f - diff.x * b.Fdr), diff.y / (1.0f - diff.y * b.Fdr), diff.z / (1.0f - diff.z * b.Fdr));
    else diff = diff / (1 - b.Fdr);
    result += diff * (kInvPi * wo.z * T12 * T21 * b.invEta2);
    return result;
}
CP_D float ma_sampleM(float v, float sinThetaI, float cosThetaI, float xi1, float xi2) { // :582-592
    float cosTheta = 1.0f + v * cr_log(xi1 + (1.0f - xi1) * cr_exp(-2.0f / v));
    float sinTheta = ma_trigInverse(cosTheta);
    float cosPhi = cr_cos(2 * kPi * xi2);
    return -cosTheta * sinThetaI + sinTheta * cosPhi * cosThetaI;
}
CP_D BsdfSampleOut ma_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    BsdfSampleOut r;
    float sinThetaI = wi.y;
    float cosThetaI = ma_trigInverse(sinThetaI);
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float v, theta; int lobe;
    float target = sx * (weightR + weightTT + weightTRT);
    if (target < weightR) { r.component = 5; v = b.vR; theta = thetaI - 2.0f * b.scaleAngle; lobe = 0; }
    else if (target < weightR + weightTT) { r.component = 6; v = b.vTT; theta = thetaI + b.scaleAngle; lobe = 1; }
    else { r.component = 7; v = b.vTRT; theta = thetaI + 4.0f * b.scaleAngle; lobe = 2; }
    float sinThetaO = ma_sampleM(v, cr_sin(theta), cr_cos(theta), sx, sy);   // one 2-D sample reused (quirk 4)
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = ma_sample_phi(b.cdf + lobe * 64 * 65, cosThetaD, sy);
    float sinPhi, cosPhi;
    cr_sincos(phi, &sinPhi, &cosPhi);
    float probSpecular = 1 - ma_T(b, wi.z);
    probSpecular = (probSpecular * b.specW) / (probSpecular * b.specW + (1 - probSpecular) * (1 - b.specW));
    if (sy < probSpecular) {
        r.wo = V3(sinPhi * cosThetaO, sinThetaO, cosPhi * cosThetaO);
        r.type = EDeltaReflection;                        // quirk 5
    } else {
        r.component = 1; r.type = EDiffuseReflection;
        r.wo = squareToCosineHemisphere(sx, sy);
    }
    r.pdf = 1.0f;                                          // pdf() == 1 (quirk 1)
    r.weight = ma_eval(b, wi, r.wo) / r.pdf;
    return r;
}

// ------------------------------------------------------------------------------------------ Marschner, "fixed" mode
// src/bsdfs/marschner.cpp (left out of the fork's build): eval :309-341 keeps the TRT lobe only, pdf :347-407, sample :421-535
// Azimuthal::pdf :91-96 + InterpolatedDistribution1D::pdf (hpp:94-101)
CP_D float mf_azimuthal_pdf(const float *__restrict__ pdfs, float phi, float cosThetaD) {
    float u = 63 * phi * (1.0f / (2.0f * kPi));
    float dist = 63 * cosThetaD;
    int d0 = clampi(int(dist), 0, 63), d1 = min(d0 + 1, 63);
    float v = clampf(dist - d0, 0.0f, 1.0f);
    const int x = int(u);
    return (__ldg(pdfs + x + d0 * 64) * (1.0f - v) + __ldg(pdfs + x + d1 * 64) * v) * float(64 * (1.0f / (2.0f * kPi)));
}
CP_D V3 mf_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    float sinThetaI = wi.y, sinThetaO = wo.y;
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = cr_atan2(wo.x, wo.z);
    if (phi < 0.0f) phi += kPi * 2.0f;
    float thetaITRT = thetaI + 4.0f * b.scaleAngle;
    float MTRT = ma_M(b.vTRT, cr_sin(thetaITRT), sinThetaO, cr_cos(thetaITRT), cosThetaO);
    const V3 zero(0.0f);      // MR = MTT = 0 in the reference (:333-334): their lobes contribute 0 * table value
    return zero * ma_azimuthal(b.tab, phi, cosThetaD) + zero * ma_azimuthal(b.tab + 4096, phi, cosThetaD) + MTRT * ma_azimuthal(b.tab + 8192, phi, cosThetaD);
}
CP_D float mf_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    float sinThetaI = wi.y, sinThetaO = wo.y;
    float cosThetaI = ma_trigInverse(sinThetaI), cosThetaO = ma_trigInverse(sinThetaO);
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = cr_atan2(wo.x, wo.z);
    if (phi < 0.0f) phi += 2.0f * kPi;
    float thetaIR = thetaI - 2.0f * b.scaleAngle, thetaITT = thetaI + b.scaleAngle, thetaITRT = thetaI + 4.0f * b.scaleAngle;
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float weightSum = weightR + weightTT + weightTRT;
    float pdfR = weightR * ma_M(b.vR, cr_sin(thetaIR), sinThetaO, cr_cos(thetaIR), cosThetaO);
    float pdfTT = weightTT * ma_M(b.vTT, cr_sin(thetaITT), sinThetaO, cr_cos(thetaITT), cosThetaO);
    float pdfTRT = weightTRT * ma_M(b.vTRT, cr_sin(thetaITRT), sinThetaO, cr_cos(thetaITRT), cosThetaO);
    return (1.0f / weightSum) * (pdfR * mf_azimuthal_pdf(b.pdfs, phi, cosThetaD) + pdfTT * mf_azimuthal_pdf(b.pdfs + 4096, phi, cosThetaD)
                                 + pdfTRT * mf_azimuthal_pdf(b.pdfs + 8192, phi, cosThetaD));
}
// xiN / xiM: the two extra sampler->next2D() draws of :473-474 (the `sample` argument of the BSDF interface is ignored)
CP_D BsdfSampleOut mf_sample(const BsdfDev &b, const V3 &wi, float xiNx, float xiNy, float xiMx, float xiMy) {
    BsdfSampleOut r; r.weight = V3(0.0f);
    float sinThetaI = wi.y;
    float cosThetaI = ma_trigInverse(sinThetaI);
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float v, theta; int lobe;
    float target = xiNx * (weightR + weightTT + weightTRT);
    if (target < weightR) { r.component = 0; v = b.vR; theta = thetaI - 2.0f * b.scaleAngle; lobe = 0; }
    else if (target < weightR + weightTT) { r.component = 1; v = b.vTT; theta = thetaI + b.scaleAngle; lobe = 1; }
    else { r.component = 2; v = b.vTRT; theta = thetaI + 4.0f * b.scaleAngle; lobe = 2; }
    float sinThetaO = ma_sampleM(v, cr_sin(theta), cr_cos(theta), xiMx, xiMy);
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = ma_sample_phi(b.cdf + lobe * 64 * 65, cosThetaD, xiNy);
    float sinPhi, cosPhi;
    cr_sincos(phi, &sinPhi, &cosPhi);
    r.wo = V3(sinPhi * cosThetaO, sinThetaO, cosPhi * cosThetaO);
    r.pdf = mf_pdf(b, wi, r.wo);
    r.type = EDeltaReflection;
    if (r.pdf <= 0 || r.pdf > 1) return r;
    r.weight = mf_eval(b, wi, r.wo) / r.pdf;
    return r;
}

// ------------------------------------------------------------------------------------------ SmoothDiffuse (+ TwoSided)
// src/bsdfs/diffuse.cpp:109-156 with a constant reflectance; src/bsdfs/twosided.cpp:101-181 when b.twoSided
CP_D V3 df_eval(const BsdfDev &b, V3 wi, V3 wo) {
    if (b.twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }
    if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
    return b.diffuse * (kInvPi * wo.z);
}
CP_D float df_pdf(const BsdfDev &b, V3 wi, V3 wo) {
    if (b.twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    return kInvPi * wo.z;
}
CP_D BsdfSampleOut df_sample(const BsdfDev &b, V3 wi, float sx, float sy) {
    BsdfSampleOut r; r.wo = V3(0.0f); r.weight = V3(0.0f); r.pdf = 0.0f; r.type = 0; r.component = -1;
    bool flipped = false;
    if (b.twoSided && wi.z < 0) { wi.z *= -1; flipped = true; }
    if (wi.z <= 0) return r;
    r.wo = squareToCosineHemisphere(sx, sy);
    r.component = 0; r.type = EDiffuseReflection;
    r.pdf = kInvPi * r.wo.z;
    r.weight = b.diffuse;
    if (flipped && !isZero(r.weight) && r.pdf != 0) { r.wo.z *= -1; r.component += 1; }
    return r;
}

// ------------------------------------------------------------------------------------------ dispatch
CP_D V3 bsdf_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    return b.kind == 0 ? kk_eval(b, wi, wo) : b.kind == 1 ? ma_eval(b, wi, wo) : b.kind == 2 ? df_eval(b, wi, wo) : mf_eval(b, wi, wo);
}
CP_D float bsdf_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    return b.kind == 0 ? kk_pdf(b, wi, wo) : b.kind == 1 ? 1.0f : b.kind == 2 ? df_pdf(b, wi, wo) : mf_pdf(b, wi, wo);
}
// true for BSDFs whose sample() pulls more numbers from the sampler than the two it is handed (fixed Marschner: 4)
CP_D bool bsdf_draws_extra(const BsdfDev &b) { return b.kind == 3; }
CP_D BsdfSampleOut bsdf_sample(const BsdfDev &b, const V3 &wi, float sx, float sy, const float4 &extra) {
    return b.kind == 0 ? kk_sample(b, wi, sx, sy) : b.kind == 1 ? ma_sample(b, wi, sx, sy) : b.kind == 2 ? df_sample(b, wi, sx, sy)
                                                                                         : mf_sample(b, wi, extra.x, extra.y, extra.z, extra.w);
}

} // namespace cp

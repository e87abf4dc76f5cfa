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
//
// Two builds of this header exist in the library (csrc/Makefile): the strict one (every call correctly rounded: bit-identical to
// the oracle) and the default, -DCP_FAST_MATH one, in which only the sensitive quantities keep their exact bits -- thetaI = asin(wi.y)
// and the sines / cosines of the three shifted lobe angles that enter M() (ma_lobe_angles below: one fp64 sincos and a few fp64
// multiply-adds instead of six fp64 libm calls), and the coordinates of the azimuthal tables (thetaO, cos thetaD, phi: the tables
// hold caustic peaks, so one ulp of a coordinate can be 1e-3 of the interpolated value) -- while every nc_* call (the exp / log
// inside M(), whose argument is O(1), the rough-transmittance warp, everything that only shapes a sampled direction) uses the
// 1-2 ulp fp32 functions.  Measured difference between the modes: ~1e-6 relative.
#pragma once
#include "cp_common.cuh"

namespace cp {

struct BsdfDev {
    int kind;            // 0 = kajiyakay, 1 = marschner (as built), 2 = diffuse (constant reflectance in `diffuse`; meshes),
                         // 3 = the unbuilt `Marschner` of src/bsdfs/marschner.cpp ("fixed" mode: TRT-only eval, real pdf),
                         // 4 = roughplastic (src/bsdfs/roughplastic.cpp; the default BSDF of the models/*/scene.xml files)
                         // 5 = thindielectric (src/bsdfs/thindielectric.cpp; specular reflectance in `specular`, transmittance in `diffuse`)
                         // 6 = marschnerdielectric (src/bsdfs/marschnerdielectric.cpp; transmittance in `specT`)
                         // 7 = plastic (src/bsdfs/plastic.cpp; fdrInt in `Fdr`)
                         // 8 = mirror (the fork's src/bsdfs/mirror.cpp: a delta reflection whose eval() is identically zero)
    int twoSided;        // kinds 2, 4, 7, 8: wrapped in `twosided` (src/bsdfs/twosided.cpp) with the same nested BRDF on both sides
    // diffuse reflectance texture of kinds 2 and 7: 0 = the constant in `diffuse`, 1 = checkerboard (src/textures/checkerboard.cpp) of
    // `diffuse` (color0) and `color1` behind Texture2D's uv transform (src/librender/texture.cpp:81-121)
    int texKind;
    V3 color1;
    float uvOffset[2], uvScale[2];
    // kajiyakay (kajiyakay.cpp:60-107) / marschner diffuse colour
    V3 diffuse, specular;
    V3 specT;            // marschnerdielectric: m_specularTransmittance
    float exponent;
    float specW;         // m_specularSamplingWeight
    // marschner (marschner_diffuse.cpp:113-160,193-247)
    float eta, invEta2, alpha, Fdr, vR, vTT, vTRT, scaleAngle;
    int nonlinear, rtSize;
    int lobeMask;        // kind 3: lobes eval() keeps (bit 0 R, 1 TT, 2 TRT): 4 = src/bsdfs/marschner.cpp as committed (:333-334), 7 = all three
    int distr, sampleVisible;   // roughplastic: microfacet distribution 0 beckmann / 1 ggx / 2 phong (exponent in `exponent`), visible-normal sampling
    const float4 *tab;   // 3 lobes x 64x64 x (r,g,b,-)
    const float *cdf;    // 3 x 64 rows x 65
    const float *sums;   // 3 x 64
    const float *pdfs;   // 3 x 64 rows x 64 (normalised row pdfs; only the fixed mode reads them)
    const float *rt;     // external rough transmittance, 1-D slice (rtSize samples over |cos|^(1/4))
    double lobeSin[3], lobeCos[3];   // sin / cos of the lobe shifts -2 s, +s, +4 s of the scale angle s (fp32 values, fp64 functions): ma_lobe_angles
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
        result += 0.15f * b.specular * ((b.exponent + 2) * kInvFourPi * nc_pow(alpha, b.exponent));
    result += b.diffuse * kInvPi;
    return result * wo.z;
}
CP_D float kk_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    float diffuseProb = kInvPi * wo.z;
    float specProb = 0.0f;
    float alpha = dot(wo, kk_reflect(wi));
    if (alpha > 0) specProb = nc_pow(alpha, b.exponent) * (b.exponent + 1.0f) / (2.0f * kPi);
    return b.specW * specProb + (1 - b.specW) * diffuseProb;
}
CP_D BsdfSampleOut kk_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    BsdfSampleOut r; r.weight = V3(0.0f); r.pdf = 0.0f; r.wo = V3(0.0f);
    bool choseSpecular = true;
    if (sx <= b.specW) sx /= b.specW;
    else { sx = (sx - b.specW) / (1 - b.specW); choseSpecular = false; }
    if (choseSpecular) {
        V3 R = kk_reflect(wi);
        float sinAlpha = sqrtf(1 - nc_pow(sy, 2 / (b.exponent + 1)));
        float cosAlpha = nc_pow(sy, 1 / (b.exponent + 1));
        float phi = (2.0f * kPi) * sx, sp, cp_;
        nc_sincos(phi, &sp, &cp_);
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
#if CP_MATH_IS_FAST
    // the same ten-term series with the reciprocal denominators 1 / (4^i (i!)^2) as constants (Horner form): the value only enters
    // log(I0) -- an absolute error of an ulp there is not amplified -- and the ten fp32 divisions were a tenth of the whole BSDF
    const float t = x * x;
    float r = 7.242258480192779e-20f;        // i = 10
    r = r * t + 2.896903392077112e-17f; r = r * t + 9.385966990329842e-15f; r = r * t + 2.4028075495244395e-12f; r = r * t + 4.709502797067901e-10f;
    r = r * t + 6.781684027777778e-08f; r = r * t + 6.781684027777777e-06f; r = r * t + 4.3402777777777775e-04f; r = r * t + 1.5625e-02f; r = r * t + 0.25f;
    return r * t + 1.0f;
#else
    float result = 1.0f, xSq = x * x, xi = xSq, denom = 4.0f;
#pragma unroll
    for (int i = 1; i <= 10; ++i) { result += xi / denom; xi *= xSq; denom *= 4.0f * float((i + 1) * (i + 1)); }
    return result;
#endif
}
CP_D float ma_logI0(float x) { // :292-299
    if (x > 12.0f) return x + 0.5f * (nc_log(1.0f / (kPi * 2.0f * x)) + 1.0f / (8.0f * x));
    return nc_log(ma_I0(x));
}
CP_D float ma_M(float v, float sinThetaI, float sinThetaO, float cosThetaI, float cosThetaO) { // :364-374
    float a = cosThetaI * cosThetaO / v, b = sinThetaI * sinThetaO / v;
    if (v < 0.1f) return nc_exp(-b + ma_logI0(a) - 1.0f / v + 0.6931f + nc_log(1.0f / (2.0f * v)));
    return nc_exp(-b) * ma_I0(a) / (2.0f * v * nc_sinh(1.0f / v));
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
#if CP_MATH_IS_FAST
    float warped = sqrtf(sqrtf(fabsf(cosTheta)));      // |cos|^(1/4): feeds a spline, not amplified
#else
    float warped = nc_pow(fabsf(cosTheta), 0.25f);
#endif
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

// Sines and cosines of the three shifted incident angles that enter M() (marschner_diffuse.cpp:433-452: thetaI - 2 s, thetaI + s,
// thetaI + 4 s for the scale angle s).  M() multiplies them by 1/v <= 400 inside an exponential, so they keep the reference's exact
// fp32 operation sequence -- fl(thetaI + shift), then a correctly rounded sin / cos of that fp32 angle -- in both math modes.  The
// strict mode calls the fp64 functions six times; the fast mode evaluates sin / cos of thetaI once in fp64 and rotates by the shift
// (whose sin / cos are per-material constants, corrected to first order for the rounding of the fp32 addition; the neglected
// second-order term is below 1e-14), which gives the same fp32 values except on fp64 double-rounding boundaries.
struct LobeAngles { float sinR, cosR, sinTT, cosTT, sinTRT, cosTRT; };
CP_D LobeAngles ma_lobe_angles(const BsdfDev &b, float thetaI) {
    const float thetaIR = thetaI - 2.0f * b.scaleAngle, thetaITT = thetaI + b.scaleAngle, thetaITRT = thetaI + 4.0f * b.scaleAngle;
    LobeAngles r;
#if CP_MATH_IS_FAST
    double s, c;
    sincos((double) thetaI, &s, &c);
    const double offR = -(double) (2.0f * b.scaleAngle), offTT = (double) b.scaleAngle, offTRT = (double) (4.0f * b.scaleAngle);
    const double eR = ((double) thetaIR - (double) thetaI) - offR, eTT = ((double) thetaITT - (double) thetaI) - offTT, eTRT = ((double) thetaITRT - (double) thetaI) - offTRT;
    double sd, cd;
    sd = b.lobeSin[0] + eR * b.lobeCos[0]; cd = b.lobeCos[0] - eR * b.lobeSin[0];
    r.sinR = (float) (s * cd + c * sd); r.cosR = (float) (c * cd - s * sd);
    sd = b.lobeSin[1] + eTT * b.lobeCos[1]; cd = b.lobeCos[1] - eTT * b.lobeSin[1];
    r.sinTT = (float) (s * cd + c * sd); r.cosTT = (float) (c * cd - s * sd);
    sd = b.lobeSin[2] + eTRT * b.lobeCos[2]; cd = b.lobeCos[2] - eTRT * b.lobeSin[2];
    r.sinTRT = (float) (s * cd + c * sd); r.cosTRT = (float) (c * cd - s * sd);
#else
    r.sinR = cr_sin(thetaIR); r.cosR = cr_cos(thetaIR);
    r.sinTT = cr_sin(thetaITT); r.cosTT = cr_cos(thetaITT);
    r.sinTRT = cr_sin(thetaITRT); r.cosTRT = cr_cos(thetaITRT);
#endif
    return r;
}

CP_D V3 ma_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    float sinThetaI = wi.y, sinThetaO = wo.y;            // the `t` axis, not the tangent (quirk 2)
    float cosThetaO = ma_trigInverse(sinThetaO);
    // exact in both math modes: thetaI feeds the lobe angles that M() amplifies; thetaO / thetaD / phi are the coordinates of the
    // azimuthal tables, whose caustic peaks turn one ulp of a coordinate into up to 1e-3 of the interpolated value
    float thetaI = cr_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float thetaO = cr_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = cr_cos(thetaD);
    float phi = cr_atan2(wo.x, wo.z);                     // depends on wo only (quirk 2)
    if (phi < 0.0f) phi += kPi * 2.0f;
    const LobeAngles la = ma_lobe_angles(b, thetaI);
    float MR = ma_M(b.vR, la.sinR, sinThetaO, la.cosR, cosThetaO);
    float MTT = ma_M(b.vTT, la.sinTT, sinThetaO, la.cosTT, cosThetaO);
    float MTRT = ma_M(b.vTRT, la.sinTRT, sinThetaO, la.cosTRT, cosThetaO);
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
    float cosTheta = 1.0f + v * nc_log(xi1 + (1.0f - xi1) * nc_exp(-2.0f / v));
    float sinTheta = ma_trigInverse(cosTheta);
    float cosPhi = nc_cos(2 * kPi * xi2);
    return -cosTheta * sinThetaI + sinTheta * cosPhi * cosThetaI;
}
CP_D BsdfSampleOut ma_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    BsdfSampleOut r;
    float sinThetaI = wi.y;
    float cosThetaI = ma_trigInverse(sinThetaI);
    float thetaI = nc_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float v, theta; int lobe;
    float target = sx * (weightR + weightTT + weightTRT);
    if (target < weightR) { r.component = 5; v = b.vR; theta = thetaI - 2.0f * b.scaleAngle; lobe = 0; }
    else if (target < weightR + weightTT) { r.component = 6; v = b.vTT; theta = thetaI + b.scaleAngle; lobe = 1; }
    else { r.component = 7; v = b.vTRT; theta = thetaI + 4.0f * b.scaleAngle; lobe = 2; }
    float sinThetaO = ma_sampleM(v, nc_sin(theta), nc_cos(theta), sx, sy);   // one 2-D sample reused (quirk 4)
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaO = nc_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = nc_cos(thetaD);
    float phi = ma_sample_phi(b.cdf + lobe * 64 * 65, cosThetaD, sy);
    float sinPhi, cosPhi;
    nc_sincos(phi, &sinPhi, &cosPhi);
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
    const LobeAngles la = ma_lobe_angles(b, thetaI);
    // MR = MTT = 0 in the reference (:333-334): their lobes contribute 0 * table value; the scene-driven variant keeps all three
    const float MR = (b.lobeMask & 1) ? ma_M(b.vR, la.sinR, sinThetaO, la.cosR, cosThetaO) : 0.0f;
    const float MTT = (b.lobeMask & 2) ? ma_M(b.vTT, la.sinTT, sinThetaO, la.cosTT, cosThetaO) : 0.0f;
    const float MTRT = (b.lobeMask & 4) ? ma_M(b.vTRT, la.sinTRT, sinThetaO, la.cosTRT, cosThetaO) : 0.0f;
    return MR * ma_azimuthal(b.tab, phi, cosThetaD) + MTT * ma_azimuthal(b.tab + 4096, phi, cosThetaD) + MTRT * ma_azimuthal(b.tab + 8192, phi, cosThetaD);
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
    const LobeAngles la = ma_lobe_angles(b, thetaI);
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float weightSum = weightR + weightTT + weightTRT;
    float pdfR = weightR * ma_M(b.vR, la.sinR, sinThetaO, la.cosR, cosThetaO);
    float pdfTT = weightTT * ma_M(b.vTT, la.sinTT, sinThetaO, la.cosTT, cosThetaO);
    float pdfTRT = weightTRT * ma_M(b.vTRT, la.sinTRT, sinThetaO, la.cosTRT, cosThetaO);
    return (1.0f / weightSum) * (pdfR * mf_azimuthal_pdf(b.pdfs, phi, cosThetaD) + pdfTT * mf_azimuthal_pdf(b.pdfs + 4096, phi, cosThetaD)
                                 + pdfTRT * mf_azimuthal_pdf(b.pdfs + 8192, phi, cosThetaD));
}
// xiN / xiM: the two extra sampler->next2D() draws of :473-474 (the `sample` argument of the BSDF interface is ignored)
CP_D BsdfSampleOut mf_sample(const BsdfDev &b, const V3 &wi, float xiNx, float xiNy, float xiMx, float xiMy) {
    BsdfSampleOut r; r.weight = V3(0.0f);
    float sinThetaI = wi.y;
    float cosThetaI = ma_trigInverse(sinThetaI);
    float thetaI = nc_asin(clampf(sinThetaI, -1.0f, 1.0f));
    float weightR = ma_weight(b.sums, cosThetaI), weightTT = ma_weight(b.sums + 64, cosThetaI), weightTRT = ma_weight(b.sums + 128, cosThetaI);
    float v, theta; int lobe;
    float target = xiNx * (weightR + weightTT + weightTRT);
    if (target < weightR) { r.component = 0; v = b.vR; theta = thetaI - 2.0f * b.scaleAngle; lobe = 0; }
    else if (target < weightR + weightTT) { r.component = 1; v = b.vTT; theta = thetaI + b.scaleAngle; lobe = 1; }
    else { r.component = 2; v = b.vTRT; theta = thetaI + 4.0f * b.scaleAngle; lobe = 2; }
    float sinThetaO = ma_sampleM(v, nc_sin(theta), nc_cos(theta), xiMx, xiMy);
    float cosThetaO = ma_trigInverse(sinThetaO);
    float thetaO = nc_asin(clampf(sinThetaO, -1.0f, 1.0f));
    float thetaD = (thetaO - thetaI) * 0.5f;
    float cosThetaD = nc_cos(thetaD);
    float phi = ma_sample_phi(b.cdf + lobe * 64 * 65, cosThetaD, xiNy);
    float sinPhi, cosPhi;
    nc_sincos(phi, &sinPhi, &cosPhi);
    r.wo = V3(sinPhi * cosThetaO, sinThetaO, cosPhi * cosThetaO);
    r.pdf = mf_pdf(b, wi, r.wo);
    r.type = EDeltaReflection;
    if (r.pdf <= 0 || r.pdf > 1) return r;
    r.weight = mf_eval(b, wi, r.wo) / r.pdf;
    return r;
}

// ------------------------------------------------------------------------------------------ RoughPlastic
// MicrofacetDistribution (isotropic): src/bsdfs/microfacet.h:184-232 eval, :238-279 sample/pdf, :284-386 sampleAll, :389-447 visible
// normals, :470-510 smithG1 / G, :555-673 sampleVisible11;  math::erf / erfinv / hypot2: src/libcore/math.cpp:25-86
// RoughPlastic eval / pdf / sample: src/bsdfs/roughplastic.cpp:325-375, 377-436, 438-494
CP_D float mfd_signum(float v) { return v < 0 ? -1.0f : (v > 0 ? 1.0f : 0.0f); }
CP_D float mfd_erfinv(float x) {
    float w = -nc_log((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
    }
    return p * x;
}
CP_D float mfd_erf(float x) {
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    const float sign = mfd_signum(x);
    x = fabsf(x);
    const float t = 1.0f / (1.0f + p * x);
    const float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * nc_exp(-x * x);
    return sign * y;
}
CP_D float mfd_hypot2(float a, float b) {
    float r;
    if (fabsf(a) > fabsf(b)) { r = b / a; r = fabsf(a) * sqrtf(1.0f + r * r); }
    else if (b != 0.0f) { r = a / b; r = fabsf(b) * sqrtf(1.0f + r * r); }
    else r = 0.0f;
    return r;
}
CP_D float mfd_eval(const BsdfDev &b, const V3 &m) {
    if (m.z <= 0) return 0.0f;
    const float alpha = b.alpha, cosTheta2 = m.z * m.z;
    const float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
    float result;
    if (b.distr == 0) result = nc_exp(-beckmannExponent) / (kPi * alpha * alpha * cosTheta2 * cosTheta2);
    else if (b.distr == 1) { const float root = (1.0f + beckmannExponent) * cosTheta2; result = 1.0f / (kPi * alpha * alpha * root * root); }
    else result = sqrtf((b.exponent + 2) * (b.exponent + 2)) * kInvTwoPi * nc_pow(m.z, b.exponent);
    if (result * m.z < 1e-20f) result = 0;
    return result;
}
CP_D float mfd_smithG1(const BsdfDev &b, const V3 &v, const V3 &m) {
    if (dot(v, m) * v.z <= 0) return 0.0f;
    const float temp = 1 - v.z * v.z;
    const float tanTheta = temp <= 0.0f ? 0.0f : fabsf(sqrtf(temp) / v.z);
    if (tanTheta == 0.0f) return 1.0f;
    if (b.distr != 1) {
        const float a = 1.0f / (b.alpha * tanTheta);
        if (a >= 1.6f) return 1.0f;
        const float aSqr = a * a;
        return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
    }
    return 2.0f / (1.0f + mfd_hypot2(1.0f, b.alpha * tanTheta));
}
CP_D float mfd_pdf(const BsdfDev &b, const V3 &wi, const V3 &m) {
    if (!b.sampleVisible) return mfd_eval(b, m) * m.z;
    if (wi.z == 0) return 0.0f;
    return mfd_smithG1(b, wi, m) * fabsf(dot(wi, m)) * mfd_eval(b, m) / fabsf(wi.z);
}
CP_D V3 mfd_sampleAll(const BsdfDev &b, float sx, float sy) {
    float cosThetaM, sinPhiM, cosPhiM;
    nc_sincos((2.0f * kPi) * sy, &sinPhiM, &cosPhiM);
    if (b.distr == 0) cosThetaM = 1.0f / sqrtf(1.0f + b.alpha * b.alpha * -nc_log(1.0f - sx));
    else if (b.distr == 1) cosThetaM = 1.0f / sqrtf(1.0f + b.alpha * b.alpha * sx / (1.0f - sx));
    else cosThetaM = nc_pow(sx, 1.0f / (b.exponent + 2.0f));
    const float sinThetaM = sqrtf(fmaxf(0.0f, 1 - cosThetaM * cosThetaM));
    return V3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
}
CP_D void mfd_sampleVisible11(const BsdfDev &b, float thetaI, float sx, float sy, float &slopeX, float &slopeY) {
    const float SQRT_PI_INV = 1 / sqrtf(kPi);
    if (b.distr == 0) {
        if (thetaI < 1e-4f) {
            const float r = sqrtf(-nc_log(1.0f - sx));
            float sinPhi, cosPhi; nc_sincos(2 * kPi * sy, &sinPhi, &cosPhi);
            slopeX = r * cosPhi; slopeY = r * sinPhi; return;
        }
        const float tanThetaI = nc_tan(thetaI), cotThetaI = 1 / tanThetaI;
        float a = -1, c = mfd_erf(cotThetaI);
        const float sample_x = fmaxf(sx, 1e-6f);
        const float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
        float bb = c - (1 + c) * nc_pow(1 - sample_x, fit);
        const float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * nc_exp(-cotThetaI * cotThetaI));
        int it = 0;
        while (++it < 10) {
            if (!(bb >= a && bb <= c)) bb = 0.5f * (a + c);
            const float invErf = mfd_erfinv(bb);
            const float value = normalization * (1 + bb + SQRT_PI_INV * tanThetaI * nc_exp(-invErf * invErf)) - sample_x;
            const float derivative = normalization * (1 - invErf * tanThetaI);
            if (fabsf(value) < 1e-5f) break;
            if (value > 0) c = bb; else a = bb;
            bb -= value / derivative;
        }
        slopeX = mfd_erfinv(bb);
        slopeY = mfd_erfinv(2.0f * fmaxf(sy, 1e-6f) - 1.0f);
    } else {
        if (thetaI < 1e-4f) {
            const float r = safe_sqrt(sx / (1 - sx));
            float sinPhi, cosPhi; nc_sincos(2 * kPi * sy, &sinPhi, &cosPhi);
            slopeX = r * cosPhi; slopeY = r * sinPhi; return;
        }
        const float tanThetaI = nc_tan(thetaI);
        const float a = 1 / tanThetaI;
        const float G1 = 2.0f / (1.0f + safe_sqrt(1.0f + 1.0f / (a * a)));
        float A = 2.0f * sx / G1 - 1.0f;
        if (fabsf(A) == 1) A -= mfd_signum(A) * kEpsilon;
        const float tmp = 1.0f / (A * A - 1.0f);
        const float B = tanThetaI;
        const float D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
        const float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
        slopeX = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
        float S;
        if (sy > 0.5f) { S = 1.0f; sy = 2.0f * (sy - 0.5f); }
        else { S = -1.0f; sy = 2.0f * (0.5f - sy); }
        const float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                        (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
        slopeY = S * z * sqrtf(1.0f + slopeX * slopeX);
    }
}
CP_D V3 mfd_sample(const BsdfDev &b, const V3 &_wi, float sx, float sy) {
    if (!b.sampleVisible) return mfd_sampleAll(b, sx, sy);
    const V3 wi = normalize(V3(b.alpha * _wi.x, b.alpha * _wi.y, _wi.z));
    float theta = 0, phi = 0;
    if (wi.z < 0.99999f) { theta = nc_acos(wi.z); phi = nc_atan2(wi.y, wi.x); }
    float sinPhi, cosPhi; nc_sincos(phi, &sinPhi, &cosPhi);
    float slx, sly;
    mfd_sampleVisible11(b, theta, sx, sy, slx, sly);
    float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
    rx *= b.alpha; ry *= b.alpha;
    const float normalization = 1.0f / sqrtf(rx * rx + ry * ry + 1.0f);
    return V3(-rx * normalization, -ry * normalization, normalization);
}

CP_D V3 rp_eval(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
    V3 result(0.0f);
    const V3 H = normalize(wo + wi);
    const float D = mfd_eval(b, H);
    const float F = fresnelDielectricExt(dot(wi, H), b.eta);
    const float G = mfd_smithG1(b, wi, H) * mfd_smithG1(b, wo, H);
    const float value = F * D * G / (4.0f * wi.z);
    result += b.specular * value;
    V3 diff = b.diffuse;
    const float T12 = ma_T(b, wi.z), T21 = ma_T(b, wo.z);
    if (b.nonlinear) diff = V3(diff.x / (1.0f - diff.x * b.Fdr), diff.y / (1.0f - diff.y * b.Fdr), diff.z / (1.0f - diff.z * b.Fdr));
    else diff = diff / (1 - b.Fdr);
    result += diff * (kInvPi * wo.z * T12 * T21 * b.invEta2);
    return result;
}
CP_D float rp_probSpecular(const BsdfDev &b, float cosThetaI) {
    const float probSpecular = 1 - ma_T(b, cosThetaI);
    return (probSpecular * b.specW) / (probSpecular * b.specW + (1 - probSpecular) * (1 - b.specW));
}
CP_D float rp_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo) {
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    const V3 H = normalize(wo + wi);
    const float probSpecular = rp_probSpecular(b, wi.z), probDiffuse = 1 - probSpecular;
    const float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
    const float prob = mfd_pdf(b, wi, H);
    float result = prob * dwh_dwo * probSpecular;
    result += probDiffuse * (kInvPi * wo.z);
    return result;
}
CP_D BsdfSampleOut rp_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    BsdfSampleOut r; r.wo = V3(0.0f); r.weight = V3(0.0f); r.pdf = 0.0f; r.type = 0; r.component = -1;
    if (wi.z <= 0) return r;
    bool choseSpecular = true;
    const float probSpecular = rp_probSpecular(b, wi.z);
    if (sy < probSpecular) sy /= probSpecular;
    else { sy = (sy - probSpecular) / (1 - probSpecular); choseSpecular = false; }
    if (choseSpecular) {
        const V3 m = mfd_sample(b, wi, sx, sy);
        r.wo = 2 * dot(wi, m) * m - wi;
        r.component = 0; r.type = EGlossyReflection;
        if (r.wo.z <= 0) return r;
    } else {
        r.component = 1; r.type = EDiffuseReflection;
        r.wo = squareToCosineHemisphere(sx, sy);
    }
    r.pdf = rp_pdf(b, wi, r.wo);
    if (r.pdf == 0) return r;
    r.weight = rp_eval(b, wi, r.wo) / r.pdf;
    return r;
}

// ------------------------------------------------------------------------------------------ textures
// Texture2D::eval (texture.cpp:112-121, no filtering: Checkerboard ignores the differentials) + Checkerboard::eval (checkerboard.cpp:65-72)
CP_D V3 bsdf_reflectance(const BsdfDev &b, float u, float v) {
    if (b.texKind == 0) return b.diffuse;
    const float uu = u * b.uvScale[0] + b.uvOffset[0], vv = v * b.uvScale[1] + b.uvOffset[1];
    int mx = ((int) (uu * 2)) % 2, my = ((int) (vv * 2)) % 2;           // math::modulo(int, 2): always-positive remainder
    if (mx < 0) mx += 2;
    if (my < 0) my += 2;
    const int x = 2 * mx - 1, y = 2 * my - 1;
    return x * y == 1 ? b.diffuse : b.color1;
}

// ------------------------------------------------------------------------------------------ SmoothDiffuse (+ TwoSided)
// src/bsdfs/diffuse.cpp:109-156; src/bsdfs/twosided.cpp:101-181 when b.twoSided
CP_D V3 df_eval(const BsdfDev &b, V3 wi, V3 wo, float u, float v) {
    if (b.twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }
    if (wi.z <= 0 || wo.z <= 0) return V3(0.0f);
    return bsdf_reflectance(b, u, v) * (kInvPi * wo.z);
}
CP_D float df_pdf(const BsdfDev &b, V3 wi, V3 wo) {
    if (b.twoSided && wi.z <= 0) { wi.z *= -1; wo.z *= -1; }
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    return kInvPi * wo.z;
}
CP_D BsdfSampleOut df_sample(const BsdfDev &b, V3 wi, float sx, float sy, float u, float v) {
    BsdfSampleOut r; r.wo = V3(0.0f); r.weight = V3(0.0f); r.pdf = 0.0f; r.type = 0; r.component = -1;
    bool flipped = false;
    if (b.twoSided && wi.z < 0) { wi.z *= -1; flipped = true; }
    if (wi.z <= 0) return r;
    r.wo = squareToCosineHemisphere(sx, sy);
    r.component = 0; r.type = EDiffuseReflection;
    r.pdf = kInvPi * r.wo.z;
    r.weight = bsdf_reflectance(b, u, v);
    if (flipped && !isZero(r.weight) && r.pdf != 0) { r.wo.z *= -1; r.component += 1; }
    return r;
}

// ------------------------------------------------------------------------------------------ SmoothPlastic
// src/bsdfs/plastic.cpp:246-281 (eval), :283-313 (pdf), :381-445 (sample with pdf): a delta reflection (component 0, discrete measure)
// over a diffuse base (component 1), front side only; typeMask = EAll, component = -1.  fdrInt rides in b.Fdr.
CP_D V3 pl_diff(const BsdfDev &b, float u, float v) {
    V3 diff = bsdf_reflectance(b, u, v);
    if (b.nonlinear) diff = V3(diff.x / (1.0f - diff.x * b.Fdr), diff.y / (1.0f - diff.y * b.Fdr), diff.z / (1.0f - diff.z * b.Fdr));
    else diff = diff / (1 - b.Fdr);
    return diff;
}
CP_D float pl_probSpecular(const BsdfDev &b, float Fi) { return (Fi * b.specW) / (Fi * b.specW + (1 - Fi) * (1 - b.specW)); }
CP_D V3 pl_eval(const BsdfDev &b, const V3 &wi, const V3 &wo, bool discrete, float u, float v) {
    if (wo.z <= 0 || wi.z <= 0) return V3(0.0f);
    const float Fi = fresnelDielectricExt(wi.z, b.eta);
    if (discrete) {
        if (fabsf(dot(kk_reflect(wi), wo) - 1) < kDeltaEpsilon) return b.specular * Fi;
        return V3(0.0f);
    }
    const float Fo = fresnelDielectricExt(wo.z, b.eta);
    return pl_diff(b, u, v) * ((kInvPi * wo.z) * b.invEta2 * (1 - Fi) * (1 - Fo));
}
CP_D float pl_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo, bool discrete) {
    if (wo.z <= 0 || wi.z <= 0) return 0.0f;
    const float ps = pl_probSpecular(b, fresnelDielectricExt(wi.z, b.eta));
    if (discrete) return fabsf(dot(kk_reflect(wi), wo) - 1) < kDeltaEpsilon ? ps : 0.0f;
    return (kInvPi * wo.z) * (1 - ps);
}
CP_D BsdfSampleOut pl_sample(const BsdfDev &b, const V3 &wi, float sx, float sy, float u, float v) {
    BsdfSampleOut r; r.wo = V3(0.0f); r.weight = V3(0.0f); r.pdf = 0.0f; r.type = 0; r.component = -1;
    if (wi.z <= 0) return r;
    const float Fi = fresnelDielectricExt(wi.z, b.eta);
    const float ps = pl_probSpecular(b, Fi);
    if (sx < ps) {
        r.component = 0; r.type = EDeltaReflection; r.wo = kk_reflect(wi); r.pdf = ps;
        r.weight = b.specular * Fi / ps;
    } else {
        r.component = 1; r.type = EDiffuseReflection;
        r.wo = squareToCosineHemisphere((sx - ps) / (1 - ps), sy);
        const float Fo = fresnelDielectricExt(r.wo.z, b.eta);
        r.pdf = (1 - ps) * (kInvPi * r.wo.z);
        r.weight = pl_diff(b, u, v) * (b.invEta2 * (1 - Fi) * (1 - Fo) / (1 - ps));
    }
    return r;
}

// ------------------------------------------------------------------------------------------ ThinDielectric / MarschnerDielectric
// src/bsdfs/thindielectric.cpp:143-245 and the fork's src/bsdfs/marschnerdielectric.cpp:245-499 (see oracle/o_bsdf.h for the
// walk through the latter: its eval() is identically zero as committed, its pdf() is the cosine density of the diffuse component,
// its sample() is a thin dielectric with probability m_specularSamplingWeight and a dead diffuse branch otherwise).
// `discrete` selects the EDiscrete measure (common.h:56-67); the path tracer only ever asks for ESolidAngle.
CP_D float thin_slab_reflectance(float cosThetaI, float eta) {          // R' = R + TRT + TR^3T + ...
    float R = fresnelDielectricExt(cosThetaI, eta), T = 1 - R;
    if (R < 1) R += T * T * R / (1 - R * R);
    return R;
}
CP_D V3 td_eval(const BsdfDev &b, const V3 &wi, const V3 &wo, bool discrete) {
    const float R = thin_slab_reflectance(fabsf(wi.z), b.eta);
    if (wi.z * wo.z >= 0) {
        if (!discrete || fabsf(dot(kk_reflect(wi), wo) - 1) > kDeltaEpsilon) return V3(0.0f);
        return b.specular * R;
    }
    if (!discrete || fabsf(dot(V3(-wi.x, -wi.y, -wi.z), wo) - 1) > kDeltaEpsilon) return V3(0.0f);
    return b.diffuse * (1 - R);
}
CP_D float td_pdf(const BsdfDev &b, const V3 &wi, const V3 &wo, bool discrete) {
    const float R = thin_slab_reflectance(fabsf(wi.z), b.eta);
    if (wi.z * wo.z >= 0) {
        if (!discrete || fabsf(dot(kk_reflect(wi), wo) - 1) > kDeltaEpsilon) return 0.0f;
        return R;
    }
    if (!discrete || fabsf(dot(V3(-wi.x, -wi.y, -wi.z), wo) - 1) > kDeltaEpsilon) return 0.0f;
    return 1 - R;
}
CP_D BsdfSampleOut thin_sample(const V3 &wi, float eta, const V3 &specR, const V3 &specT, float sx) {
    BsdfSampleOut r;
    const float R = thin_slab_reflectance(fabsf(wi.z), eta);
    if (sx <= R) { r.component = 0; r.type = EDeltaReflection; r.wo = kk_reflect(wi); r.pdf = R; r.weight = specR; }
    else { r.component = 1; r.type = ENull; r.wo = V3(-wi.x, -wi.y, -wi.z); r.pdf = 1 - R; r.weight = specT; }
    return r;
}
CP_D float md_pdf(const V3 &wi, const V3 &wo, bool discrete) {
    if (discrete || wi.z <= 0 || wo.z <= 0) return 0.0f;
    return kInvPi * wo.z;
}
CP_D BsdfSampleOut md_sample(const BsdfDev &b, const V3 &wi, float sx, float sy) {
    if (sx <= b.specW) return thin_sample(wi, b.eta, b.specular, b.specT, sx / b.specW);
    sx = (sx - b.specW) / (1 - b.specW);
    BsdfSampleOut r; r.weight = V3(0.0f);
    r.wo = squareToCosineHemisphere(sx, sy);
    r.component = 2; r.type = EDiffuseReflection;
    r.pdf = md_pdf(wi, r.wo, false);
    return r;                                           // eval() / pdf == 0: the path ends here
}

// ------------------------------------------------------------------------------------------ dispatch
// (u, v) = its.uv, only read by the textured kinds.  `twosided` around roughplastic / plastic is applied here (twosided.cpp:101-181);
// the diffuse kind handles its own flag.
CP_D bool bsdf_wrapped(const BsdfDev &b) { return b.twoSided && (b.kind == 4 || b.kind == 7 || b.kind == 8); }
CP_D V3 bsdf_eval(const BsdfDev &b, V3 wi, V3 wo, bool discrete = false, float u = 0.0f, float v = 0.0f) {
    if (bsdf_wrapped(b) && !(wi.z > 0)) { wi.z *= -1; wo.z *= -1; }
    if (b.kind == 5) return td_eval(b, wi, wo, discrete);
    if (b.kind == 7) return pl_eval(b, wi, wo, discrete, u, v);
    if (b.kind == 6 || b.kind == 8 || discrete) return V3(0.0f);         // mirror.cpp:233-247: zero in both measures, as committed
    return b.kind == 0 ? kk_eval(b, wi, wo) : b.kind == 1 ? ma_eval(b, wi, wo) : b.kind == 2 ? df_eval(b, wi, wo, u, v) : b.kind == 3 ? mf_eval(b, wi, wo) : rp_eval(b, wi, wo);
}
CP_D float bsdf_pdf(const BsdfDev &b, V3 wi, V3 wo, bool discrete = false) {
    if (bsdf_wrapped(b) && !(wi.z > 0)) { wi.z *= -1; wo.z *= -1; }
    if (b.kind == 5) return td_pdf(b, wi, wo, discrete);
    if (b.kind == 6) return md_pdf(wi, wo, discrete);
    if (b.kind == 7) return pl_pdf(b, wi, wo, discrete);
    if (b.kind == 8) return discrete ? 1.0f : 0.0f;                       // mirror.cpp:249-254
    if (discrete) return 0.0f;
    return b.kind == 0 ? kk_pdf(b, wi, wo) : b.kind == 1 ? 1.0f : b.kind == 2 ? df_pdf(b, wi, wo) : b.kind == 3 ? mf_pdf(b, wi, wo) : rp_pdf(b, wi, wo);
}
// true for BSDFs whose sample() pulls more numbers from the sampler than the two it is handed (fixed Marschner: 4)
CP_D bool bsdf_draws_extra(const BsdfDev &b) { return b.kind == 3; }
// BSDF::getType() & ESmooth (bsdf.h:278): everything except the thin dielectric, whose two components are discrete
CP_D bool bsdf_has_smooth(const BsdfDev &b) { return b.kind != 5 && b.kind != 8; }
// eval() in the solid-angle measure is identically zero: emitter samples are drawn and counted, but can never contribute
CP_D bool bsdf_eval_is_zero(const BsdfDev &b) { return b.kind == 6; }
CP_D BsdfSampleOut bsdf_sample(const BsdfDev &b, V3 wi, float sx, float sy, const float4 &extra, float u = 0.0f, float v = 0.0f) {
    if (b.kind == 5) return thin_sample(wi, b.eta, b.specular, b.diffuse, sx);
    if (b.kind == 6) return md_sample(b, wi, sx, sy);
    bool flipped = false;
    if (bsdf_wrapped(b) && wi.z < 0) { wi.z *= -1; flipped = true; }
    if (b.kind == 8) {                                                   // mirror.cpp:256-276
        BsdfSampleOut m; m.wo = V3(0.0f); m.weight = V3(0.0f); m.pdf = 0.0f; m.type = 0; m.component = -1;
        if (wi.z > 0) { m.component = 0; m.type = EDeltaReflection; m.wo = kk_reflect(wi); m.pdf = 1.0f; m.weight = b.specular; }
        if (flipped && !isZero(m.weight) && m.pdf != 0) { m.wo.z *= -1; m.component += 1; }
        return m;
    }
    BsdfSampleOut r = b.kind == 0 ? kk_sample(b, wi, sx, sy) : b.kind == 1 ? ma_sample(b, wi, sx, sy) : b.kind == 2 ? df_sample(b, wi, sx, sy, u, v)
         : b.kind == 3 ? mf_sample(b, wi, extra.x, extra.y, extra.z, extra.w) : b.kind == 7 ? pl_sample(b, wi, sx, sy, u, v) : rp_sample(b, wi, sx, sy);
    if (flipped && !isZero(r.weight) && r.pdf != 0) { r.wo.z *= -1; r.component += 2; }
    return r;
}

} // namespace cp

// cp_env.cuh -- environment-map emitter on the device (sm_100a).
//
// Replaces (reference file:line):
//   EnvironmentMap::evalEnvironment          src/emitters/envmap.cpp:380-410
//   MIPMap::evalTexel / evalBilinear / eval   include/mitsuba/render/mipmap.h:503-596, 629-700
//   sampleDirect / internalSampleDirection    src/emitters/envmap.cpp:516-543, 567-602, sampleReuse :657-662
//   pdfDirect / internalPdfDirection          src/emitters/envmap.cpp:545-556, 603-635
//   fillDirectSamplingRecord                  src/emitters/envmap.cpp:358-374
// The 512x256 texel array (2 MB as float4) and the CDFs (0.5 MB) stay L2-resident.
#pragma once
#include "cp_scene.cuh"

namespace cp {

CP_D V3 mul3(const float *m, const V3 &v) {
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}

// mipmap.h:503-560, level 0, bcu = ERepeat, bcv = EClamp (envmap.cpp:167-170)
CP_D V3 env_texel(const EnvDev &E, int x, int y) {
    if (x < 0 || x >= E.w) { int r = x % E.w; x = r < 0 ? r + E.w : r; }
    if (y < 0 || y >= E.h) y = clampi(y, 0, E.h - 1);
    float4 t = __ldg(E.texels + (size_t) y * E.w + x);
    return V3(t.x, t.y, t.z);
}
CP_D V3 env_bilinear(const EnvDev &E, float uvx, float uvy) {
    if (!isfinite(uvx) || !isfinite(uvy)) return V3(0.0f);
    float u = uvx * E.w - 0.5f, v = uvy * E.h - 0.5f;
    int xPos = (int) floorf(u), yPos = (int) floorf(v);
    float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    return env_texel(E, xPos, yPos) * dx2 * dy2 + env_texel(E, xPos, yPos + 1) * dx2 * dy1
         + env_texel(E, xPos + 1, yPos) * dx1 * dy2 + env_texel(E, xPos + 1, yPos + 1) * dx1 * dy1;
}

// evalEnvironment for rays without differentials (every bounce ray: Ray(...) clears them, ray.h:196-208)
CP_D V3 env_eval(const EnvDev &E, const V3 &d) {
    V3 v = mul3(E.toLocal, d);
    float uvx = cr_atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
    return env_bilinear(E, uvx, uvy) * E.scale;
}
// evalEnvironment for camera rays (EWA filter type, maxAnisotropy 10).  Footprints below one texel resolve
// to evalBilinear(0, uv) in every branch of MIPMap::eval; larger footprints would need the Lanczos MIP
// pyramid, which is not built -- they are counted in `unsupported` and answered at level 0.
CP_D V3 env_eval_filtered(const EnvDev &E, const V3 &d, const V3 &rxDir, const V3 &ryDir, unsigned long long *unsupported) {
    V3 v = mul3(E.toLocal, d);
    float uvx = cr_atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
    V3 dvdx = mul3(E.toLocal, rxDir) - v, dvdy = mul3(E.toLocal, ryDir) - v;
    float t1 = kInvTwoPi / (v.x * v.x + v.z * v.z), t2 = -kInvPi / fmaxf(safe_sqrt(1.0f - v.y * v.y), kEpsilon);
    float du0 = t1 * (dvdx.z * v.x - dvdx.x * v.z) * E.w, dv0 = t2 * dvdx.y * E.h;
    float du1 = t1 * (dvdy.z * v.x - dvdy.x * v.z) * E.w, dv1 = t2 * dvdy.y * E.h;
    float A = dv0 * dv0 + dv1 * dv1, B = -2.0f * (du0 * dv0 + du1 * dv1), C = du0 * du0 + du1 * du1, F = A * C - B * B * 0.25f;
    float root = cr_hypot(A - C, B), Aprime = 0.5f * (A + C - root), Cprime = 0.5f * (A + C + root);
    float majorRadius = Aprime != 0 ? sqrtf(F / Aprime) : 0, minorRadius = Cprime != 0 ? sqrtf(F / Cprime) : 0;
    bool level0;
    if (!(minorRadius > 0) || !(majorRadius > 0) || F < 0) level0 = floorf(cr_log2(fmaxf(majorRadius, kEpsilon))) < 0;
    else level0 = majorRadius < 1;
    if (!level0 && unsupported) atomicAdd(unsupported, 1ull);
    return env_bilinear(E, uvx, uvy) * E.scale;
}

// envmap.cpp:657-662 -- std::lower_bound over cdf[0..size] then sample reuse
CP_D uint32_t env_sample_reuse(const float *__restrict__ cdf, uint32_t size, float &sample) {
    uint32_t lo = 0, len = size + 1;
    while (len > 0) {                                   // first index with cdf[idx] >= sample
        uint32_t half = len >> 1, mid = lo + half;
        if (__ldg(cdf + mid) < sample) { lo = mid + 1; len -= half + 1; } else len = half;
    }
    int idx = (int) lo - 1;
    uint32_t index = (uint32_t) min(max(idx, 0), (int) size - 1);
    float c0 = __ldg(cdf + index), c1 = __ldg(cdf + index + 1);
    sample = (sample - c0) / (c1 - c0);
    return index;
}

struct EnvSample { V3 value, d; float dist, pdf; };

// sampleDirect: value is already divided by the pdf; pdf == 0 signals "no sample" (scene.cpp:838-852)
CP_D EnvSample env_sample_direct(const EnvDev &E, const V3 &ref, float sx, float sy) {
    EnvSample r; r.pdf = 0; r.value = V3(0.0f); r.d = V3(0.0f); r.dist = 0;
    uint32_t row = env_sample_reuse(E.cdfRows, (uint32_t) E.h, sy);
    uint32_t col = env_sample_reuse(E.cdfCols + (size_t) row * (E.w + 1), (uint32_t) E.w, sx);
    float posx = (float) col + intervalToTent(sx), posy = (float) row + intervalToTent(sy);
    int xPos = (int) floorf(posx), yPos = (int) floorf(posy);
    float dx1 = posx - xPos, dx2 = 1.0f - dx1, dy1 = posy - yPos, dy2 = 1.0f - dy1;
    V3 value1 = env_texel(E, xPos, yPos) * dx2 * dy2 + env_texel(E, xPos + 1, yPos) * dx1 * dy2;
    V3 value2 = env_texel(E, xPos, yPos + 1) * dx2 * dy1 + env_texel(E, xPos + 1, yPos + 1) * dx1 * dy1;
    V3 value = (value1 + value2) * E.scale;
    float pdf = (luminance(value1) * __ldg(E.rowWeights + clampi(yPos, 0, E.h - 1)) +
                 luminance(value2) * __ldg(E.rowWeights + clampi(yPos + 1, 0, E.h - 1))) * E.normalization;
    float sinPhi, cosPhi, sinTheta, cosTheta;
    cr_sincos(E.pixelSizeX * (posx + 0.5f), &sinPhi, &cosPhi);
    cr_sincos(E.pixelSizeY * (posy + 0.5f), &sinTheta, &cosTheta);
    V3 d(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
    pdf /= fmaxf(fabsf(sinTheta), kEpsilon);
    V3 dw = mul3(E.toWorld, d);
    // intersect the scene bounding sphere (bsphere.h:88-95)
    V3 o = ref - V3(E.bsCenter[0], E.bsCenter[1], E.bsCenter[2]);
    float nearT, farT;
    bool ok = solveQuadratic(dot(dw, dw), 2 * dot(o, dw), dot(o, o) - E.bsRadius * E.bsRadius, nearT, farT);
    if (isZero(value) || pdf == 0 || !ok || nearT >= 0 || farT <= 0) return r;
    r.pdf = pdf; r.dist = farT; r.d = dw; r.value = value / pdf;
    return r;
}

CP_D float env_pdf_direct(const EnvDev &E, const V3 &dWorld) {
    V3 d = mul3(E.toLocal, dWorld);
    float uvx = cr_atan2(d.x, -d.z) * kInvTwoPi, uvy = safe_acos(d.y) * kInvPi;
    if (!isfinite(uvx) || !isfinite(uvy)) return 0.0f;
    float u = uvx * E.w - 0.5f, v = uvy * E.h - 0.5f;
    int xPos = (int) floorf(u), yPos = (int) floorf(v);
    float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    V3 value1 = env_texel(E, xPos, yPos) * dx2 * dy2 + env_texel(E, xPos + 1, yPos) * dx1 * dy2;
    V3 value2 = env_texel(E, xPos, yPos + 1) * dx2 * dy1 + env_texel(E, xPos + 1, yPos + 1) * dx1 * dy1;
    float sinTheta = safe_sqrt(1 - d.y * d.y);
    return (luminance(value1) * __ldg(E.rowWeights + clampi(yPos, 0, E.h - 1)) +
            luminance(value2) * __ldg(E.rowWeights + clampi(yPos + 1, 0, E.h - 1))) * E.normalization / fmaxf(fabsf(sinTheta), kEpsilon);
}

// fillDirectSamplingRecord: false => the BSDF-sampled escape is dropped (path.cpp:240-241)
CP_D bool env_fill_direct(const EnvDev &E, const V3 &ro, const V3 &rd) {
    V3 o = ro - V3(E.bsCenter[0], E.bsCenter[1], E.bsCenter[2]);
    float nearT, farT;
    if (!solveQuadratic(dot(rd, rd), 2 * dot(o, rd), dot(o, o) - E.bsRadius * E.bsRadius, nearT, farT) || nearT > 0 || farT < 0) return false;
    return true;
}

} // namespace cp

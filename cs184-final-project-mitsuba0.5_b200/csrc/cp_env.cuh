// cp_env.cuh -- environment-map emitter on the device (sm_100a).
//
// Replaces (reference file:line):
//   EnvironmentMap::evalEnvironment          src/emitters/envmap.cpp:380-410
//   MIPMap::evalTexel / evalBilinear / eval / evalEWA   include/mitsuba/render/mipmap.h:503-596, 629-725, 760-836
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
    float uvx = nc_atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
    return env_bilinear(E, uvx, uvy) * E.scale;
}
// math::log2 of the reference (src/libcore/math.cpp:103-106): natural logarithm times the fp32 constant 1 / log(2.0f)
CP_D float env_log2(float x) { return nc_log(x) * (1.0f / 0.6931471805599453f); }

// ---- MIP levels (TMIPMap, include/mitsuba/render/mipmap.h): level 0 = E.texels, level l >= 1 = E.mipTexels + offset[l]
CP_D V3 env_texel_level(const EnvDev &E, int level, int x, int y) {                      // evalTexel :503-560, bcu = ERepeat, bcv = EClamp
    if (level == 0) return env_texel(E, x, y);
    const int w = E.mip->w[level], h = E.mip->h[level];
    if (x < 0 || x >= w) { int r = x % w; x = r < 0 ? r + w : r; }
    if (y < 0 || y >= h) y = clampi(y, 0, h - 1);
    const float4 t = __ldg(E.mipTexels + E.mip->offset[level] + (size_t) y * w + x);
    return V3(t.x, t.y, t.z);
}
CP_D V3 env_box_level(const EnvDev &E, int level, float uvx, float uvy) {               // evalBox :563-566
    const int w = level == 0 ? E.w : E.mip->w[level], h = level == 0 ? E.h : E.mip->h[level];
    return env_texel_level(E, level, (int) floorf(uvx * w), (int) floorf(uvy * h));
}
CP_D V3 env_bilinear_level(const EnvDev &E, int level, float uvx, float uvy) {          // evalBilinear :572-596
    if (!isfinite(uvx) || !isfinite(uvy)) return V3(0.0f);
    const int levels = E.mip ? E.mip->levels : 1;
    if (level >= levels) return env_box_level(E, levels - 1, uvx, uvy);
    if (level == 0) return env_bilinear(E, uvx, uvy);
    const int w = E.mip->w[level], h = E.mip->h[level];
    float u = uvx * w - 0.5f, v = uvy * h - 0.5f;
    int xPos = (int) floorf(u), yPos = (int) floorf(v);
    float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    return env_texel_level(E, level, xPos, yPos) * dx2 * dy2 + env_texel_level(E, level, xPos, yPos + 1) * dx2 * dy1
         + env_texel_level(E, level, xPos + 1, yPos) * dx1 * dy2 + env_texel_level(E, level, xPos + 1, yPos + 1) * dx1 * dy1;
}
// evalEWA :760-836: elliptically weighted average over the texels inside the ellipse A u^2 + B u v + C v^2 < 1 (level-0 texel units)
CP_D V3 env_ewa_level(const EnvDev &E, int level, float uvx, float uvy, float A, float B, float C) {
    if (!isfinite(A + B + C + uvx + uvy)) return V3(0.0f);
    const int levels = E.mip->levels;
    if (level >= levels) return env_box_level(E, levels - 1, uvx, uvy);
    const int w = level == 0 ? E.w : E.mip->w[level], h = level == 0 ? E.h : E.mip->h[level];
    const float u = uvx * w - 0.5f, v = uvy * h - 0.5f;
    const float rx = E.mip->ratioX[level], ry = E.mip->ratioY[level];
    A /= rx * rx; B /= rx * ry; C /= ry * ry;
    const float invDet = 1.0f / (-B * B + 4.0f * A * C), deltaU = 2.0f * sqrtf(C * invDet), deltaV = 2.0f * sqrtf(A * invDet);
    const int u0 = (int) ceilf(u - deltaU), u1 = (int) floorf(u + deltaU), v0 = (int) ceilf(v - deltaV), v1 = (int) floorf(v + deltaV);
    const float As = A * 64, Bs = B * 64, Cs = C * 64;
    V3 result(0.0f);
    float denominator = 0.0f;
    const float ddq = 2 * As, uu0 = (float) u0 - u;
    for (int vt = v0; vt <= v1; ++vt) {
        const float vv = (float) vt - v;
        float q = As * uu0 * uu0 + (Bs * uu0 + Cs * vv) * vv;
        float dq = As * (2 * uu0 + 1) + Bs * vv;
        for (int ut = u0; ut <= u1; ++ut) {
            if (q < 64.0f) {
                const uint32_t qi = (uint32_t) q;
                if (qi < 64u) {
                    const float weight = E.mip->lut[(int) q];
                    result += env_texel_level(E, level, ut, vt) * weight;
                    denominator += weight;
                }
            }
            q += dq; dq += ddq;
        }
    }
    if (denominator == 0) return env_bilinear_level(E, level, uvx, uvy);
    return result / denominator;
}
// evalEnvironment for camera rays: MIPMap::eval with the EWA filter type and maxAnisotropy 10 (mipmap.h:629-725, envmap.cpp:150-152,391-407)
CP_D V3 env_eval_filtered(const EnvDev &E, const V3 &d, const V3 &rxDir, const V3 &ryDir, unsigned long long *unsupported) {
    V3 v = mul3(E.toLocal, d);
    float uvx = nc_atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
    V3 dvdx = mul3(E.toLocal, rxDir) - v, dvdy = mul3(E.toLocal, ryDir) - v;
    float t1 = kInvTwoPi / (v.x * v.x + v.z * v.z), t2 = -kInvPi / fmaxf(safe_sqrt(1.0f - v.y * v.y), kEpsilon);
    float du0 = t1 * (dvdx.z * v.x - dvdx.x * v.z) * E.w, dv0 = t2 * dvdx.y * E.h;
    float du1 = t1 * (dvdy.z * v.x - dvdy.x * v.z) * E.w, dv1 = t2 * dvdy.y * E.h;
    float A = dv0 * dv0 + dv1 * dv1, B = -2.0f * (du0 * dv0 + du1 * dv1), C = du0 * du0 + du1 * du1, F = A * C - B * B * 0.25f;
    float root = mfd_hypot2(A - C, B), Aprime = 0.5f * (A + C - root), Cprime = 0.5f * (A + C + root);
    float majorRadius = Aprime != 0 ? sqrtf(F / Aprime) : 0, minorRadius = Cprime != 0 ? sqrtf(F / Cprime) : 0;
    if (!(minorRadius > 0) || !(majorRadius > 0) || F < 0) {
        // degenerate footprint: trilinear interpolation, preferring blurring over aliasing (:659-674)
        const float level = env_log2(fmaxf(majorRadius, kEpsilon));
        const int ilevel = (int) floorf(level);
        if (ilevel < 0) return env_bilinear(E, uvx, uvy) * E.scale;
        if (!E.mip) { if (unsupported) atomicAdd(unsupported, 1ull); return env_bilinear(E, uvx, uvy) * E.scale; }
        const float a = level - ilevel;
        return (env_bilinear_level(E, ilevel, uvx, uvy) * (1.0f - a) + env_bilinear_level(E, ilevel + 1, uvx, uvy) * a) * E.scale;
    }
    const float maxAnisotropy = 10.0f;
    if (minorRadius * maxAnisotropy < majorRadius) {          // too skinny: enlarge the minor radius (:678-701)
        minorRadius = majorRadius / maxAnisotropy;
        const float theta = 0.5f * (float) atan((double) (B / (A - C)));
        float sinTheta, cosTheta;
        nc_sincos(theta, &sinTheta, &cosTheta);
        const float a2 = majorRadius * majorRadius, b2 = minorRadius * minorRadius, sinTheta2 = sinTheta * sinTheta, cosTheta2 = cosTheta * cosTheta,
                    sin2Theta = 2 * sinTheta * cosTheta;
        A = a2 * cosTheta2 + b2 * sinTheta2; B = (a2 - b2) * sin2Theta; C = a2 * sinTheta2 + b2 * cosTheta2; F = a2 * b2;
    }
    const float scale = 1.0f / F;
    A *= scale; B *= scale; C *= scale;
    const float level = fmaxf(0.0f, env_log2(minorRadius));
    const int ilevel = (int) level;
    const float a = level - ilevel;
    if (majorRadius < 1 || !(A > 0 && C > 0)) {
        if (ilevel == 0 || !E.mip) { if (ilevel != 0 && unsupported) atomicAdd(unsupported, 1ull); return env_bilinear(E, uvx, uvy) * E.scale; }
        return env_bilinear_level(E, ilevel, uvx, uvy) * E.scale;
    }
    if (!E.mip) { if (unsupported) atomicAdd(unsupported, 1ull); return env_bilinear(E, uvx, uvy) * E.scale; }
    return (env_ewa_level(E, ilevel, uvx, uvy, A, B, C) * (1.0f - a) + env_ewa_level(E, ilevel + 1, uvx, uvy, A, B, C) * a) * E.scale;
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
    nc_sincos(E.pixelSizeX * (posx + 0.5f), &sinPhi, &cosPhi);
    nc_sincos(E.pixelSizeY * (posy + 0.5f), &sinTheta, &cosTheta);
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
    float uvx = nc_atan2(d.x, -d.z) * kInvTwoPi, uvy = safe_acos(d.y) * kInvPi;
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

// oracle/o_env.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Environment emitter of the hair scenes:
//   EnvironmentMap  src/emitters/envmap.cpp:260-329 (CDFs), :358-374, :380-410 (eval), :516-662 (sampling/pdf)
//   MIPMap  include/mitsuba/render/mipmap.h:180-302 (pyramid: progressive 2-lobed-Lanczos downsampling, EWA weight table),
//           :503-596 (texel / box / bilinear lookups), :629-725 (eval: ellipse from the uv Jacobian, anisotropy clamp, level choice),
//           :760-836 (evalEWA); Resampler src: include/mitsuba/core/rfilter.h:122-170,216-262,436-457; src/libcore/bitmap.cpp:2230-2328;
//           src/rfilters/lanczos.cpp:42-55; math::log2 / hypot2 src/libcore/math.cpp:74-86,103-106
//   SunSkyEmitter bake  src/emitters/sunsky.cpp:100-236, src/emitters/sky.cpp:219-256,413-447,
//                       src/emitters/sunsky/sunmodel.h:90-105,206-222,316-371
// The Hosek-Wilkie sky model itself is NOT restated: the oracle calls the reference's own
// src/emitters/sunsky/skymodel.cpp, compiled into oracle/_ref/libref_pieces.so (see oracle/Makefile).
#pragma once
#include "o_math.h"
#include <cstdio>
#include <string>
#include <stdexcept>
#include "o_hair.h"

namespace orc {

struct BSphere { V3 center; float radius = 0; };
// include/mitsuba/core/bsphere.h:88-95
static inline bool bsphereRayIntersect(const BSphere &s, const V3 &ro, const V3 &rd, float &nearHit, float &farHit) {
    V3 o = ro - s.center;
    float A = dot(rd, rd), B = 2 * dot(o, rd), C = dot(o, o) - s.radius * s.radius;
    return solveQuadratic(A, B, C, nearHit, farHit);
}

struct EnvMap {
    int w = 0, h = 0;
    std::vector<uint16_t> texels; // half RGB (envmap.cpp:102-103)
    std::vector<float> cdfCols, cdfRows, rowWeights;
    float normalization = 0, scale = 1;
    float pixelSizeX = 0, pixelSizeY = 0;
    M44 toWorld = M44::identity(), toLocal = M44::identity();
    BSphere sceneBSphere;
    mutable long unsupportedFiltered = 0; // kept for the statistics interface: always 0 since the pyramid exists
    // MIP pyramid: level 0 = `texels`; levels 1.. half-quantised like level 0 (TMIPMap<Spectrum, SpectrumHalf>)
    struct Level { int w = 0, h = 0; std::vector<uint16_t> t; };
    std::vector<Level> levels;            // levels[0] is left empty (w, h only)
    float weightLut[64];

    // LanczosSincFilter::eval with 2 lobes (lanczos.cpp:42-55)
    static float lanczos(float x) {
        const float radius = 2.0f;
        x = std::abs(x);
        if (x < kEpsilon) return 1.0f;
        else if (x > radius) return 0.0f;
        float x1 = kPi * x, x2 = x1 / radius;
        return (cr::sin(x1) * cr::sin(x2)) / (x1 * x2);
    }
    // Resampler<float>(rfilter, bc, sourceRes, targetRes) for targetRes < sourceRes + resampleAndClamp(min 0, max inf) over one line
    static void resampleLine(const float *source, size_t sourceStride, int sourceRes, float *target, size_t targetStride, int targetRes, bool repeat) {
        const float scale = (float) sourceRes / (float) targetRes, invScale = 1 / scale, filterRadius = 2.0f * scale;
        const int taps = ceilToInt(filterRadius * 2);
        std::vector<float> weights(taps);
        for (int i = 0; i < targetRes; i++) {
            float center = (i + 0.5f) / targetRes * sourceRes;
            int start = floorToInt(center - filterRadius + 0.5f);
            float sum = 0;
            for (int j = 0; j < taps; j++) {
                float pos = start + j + 0.5f - center;
                float weight = lanczos(pos * invScale);
                weights[j] = weight; sum += weight;
            }
            float normalization = 1.0f / sum;
            for (int j = 0; j < taps; j++) weights[j] = weights[j] * normalization;
            for (int ch = 0; ch < 3; ++ch) {
                float result = 0;
                for (int j = 0; j < taps; ++j) {
                    int pos = start + j;
                    if (pos < 0 || pos >= sourceRes) pos = repeat ? modulo(pos, sourceRes) : clampi(pos, 0, sourceRes - 1);
                    result += source[sourceStride * 3 * (size_t) pos + ch] * weights[j];
                }
                target[targetStride * 3 * (size_t) i + ch] = std::min(kInf, std::max(0.0f, result));
            }
        }
    }
    // mipmap.h:245-271 + 296-302: every level from the fp32 image of the level above it
    void buildPyramid(std::vector<float> img) {
        levels.clear(); levels.push_back(Level()); levels[0].w = w; levels[0].h = h;
        int sx = w, sy = h;
        while (sx > 1 || sy > 1) {
            const int tx = std::max(1, (sx + 1) / 2), ty = std::max(1, (sy + 1) / 2);
            std::vector<float> tmp, out((size_t) 3 * tx * ty);
            const std::vector<float> *src = &img;
            if (sx != tx) {            // horizontal pass, ERepeat
                tmp.resize((size_t) 3 * tx * sy);
                for (int y = 0; y < sy; ++y) resampleLine(img.data() + (size_t) 3 * y * sx, 1, sx, tmp.data() + (size_t) 3 * y * tx, 1, tx, true);
                src = &tmp;
            }
            if (sy != ty) {            // vertical pass, EClamp
                for (int x = 0; x < tx; ++x) resampleLine(src->data() + (size_t) 3 * x, (size_t) tx, sy, out.data() + (size_t) 3 * x, (size_t) tx, ty, false);
            } else out = *src;
            Level L; L.w = tx; L.h = ty; L.t.resize(out.size());
            for (size_t i = 0; i < out.size(); ++i) L.t[i] = float_to_half(out[i]);
            levels.push_back(std::move(L));
            img.swap(out); sx = tx; sy = ty;
        }
        for (int i = 0; i < 64; ++i) { float r2 = (float) i / (float) (64 - 1); weightLut[i] = cr::exp(-2.0f * r2) - cr::exp(-2.0f); }
    }
    V3 evalTexel(int level, int x, int y) const {
        if (level == 0) return evalTexel(x, y);
        const Level &L = levels[level];
        if (x < 0 || x >= L.w) x = modulo(x, L.w);
        if (y < 0 || y >= L.h) y = clampi(y, 0, L.h - 1);
        const uint16_t *p = &L.t[3 * ((size_t) y * L.w + x)];
        return V3(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]));
    }
    V3 evalBox(int level, float uvx, float uvy) const { return evalTexel(level, floorToInt(uvx * levels[level].w), floorToInt(uvy * levels[level].h)); }
    V3 evalBilinear(int level, float uvx, float uvy) const {
        if (!std::isfinite(uvx) || !std::isfinite(uvy)) return V3(0.0f);
        if (level >= (int) levels.size()) return evalBox((int) levels.size() - 1, uvx, uvy);
        const int lw = levels[level].w, lh = levels[level].h;
        float u = uvx * lw - 0.5f, v = uvy * lh - 0.5f;
        int xPos = floorToInt(u), yPos = floorToInt(v);
        float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
        return evalTexel(level, xPos, yPos) * dx2 * dy2 + evalTexel(level, xPos, yPos + 1) * dx2 * dy1
             + evalTexel(level, xPos + 1, yPos) * dx1 * dy2 + evalTexel(level, xPos + 1, yPos + 1) * dx1 * dy1;
    }
    // mipmap.h:760-836
    V3 evalEWA(int level, float uvx, float uvy, float A, float B, float C) const {
        if (!std::isfinite(A + B + C + uvx + uvy)) return V3(0.0f);
        if (level >= (int) levels.size()) return evalBox((int) levels.size() - 1, uvx, uvy);
        const int lw = levels[level].w, lh = levels[level].h;
        float u = uvx * lw - 0.5f, v = uvy * lh - 0.5f;
        const float ratioX = (float) lw / (float) w, ratioY = (float) lh / (float) h;
        A /= ratioX * ratioX; B /= ratioX * ratioY; C /= ratioY * ratioY;
        float invDet = 1.0f / (-B * B + 4.0f * A * C), deltaU = 2.0f * std::sqrt(C * invDet), deltaV = 2.0f * std::sqrt(A * invDet);
        int u0 = ceilToInt(u - deltaU), u1 = floorToInt(u + deltaU), v0 = ceilToInt(v - deltaV), v1 = floorToInt(v + deltaV);
        float As = A * 64, Bs = B * 64, Cs = C * 64;
        V3 result(0.0f); float denominator = 0.0f;
        float ddq = 2 * As, uu0 = (float) u0 - u;
        for (int vt = v0; vt <= v1; ++vt) {
            const float vv = (float) vt - v;
            float q = As * uu0 * uu0 + (Bs * uu0 + Cs * vv) * vv;
            float dq = As * (2 * uu0 + 1) + Bs * vv;
            for (int ut = u0; ut <= u1; ++ut) {
                if (q < 64.0f) {
                    uint32_t qi = (uint32_t) q;
                    if (qi < 64) { const float weight = weightLut[(int) q]; result = result + evalTexel(level, ut, vt) * weight; denominator += weight; }
                }
                q += dq; dq += ddq;
            }
        }
        if (denominator == 0) return evalBilinear(level, uvx, uvy);
        return result / denominator;
    }

    V3 texel(int x, int y) const { const uint16_t *p = &texels[3 * ((size_t) y * w + x)]; return V3(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2])); }
    // mipmap.h:503-560 with bcu=ERepeat, bcv=EClamp (envmap.cpp:167-170)
    V3 evalTexel(int x, int y) const {
        if (x < 0 || x >= w) x = modulo(x, w);
        if (y < 0 || y >= h) y = clampi(y, 0, h - 1);
        return texel(x, y);
    }
    // mipmap.h:572-596
    V3 evalBilinear(float uvx, float uvy) const {
        if (!std::isfinite(uvx) || !std::isfinite(uvy)) return V3(0.0f);
        float u = uvx * w - 0.5f, v = uvy * h - 0.5f;
        int xPos = floorToInt(u), yPos = floorToInt(v);
        float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
        return evalTexel(xPos, yPos) * dx2 * dy2 + evalTexel(xPos, yPos + 1) * dx2 * dy1
             + evalTexel(xPos + 1, yPos) * dx1 * dy2 + evalTexel(xPos + 1, yPos + 1) * dx1 * dy1;
    }

    // ctor (fp32 bitmap -> half) + configure() envmap.cpp:260-329
    void init(const float *rgb, int w_, int h_, const M44 &toWorld_, float scale_) {
        w = w_; h = h_; scale = scale_; toWorld = toWorld_;
        if (!invert(toWorld, toLocal)) throw std::runtime_error("oracle: singular envmap transform");
        texels.resize((size_t) 3 * w * h);
        for (size_t i = 0; i < texels.size(); ++i) texels[i] = float_to_half(std::max(rgb[i], 0.0f)); // mipmap.h:232-240 clamps negatives
        { std::vector<float> base(rgb, rgb + (size_t) 3 * w * h); for (float &v : base) v = std::max(v, 0.0f); buildPyramid(std::move(base)); }
        cdfCols.assign((size_t) (w + 1) * h, 0.0f); cdfRows.assign(h + 1, 0.0f); rowWeights.assign(h, 0.0f);
        size_t colPos = 0, rowPos = 0;
        float rowSum = 0.0f;
        cdfRows[rowPos++] = 0;
        for (int y = 0; y < h; ++y) {
            float colSum = 0;
            cdfCols[colPos++] = 0;
            for (int x = 0; x < w; ++x) {
                colSum += luminance(texel(x, y));
                cdfCols[colPos++] = colSum;
            }
            float norm = 1.0f / colSum;
            for (int x = 1; x < w; ++x) cdfCols[colPos - x - 1] *= norm;
            cdfCols[colPos - 1] = 1.0f;
            float weight = cr::sin((y + 0.5f) * kPi / h);
            rowWeights[y] = weight;
            rowSum += colSum * weight;
            cdfRows[rowPos++] = rowSum;
        }
        float norm = 1.0f / rowSum;
        for (int y = 1; y < h; ++y) cdfRows[rowPos - y - 1] *= norm;
        cdfRows[rowPos - 1] = 1.0f;
        if (rowSum == 0 || !std::isfinite(rowSum)) throw std::runtime_error("oracle: invalid environment map");
        normalization = 1.0f / (rowSum * (2 * kPi / w) * (kPi / h));
        pixelSizeX = 2 * kPi / w; pixelSizeY = kPi / h;
    }
    // envmap.cpp:331-341 (scene bounding sphere x1.5)
    void setSceneBounds(const AABB &sceneAABB) {
        V3 c = sceneAABB.center();
        sceneBSphere.center = c;
        sceneBSphere.radius = std::max(kEpsilon, length(c - sceneAABB.mx) * 1.5f);
    }

    // envmap.cpp:380-410.  hasDiff: camera rays carry differentials (rx/ry directions)
    V3 evalEnvironment(const V3 &d, bool hasDiff = false, const V3 &rxDir = V3(), const V3 &ryDir = V3()) const {
        V3 v = xfmVector(toLocal, d);
        float uvx = cr::atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
        V3 value;
        if (!hasDiff) {
            value = evalBilinear(uvx, uvy);
        } else {
            V3 dvdx = xfmVector(toLocal, rxDir) - v, dvdy = xfmVector(toLocal, ryDir) - v;
            float t1 = kInvTwoPi / (v.x * v.x + v.z * v.z),
                  t2 = -kInvPi / std::max(safe_sqrt(1.0f - v.y * v.y), kEpsilon);
            float d0x = t1 * (dvdx.z * v.x - dvdx.x * v.z), d0y = t2 * dvdx.y;
            float d1x = t1 * (dvdy.z * v.x - dvdy.x * v.z), d1y = t2 * dvdy.y;
            value = evalFiltered(uvx, uvy, d0x, d0y, d1x, d1y);
        }
        return value * scale;
    }
    // MIPMap::eval, mipmap.h:629-725 (EWA filter type, maxAnisotropy 10)
    static float log2f32(float x) { const float invLn2 = 1.0f / cr::log(2.0f); return cr::log(x) * invLn2; }       // math.cpp:103-106
    static float hypot2(float a, float b) {                                                                         // math.cpp:74-86
        float r;
        if (std::abs(a) > std::abs(b)) { r = b / a; r = std::abs(a) * std::sqrt(1.0f + r * r); }
        else if (b != 0.0f) { r = a / b; r = std::abs(b) * std::sqrt(1.0f + r * r); }
        else r = 0.0f;
        return r;
    }
    V3 evalFiltered(float uvx, float uvy, float d0x, float d0y, float d1x, float d1y) const {
        float du0 = d0x * w, dv0 = d0y * h, du1 = d1x * w, dv1 = d1y * h;
        float A = dv0 * dv0 + dv1 * dv1, B = -2.0f * (du0 * dv0 + du1 * dv1), C = du0 * du0 + du1 * du1, F = A * C - B * B * 0.25f;
        float root = hypot2(A - C, B), Aprime = 0.5f * (A + C - root), Cprime = 0.5f * (A + C + root);
        float majorRadius = Aprime != 0 ? std::sqrt(F / Aprime) : 0, minorRadius = Cprime != 0 ? std::sqrt(F / Cprime) : 0;
        if (!(minorRadius > 0) || !(majorRadius > 0) || F < 0) {
            float level = log2f32(std::max(majorRadius, kEpsilon));
            int ilevel = floorToInt(level);
            if (ilevel < 0) return evalBilinear(0, uvx, uvy);
            float a = level - ilevel;
            return evalBilinear(ilevel, uvx, uvy) * (1.0f - a) + evalBilinear(ilevel + 1, uvx, uvy) * a;
        }
        const float maxAnisotropy = 10.0f;
        if (minorRadius * maxAnisotropy < majorRadius) {
            minorRadius = majorRadius / maxAnisotropy;
            float theta = 0.5f * (float) std::atan((double) (B / (A - C))), sinTheta, cosTheta;
            cr::sincos(theta, &sinTheta, &cosTheta);
            float a2 = majorRadius * majorRadius, b2 = minorRadius * minorRadius, sinTheta2 = sinTheta * sinTheta, cosTheta2 = cosTheta * cosTheta,
                  sin2Theta = 2 * sinTheta * cosTheta;
            A = a2 * cosTheta2 + b2 * sinTheta2; B = (a2 - b2) * sin2Theta; C = a2 * sinTheta2 + b2 * cosTheta2; F = a2 * b2;
        }
        float scaleF = 1.0f / F;
        A *= scaleF; B *= scaleF; C *= scaleF;
        float level = std::max(0.0f, log2f32(minorRadius));
        int ilevel = (int) level;
        float a = level - ilevel;
        if (majorRadius < 1 || !(A > 0 && C > 0)) return evalBilinear(ilevel, uvx, uvy);
        return evalEWA(ilevel, uvx, uvy, A, B, C) * (1.0f - a) + evalEWA(ilevel + 1, uvx, uvy, A, B, C) * a;
    }

    // envmap.cpp:657-662
    static uint32_t sampleReuse(const float *cdf, uint32_t size, float &sample) {
        const float *entry = std::lower_bound(cdf, cdf + size + 1, sample);
        uint32_t index = std::min((uint32_t) std::max((ptrdiff_t) 0, entry - cdf - 1), size - 1);
        sample = (sample - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
    // envmap.cpp:567-602
    void internalSampleDirection(float sx, float sy, V3 &d, V3 &value, float &pdf) const {
        uint32_t row = sampleReuse(cdfRows.data(), h, sy);
        uint32_t col = sampleReuse(cdfCols.data() + (size_t) row * (w + 1), w, sx);
        float posx = (float) col + intervalToTent(sx), posy = (float) row + intervalToTent(sy);
        int xPos = floorToInt(posx), yPos = floorToInt(posy);
        float dx1 = posx - xPos, dx2 = 1.0f - dx1, dy1 = posy - yPos, dy2 = 1.0f - dy1;
        V3 value1 = evalTexel(xPos, yPos) * dx2 * dy2 + evalTexel(xPos + 1, yPos) * dx1 * dy2;
        V3 value2 = evalTexel(xPos, yPos + 1) * dx2 * dy1 + evalTexel(xPos + 1, yPos + 1) * dx1 * dy1;
        value = (value1 + value2) * scale;
        pdf = (luminance(value1) * rowWeights[clampi(yPos, 0, h - 1)] + luminance(value2) * rowWeights[clampi(yPos + 1, 0, h - 1)]) * normalization;
        float phi = pixelSizeX * (posx + 0.5f), theta = pixelSizeY * (posy + 0.5f);
        float sinPhi = cr::sin(phi), cosPhi = cr::cos(phi), sinTheta = cr::sin(theta), cosTheta = cr::cos(theta);
        d = V3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
        pdf /= std::max(std::abs(sinTheta), kEpsilon);
    }
    // envmap.cpp:603-635
    float internalPdfDirection(const V3 &d) const {
        float uvx = cr::atan2(d.x, -d.z) * kInvTwoPi, uvy = safe_acos(d.y) * kInvPi;
        if (!std::isfinite(uvx) || !std::isfinite(uvy)) return 0.0f;
        float u = uvx * w - 0.5f, v = uvy * h - 0.5f;
        int xPos = floorToInt(u), yPos = floorToInt(v);
        float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
        V3 value1 = evalTexel(xPos, yPos) * dx2 * dy2 + evalTexel(xPos + 1, yPos) * dx1 * dy2;
        V3 value2 = evalTexel(xPos, yPos + 1) * dx2 * dy1 + evalTexel(xPos + 1, yPos + 1) * dx1 * dy1;
        float sinTheta = safe_sqrt(1 - d.y * d.y);
        return (luminance(value1) * rowWeights[clampi(yPos, 0, h - 1)] + luminance(value2) * rowWeights[clampi(yPos + 1, 0, h - 1)])
               * normalization / std::max(std::abs(sinTheta), kEpsilon);
    }
    struct DirectSample { V3 value; V3 d; float dist = 0; float pdf = 0; };
    // envmap.cpp:516-543 (returns value/pdf; pdf==0 means "no sample")
    DirectSample sampleDirect(const V3 &ref, float sx, float sy) const {
        DirectSample r;
        V3 value, d; float pdf;
        internalSampleDirection(sx, sy, d, value, pdf);
        V3 dw = xfmVector(toWorld, d);
        float nearT, farT;
        if (isZero(value) || pdf == 0 || !bsphereRayIntersect(sceneBSphere, ref, dw, nearT, farT) || nearT >= 0 || farT <= 0) {
            r.pdf = 0; r.value = V3(0.0f); return r;
        }
        r.pdf = pdf; r.dist = farT; r.d = dw; r.value = value / pdf;
        return r;
    }
    float pdfDirect(const V3 &dWorld) const { return internalPdfDirection(xfmVector(toLocal, dWorld)); } // envmap.cpp:545-556 (ESolidAngle)
    // envmap.cpp:358-374: false => the integrator drops the sample (path.cpp:240-241)
    bool fillDirectSamplingRecord(const V3 &o, const V3 &d) const {
        float nearT, farT;
        if (!bsphereRayIntersect(sceneBSphere, o, d, nearT, farT) || nearT > 0 || farT < 0) return false;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// sunsky bake (host-side setup in the reference).  Needs oracle/_ref/libref_pieces.so.
// ---------------------------------------------------------------------------------------------
struct RefPieces {
    void *(*sky_alloc)(double, double, double) = nullptr;        // arhosek_rgb_skymodelstate_alloc_init
    double (*sky_radiance)(void *, double, double, int) = nullptr; // arhosek_tristim_skymodel_radiance
    void (*sky_free)(void *) = nullptr;
    const float *cie_wl = nullptr, *cie_x = nullptr, *cie_y = nullptr, *cie_z = nullptr; int cie_n = 0;
};

// piecewise-linear spectrum (src/libcore/spectrum.cpp:687-714 InterpolatedSpectrum::eval)
struct InterpSpectrum {
    std::vector<float> wl, val;
    InterpSpectrum(const float *w, const float *v, size_t n) : wl(w, w + n), val(v, v + n) {}
    float eval(float lambda) const {
        if (wl.size() < 2 || lambda < wl.front() || lambda > wl.back()) return 0.0f;
        auto r = std::equal_range(wl.begin(), wl.end(), lambda);
        size_t idx1 = r.first - wl.begin(), idx2 = r.second - wl.begin();
        if (idx1 == idx2) {
            float a = wl[idx1 - 1], b = wl[idx1], fa = val[idx1 - 1], fb = val[idx1];
            float t = (lambda - a) / (b - a);
            return (1.0f - t) * fa + t * fb; // math::lerp
        }
        return val[idx1];
    }
};
// Average of the product of two piecewise-linear spectra over [a,b].  The reference integrates with an
// adaptive Gauss-Lobatto rule to 1e-4 relative accuracy (spectrum.cpp:546-568); the oracle integrates
// the piecewise-quadratic product exactly (Simpson per linear piece), i.e. agrees within that tolerance.
static inline double averageProduct(const InterpSpectrum &f, const InterpSpectrum &g, float a, float b) {
    std::vector<float> knots;
    for (float x : f.wl) if (x > a && x < b) knots.push_back(x);
    for (float x : g.wl) if (x > a && x < b) knots.push_back(x);
    knots.push_back(a); knots.push_back(b);
    std::sort(knots.begin(), knots.end());
    knots.erase(std::unique(knots.begin(), knots.end()), knots.end());
    double sum = 0;
    auto ev = [&](const InterpSpectrum &s, double x, double lo, double hi) { // linear within [lo,hi], one-sided at knots
        double flo = s.eval((float) lo), fhi = s.eval((float) hi);
        // handle support edges: eval() returns 0 outside [front, back]; inside a piece it is linear
        return flo + (fhi - flo) * (x - lo) / (hi - lo);
    };
    for (size_t i = 0; i + 1 < knots.size(); ++i) {
        double lo = knots[i], hi = knots[i + 1], mid = 0.5 * (lo + hi);
        double p0 = (double) f.eval((float) lo) * g.eval((float) lo), p2 = (double) f.eval((float) hi) * g.eval((float) hi);
        double p1 = ev(f, mid, lo, hi) * ev(g, mid, lo, hi);
        sum += (hi - lo) / 6.0 * (p0 + 4 * p1 + p2);
    }
    return sum / ((double) b - a);
}

struct SunSkyParams {
    float turbidity = 3.0f, albedo = 0.2f, skyScale = 1.0f, sunScale = 1.0f, sunRadiusScale = 1.0f, stretch = 1.0f;
    V3 sunDirection = V3(0, 1, 0);
    int resolution = 512;
};

// sunmodel.h:316-371 (tables :260-314 are physical constants from Preetham et al.)
static inline V3 computeSunRadiance(float theta, float turbidity, const RefPieces &ref) {
    static const float k_oWavelengths[64] = {300,305,310,315,320,325,330,335,340,345,350,355,445,450,455,460,465,470,475,480,485,490,495,500,505,510,515,520,525,530,535,540,545,550,555,560,565,570,575,580,585,590,595,600,605,610,620,630,640,650,660,670,680,690,700,710,720,730,740,750,760,770,780,790};
    static const float k_oAmplitudes[65] = {10.0f,4.8f,2.7f,1.35f,.8f,.380f,.160f,.075f,.04f,.019f,.007f,.0f,.003f,.003f,.004f,.006f,.008f,.009f,.012f,.014f,.017f,.021f,.025f,.03f,.035f,.04f,.045f,.048f,.057f,.063f,.07f,.075f,.08f,.085f,.095f,.103f,.110f,.12f,.122f,.12f,.118f,.115f,.12f,.125f,.130f,.12f,.105f,.09f,.079f,.067f,.057f,.048f,.036f,.028f,.023f,.018f,.014f,.011f,.010f,.009f,.007f,.004f,.0f,.0f};
    static const float k_gWavelengths[4] = {759,760,770,771};
    static const float k_gAmplitudes[4] = {0,3.0f,0.210f,0};
    static const float k_waWavelengths[13] = {689,690,700,710,720,730,740,750,760,770,780,790,800};
    static const float k_waAmplitudes[13] = {0,0.160e-1f,0.240e-1f,0.125e-1f,0.100e+1f,0.870f,0.610e-1f,0.100e-2f,0.100e-4f,0.100e-4f,0.600e-3f,0.175e-1f,0.360e-1f};
    static const float solWavelengths[38] = {380,390,400,410,420,430,440,450,460,470,480,490,500,510,520,530,540,550,560,570,580,590,600,610,620,630,640,650,660,670,680,690,700,710,720,730,740,750};
    static const float solAmplitudes[38] = {16559.0f,16233.7f,21127.5f,25888.2f,25829.1f,24232.3f,26760.5f,29658.3f,30545.4f,30057.5f,30663.7f,28830.4f,28712.1f,27825.0f,27100.6f,27233.6f,26361.3f,25503.8f,25060.2f,25311.6f,25355.9f,25134.2f,24631.5f,24173.2f,23685.3f,23212.1f,22827.7f,22339.8f,21970.2f,21526.7f,21097.9f,20728.3f,20240.4f,19870.8f,19427.2f,19072.4f,18628.9f,18259.2f};
    InterpSpectrum k_oCurve(k_oWavelengths, k_oAmplitudes, 64), k_gCurve(k_gWavelengths, k_gAmplitudes, 4),
                   k_waCurve(k_waWavelengths, k_waAmplitudes, 13), solCurve(solWavelengths, solAmplitudes, 38);
    float data[91], wavelengths[91];
    float beta = 0.04608365822050f * turbidity - 0.04586025928522f;
    float m = 1.0f / (cr::cos(theta) + 0.15f * cr::pow(93.885f - theta / kPi * 180.0f, -1.253f));
    float lambda = 350;
    for (int i = 0; i < 91; i++, lambda += 5) {
        float tauR = cr::exp(-m * 0.008735f * cr::pow(lambda / 1000.0f, (float) -4.08));
        const float alpha = 1.3f;
        float tauA = cr::exp(-m * beta * cr::pow(lambda / 1000.0f, -alpha));
        const float lOzone = .35f;
        float tauO = cr::exp(-m * k_oCurve.eval(lambda) * lOzone);
        float tauG = cr::exp(-1.41f * k_gCurve.eval(lambda) * m / cr::pow(1 + 118.93f * k_gCurve.eval(lambda) * m, 0.45f));
        const float w = 2.0;
        float tauWA = cr::exp(-0.2385f * k_waCurve.eval(lambda) * w * m / cr::pow(1 + 20.07f * k_waCurve.eval(lambda) * w * m, 0.45f));
        data[i] = solCurve.eval(lambda) * tauR * tauA * tauO * tauG * tauWA;
        wavelengths[i] = lambda;
    }
    // Spectrum::fromContinuousSpectrum (src/libcore/spectrum.cpp:172-185) + fromXYZ (:222-227)
    InterpSpectrum interpolated(wavelengths, data, 91);
    InterpSpectrum cx(ref.cie_wl, ref.cie_x, ref.cie_n), cy(ref.cie_wl, ref.cie_y, ref.cie_n), cz(ref.cie_wl, ref.cie_z, ref.cie_n);
    float start = ref.cie_wl[0], end = ref.cie_wl[ref.cie_n - 1];
    std::vector<float> ones(ref.cie_n, 1.0f);
    InterpSpectrum unit(ref.cie_wl, ones.data(), ref.cie_n);
    float X = (float) averageProduct(interpolated, cx, start, end);
    float Y = (float) averageProduct(interpolated, cy, start, end);
    float Z = (float) averageProduct(interpolated, cz, start, end);
    float normalization = 1.0f / (float) averageProduct(unit, cy, start, end);
    X *= normalization; Y *= normalization; Z *= normalization;
    V3 rgb(3.240479f * X + -1.537150f * Y + -0.498535f * Z,
           -0.969256f * X + 1.875991f * Y + 0.041556f * Z,
           0.055648f * X + -0.204043f * Y + 1.057311f * Z);
    return V3(std::max(rgb.x, 0.0f), std::max(rgb.y, 0.0f), std::max(rgb.z, 0.0f));
}

// sunsky.cpp:100-216 -> fp32 RGB lat-long bitmap (resolution x resolution/2), toWorld = identity
static inline void bakeSunSky(const SunSkyParams &P, const RefPieces &ref, std::vector<float> &rgb, int &W, int &H) {
    W = P.resolution; H = P.resolution / 2;
    rgb.assign((size_t) 3 * W * H, 0.0f);
    // sunmodel.h:98-105,206-208 (fromSphere of the normalised direction)
    V3 sd = normalize(P.sunDirection);
    float sunAzimuth = cr::atan2(sd.x, -sd.z), sunElevation = safe_acos(sd.y);
    if (sunAzimuth < 0) sunAzimuth += 2 * kPi;
    // sky.cpp:219-256
    float sunElev = 0.5f * kPi - sunElevation;
    if (sunElev < 0) throw std::runtime_error("The sun is below the horizon -- this is not supported by the sky model.");
    void *state[3];
    for (int i = 0; i < 3; ++i) state[i] = ref.sky_alloc(P.turbidity, P.albedo, sunElev);
    float factorX = (2 * kPi) / W, factorY = kPi / H;
    for (int y = 0; y < H; ++y) {
        float theta0 = (y + .5f) * factorY;
        for (int x = 0; x < W; ++x) {
            float phi0 = (x + .5f) * factorX;
            // sunsky.cpp:140-145: ray direction toSphere(theta,phi); sky.cpp:392 fromSphere() of it
            float st = cr::sin(theta0), ct = cr::cos(theta0), sp = cr::sin(phi0), cp = cr::cos(phi0);
            V3 d(sp * st, ct, -cp * st);
            float azimuth = cr::atan2(d.x, -d.z), elevation = safe_acos(d.y);
            if (azimuth < 0) azimuth += 2 * kPi;
            // sky.cpp:413-447 getSkyRadiance
            float theta = elevation / P.stretch;
            V3 result(0.0f);
            if (!(cr::cos(theta) <= 0)) {
                float cosGamma = cr::cos(theta) * cr::cos(sunElevation) + cr::sin(theta) * cr::sin(sunElevation) * cr::cos(azimuth - sunAzimuth);
                float gamma = safe_acos(cosGamma);
                for (int i = 0; i < 3; i++)
                    result[i] = std::max((float) (ref.sky_radiance(state[i], theta, gamma, i) / 106.856980), 0.0f);
                result = result * P.skyScale;
            }
            float *t = &rgb[3 * ((size_t) y * W + x)];
            t[0] = result.x; t[1] = result.y; t[2] = result.z;
        }
    }
    for (int i = 0; i < 3; ++i) ref.sky_free(state[i]);
    // sun disc: sunsky.cpp:163-216
    V3 sunRadiance = computeSunRadiance(sunElevation, P.turbidity, ref) * P.sunScale;
    float sElev = sunElevation * P.stretch;
    float sst = cr::sin(sElev), sct = cr::cos(sElev), ssp = cr::sin(sunAzimuth), scp = cr::cos(sunAzimuth);
    Frame sunFrame(V3(ssp * sst, sct, -scp * sst));
    float theta = (0.5358f * 0.5f) * (kPi / 180.0f); // SUN_APP_RADIUS = 0.5358
    if (P.sunRadiusScale == 0) throw std::runtime_error("oracle: sunRadiusScale=0 (directional sun) is not on this path");
    size_t pixelCount = (size_t) P.resolution * P.resolution / 2;
    float cosTheta = cr::cos(theta * P.sunRadiusScale);
    float coveredPortion = 0.5f * (1 - cosTheta);
    size_t nSamples = (size_t) std::max(100.0f, (pixelCount * coveredPortion * 1000));
    float fx = W / (2 * kPi), fy = H / kPi;
    V3 value = sunRadiance * (2 * kPi * (1 - cr::cos(theta))) * (float) (W * H) / (2 * kPi * kPi * nSamples);
    for (size_t i = 0; i < nSamples; ++i) {
        // qmc.h:43-60,82-87,115-120 sample02: (van der Corput, Sobol' dim 2)
        uint32_t n = (uint32_t) i, rv = __builtin_bswap32(n);
        rv = ((rv & 0x0f0f0f0f) << 4) | ((rv & 0xf0f0f0f0) >> 4);
        rv = ((rv & 0x33333333) << 2) | ((rv & 0xcccccccc) >> 2);
        rv = ((rv & 0x55555555) << 1) | ((rv & 0xaaaaaaaa) >> 1);
        float s0 = (float) (rv >> 8) / (float) (1U << 24);
        uint32_t scramble = 0;
        for (uint32_t v = 1U << 31, nn = n; nn != 0; nn >>= 1, v ^= v >> 1) if (nn & 1) scramble ^= v;
        float s1 = (float) scramble / (float) (1ULL << 32);
        V3 dir = sunFrame.toWorld(squareToUniformCone(cosTheta, s0, s1));
        float sinTheta = safe_sqrt(1 - dir.y * dir.y);
        float az = cr::atan2(dir.x, -dir.z), el = safe_acos(dir.y);
        if (az < 0) az += 2 * kPi;
        int px = std::min(std::max(0, (int) (az * fx)), W - 1), py = std::min(std::max(0, (int) (el * fy)), H - 1);
        V3 add = value / std::max(1e-3f, sinTheta);
        float *t = &rgb[3 * ((size_t) py * W + px)];
        t[0] += add.x; t[1] += add.y; t[2] += add.z;
    }
}

// ---------------------------------------------------------------------------------------------
// Radiance RGBE reader -- Bitmap::readRGBE (src/libcore/bitmap.cpp:3590-3678), RGBE_ToFloat (:3522-3530), flat reads (:3579-3586),
// line reads as Stream::readLine (src/libcore/stream.cpp:392-414: CR dropped, LF ends the line).  Written against a FILE* the way
// the reference walks its Stream; returns top-down RGB triples.
// ---------------------------------------------------------------------------------------------
struct RGBEImage { int w = 0, h = 0; std::vector<float> rgb; };

static inline RGBEImage readRGBE(const char *path) {
    FILE *fp = std::fopen(path, "rb");
    if (!fp) throw std::runtime_error(std::string("Environment map file \"") + path + "\" could not be found!");
    struct Closer { FILE *f; ~Closer() { std::fclose(f); } } closer{fp};
    auto readLine = [&]() {
        std::string r; int c;
        while ((c = std::fgetc(fp)) != EOF) { if (c == 10) return r; if (c != 13) r.push_back((char) c); }
        if (r.empty()) throw std::runtime_error("readRGBE(): unexpected end of file");
        return r;
    };
    auto readBytes = [&](uint8_t *dst, size_t n) { if (std::fread(dst, 1, n, fp) != n) throw std::runtime_error("readRGBE(): unexpected end of file"); };
    auto toFloat = [](const uint8_t *q, float *out) {
        if (q[3]) { float f = std::ldexp(1.0f, (int) q[3] - (128 + 8)); out[0] = q[0] * f; out[1] = q[1] * f; out[2] = q[2] * f; }
        else { out[0] = out[1] = out[2] = 0.0f; }
    };
    std::string line = readLine();
    if (line.size() < 2 || line[0] != '#' || line[1] != '?') throw std::runtime_error("readRGBE(): Invalid header!");
    RGBEImage img; bool ok = false;
    while (true) {
        line = readLine();
        if (line.compare(0, 22, "FORMAT=32-bit_rle_rgbe") == 0) ok = true;
        if (line.compare(0, 3, "-Y ") == 0) {
            if (std::sscanf(line.c_str(), "-Y %i +X %i", &img.h, &img.w) < 2) throw std::runtime_error("readRGBE(): parser error!");
            break;
        }
    }
    if (!ok) throw std::runtime_error("readRGBE(): invalid format!");
    if (img.w <= 0 || img.h <= 0) throw std::runtime_error("readRGBE(): invalid image size!");
    const size_t total = (size_t) img.w * img.h;
    img.rgb.resize(3 * total);
    size_t done = 0;                                        // pixels written so far
    auto readFlat = [&](size_t n) { uint8_t q[4]; for (size_t k = 0; k < n; ++k) { readBytes(q, 4); toFloat(q, &img.rgb[3 * done]); ++done; } };
    if (img.w < 8 || img.w > 0x7fff) { readFlat(total); return img; }
    std::vector<uint8_t> chan[4];
    for (auto &c : chan) c.resize((size_t) img.w);
    for (int y = 0; y < img.h; ++y) {
        uint8_t q[4];
        readBytes(q, 4);
        if (q[0] != 2 || q[1] != 2 || (q[2] & 0x80)) {       // a flat file: this pixel and total-1 more from the current position
            toFloat(q, &img.rgb[3 * done]); ++done;
            const size_t room = total - done;
            readFlat(std::min(room, total - 1));
            return img;
        }
        if ((((int) q[2]) << 8 | q[3]) != img.w) throw std::runtime_error("readRGBE(): wrong scanline width!");
        for (int c = 0; c < 4; ++c) {
            size_t x = 0;
            while (x < (size_t) img.w) {
                uint8_t b[2];
                readBytes(b, 2);
                size_t count = b[0] > 128 ? b[0] - 128 : b[0];
                if (count == 0 || count > (size_t) img.w - x) throw std::runtime_error("readRGBE(): bad scanline data!");
                if (b[0] > 128) { for (size_t k = 0; k < count; ++k) chan[c][x++] = b[1]; }
                else { chan[c][x++] = b[1]; if (count > 1) { readBytes(&chan[c][x], count - 1); x += count - 1; } }
            }
        }
        for (int x = 0; x < img.w; ++x) {
            const uint8_t px[4] = {chan[0][x], chan[1][x], chan[2][x], chan[3][x]};
            toFloat(px, &img.rgb[3 * done]); ++done;
        }
    }
    return img;
}

} // namespace orc

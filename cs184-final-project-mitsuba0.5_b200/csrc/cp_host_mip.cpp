// cp_host_mip.cpp -- MIP pyramid of the environment map (host side, once per scene).
//
// Replaces (reference file:line):
//   TMIPMap constructor, steps 1 and 3        include/mitsuba/render/mipmap.h:180-192, 245-271 (levels, progressive downsampling)
//   Bitmap::resample -> resample<float>        src/libcore/bitmap.cpp:2230-2328 (X pass into a temporary, then Y pass; clamped to [0, inf))
//   Resampler<float>, resampling mode          include/mitsuba/core/rfilter.h:122-170, 216-262, 436-457
//   LanczosSincFilter::eval, 2 lobes           src/rfilters/lanczos.cpp:42-55 (the filter EnvironmentMap asks for, src/emitters/envmap.cpp:160-166)
//   EWA weight table                           mipmap.h:296-302 (math::fastexp is the fp64 exp rounded once on Linux/x86-64, math.h:175-199)
// Each level is resampled from the fp32 image of the level before (not from its half-quantised texels), boundary conditions
// ERepeat along u and EClamp along v (envmap.cpp:178-179); the device stores every level half-quantised like level 0.
#include "cp_host.h"
#include <cmath>
#include <algorithm>

namespace cp {

namespace {
float lanczos2(float x) {
    const float radius = 2.0f;
    x = std::fabs(x);
    if (x < kEpsilon) return 1.0f;
    else if (x > radius) return 0.0f;
    const float x1 = kPi * x, x2 = x1 / radius;
    return (cr_sin(x1) * cr_sin(x2)) / (x1 * x2);
}
int modulo(int a, int b) { const int r = a % b; return r < 0 ? r + b : r; }

// one Resampler (sourceRes -> targetRes < sourceRes) applied to `count` lines of `channels`-interleaved samples
struct Resampler1D {
    int sourceRes, targetRes, taps; bool repeat;
    std::vector<int> start; std::vector<float> weights;
    Resampler1D(int src, int tgt, bool repeat_) : sourceRes(src), targetRes(tgt), repeat(repeat_) {
        float filterRadius = 2.0f, scale = 1.0f, invScale = 1.0f;
        if (targetRes < sourceRes) { scale = (float) sourceRes / (float) targetRes; invScale = 1 / scale; filterRadius *= scale; }
        taps = (int) std::ceil(filterRadius * 2);
        start.resize(targetRes); weights.resize((size_t) taps * targetRes);
        for (int i = 0; i < targetRes; i++) {
            const float center = (i + 0.5f) / targetRes * sourceRes;
            start[i] = (int) std::floor(center - filterRadius + 0.5f);
            float sum = 0;
            for (int j = 0; j < taps; j++) {
                const float pos = start[i] + j + 0.5f - center;
                const float weight = lanczos2(pos * invScale);
                weights[(size_t) i * taps + j] = weight;
                sum += weight;
            }
            const float normalization = 1.0f / sum;
            for (int j = 0; j < taps; j++) weights[(size_t) i * taps + j] = weights[(size_t) i * taps + j] * normalization;
        }
    }
    // resampleAndClamp with min = 0, max = +inf; strides in samples
    void run(const float *source, size_t sourceStride, float *target, size_t targetStride, int channels) const {
        for (int i = 0; i < targetRes; ++i) {
            for (int ch = 0; ch < channels; ++ch) {
                float result = 0;
                for (int j = 0; j < taps; ++j) {
                    int pos = start[i] + j;
                    if (pos < 0 || pos >= sourceRes) pos = repeat ? modulo(pos, sourceRes) : std::min(std::max(pos, 0), sourceRes - 1);
                    result += source[sourceStride * channels * (size_t) pos + ch] * weights[(size_t) i * taps + j];
                }
                target[targetStride * channels * (size_t) i + ch] = std::min(INFINITY, std::max(0.0f, result));
            }
        }
    }
};
}

// levels[0] = the input clamped to >= 0; levels[l] = levels[l-1] downsampled to max(1, (size + 1) / 2) per axis, until 1 x 1
void build_env_pyramid(const float *rgb, int w, int h, std::vector<EnvMipLevel> &levels, float lut[64]) {
    levels.clear();
    EnvMipLevel base; base.w = w; base.h = h; base.rgb.assign(rgb, rgb + (size_t) 3 * w * h);
    for (float &v : base.rgb) v = std::max(v, 0.0f);                          // mipmap.h:232-240 (negative texels are clamped before the pyramid is built)
    levels.push_back(std::move(base));
    int sx = w, sy = h;
    while (sx > 1 || sy > 1) {
        const EnvMipLevel &src = levels.back();
        const int tx = std::max(1, (sx + 1) / 2), ty = std::max(1, (sy + 1) / 2);
        EnvMipLevel dst; dst.w = tx; dst.h = ty; dst.rgb.assign((size_t) 3 * tx * ty, 0.0f);
        const float *cur = src.rgb.data();
        std::vector<float> temp;
        if (sx != tx) {                                                       // X pass: ERepeat
            const Resampler1D r(sx, tx, true);
            float *out = dst.rgb.data();
            if (sy != ty) { temp.assign((size_t) 3 * tx * sy, 0.0f); out = temp.data(); }
            for (int y = 0; y < sy; ++y) r.run(cur + (size_t) 3 * y * sx, 1, out + (size_t) 3 * y * tx, 1, 3);
            cur = out;
        }
        if (sy != ty) {                                                       // Y pass: EClamp
            const Resampler1D r(sy, ty, false);
            for (int x = 0; x < tx; ++x) r.run(cur + (size_t) 3 * x, (size_t) tx, dst.rgb.data() + (size_t) 3 * x, (size_t) tx, 3);
        } else if (sx == tx) dst.rgb = src.rgb;
        levels.push_back(std::move(dst));
        sx = tx; sy = ty;
    }
    for (int i = 0; i < 64; ++i) {
        const float r2 = (float) i / (float) (64 - 1);
        lut[i] = cr_exp(-2.0f * r2) - cr_exp(-2.0f);
    }
}

} // namespace cp

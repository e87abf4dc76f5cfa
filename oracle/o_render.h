// oracle/o_render.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
//   PerspectiveCamera   src/sensors/perspective.cpp:126-173,271-298; src/librender/sensor.cpp:156-160,225-300
//   renderBlock         src/librender/integrator.cpp:140-188
//   MIPathTracer::Li    src/integrators/path/path.cpp:119-300
//   Scene::sampleEmitterDirect  src/librender/scene.cpp:828-853
//   ImageBlock::put     include/mitsuba/render/imageblock.h:124-186; filter table src/libcore/rfilter.cpp:37-55
//   develop             src/libcore/fmtconv.cpp:955-1056 (value / weight)
// The sampler plugins are replaced by the counter-based Philox stream described in o_math.h.
#pragma once
#include "o_math.h"
#include "o_hair.h"
#include "o_bsdf.h"
#include "o_env.h"
#include <thread>
#include <atomic>
#include <mutex>

namespace orc {

struct Camera {
    M44 toWorld = M44::identity();
    float xfov = 35, nearClip = 1e-2f, farClip = 1e4f;
    int filmW = 0, filmH = 0;
    float aspect = 1, invResX = 0, invResY = 0;
    M44 sampleToCamera;
    V3 dx, dy;

    // perspective.cpp:126-160 (no crop window) with the fp32 Transform algebra of the reference: every factory returns the matrix AND its
    // inverse (transform.cpp:33-63; Transform::perspective inverts numerically, :99-123 + matrix.inl:138-193), a product carries
    // (A.m * B.m, B.inv * A.inv) (transform.cpp:28-31), and sampleToCamera is the inverse member of
    //   scale(1/relSize) * translate(-relOffset) * scale(-0.5, -0.5 aspect, 1) * translate(-1, -1/aspect, 0) * perspective(xfov, near, far)
    // evaluated left to right.  The first two factors are exact identities without a crop window.
    void configure() {
        aspect = filmW / (float) filmH;                                    // sensor.cpp:101-102
        invResX = (float) 1 / (float) filmW; invResY = (float) 1 / (float) filmH;
        float recip = 1.0f / (farClip - nearClip);
        float cot = 1.0f / std::tan((xfov / 2.0f) * (kPi / 180.0f));      // libm tanf like the reference; the fov is a scene constant
        M44 persp; std::memset(persp.m, 0, sizeof(persp.m));
        persp.m[0][0] = cot; persp.m[1][1] = cot; persp.m[2][2] = farClip * recip; persp.m[2][3] = -nearClip * farClip * recip; persp.m[3][2] = 1;
        M44 perspInv;
        if (!invert(persp, perspInv)) throw std::runtime_error("oracle: singular camera transform");
        const float sx = -0.5f, sy = -0.5f * aspect, tx = -1.0f, ty = -1.0f / aspect;
        M44 scInv = M44::identity(); scInv.m[0][0] = 1.0f / sx; scInv.m[1][1] = 1.0f / sy; scInv.m[2][2] = 1.0f / 1.0f;
        M44 trInv = M44::identity(); trInv.m[0][3] = -tx; trInv.m[1][3] = -ty; trInv.m[2][3] = -0.0f;
        sampleToCamera = mul(perspInv, mul(trInv, scInv));
        dx = xfmPoint(sampleToCamera, V3(invResX, 0, 0)) - xfmPoint(sampleToCamera, V3(0.0f));
        dy = xfmPoint(sampleToCamera, V3(0, invResY, 0)) - xfmPoint(sampleToCamera, V3(0.0f));
    }
    // perspective.cpp:271-298 (time fixed at 0)
    void sampleRayDifferential(float px, float py, Ray &ray, V3 &rxDir, V3 &ryDir) const {
        V3 nearP = xfmPoint(sampleToCamera, V3(px * invResX, py * invResY, 0.0f));
        V3 d = normalize(nearP);
        float invZ = 1.0f / d.z;
        V3 o = xfmPoint(toWorld, V3(0.0f));
        ray = Ray(o, xfmVector(toWorld, d), nearClip * invZ, farClip * invZ);
        rxDir = xfmVector(toWorld, normalize(nearP + dx));
        ryDir = xfmVector(toWorld, normalize(nearP + dy));
    }
};

// rfilter.cpp:37-55 + rfilter.h:76-77, tent src/rfilters/tent.cpp:42-44, box/gaussian for completeness
struct ReconFilter {
    enum { RES = 31 };
    int type = 0; // 0 tent, 1 box, 2 gaussian
    float radius = 1.0f, stddev = 0.5f;
    float values[RES + 1];
    float scaleFactor = 0;
    float evalFilter(float x) const {
        if (type == 0) return std::max(0.0f, 1.0f - std::abs(x / radius));
        if (type == 1) return std::abs(x) <= radius ? 1.0f : 0.0f; // src/rfilters/box.cpp
        float alpha = -1.0f / (2.0f * stddev * stddev); // src/rfilters/gaussian.cpp
        return std::max(0.0f, cr::exp(alpha * x * x) - cr::exp(alpha * radius * radius));
    }
    void configure() {
        if (type == 1) radius = 0.5f + 1e-5f;          // box.cpp:38: props.getFloat("radius", 0.5f) + 1e-5f
        if (type == 2) radius = 4 * stddev;
        float sum = 0.0f;
        for (size_t i = 0; i < RES; ++i) { float v = evalFilter((radius * i) / RES); values[i] = v; sum += v; }
        values[RES] = 0.0f;
        scaleFactor = RES / radius;
        sum *= 2 * radius / RES;
        float normalization = 1.0f / sum;
        for (size_t i = 0; i < RES; ++i) values[i] *= normalization;
    }
    float evalDiscretized(float x) const { return values[std::min((int) std::abs(x * scaleFactor), (int) RES)]; }
};

// Film = full-size 5-channel accumulation buffer (R,G,B,alpha,weight), no border (ldrfilm.cpp:226-228)
struct Film {
    int w = 0, h = 0;
    std::vector<float> data;
    void init(int w_, int h_) { w = w_; h = h_; data.assign((size_t) 5 * w * h, 0.0f); }
    // imageblock.h:144-186 with offset=0, borderSize=0, size=(w,h)
    bool put(const ReconFilter &f, float posx, float posy, const V3 &spec, float alpha) {
        float value[5] = {spec.x, spec.y, spec.z, alpha, 1.0f};
        for (int i = 0; i < 5; ++i) if (!std::isfinite(value[i]) || value[i] < 0) return false;
        const float px = posx - 0.5f, py = posy - 0.5f, r = f.radius;
        int minx = std::max((int) std::ceil(px - r), 0), miny = std::max((int) std::ceil(py - r), 0);
        int maxx = std::min((int) std::floor(px + r), w - 1), maxy = std::min((int) std::floor(py + r), h - 1);
        float wx[16], wy[16];
        for (int x = minx, idx = 0; x <= maxx; ++x) wx[idx++] = f.evalDiscretized(x - px);
        for (int y = miny, idx = 0; y <= maxy; ++y) wy[idx++] = f.evalDiscretized(y - py);
        for (int y = miny, yr = 0; y <= maxy; ++y, ++yr) {
            float *dest = &data[((size_t) y * w + minx) * 5];
            for (int x = minx, xr = 0; x <= maxx; ++x, ++xr) {
                const float weight = wx[xr] * wy[yr];
                for (int k = 0; k < 5; ++k) *dest++ += weight * value[k];
            }
        }
        return true;
    }
};

struct BSDFAny {
    int kind = 0; // 0 KajiyaKay, 1 Marschner (as built), 2 SmoothDiffuse (meshes), 3 Marschner "fixed" (unbuilt marschner.cpp)
    KajiyaKay kk;
    std::shared_ptr<Marschner> ma;
    SmoothDiffuse df;
    std::shared_ptr<MarschnerFixed> mf;
    std::shared_ptr<RoughPlastic> rp;      // kind 4: `roughplastic` (default BSDF of the models/*/scene.xml files)
    ThinDielectric td;                     // kind 5: `thindielectric` (models/straight-hair/scene_thindielectric.xml)
    MarschnerDielectric md;                // kind 6: `marschnerdielectric` (models/straight-hair/scene_dielectric*.xml)
    SmoothPlastic pl;                      // kind 7: `plastic` (models/teapot/scene.xml:31-38)
    Mirror mi;                             // kind 8: `mirror` (the fork's plugin, models/teapot/mirror_scene.xml:32-36)
    // `twosided` (src/bsdfs/twosided.cpp:101-181) around a roughplastic / plastic with the same nested BRDF on both sides; the diffuse
    // kind keeps its own flag (SmoothDiffuse::twoSided)
    bool twoSided = false;
    int componentCount() const { return (kind == 2 || kind == 8) ? 1 : 2; }
    // measure: ESolidAngle unless `discrete` (only the dielectric kinds and `plastic` have discrete components); (u, v) = its.uv
    V3 evalOne(const V3 &wi, const V3 &wo, bool discrete, float u, float v) const {
        if (kind == 5) return td.eval(wi, wo, discrete);
        if (kind == 6) return md.eval(wi, wo, discrete);
        if (kind == 7) return pl.eval(wi, wo, discrete, u, v);
        if (kind == 8) return mi.eval(wi, wo, discrete);
        if (discrete) return V3(0.0f);
        return kind == 0 ? kk.eval(wi, wo) : kind == 1 ? ma->eval(wi, wo) : kind == 2 ? df.eval(wi, wo, u, v) : kind == 3 ? mf->eval(wi, wo) : rp->eval(wi, wo);
    }
    float pdfOne(const V3 &wi, const V3 &wo, bool discrete) const {
        if (kind == 5) return td.pdf(wi, wo, discrete);
        if (kind == 6) return md.pdf(wi, wo, discrete);
        if (kind == 7) return pl.pdf(wi, wo, discrete);
        if (kind == 8) return mi.pdf(wi, wo, discrete);
        if (discrete) return 0.0f;
        return kind == 0 ? kk.pdf(wi, wo) : kind == 1 ? ma->pdf(wi, wo) : kind == 2 ? df.pdf(wi, wo) : kind == 3 ? mf->pdf(wi, wo) : rp->pdf(wi, wo);
    }
    BSDFSample sampleOne(const V3 &wi, float sx, float sy, const float *extra, float u, float v) const {
        if (kind == 5) return td.sample(wi, sx, sy);
        if (kind == 6) return md.sample(wi, sx, sy);
        if (kind == 7) return pl.sample(wi, sx, sy, u, v);
        if (kind == 8) return mi.sample(wi, sx, sy);
        return kind == 0 ? kk.sample(wi, sx, sy) : kind == 1 ? ma->sample(wi, sx, sy) : kind == 2 ? df.sample(wi, sx, sy, u, v)
             : kind == 3 ? mf->sample(wi, extra[0], extra[1], extra[2], extra[3]) : rp->sample(wi, sx, sy);
    }
    V3 eval(V3 wi, V3 wo, bool discrete = false, float u = 0, float v = 0) const {
        if (twoSided && !(wi.z > 0)) { wi.z *= -1; wo.z *= -1; }           // twosided.cpp:101-115
        return evalOne(wi, wo, discrete, u, v);
    }
    float pdf(V3 wi, V3 wo, bool discrete = false) const {
        if (twoSided && !(wi.z > 0)) { wi.z *= -1; wo.z *= -1; }           // twosided.cpp:117-130
        return pdfOne(wi, wo, discrete);
    }
    // `extra` = four more uniform numbers: only the fixed Marschner draws them (two sampler->next2D() calls inside its sample())
    BSDFSample sample(V3 wi, float sx, float sy, const float *extra, float u = 0, float v = 0) const {
        bool flipped = false;
        if (twoSided && wi.z < 0) { wi.z *= -1; flipped = true; }           // twosided.cpp:162-181
        BSDFSample r = sampleOne(wi, sx, sy, extra, u, v);
        if (flipped && !isZero(r.weight) && r.pdf != 0) { r.wo.z *= -1; r.sampledComponent += componentCount(); }
        return r;
    }
    // BSDF::getType() & ESmooth: the thin dielectric has only discrete components (path.cpp:174-175 then skips emitter sampling)
    bool hasSmooth() const { return kind != 5 && kind != 8; }
    bool drawsExtra() const { return kind == 3; }
};

struct RenderStats { std::atomic<uint64_t> rays{0}, shadowRays{0}, paths{0}, pathLength{0}, dropped{0}; };

struct Scene {
    Geometry geo;
    std::vector<BSDFAny> bsdfs;
    EnvMap env; bool hasEnv = false;
    Camera cam;
    ReconFilter filter;
    bool filmHasAlpha = false;
    int maxDepth = -1, rrDepth = 5; bool strictNormals = false, hideEmitters = false;
    uint64_t seed = 0;
    RenderStats stats;
    // sampler: 0 = the counter-based Philox stream (default), 1 = the reference's `sobol` sampler (src/samplers/sobol.cpp), consumed in the
    // order renderBlock and MIPathTracer::Li call next2D / next1D
    int samplerKind = 0; uint64_t sobolScramble = 0;
    std::shared_ptr<SobolTables> sobolTables;
    SobolSampler makeSobol() const { SobolSampler sb; sb.configure(sobolTables.get(), sobolScramble, cam.filmW, cam.filmH); return sb; }

    void finalize() {
        geo.finalize();
        cam.configure();
        filter.configure();
        if (hasEnv) {
            // scene.cpp:387-413: kd-tree bounds + sensor position (+ the envmap's own AABB = centre point)
            AABB a = geo.aabb;
            a.expand(xfmPoint(cam.toWorld, V3(0.0f)));
            env.setSceneBounds(a);
        }
    }

    static float miWeight(float pdfA, float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); } // path.cpp:296-300

    // path.cpp:119-294.  `alpha` mirrors RadianceQueryRecord::alpha (records.inl:117-144)
    V3 Li(Ray ray, const V3 &rxDir, const V3 &ryDir, uint32_t pix, uint32_t samp, float &alpha, int *depthOut = nullptr, SobolSampler *sob = nullptr) {
        const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
        V3 Li(0.0f);
        bool scattered = false;
        Intersection its;
        stats.rays++;
        geo.rayIntersect(ray, its);
        alpha = its.valid ? 1.0f : 0.0f;
        ray.mint = kEpsilon;
        bool cameraRay = true;
        V3 throughput(1.0f);
        float eta = 1.0f;
        int depth = 1;
        bool emitted = true; // ERadiance includes EEmittedRadiance until the first bounce sets ERadianceNoEmission
        while (depth <= maxDepth || maxDepth < 0) {
            if (!its.valid) {
                if (emitted && (!hideEmitters || scattered) && hasEnv)
                    Li += throughput * env.evalEnvironment(ray.d, cameraRay, rxDir, ryDir);
                break;
            }
            const BSDFAny &bsdf = bsdfs[geo.shapes[its.shape].bsdf];
            if ((depth >= maxDepth && maxDepth > 0) || (strictNormals && dot(ray.d, its.geoFrame.n) * its.wi.z >= 0))
                break;
            Philox4 u = philox4x32_10(pix, samp, (uint32_t) depth, 0, k0, k1);
            /* direct illumination sampling, only for BSDFs with a smooth component (path.cpp:174-175) */
            float e0 = u32_to_unit(u.v[0]), e1 = u32_to_unit(u.v[1]);
            if (sob && bsdf.hasSmooth()) sob->next2D(e0, e1);                          // rRec.nextSample2D(), path.cpp:179: an argument of sampleEmitterDirect, drawn with or without emitters
            if (hasEnv && bsdf.hasSmooth()) {
                EnvMap::DirectSample ds = env.sampleDirect(its.p, e0, e1);
                V3 value(0.0f);
                if (ds.pdf != 0) {
                    Ray shadow(its.p, ds.d, kEpsilon, ds.dist * (1 - kShadowEpsilon));
                    stats.shadowRays++;
                    if (!geo.rayOccluded(shadow)) value = ds.value;
                }
                if (!isZero(value)) {
                    V3 wo = its.shFrame.toLocal(ds.d);
                    const V3 bsdfVal = bsdf.eval(its.wi, wo, false, its.u, its.v);
                    if (!isZero(bsdfVal) && (!strictNormals || dot(its.geoFrame.n, ds.d) * wo.z > 0)) {
                        float bsdfPdf = bsdf.pdf(its.wi, wo);
                        float weight = miWeight(ds.pdf, bsdfPdf);
                        Li += throughput * value * bsdfVal * weight;
                    }
                }
            }
            /* BSDF sampling */
            float extra[4] = {0, 0, 0, 0};
            float b0 = u32_to_unit(u.v[2]), b1 = u32_to_unit(u.v[3]);
            if (sob) sob->next2D(b0, b1);                                              // the `sample` argument of bsdf->sample(), path.cpp:210
            if (bsdf.drawsExtra()) {                       // counter stream 2 of this vertex (stream 0: emitter + BSDF sample, 1: roulette)
                Philox4 ue = philox4x32_10(pix, samp, (uint32_t) depth, 2, k0, k1);
                for (int k = 0; k < 4; ++k) extra[k] = u32_to_unit(ue.v[k]);
                if (sob) { sob->next2D(extra[0], extra[1]); sob->next2D(extra[2], extra[3]); }   // bRec.sampler->next2D() twice inside sample(), marschner.cpp:473-474
            }
            BSDFSample bs = bsdf.sample(its.wi, b0, b1, extra, its.u, its.v);
            if (isZero(bs.weight)) break;
            scattered |= bs.sampledType != ENull;
            const V3 wo = its.shFrame.toWorld(bs.wo);
            float woDotGeoN = dot(its.geoFrame.n, wo);
            if (strictNormals && woDotGeoN * bs.wo.z <= 0) break;
            bool hitEmitter = false;
            V3 value;
            ray = Ray(its.p, wo);
            cameraRay = false;
            stats.rays++;
            if (geo.rayIntersect(ray, its)) {
                /* neither hair nor the meshes of this path are emitters */
            } else {
                if (hasEnv) {
                    if (hideEmitters && !scattered) break;
                    value = env.evalEnvironment(ray.d);
                    if (!env.fillDirectSamplingRecord(ray.o, ray.d)) break;
                    hitEmitter = true;
                } else break;
            }
            throughput *= bs.weight;
            eta *= bs.eta;
            if (hitEmitter) {
                const float lumPdf = (!(bs.sampledType & EDelta)) ? env.pdfDirect(ray.d) : 0;
                Li += throughput * value * miWeight(bs.pdf, lumPdf);
            }
            if (!its.valid) break;
            emitted = false;
            if (depth++ >= rrDepth) {
                float q = std::min(maxc(throughput) * eta * eta, 0.95f);
                Philox4 ur = philox4x32_10(pix, samp, (uint32_t) (depth - 1), 1, k0, k1);
                const float rr = sob ? sob->next1D() : u32_to_unit(ur.v[0]);          // rRec.nextSample1D(), path.cpp:284
                if (rr >= q) break;
                throughput = throughput / q;
            }
        }
        stats.paths++; stats.pathLength += depth;
        if (depthOut) *depthOut = depth;
        return Li;
    }

    // integrator.cpp:140-188 for one (pixel, sample)
    bool renderSample(Film &film, uint32_t x, uint32_t y, uint32_t samp, uint32_t spp, V3 *LiOut = nullptr, float *posOut = nullptr) {
        const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
        uint32_t pix = y * (uint32_t) cam.filmW + x;
        Philox4 u = philox4x32_10(pix, samp, 0, 0, k0, k1);
        float px = (float) x + u32_to_unit(u.v[0]), py = (float) y + u32_to_unit(u.v[1]);
        SobolSampler sob;
        if (samplerKind == 1) {                                                    // sampler->generate(offset) ... advance() up to this sample (integrator.cpp:167-185)
            sob = makeSobol(); sob.px = (int) x; sob.py = (int) y; sob.setSampleIndex(samp);
            float a, b; sob.next2D(a, b);                                          // samplePos = Point2(offset) + Vector2(rRec.nextSample2D())
            px = (float) (int) x + a; py = (float) (int) y + b;
        }
        Ray ray; V3 rx, ry;
        cam.sampleRayDifferential(px, py, ray, rx, ry);
        // RayDifferential::scaleDifferential (ray.h:160-168) with 1/sqrt(spp)
        float ds = 1.0f / std::sqrt((float) spp);
        rx = ray.d + (rx - ray.d) * ds; ry = ray.d + (ry - ray.d) * ds;
        float alpha;
        V3 L = Li(ray, rx, ry, pix, samp, alpha, nullptr, samplerKind == 1 ? &sob : nullptr);
        if (!filmHasAlpha) alpha = 1.0f; // integrator.cpp:157-160 + RadianceQueryRecord::newQuery alpha=1
        if (LiOut) *LiOut = L;
        if (posOut) { posOut[0] = px; posOut[1] = py; }
        bool ok = film.put(filter, px, py, L, alpha);
        if (!ok) stats.dropped++;
        return ok;
    }

    // Multithreaded tile loop (renderproc.cpp:68-86: 32x32 tiles, private block merged under a mutex).
    // Renders sample indices [sBegin, sEnd) of `spp` for all pixels; tile-private accumulation
    // is emulated by a per-thread film merged at the end (sum order differs from the reference's tile order,
    // which is itself non-deterministic across threads).
    void render(Film &film, uint32_t spp, uint32_t sBegin, uint32_t sEnd, int nThreads) {
        const int W = cam.filmW, H = cam.filmH, TS = 32;
        const int tx = (W + TS - 1) / TS, ty = (H + TS - 1) / TS;
        std::atomic<int> next{0};
        std::mutex mtx;
        auto worker = [&]() {
            Film local; local.init(W, H);
            for (;;) {
                int t = next++;
                if (t >= tx * ty) break;
                int x0 = (t % tx) * TS, y0 = (t / tx) * TS;
                for (int y = y0; y < std::min(y0 + TS, H); ++y)
                    for (int x = x0; x < std::min(x0 + TS, W); ++x)
                        for (uint32_t s = sBegin; s < sEnd; ++s)
                            renderSample(local, (uint32_t) x, (uint32_t) y, s, spp);
            }
            std::lock_guard<std::mutex> g(mtx);
            for (size_t i = 0; i < film.data.size(); ++i) film.data[i] += local.data[i];
        };
        std::vector<std::thread> th;
        for (int i = 0; i < nThreads; ++i) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
};

} // namespace orc

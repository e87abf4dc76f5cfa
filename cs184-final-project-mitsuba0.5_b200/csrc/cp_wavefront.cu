// cp_wavefront.cu -- wavefront restructuring of MIPathTracer::Li (sm_100a).
//
// Replaces (reference file:line):
//   SamplingIntegrator::renderBlock      src/librender/integrator.cpp:140-188   (pixel x spp loop, jitter, put)
//   PerspectiveCameraImpl::sampleRayDifferential  src/sensors/perspective.cpp:271-298
//   MIPathTracer::Li                     src/integrators/path/path.cpp:119-300  (NEE + MIS + BSDF sampling + RR)
//   Scene::sampleEmitterDirect           src/librender/scene.cpp:828-853        (shadow ray [Epsilon, dist(1-ShadowEpsilon)])
//   ImageBlock::put                      include/mitsuba/render/imageblock.h:144-186 (discretised filter splat)
//
// One iteration of the reference's while-loop is cut at the ray cast: k_shade() first finishes the previous
// iteration for its path (environment MIS term on a miss, Russian roulette) and then runs the head of the next
// one (termination tests, emitter sampling, BSDF sampling).  Stages per bounce:
//     k_intersect (closest hit)  ->  k_shade  ->  k_shadow (any hit, adds the emitter sample)  -> swap queues
// Path state lives in HBM as SoA float4 streams; survivors are appended densely (warp-aggregated atomics) to
// the other half of a ping-pong pair, so every stage reads and writes coalesced 16-byte vectors.  Radiance is
// accumulated per path in a wave-sized array and splatted once per path by k_splat.
#include "cp_host.h"
#include "cp_traverse.cuh"
#include "cp_env.cuh"
#include "cp_camera.cuh"
#include "cp_wavefront.h"

namespace cp {

// flags word: bits 0..15 depth, bit 16 first (camera) segment, bit 17 last BSDF sample was EDelta, bit 18 invalid (padding pixel)
enum : uint32_t { F_FIRST = 1u << 16, F_DELTA = 1u << 17, F_INVALID = 1u << 18 };

__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!pred) return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t) __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// wave-local path index -> (pixel x, pixel y, sample index).  Pixels are enumerated in 8x8 tiles so that a warp
// covers an 8x4 block of the image (coherent camera rays, film atomics spread over 4 rows).
__device__ __forceinline__ bool path_to_pixel(const WaveParams &wp, uint64_t g, uint32_t &x, uint32_t &y, uint32_t &samp) {
    const uint64_t s = g / wp.pixPadded;
    const uint32_t rank = (uint32_t) (g - s * wp.pixPadded);
    const uint32_t tile = rank >> 6, within = rank & 63u;
    x = (tile % wp.tilesX) * 8u + (within & 7u);
    y = (tile / wp.tilesX) * 8u + (within >> 3);
    samp = wp.sampleBegin + (uint32_t) s;
    return x < wp.filmW && y < wp.filmH;
}

__global__ void __launch_bounds__(256) k_raygen(SceneDev S, WaveParams wp, PathQueue q, float4 *liAcc, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x, y, samp;
    const bool valid = path_to_pixel(wp, wp.waveBase + i, x, y, samp);
    liAcc[i] = make_float4(0, 0, 0, 1.0f);
    if (!valid) {
        q.ro[i] = make_float4(0, 0, 0, 1.0f); q.rd[i] = make_float4(0, 0, 1, 0.0f);
        q.thr[i] = make_float4(0, 0, 0, 0); q.id[i] = make_uint2(i, F_INVALID | F_FIRST | 1u);
        return;
    }
    const Philox4 u = philox4x32_10(y * wp.filmW + x, samp, 0u, 0u, wp.seedLo, wp.seedHi);
    const float px = (float) x + u32_to_unit(u.v[0]), py = (float) y + u32_to_unit(u.v[1]);
    const CameraRay r = camera_ray(S.cam, px, py, wp.diffScale);
    q.ro[i] = make_float4(r.o.x, r.o.y, r.o.z, r.mint);
    q.rd[i] = make_float4(r.d.x, r.d.y, r.d.z, r.maxt);
    q.thr[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    q.id[i] = make_uint2(i, F_FIRST | 1u);                 // depth starts at 1 (integrator.h:218-224)
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_intersect(SceneDev S, PathQueue q, uint32_t n, float4 *hitPT, uint32_t *hitPrim,
                                                   unsigned long long *stats, int *errFlag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 ro = q.ro[i], rd = q.rd[i];
    const uint2 id = q.id[i];
    RayHit h; h.t = CP_INF; h.gv = 0xffffffffu; h.p = V3(0.0f);
    uint32_t nv = 0, np = 0; int ovf = 0;
    if (!(id.y & F_INVALID))
        traverse<false, STATS>(S, V3(ro.x, ro.y, ro.z), V3(rd.x, rd.y, rd.z), ro.w, rd.w, h, nv, np, ovf);
    hitPT[i] = make_float4(h.p.x, h.p.y, h.p.z, h.t);
    hitPrim[i] = h.gv;
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 0, (unsigned long long) nv); atomicAdd(stats + 1, (unsigned long long) np); }
}

__device__ __forceinline__ float mi_weight(float pdfA, float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); } // path.cpp:296-300

__global__ void __launch_bounds__(128) k_shade(SceneDev S, WaveParams wp, PathQueue in, uint32_t n, const float4 *__restrict__ hitPT,
                                               const uint32_t *__restrict__ hitPrim, PathQueue out, ShadowQueue sq, float4 *liAcc,
                                               uint32_t *counters, unsigned long long *unsupportedLookups) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool survive = false, wantShadow = false;
    float4 nro, nrd, nthr; uint2 nid;
    float4 so, sd, sc;
    if (i < n) {
        const float4 ro4 = in.ro[i], rd4 = in.rd[i], thr4 = in.thr[i];
        const uint2 id = in.id[i];
        const uint32_t pathId = id.x, flags = id.y;
        do {
            if (flags & F_INVALID) break;
            const V3 rayO(ro4.x, ro4.y, ro4.z), rayD(rd4.x, rd4.y, rd4.z);
            V3 thr(thr4.x, thr4.y, thr4.z);
            const uint32_t gv = hitPrim[i];
            const bool first = flags & F_FIRST;
            int depth = (int) (flags & 0xffffu);
            uint32_t x, y, samp;
            path_to_pixel(wp, wp.waveBase + pathId, x, y, samp);
            const uint32_t pix = y * wp.filmW + x;

            if (gv == 0xffffffffu) {
                // ---- the ray escaped
                float4 acc = liAcc[pathId];
                if (first) {
                    acc.w = S.film.hasAlpha ? 0.0f : 1.0f;                      // records.inl:117-135
                    if (S.env.present && !S.integ.hideEmitters) {               // path.cpp:136-143
                        const Philox4 u = philox4x32_10(pix, samp, 0u, 0u, wp.seedLo, wp.seedHi);
                        const CameraRay cr = camera_ray(S.cam, (float) x + u32_to_unit(u.v[0]), (float) y + u32_to_unit(u.v[1]), wp.diffScale);
                        V3 L = thr * env_eval_filtered(S.env, rayD, cr.rx, cr.ry, unsupportedLookups);
                        acc.x += L.x; acc.y += L.y; acc.z += L.z;
                    }
                } else if (S.env.present) {                                     // path.cpp:233-264
                    const V3 value = env_eval(S.env, rayD);
                    if (env_fill_direct(S.env, rayO, rayD)) {
                        const float lumPdf = (flags & F_DELTA) ? 0.0f : env_pdf_direct(S.env, rayD);
                        V3 L = thr * value * mi_weight(thr4.w, lumPdf);
                        acc.x += L.x; acc.y += L.y; acc.z += L.z;
                    }
                }
                liAcc[pathId] = acc;
                break;
            }
            if (!first) {
                // ---- tail of the previous iteration: Russian roulette (path.cpp:276-286); eta == 1 for both hair BSDFs
                if (depth >= S.integ.rrDepth) {
                    const float qv = fminf(maxc(thr), 0.95f);
                    const Philox4 ur = philox4x32_10(pix, samp, (uint32_t) depth, 1u, wp.seedLo, wp.seedHi);
                    if (u32_to_unit(ur.v[0]) >= qv) break;
                    thr = thr / qv;
                }
                depth++;
                if (!(depth <= S.integ.maxDepth || S.integ.maxDepth < 0)) break;  // loop condition path.cpp:135
            }
            // ---- intersection record
            const float4 hp = hitPT[i];
            const float4 v1 = __ldg(S.vtx + gv), v2 = __ldg(S.vtx + gv + 1);
            const ShapeDev &shape = S.shapes[vtx_shape(v1)];
            HitRecord rec;
            fill_intersection(v1, v2, shape.radius, V3(hp.x, hp.y, hp.z), rayD, rec);
            if ((depth >= S.integ.maxDepth && S.integ.maxDepth > 0) ||
                (S.integ.strictNormals && dot(rayD, rec.geoN) * rec.wi.z >= 0)) break;   // path.cpp:156-165
            const BsdfDev &bsdf = S.bsdfs[shape.bsdf];
            const Philox4 u = philox4x32_10(pix, samp, (uint32_t) depth, 0u, wp.seedLo, wp.seedHi);
            // ---- emitter sampling (path.cpp:174-200)
            if (S.env.present) {
                const EnvSample es = env_sample_direct(S.env, rec.p, u32_to_unit(u.v[0]), u32_to_unit(u.v[1]));
                if (es.pdf != 0) {
                    // The shadow ray is always traced when pdf != 0 (scene.cpp:838-845); its contribution may be zero.
                    V3 contrib(0.0f);
                    if (!isZero(es.value)) {
                        const V3 wo = rec.sh.toLocal(es.d);
                        const V3 bsdfVal = bsdf_eval(bsdf, rec.wi, wo);
                        if (!isZero(bsdfVal) && (!S.integ.strictNormals || dot(rec.geoN, es.d) * wo.z > 0)) {
                            const float bsdfPdf = bsdf_pdf(bsdf, rec.wi, wo);
                            contrib = thr * es.value * bsdfVal * mi_weight(es.pdf, bsdfPdf);
                        }
                    }
                    wantShadow = true;
                    so = make_float4(rec.p.x, rec.p.y, rec.p.z, kEpsilon);
                    sd = make_float4(es.d.x, es.d.y, es.d.z, es.dist * (1 - kShadowEpsilon));
                    sc = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(pathId));
                }
            }
            // ---- BSDF sampling (path.cpp:207-226)
            const BsdfSampleOut bs = bsdf_sample(bsdf, rec.wi, u32_to_unit(u.v[2]), u32_to_unit(u.v[3]));
            if (isZero(bs.weight)) break;
            const V3 wo = rec.sh.toWorld(bs.wo);
            if (S.integ.strictNormals && dot(rec.geoN, wo) * bs.wo.z <= 0) break;
            thr = thr * bs.weight;
            nro = make_float4(rec.p.x, rec.p.y, rec.p.z, kEpsilon);
            nrd = make_float4(wo.x, wo.y, wo.z, CP_INF);
            nthr = make_float4(thr.x, thr.y, thr.z, bs.pdf);
            nid = make_uint2(pathId, (uint32_t) depth | ((bs.type & EDeltaReflection) ? F_DELTA : 0u));
            survive = true;
        } while (false);
    }
    const uint32_t oi = warp_append(counters + 0, survive);
    if (survive) { out.ro[oi] = nro; out.rd[oi] = nrd; out.thr[oi] = nthr; out.id[oi] = nid; }
    const uint32_t si = warp_append(counters + 1, wantShadow);
    if (wantShadow) { sq.o[si] = so; sq.d[si] = sd; sq.c[si] = sc; }
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_shadow(SceneDev S, ShadowQueue sq, uint32_t n, float4 *liAcc, unsigned long long *stats, int *errFlag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = sq.o[i], d = sq.d[i], c = sq.c[i];
    RayHit h; uint32_t nv = 0, np = 0; int ovf = 0;
    const bool occluded = traverse<true, STATS>(S, V3(o.x, o.y, o.z), V3(d.x, d.y, d.z), o.w, d.w, h, nv, np, ovf);
    if (!occluded && (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f)) {
        const uint32_t pathId = __float_as_uint(c.w);      // one shadow ray per path and bounce: no atomics needed
        float4 acc = liAcc[pathId];
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
        liAcc[pathId] = acc;
    }
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 2, (unsigned long long) nv); atomicAdd(stats + 3, (unsigned long long) np); }
}

// imageblock.h:144-186 with offset 0 / border 0 (the film itself, ldrfilm.cpp:226-228): 5 channels R,G,B,alpha,weight
__device__ __forceinline__ void film_put(const FilmDev &F, float *film, int W, int H, float posx, float posy, const V3 &spec, float alpha,
                                         unsigned long long *dropped) {
    const float value[5] = {spec.x, spec.y, spec.z, alpha, 1.0f};
#pragma unroll
    for (int k = 0; k < 5; ++k) if (!isfinite(value[k]) || value[k] < 0) { if (dropped) atomicAdd(dropped, 1ull); return; }
    const float px = posx - 0.5f, py = posy - 0.5f, r = F.filterRadius;
    const int minx = max((int) ceilf(px - r), 0), miny = max((int) ceilf(py - r), 0);
    const int maxx = min((int) floorf(px + r), W - 1), maxy = min((int) floorf(py + r), H - 1);
    for (int yy = miny; yy <= maxy; ++yy) {
        const float wy = F.filterValues[min((int) fabsf((yy - py) * F.filterScale), 31)];
        for (int xx = minx; xx <= maxx; ++xx) {
            const float wgt = F.filterValues[min((int) fabsf((xx - px) * F.filterScale), 31)] * wy;
            if (wgt == 0.0f) continue;                      // adding +0 is a no-op in the reference as well
            float *dest = film + ((size_t) yy * W + xx) * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) atomicAdd(dest + k, wgt * value[k]);
        }
    }
}

__global__ void __launch_bounds__(256) k_splat(SceneDev S, WaveParams wp, const float4 *__restrict__ liAcc, uint32_t n, float *film,
                                               unsigned long long *dropped) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x, y, samp;
    if (!path_to_pixel(wp, wp.waveBase + i, x, y, samp)) return;
    const Philox4 u = philox4x32_10(y * wp.filmW + x, samp, 0u, 0u, wp.seedLo, wp.seedHi);
    const float4 acc = liAcc[i];
    film_put(S.film, film, (int) wp.filmW, (int) wp.filmH, (float) x + u32_to_unit(u.v[0]), (float) y + u32_to_unit(u.v[1]),
             V3(acc.x, acc.y, acc.z), acc.w, dropped);
}

// Parity hook for F1: splat explicit samples
__global__ void k_splat_batch(SceneDev S, const float *__restrict__ pos, const float *__restrict__ rgb, const float *__restrict__ alpha,
                              uint64_t n, float *film, int W, int H) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    film_put(S.film, film, W, H, pos[2 * i], pos[2 * i + 1], V3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), alpha[i], nullptr);
}

// ------------------------------------------------------------------------------------------ host driver
#define CKW(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)

bool Wavefront::reserve(uint32_t waveSize, std::string &err) {
    if (waveSize <= capacity) return true;
    release();
    for (int k = 0; k < 2; ++k) {
        CKW(cudaMalloc(&q[k].ro, sizeof(float4) * (size_t) waveSize)); CKW(cudaMalloc(&q[k].rd, sizeof(float4) * (size_t) waveSize));
        CKW(cudaMalloc(&q[k].thr, sizeof(float4) * (size_t) waveSize)); CKW(cudaMalloc(&q[k].id, sizeof(uint2) * (size_t) waveSize));
    }
    CKW(cudaMalloc(&sq.o, sizeof(float4) * (size_t) waveSize)); CKW(cudaMalloc(&sq.d, sizeof(float4) * (size_t) waveSize));
    CKW(cudaMalloc(&sq.c, sizeof(float4) * (size_t) waveSize));
    CKW(cudaMalloc(&hitPT, sizeof(float4) * (size_t) waveSize)); CKW(cudaMalloc(&hitPrim, sizeof(uint32_t) * (size_t) waveSize));
    CKW(cudaMalloc(&liAcc, sizeof(float4) * (size_t) waveSize));
    CKW(cudaMalloc(&counters, sizeof(uint32_t) * 4));
    CKW(cudaMalloc(&stats, sizeof(unsigned long long) * 8));
    CKW(cudaMalloc(&errFlag, sizeof(int)));
    CKW(cudaMallocHost(&hCounters, sizeof(uint32_t) * 4));
    capacity = waveSize;
    return true;
}
void Wavefront::release() {
    for (int k = 0; k < 2; ++k) { cudaFree(q[k].ro); cudaFree(q[k].rd); cudaFree(q[k].thr); cudaFree(q[k].id); q[k] = PathQueue(); }
    cudaFree(sq.o); cudaFree(sq.d); cudaFree(sq.c); sq = ShadowQueue();
    cudaFree(hitPT); cudaFree(hitPrim); cudaFree(liAcc); cudaFree(counters); cudaFree(stats); cudaFree(errFlag);
    if (hCounters) cudaFreeHost(hCounters);
    hitPT = nullptr; hitPrim = nullptr; liAcc = nullptr; counters = nullptr; stats = nullptr; errFlag = nullptr; hCounters = nullptr;
    capacity = 0;
}

// Accumulates sample indices [sampleBegin, sampleEnd) of `spp` for every pixel into `d_film` (5 x W x H fp32).
bool Wavefront::render(const SceneDev &S, uint32_t spp, uint64_t seed, uint32_t sampleBegin, uint32_t sampleEnd, float *d_film,
                       uint32_t waveSize, bool collectStats, cudaStream_t stream, RenderStats &rs, std::string &err) {
    if (sampleEnd <= sampleBegin) return true;
    if (!reserve(waveSize, err)) return false;
    WaveParams wp;
    wp.filmW = (uint32_t) S.cam.filmW; wp.filmH = (uint32_t) S.cam.filmH;
    wp.tilesX = (wp.filmW + 7) / 8;
    const uint32_t tilesY = (wp.filmH + 7) / 8;
    wp.pixPadded = (uint64_t) wp.tilesX * tilesY * 64;
    wp.sampleBegin = sampleBegin;
    wp.seedLo = (uint32_t) seed; wp.seedHi = (uint32_t) (seed >> 32);
    wp.diffScale = 1.0f / sqrtf((float) spp);
    const uint64_t total = wp.pixPadded * (uint64_t) (sampleEnd - sampleBegin);
    CKW(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 8, stream));
    CKW(cudaMemsetAsync(errFlag, 0, sizeof(int), stream));
    for (uint64_t base = 0; base < total; base += waveSize) {
        const uint32_t n = (uint32_t) std::min<uint64_t>(waveSize, total - base);
        wp.waveBase = base;
        k_raygen<<<(n + 255) / 256, 256, 0, stream>>>(S, wp, q[0], liAcc, n);
        rs.launches++;
        uint32_t nActive = n; int cur = 0;
        while (nActive > 0) {
            CKW(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 4, stream));
            if (collectStats) k_intersect<true><<<(nActive + 127) / 128, 128, 0, stream>>>(S, q[cur], nActive, hitPT, hitPrim, stats, errFlag);
            else k_intersect<false><<<(nActive + 127) / 128, 128, 0, stream>>>(S, q[cur], nActive, hitPT, hitPrim, stats, errFlag);
            k_shade<<<(nActive + 127) / 128, 128, 0, stream>>>(S, wp, q[cur], nActive, hitPT, hitPrim, q[cur ^ 1], sq, liAcc, counters, stats + 4);
            CKW(cudaMemcpyAsync(hCounters, counters, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, stream));
            CKW(cudaStreamSynchronize(stream));
            rs.launches += 2; rs.rays += nActive; rs.bounces++;
            const uint32_t nNext = hCounters[0], nShadow = hCounters[1];
            if (nShadow) {
                if (collectStats) k_shadow<true><<<(nShadow + 127) / 128, 128, 0, stream>>>(S, sq, nShadow, liAcc, stats, errFlag);
                else k_shadow<false><<<(nShadow + 127) / 128, 128, 0, stream>>>(S, sq, nShadow, liAcc, stats, errFlag);
                rs.launches++; rs.shadowRays += nShadow;
            }
            cur ^= 1; nActive = nNext;
        }
        k_splat<<<(n + 255) / 256, 256, 0, stream>>>(S, wp, liAcc, n, d_film, stats + 5);
        rs.launches++;
        rs.paths += n;
    }
    unsigned long long hs[8]; int herr = 0;
    CKW(cudaMemcpyAsync(hs, stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
    CKW(cudaMemcpyAsync(&herr, errFlag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKW(cudaStreamSynchronize(stream));
    CKW(cudaGetLastError());
    if (herr) { err = "BVH traversal stack overflow"; return false; }
    rs.nodesVisited += hs[0] + hs[2]; rs.primsTested += hs[1] + hs[3]; rs.unsupportedLookups += hs[4]; rs.droppedSamples += hs[5];
    // the invalid padding "paths" (image sizes that are not multiples of 8) are not camera paths
    const uint64_t realPix = (uint64_t) wp.filmW * wp.filmH, padPix = wp.pixPadded - realPix;
    rs.paths -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    rs.rays -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    return true;
}

bool splat_batch(const SceneDev &S, const float *d_pos, const float *d_rgb, const float *d_alpha, uint64_t n, float *d_film, cudaStream_t stream, std::string &err) {
    if (n == 0) return true;
    k_splat_batch<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(S, d_pos, d_rgb, d_alpha, n, d_film, S.cam.filmW, S.cam.filmH);
    CKW(cudaStreamSynchronize(stream));
    CKW(cudaGetLastError());
    return true;
}

} // namespace cp

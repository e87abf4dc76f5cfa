// cp_wavefront.cu -- wavefront restructuring of MIPathTracer::Li: ray-casting stages and the host driver (sm_100a).
//
// Replaces (reference file:line):
//   SamplingIntegrator::render / renderBlock   src/librender/integrator.cpp:95-188   (pixel x spp loop)
//   MIPathTracer::Li                            src/integrators/path/path.cpp:119-300 (the bounce loop, cut at the ray casts)
//   Scene::rayIntersect / shadow-ray query      src/librender/skdtree.cpp:112-142, 207-226
//
// One iteration of the reference's while-loop is cut at the ray cast: k_shade() (cp_shade.cu) first finishes the previous
// iteration for its path (environment MIS term on a miss, Russian roulette) and then runs the head of the next one
// (termination tests, emitter sampling, BSDF sampling).  Stages per bounce:
//     k_intersect (closest hit)  ->  k_shade  ->  k_shadow (any hit, adds the emitter sample)  -> swap queues
// Path state lives in HBM as SoA float4 streams; survivors are appended densely (warp-aggregated atomics) to the other
// half of a ping-pong pair, so every stage reads and writes coalesced 16-byte vectors.  Radiance is accumulated per path in
// a wave-sized array and splatted once per path by k_splat.  This translation unit keeps FMA contraction (the FP64
// cylinder test and the slab tests profit from it); the shading stages live in cp_shade.cu, built with -fmad=false.
#include "cp_host.h"
#include "cp_traverse.cuh"
#include "cp_wavefront.h"
#include <mutex>
#include <vector>
#include <cub/cub.cuh>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

namespace cp {

// ray source / sink of the closest-hit stage: the active path queue -> dense hit records
// `perm` (optional) is the coherence order of the rays: slot k of the persistent counter traces queue entry perm[k]; results are
// written back at the queue index, so the shading stage keeps reading and writing dense, unpermuted streams.
struct PathIO {
    PathQueue q; float4 *hitPT; uint32_t *hitPrim; const uint32_t *perm;
    CP_D uint32_t slot(uint32_t i) const { return perm ? __ldg(perm + i) : i; }
    CP_D bool load(uint32_t k, V3 &o, V3 &d, float &mint, float &maxt) const {
        const uint32_t i = slot(k);
        const float4 ro = q.ro[i], rd = q.rd[i];
        o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z); mint = ro.w; maxt = rd.w;
        return !(q.id[i].y & F_INVALID);
    }
    CP_D void store(uint32_t k, bool, const RayHit &h) const { const uint32_t i = slot(k); hitPT[i] = make_float4(h.p.x, h.p.y, h.p.z, h.t); hitPrim[i] = h.gv; }
};
// shadow stage: unoccluded rays add their emitter sample to the owning path (one shadow ray per path and bounce: no atomics)
struct ShadowIO {
    ShadowQueue sq; float4 *liAcc; const uint32_t *perm;
    CP_D uint32_t slot(uint32_t i) const { return perm ? __ldg(perm + i) : i; }
    CP_D bool load(uint32_t k, V3 &o, V3 &d, float &mint, float &maxt) const {
        const uint32_t i = slot(k);
        const float4 ro = sq.o[i], rd = sq.d[i];
        o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z); mint = ro.w; maxt = rd.w;
        return true;
    }
    CP_D void store(uint32_t k, bool occluded, const RayHit &) const {
        if (occluded) return;
        const float4 c = sq.c[slot(k)];
        if (c.x == 0.0f && c.y == 0.0f && c.z == 0.0f) return;
        const uint32_t pathId = __float_as_uint(c.w);
        float4 acc = liAcc[pathId];
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
        liAcc[pathId] = acc;
    }
};

template <bool STATS, bool MESH>
__global__ void __launch_bounds__(128, CP_MIN_BLOCKS) k_intersect(SceneDev S, PathIO io, uint32_t n, uint32_t *rayCounter, unsigned long long *stats, int *errFlag) {
    TraceCounters tc = {0, 0, 0}; int ovf = 0;
    trace_persistent<false, STATS, MESH>(S, io, n, rayCounter, tc, ovf);
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 0, tc.nodes); atomicAdd(stats + 1, tc.prims); atomicAdd(stats + 6, tc.fullTests); }
}
template <bool STATS, bool MESH>
__global__ void __launch_bounds__(128, CP_MIN_BLOCKS) k_shadow(SceneDev S, ShadowIO io, uint32_t n, uint32_t *rayCounter, unsigned long long *stats, int *errFlag) {
    TraceCounters tc = {0, 0, 0}; int ovf = 0;
    trace_persistent<true, STATS, MESH>(S, io, n, rayCounter, tc, ovf);
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 2, tc.nodes); atomicAdd(stats + 3, tc.prims); atomicAdd(stats + 7, tc.fullTests); }
}

// Coherence keys: 27-bit Morton code of the ray origin inside the scene bounds, then the direction octant.  Rays that start
// close together (and head the same way) end up in the same warp, walk the same part of the tree and hit in L1/L2.
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_ray_keys(const float4 *__restrict__ ro, const float4 *__restrict__ rd, uint32_t n, float3 smin, float3 sinv, uint32_t *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = ro[i], d = rd[i];
    const uint32_t x = (uint32_t) fminf(fmaxf((o.x - smin.x) * sinv.x, 0.0f), 511.0f);
    const uint32_t y = (uint32_t) fminf(fmaxf((o.y - smin.y) * sinv.y, 0.0f), 511.0f);
    const uint32_t z = (uint32_t) fminf(fmaxf((o.z - smin.z) * sinv.z, 0.0f), 511.0f);
    const uint32_t oct = (d.x < 0 ? 1u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 4u : 0u);
    keys[i] = (((spread10(x) << 2) | (spread10(y) << 1) | spread10(z)) << 3) | oct;
    vals[i] = i;
}

// persistent launch: enough CTAs to fill the machine, never more than the work needs
static unsigned persistent_grid(const void *kernel, uint32_t n) {
    static int numSMs = 0;
    if (!numSMs) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev); }
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, 128, 0);
    if (perSM < 1) perSM = 1;
    const unsigned full = (unsigned) (numSMs * perSM), need = (n + 127u) / 128u;
    return need < full ? need : full;
}

// ------------------------------------------------------------------------------------------ host driver
#define CKW(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)

static std::mutex g_pinnedMutex;
static std::vector<uint32_t *> g_pinnedFree;

bool Wavefront::reserve(uint32_t waveSize, cudaStream_t stream, std::string &err) {
    if (waveSize <= capacity) return true;
    release();
    allocStream = stream;
    for (int k = 0; k < 2; ++k) {
        CKW(dev_alloc(&q[k].ro, sizeof(float4) * (size_t) waveSize)); CKW(dev_alloc(&q[k].rd, sizeof(float4) * (size_t) waveSize));
        CKW(dev_alloc(&q[k].thr, sizeof(float4) * (size_t) waveSize)); CKW(dev_alloc(&q[k].id, sizeof(uint2) * (size_t) waveSize));
    }
    CKW(dev_alloc(&sq.o, sizeof(float4) * (size_t) waveSize)); CKW(dev_alloc(&sq.d, sizeof(float4) * (size_t) waveSize));
    CKW(dev_alloc(&sq.c, sizeof(float4) * (size_t) waveSize));
    CKW(dev_alloc(&hitPT, sizeof(float4) * (size_t) waveSize)); CKW(dev_alloc(&hitPrim, sizeof(uint32_t) * (size_t) waveSize));
    CKW(dev_alloc(&liAcc, sizeof(float4) * (size_t) waveSize));
    CKW(dev_alloc(&counters, sizeof(uint32_t) * 8));
    CKW(dev_alloc(&stats, sizeof(unsigned long long) * 8));
    CKW(dev_alloc(&errFlag, sizeof(int)));
    for (int k = 0; k < 2; ++k) {
        CKW(dev_alloc(&sortKeys[k], sizeof(uint32_t) * (size_t) waveSize));
        CKW(dev_alloc(&sortVals[k], sizeof(uint32_t) * (size_t) waveSize));
    }
    sortTempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortTempBytes, sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], (int) waveSize, 0, 30, stream);
    CKW(dev_alloc(&sortTemp, sortTempBytes));
    {   // page-locked read-back slot: recycled through a process-wide free list (cudaMallocHost / cudaFreeHost cost up to a millisecond and synchronise)
        std::lock_guard<std::mutex> g(g_pinnedMutex);
        if (!g_pinnedFree.empty()) { hCounters = g_pinnedFree.back(); g_pinnedFree.pop_back(); }
    }
    if (!hCounters) CKW(cudaMallocHost(&hCounters, sizeof(uint32_t) * 8));
    capacity = waveSize;
    return true;
}
// sorts ray indices by coherence key; returns the permutation (device pointer) or nullptr when sorting is off / not worthwhile
const uint32_t *Wavefront::coherence_order(const SceneDev &S, const float4 *ro, const float4 *rd, uint32_t n, cudaStream_t stream) {
    if (!sortRays || n < 65536u) return nullptr;
    const float3 smin = make_float3(S.sceneMin[0], S.sceneMin[1], S.sceneMin[2]);
    const float3 sinv = make_float3(512.0f / fmaxf(S.sceneMax[0] - S.sceneMin[0], 1e-20f), 512.0f / fmaxf(S.sceneMax[1] - S.sceneMin[1], 1e-20f),
                                    512.0f / fmaxf(S.sceneMax[2] - S.sceneMin[2], 1e-20f));
    k_ray_keys<<<(n + 255) / 256, 256, 0, stream>>>(ro, rd, n, smin, sinv, sortKeys[0], sortVals[0]);
    size_t need = sortTempBytes;
    cub::DeviceRadixSort::SortPairs(sortTemp, need, sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], (int) n, 0, 30, stream);
    return sortVals[1];
}

void Wavefront::release() {
    void *ptrs[] = {sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], sortTemp, q[0].ro, q[0].rd, q[0].thr, q[0].id, q[1].ro, q[1].rd, q[1].thr, q[1].id, sq.o, sq.d, sq.c, hitPT, hitPrim, liAcc, counters, stats, errFlag};
    if (capacity) cudaDeviceSynchronize();          // the last render may have run on a caller's stream
    for (void *p : ptrs) if (p) dev_free(p);
    q[0] = PathQueue(); q[1] = PathQueue(); sq = ShadowQueue();
    sortKeys[0] = sortKeys[1] = sortVals[0] = sortVals[1] = nullptr; sortTemp = nullptr;
    if (hCounters) { std::lock_guard<std::mutex> g(g_pinnedMutex); g_pinnedFree.push_back(hCounters); }
    hitPT = nullptr; hitPrim = nullptr; liAcc = nullptr; counters = nullptr; stats = nullptr; errFlag = nullptr; hCounters = nullptr;
    capacity = 0;
}

// Accumulates sample indices [sampleBegin, sampleEnd) of `spp` for every pixel into `d_film` (5 x W x H fp32).
bool Wavefront::render(const SceneDev &S, uint32_t spp, uint64_t seed, uint32_t sampleBegin, uint32_t sampleEnd, float *d_film,
                       uint32_t waveSize, bool collectStats, bool profileStages, cudaStream_t stream, RenderStats &rs, std::string &err) {
    if (sampleEnd <= sampleBegin) return true;
    {   // never allocate more queue slots than there are paths
        const uint64_t tiles = (uint64_t) ((S.cam.filmW + 7) / 8) * ((S.cam.filmH + 7) / 8) * 64ull * (sampleEnd - sampleBegin);
        if (tiles < waveSize) waveSize = (uint32_t) std::max<uint64_t>(tiles, 1024);
    }
    const bool hasMesh = S.mesh.triCount > 0;
    const bool trace = getenv("CUDAPATH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tr0 = now();
    if (!reserve(waveSize, stream, err)) return false;
    if (trace) { cudaStreamSynchronize(stream); fprintf(stderr, "[cudapath] queue reserve (%u paths): %.3f s\n", waveSize, now() - tr0); }
    // optional per-launch stage timing: one event pair per launch, resolved after the last wave
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> spans;
    auto begin = [&](int stage) { if (!profileStages) return; Span sp; cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); sp.stage = stage; cudaEventRecord(sp.a, stream); spans.push_back(sp); };
    auto end = [&]() { if (profileStages) cudaEventRecord(spans.back().b, stream); };
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
        begin(3); launch_raygen(S, wp, q[0], liAcc, n, stream); end();
        rs.launches++;
        uint32_t nActive = n; int cur = 0, bounce = 0;
        while (nActive > 0) {
            CKW(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 8, stream));
            begin(0);
            {
                // camera rays leave raygen in pixel order (already coherent); later bounces are re-ordered
                const uint32_t *perm = bounce > 0 ? coherence_order(S, q[cur].ro, q[cur].rd, nActive, stream) : nullptr;
                PathIO io{q[cur], hitPT, hitPrim, perm};
#define CP_LAUNCH_TRACE(K, ST, ME, N, CTR) K<ST, ME><<<persistent_grid((const void *) K<ST, ME>, N), 128, 0, stream>>>(S, io, N, CTR, stats, errFlag)
                if (hasMesh) { if (collectStats) CP_LAUNCH_TRACE(k_intersect, true, true, nActive, counters + 2); else CP_LAUNCH_TRACE(k_intersect, false, true, nActive, counters + 2); }
                else { if (collectStats) CP_LAUNCH_TRACE(k_intersect, true, false, nActive, counters + 2); else CP_LAUNCH_TRACE(k_intersect, false, false, nActive, counters + 2); }
            }
            end();
            begin(1);
            launch_shade(S, wp, q[cur], nActive, hitPT, hitPrim, q[cur ^ 1], sq, liAcc, counters, stats + 4, stream);
            end();
            CKW(cudaMemcpyAsync(hCounters, counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, stream));
            CKW(cudaStreamSynchronize(stream));
            rs.launches += 2; rs.rays += nActive; rs.bounces++;
            const uint32_t nNext = hCounters[0], nShadow = hCounters[1];
            rs.shadowRays += hCounters[4];              // shadow rays the reference traces although they cannot contribute (see k_shade): counted, not traced
            if (nShadow) {
                begin(2);
                ShadowIO io{sq, liAcc, coherence_order(S, sq.o, sq.d, nShadow, stream)};
                if (hasMesh) { if (collectStats) CP_LAUNCH_TRACE(k_shadow, true, true, nShadow, counters + 3); else CP_LAUNCH_TRACE(k_shadow, false, true, nShadow, counters + 3); }
                else { if (collectStats) CP_LAUNCH_TRACE(k_shadow, true, false, nShadow, counters + 3); else CP_LAUNCH_TRACE(k_shadow, false, false, nShadow, counters + 3); }
                end();
                rs.launches++; rs.shadowRays += nShadow; rs.shadowRaysTraced += nShadow;
            }
            cur ^= 1; nActive = nNext; bounce++;
            if (cancelRequested.load(std::memory_order_relaxed)) {      // the unfinished wave is dropped: the film holds whole waves only
                cudaStreamSynchronize(stream);
                for (auto &sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
                err = "render cancelled";
                return false;
            }
        }
        begin(4); launch_splat(S, wp, liAcc, n, d_film, stats + 5, stream); end();
        rs.launches++;
        rs.paths += n;
        if (progress) progress(progressUser, std::min<uint64_t>(base + n, total), total);
    }
    unsigned long long hs[8]; int herr = 0;
    CKW(cudaMemcpyAsync(hs, stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
    CKW(cudaMemcpyAsync(&herr, errFlag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKW(cudaStreamSynchronize(stream));
    CKW(cudaGetLastError());
    if (herr) { err = "BVH traversal stack overflow"; return false; }
    for (auto &sp : spans) {
        float ms = 0; cudaEventElapsedTime(&ms, sp.a, sp.b);
        rs.stageMs[sp.stage] += ms; rs.stageLaunches[sp.stage]++;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    rs.nodesVisited += hs[0]; rs.primsTested += hs[1]; rs.shadowNodesVisited += hs[2]; rs.shadowPrimsTested += hs[3]; rs.unsupportedLookups += hs[4]; rs.droppedSamples += hs[5];
    rs.fullTests += hs[6]; rs.shadowFullTests += hs[7];
    // the invalid padding "paths" (image sizes that are not multiples of 8) are not camera paths
    const uint64_t realPix = (uint64_t) wp.filmW * wp.filmH, padPix = wp.pixPadded - realPix;
    rs.paths -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    rs.rays -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    return true;
}

} // namespace cp

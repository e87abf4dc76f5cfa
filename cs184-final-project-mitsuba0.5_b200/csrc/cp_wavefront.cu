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

// Ray source / sink of the trace stage.  Entry k of the launch is, in coherence order (`perm`) or in queue order, either a
// closest-hit ray of the active path queue (result: dense hit records at the queue index) or a shadow ray of the previous bounce
// (result: an unoccluded ray adds its emitter sample to the owning path; one shadow ray per path and bounce, so no atomics).
// Results are written back at the queue index, so the shading stage keeps reading and writing dense, unpermuted streams.
struct WaveIO {
    PathQueue q; float4 *hitPT; uint32_t *hitPrim;
    ShadowQueue sq; float4 *liAcc;
    const uint32_t *perm;          // sorted entries: bit 31 = shadow ray, bits 0..30 = queue index; nullptr = closest rays first, then shadow rays
    uint32_t nClosest;
    CP_D bool load(uint32_t k, V3 &o, V3 &d, float &mint, float &maxt, bool &any, uint32_t &slot) const {
        uint32_t e;
        if (perm) e = __ldg(perm + k); else e = k < nClosest ? k : (0x80000000u | (k - nClosest));
        slot = e;
        const uint32_t i = e & 0x7fffffffu;
        if (e & 0x80000000u) {
            const float4 ro = sq.o[i], rd = sq.d[i];
            o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z); mint = ro.w; maxt = rd.w; any = true;
            return true;
        }
        const float4 ro = q.ro[i], rd = q.rd[i];
        o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z); mint = ro.w; maxt = rd.w; any = false;
        return !(q.id[i].y & F_INVALID);
    }
    CP_D void store(uint32_t e, bool any, bool hit, const RayHit &h) const {
        const uint32_t i = e & 0x7fffffffu;
        if (!any) { hitPT[i] = make_float4(h.p.x, h.p.y, h.p.z, h.t); hitPrim[i] = h.gv; return; }
        if (hit) return;
        const float4 c = sq.c[i];
        if (c.x == 0.0f && c.y == 0.0f && c.z == 0.0f) return;
        const uint32_t pathId = __float_as_uint(c.w);
        float4 acc = liAcc[pathId];
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
        liAcc[pathId] = acc;
    }
};

// prevSlot[0] closest-hit rays, prevSlot[1] shadow rays waiting in the queues; slot[2] = work counter of this launch
template <bool STATS, bool MESH>
__global__ void __launch_bounds__(CP_TRACE_THREADS, CP_MIN_BLOCKS) k_trace(SceneDev S, WaveIO io, const uint32_t *__restrict__ prevSlot, uint32_t *rayCounter,
                                                                          unsigned long long *stats, int *errFlag) {
    TraceCounters tc[2] = {{0, 0, 0}, {0, 0, 0}}; int ovf = 0;
    io.nClosest = prevSlot[0];
    const uint32_t n = io.nClosest + prevSlot[1];
    if (blockIdx.x * CP_TRACE_THREADS >= n) return;     // more CTAs than work: the CTAs below this one pull everything
    trace_persistent<TRACE_MIXED, STATS, MESH>(S, io, n, rayCounter, tc, ovf);
    if (ovf) *errFlag = 1;
    if (STATS) {
        atomicAdd(stats + 0, tc[0].nodes); atomicAdd(stats + 1, tc[0].prims); atomicAdd(stats + 6, tc[0].fullTests);
        atomicAdd(stats + 2, tc[1].nodes); atomicAdd(stats + 3, tc[1].prims); atomicAdd(stats + 7, tc[1].fullTests);
    }
}

// Coherence keys: 27-bit Morton code of the ray origin inside the scene bounds, then the direction octant.  Rays that start
// close together (and head the same way) end up in the same warp, walk the same part of the tree and hit in L1/L2.  A shadow
// ray and the next closest-hit ray of a path start at the same hit point, so one sort serves both.  The grid covers upper bounds
// of the two queue lengths (the host runs ahead of the device); entries beyond the real lengths get a key above all real keys.
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_ray_keys(PathQueue q, ShadowQueue sq, const uint32_t *__restrict__ prevSlot, uint32_t ubClosest, uint32_t ubTotal, float3 smin, float3 sinv,
                           uint32_t *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ubTotal) return;
    const bool shadow = i >= ubClosest;
    const uint32_t j = shadow ? i - ubClosest : i;
    vals[i] = (shadow ? 0x80000000u : 0u) | j;
    if (j >= prevSlot[shadow ? 1 : 0]) { keys[i] = 0x40000000u; return; }
    const float4 o = shadow ? sq.o[j] : q.ro[j], d = shadow ? sq.d[j] : q.rd[j];
    const uint32_t x = (uint32_t) fminf(fmaxf((o.x - smin.x) * sinv.x, 0.0f), 511.0f);
    const uint32_t y = (uint32_t) fminf(fmaxf((o.y - smin.y) * sinv.y, 0.0f), 511.0f);
    const uint32_t z = (uint32_t) fminf(fmaxf((o.z - smin.z) * sinv.z, 0.0f), 511.0f);
    const uint32_t oct = (d.x < 0 ? 1u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 4u : 0u);
    keys[i] = (((spread10(x) << 2) | (spread10(y) << 1) | spread10(z)) << 3) | oct;
}

// persistent launch: enough CTAs to fill the machine, never more than the work needs
static unsigned persistent_grid(const void *kernel) {
    int dev = 0, numSMs = 0, perSM = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, CP_TRACE_THREADS, 0);
    if (perSM < 1) perSM = 1;
    return (unsigned) (numSMs * perSM);
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
    CKW(dev_alloc(&counters, sizeof(uint32_t) * 8 * (CP_CTR_RING + 1)));
    CKW(dev_alloc(&stats, sizeof(unsigned long long) * 8));
    CKW(dev_alloc(&errFlag, sizeof(int)));
    // one sort covers the closest-hit rays of a bounce and the shadow rays of the bounce before: 2 x waveSize entries
    for (int k = 0; k < 2; ++k) {
        CKW(dev_alloc(&sortKeys[k], sizeof(uint32_t) * 2 * (size_t) waveSize));
        CKW(dev_alloc(&sortVals[k], sizeof(uint32_t) * 2 * (size_t) waveSize));
    }
    sortTempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortTempBytes, sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], (int) std::min<uint64_t>(2ull * waveSize, 0x7fffffffull), 0, 31, stream);
    CKW(dev_alloc(&sortTemp, sortTempBytes));
    {   // page-locked read-back slots: recycled through a process-wide free list (cudaMallocHost / cudaFreeHost cost up to a millisecond and synchronise)
        std::lock_guard<std::mutex> g(g_pinnedMutex);
        if (!g_pinnedFree.empty()) { hCounters = g_pinnedFree.back(); g_pinnedFree.pop_back(); }
    }
    if (!hCounters) CKW(cudaMallocHost(&hCounters, sizeof(uint32_t) * 8 * CP_CTR_RING));
    for (int k = 0; k < CP_CTR_RING; ++k) if (!slotEvent[k]) CKW(cudaEventCreateWithFlags(&slotEvent[k], cudaEventDisableTiming));
    capacity = waveSize;
    return true;
}
// Sorts the waiting rays (closest-hit rays of q, shadow rays of sq) by coherence key; returns the sorted entry list (device pointer)
// or nullptr when sorting is off / not worthwhile.  ubClosest / ubShadow bound the queue lengths, which only the device knows.
const uint32_t *Wavefront::coherence_order(const SceneDev &S, const PathQueue &pq, const ShadowQueue &shq, const uint32_t *prevSlot, uint32_t ubClosest, uint32_t ubShadow,
                                           cudaStream_t stream) {
    const uint64_t total = (uint64_t) ubClosest + ubShadow;
    if (!sortRays || total < 65536u || total > 0x7fffffffull) return nullptr;
    const float3 smin = make_float3(S.sceneMin[0], S.sceneMin[1], S.sceneMin[2]);
    const float3 sinv = make_float3(512.0f / fmaxf(S.sceneMax[0] - S.sceneMin[0], 1e-20f), 512.0f / fmaxf(S.sceneMax[1] - S.sceneMin[1], 1e-20f),
                                    512.0f / fmaxf(S.sceneMax[2] - S.sceneMin[2], 1e-20f));
    k_ray_keys<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(pq, shq, prevSlot, ubClosest, (uint32_t) total, smin, sinv, sortKeys[0], sortVals[0]);
    size_t need = sortTempBytes;
    cub::DeviceRadixSort::SortPairs(sortTemp, need, sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], (int) total, 0, 31, stream);
    return sortVals[1];
}

void Wavefront::release() {
    void *ptrs[] = {sortKeys[0], sortKeys[1], sortVals[0], sortVals[1], sortTemp, q[0].ro, q[0].rd, q[0].thr, q[0].id, q[1].ro, q[1].rd, q[1].thr, q[1].id, sq.o, sq.d, sq.c, hitPT, hitPrim, liAcc, counters, stats, errFlag};
    if (capacity) cudaDeviceSynchronize();          // the last render may have run on a caller's stream
    for (void *p : ptrs) if (p) dev_free(p);
    q[0] = PathQueue(); q[1] = PathQueue(); sq = ShadowQueue();
    sortKeys[0] = sortKeys[1] = sortVals[0] = sortVals[1] = nullptr; sortTemp = nullptr;
    if (hCounters) { std::lock_guard<std::mutex> g(g_pinnedMutex); g_pinnedFree.push_back(hCounters); }
    for (int k = 0; k < CP_CTR_RING; ++k) if (slotEvent[k]) { cudaEventDestroy(slotEvent[k]); slotEvent[k] = nullptr; }
    hitPT = nullptr; hitPrim = nullptr; liAcc = nullptr; counters = nullptr; stats = nullptr; errFlag = nullptr; hCounters = nullptr;
    capacity = 0;
}

// Accumulates sample indices [sampleBegin, sampleEnd) of `spp` for every pixel into `d_film` (5 x W x H fp32).
//
// Bounce b of a wave is three launches: [coherence sort] -> k_trace (closest-hit rays of bounce b + shadow rays of bounce b-1) -> k_shade.
// All of them read the queue lengths from the counter slot of bounce b-1 on the device.  The host launches bounce b as soon as it
// has seen the counters of bounce b-2 (which bound those of b-1: a path queue only shrinks, and a path emits at most one shadow
// ray), so the device always has the next bounce queued while the host learns how the last but one ended; the wave is over when a
// bounce leaves no path alive -- its shadow rays are traced by the launch that is already queued.
bool Wavefront::render(const SceneDev &scene, uint32_t spp, uint64_t seed, uint32_t sampleBegin, uint32_t sampleEnd, float *d_film,
                       uint32_t waveSize, bool collectStats, bool profileStages, cudaStream_t stream, RenderStats &rs, std::string &err) {
    if (sampleEnd <= sampleBegin) return true;
    SceneDev S = scene;                  // the copy handed to the kernels (S.sobol.err points at this wavefront's error word once it exists)
    WaveParams wp;
    wp.filmW = (uint32_t) S.cam.filmW; wp.filmH = (uint32_t) S.cam.filmH;
    wp.tilesX = (wp.filmW + 7) / 8;
    uint64_t realPix = (uint64_t) wp.filmW * wp.filmH;           // pixels this context renders
    if (shardCount <= 1) wp.pixPadded = (uint64_t) wp.tilesX * ((wp.filmH + 7) / 8) * 64;
    else {
        wp.blockShift = shardBlockShift;
        const uint32_t bpx = 8u << wp.blockShift;               // pixels per block side
        const uint32_t blocksX = (wp.filmW + bpx - 1) / bpx, blocksY = (wp.filmH + bpx - 1) / bpx;
        wp.shardIndex = shardIndex; wp.shardCount = shardCount;
        wp.cellW = shardCount; wp.cellH = 1;                     // cell shape: as square as the shard count allows (8 -> 4 x 2, 4 -> 2 x 2, 6 -> 3 x 2)
        for (uint32_t hgt = 2; hgt * hgt <= shardCount; ++hgt) if (shardCount % hgt == 0) { wp.cellH = hgt; wp.cellW = shardCount / hgt; }
        wp.cellsPerRow = (blocksX + wp.cellW - 1) / wp.cellW;
        const uint32_t cellRows = (blocksY + wp.cellH - 1) / wp.cellH;
        wp.pixPadded = (uint64_t) wp.cellsPerRow * cellRows * (uint64_t) bpx * bpx;
        realPix = 0;
        for (uint32_t cy = 0; cy < cellRows; ++cy) for (uint32_t cx = 0; cx < wp.cellsPerRow; ++cx) {
            const uint32_t slot = (shardIndex + cx + 3u * cy) % shardCount;
            const uint32_t bx = cx * wp.cellW + slot % wp.cellW, by = cy * wp.cellH + slot / wp.cellW;
            if (bx >= blocksX || by >= blocksY) continue;
            realPix += (uint64_t) std::min(bpx, wp.filmW - bx * bpx) * std::min(bpx, wp.filmH - by * bpx);
        }
        if (wp.pixPadded == 0 || realPix == 0) return true;     // this shard owns no pixel
    }
    {   // never allocate more queue slots than there are paths
        const uint64_t tiles = wp.pixPadded * (sampleEnd - sampleBegin);
        if (tiles < waveSize) waveSize = (uint32_t) std::max<uint64_t>(tiles, 1024);
    }
    if (waveSize > 0x3fffffffu) waveSize = 0x3fffffffu;
    const bool hasMesh = S.mesh.triCount + S.mesh.rectCount > 0;
    const bool trace = getenv("CUDAPATH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tr0 = now();
    if (!reserve(waveSize, stream, err)) return false;
    S.sobol.err = errFlag;
    if (trace) { cudaStreamSynchronize(stream); fprintf(stderr, "[cudapath] queue reserve (%u paths): %.3f s\n", waveSize, now() - tr0); }
    // optional per-launch stage timing: one event pair per launch, resolved after the last wave
    struct Span { cudaEvent_t a, b; int stage; double tHost; };
    std::vector<Span> spans;
    auto begin = [&](int stage) { if (!profileStages) return; Span sp; cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); sp.stage = stage; sp.tHost = now(); cudaEventRecord(sp.a, stream); spans.push_back(sp); };
    auto end = [&]() { if (profileStages) cudaEventRecord(spans.back().b, stream); };
    auto dropSpans = [&]() { for (auto &sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); } spans.clear(); };
    // the trace kernel variant of this scene and its machine-filling grid
    void (*traceKernel)(SceneDev, WaveIO, const uint32_t *, uint32_t *, unsigned long long *, int *) =
        hasMesh ? (collectStats ? k_trace<true, true> : k_trace<false, true>) : (collectStats ? k_trace<true, false> : k_trace<false, false>);
    const unsigned fullGrid = persistent_grid((const void *) traceKernel);
    wp.sampleBegin = sampleBegin;
    wp.seedLo = (uint32_t) seed; wp.seedHi = (uint32_t) (seed >> 32);
    wp.diffScale = 1.0f / sqrtf((float) spp);
    const uint64_t total = wp.pixPadded * (uint64_t) (sampleEnd - sampleBegin);
    CKW(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 8, stream));
    CKW(cudaMemsetAsync(errFlag, 0, sizeof(int), stream));
    uint32_t *initSlot = counters + 8 * CP_CTR_RING;
    for (uint64_t base = 0; base < total; base += waveSize) {
        const uint32_t n = (uint32_t) std::min<uint64_t>(waveSize, total - base);
        wp.waveBase = base;
        begin(3); launch_raygen(S, wp, q[0], liAcc, n, initSlot, stream); end();
        rs.launches++; rs.rays += n;
        // host view of the queue lengths: exact for the bounces whose counters have been read, upper bounds for the rest
        uint32_t ubClosest = n, ubShadow = 0;
        int cur = 0, lastRead = -1, launched = -1;
        auto readSlot = [&](int b) -> const uint32_t * {      // counters of bounce b
            if (cudaEventQuery(slotEvent[b % CP_CTR_RING]) != cudaSuccess) { cudaEventSynchronize(slotEvent[b % CP_CTR_RING]); rs.hostSyncs++; }
            const uint32_t *h = hCounters + 8 * (b % CP_CTR_RING);
            rs.rays += h[0]; rs.shadowRays += h[1] + h[4]; rs.shadowRaysTraced += h[1];   // h[4]: shadow rays the reference traces although they cannot contribute (see k_shade): counted, not traced
            return h;
        };
        for (int bounce = 0; ; ++bounce) {
            if (bounce >= 1) {
                // Large bounces (milliseconds of device time) are sized exactly: waiting for the counters of the bounce before costs
                // less than sorting the slack.  Small ones run ahead on the bounds of the bounce before that.
                const int want = (bounce == 1 || ubClosest > runAheadMax) ? bounce - 1 : bounce - 2;     // bounce 1: the bound is every camera path, 3-4x the paths that hit something
                bool stop = false;
                while (lastRead < want) {
                    const uint32_t *h = readSlot(++lastRead);
                    if (lastRead == bounce - 1) { ubClosest = h[0]; ubShadow = h[1]; stop = h[0] == 0u && h[1] == 0u; }
                    else { ubClosest = h[0]; ubShadow = h[0]; stop = h[0] == 0u; }       // bounce-1 is queued and will trace the shadow rays left over
                }
                if (stop) break;
            }
            uint32_t *slot = counters + 8 * (bounce % CP_CTR_RING);
            const uint32_t *prevSlot = bounce == 0 ? initSlot : counters + 8 * ((bounce - 1) % CP_CTR_RING);
            CKW(cudaMemsetAsync(slot, 0, sizeof(uint32_t) * 8, stream));
            // camera rays leave raygen in pixel order (already coherent); later bounces are re-ordered
            const uint32_t *perm = nullptr;
            if (bounce > 0) { begin(2); perm = coherence_order(S, q[cur], sq, prevSlot, ubClosest, ubShadow, stream); end(); if (perm) rs.launches += 2; }
            begin(0);
            {
                WaveIO io{q[cur], hitPT, hitPrim, sq, liAcc, perm, 0u};
                const uint64_t work = (uint64_t) ubClosest + ubShadow;
                const unsigned need = (unsigned) std::min<uint64_t>((work + CP_TRACE_THREADS - 1) / CP_TRACE_THREADS, fullGrid);
                traceKernel<<<need ? need : 1u, CP_TRACE_THREADS, 0, stream>>>(S, io, prevSlot, slot + 2, stats, errFlag);
            }
            end();
            begin(1);
            if (ubClosest) (fastMath ? launch_shade_fast : launch_shade)(S, wp, q[cur], prevSlot, ubClosest, hitPT, hitPrim, q[cur ^ 1], sq, liAcc, slot, stats + 4, stream);
            end();
            CKW(cudaMemcpyAsync(hCounters + 8 * (bounce % CP_CTR_RING), slot, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, stream));
            CKW(cudaEventRecord(slotEvent[bounce % CP_CTR_RING], stream));
            rs.launches += ubClosest ? 2 : 1; rs.bounces++;
            launched = bounce;
            cur ^= 1;
            if (cancelRequested.load(std::memory_order_relaxed)) {      // the unfinished wave is dropped: the film holds whole waves only
                cudaStreamSynchronize(stream);
                dropSpans();
                err = "render cancelled";
                return false;
            }
            if (ubClosest == 0u) break;        // that launch only traced the last shadow rays
            ubShadow = ubClosest;              // a path emits at most one shadow ray
        }
        while (lastRead < launched) readSlot(++lastRead);
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
    if (herr == 2) { dropSpans(); err = "Lookup dimension exceeds the direction number table size! You may have to reduce the 'maxDepth' parameter of your integrator."; return false; }   // sobol.cpp:222-224
    if (herr) { dropSpans(); err = "BVH traversal stack overflow"; return false; }
    for (auto &sp : spans) {
        float ms = 0; cudaEventElapsedTime(&ms, sp.a, sp.b);
        rs.stageMs[sp.stage] += ms; rs.stageLaunches[sp.stage]++;
    }
    if (trace && spans.size() > 1) {      // where the device waited: idle time between consecutive launches, and how late the host issued the later one
        static const char *names[] = {"trace", "shade", "sort", "raygen", "splat", "clear"};
        double idle = 0; int shown = 0;
        for (size_t k = 1; k < spans.size(); ++k) {
            float gap = 0, sinceStart = 0; cudaEventElapsedTime(&gap, spans[k - 1].b, spans[k].a); cudaEventElapsedTime(&sinceStart, spans[0].a, spans[k - 1].b);
            if (gap < 0.05f) continue;
            idle += gap;
            if (gap >= 0.5f && shown++ < 24)
                fprintf(stderr, "[cudapath] idle %.2f ms before launch %zu (%s after %s); the device was free at %.2f ms, the host issued the launch at %.2f ms\n", gap, k,
                        names[spans[k].stage], names[spans[k - 1].stage], sinceStart, (spans[k].tHost - spans[0].tHost) * 1e3);
        }
        float first2last = 0; cudaEventElapsedTime(&first2last, spans.front().a, spans.back().b);
        fprintf(stderr, "[cudapath] profiled render: %zu launches, device idle between launches %.2f ms, first launch to last %.2f ms; host: entry to first launch %.2f ms, "
                        "first launch to last issued %.2f ms, last issued to synchronised %.2f ms\n", spans.size(), idle, first2last,
                (spans.front().tHost - tr0) * 1e3, (spans.back().tHost - spans.front().tHost) * 1e3, (now() - spans.back().tHost) * 1e3);
    }
    dropSpans();
    rs.nodesVisited += hs[0]; rs.primsTested += hs[1]; rs.shadowNodesVisited += hs[2]; rs.shadowPrimsTested += hs[3]; rs.unsupportedLookups += hs[4]; rs.droppedSamples += hs[5];
    rs.fullTests += hs[6]; rs.shadowFullTests += hs[7];
    // the invalid padding "paths" (image sizes that are not multiples of 8) are not camera paths
    const uint64_t padPix = wp.pixPadded - realPix;
    rs.paths -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    rs.rays -= padPix * (uint64_t) (sampleEnd - sampleBegin);
    return true;
}

} // namespace cp

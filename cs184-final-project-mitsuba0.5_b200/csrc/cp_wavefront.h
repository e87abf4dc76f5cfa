// cp_wavefront.h -- queues and host driver of the wavefront path tracer
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <string>
#include "cp_scene.cuh"

namespace cp {

// SoA path queue (one float4 / uint2 stream per field, 56 B per path)
struct PathQueue {
    float4 *ro = nullptr;   // ray origin xyz, mint
    float4 *rd = nullptr;   // ray direction xyz, maxt
    float4 *thr = nullptr;  // throughput rgb, pdf of the BSDF sample that produced the ray
    uint2 *id = nullptr;    // wave-local path index, flags | depth
};
struct ShadowQueue {
    float4 *o = nullptr;    // origin xyz, mint
    float4 *d = nullptr;    // direction xyz, maxt
    float4 *c = nullptr;    // contribution rgb if unoccluded, path index (bits)
};
struct WaveParams {
    uint64_t waveBase = 0, pixPadded = 0;
    uint32_t filmW = 0, filmH = 0, tilesX = 0, sampleBegin = 0, seedLo = 0, seedHi = 0;
    float diffScale = 1.0f;
    // pixel shard (cudapath_set_pixel_shard): this context renders the pixel blocks owned by shard `shardIndex` of `shardCount`; a block is
    // (8 << blockShift)^2 pixels = 4^blockShift tiles of 8x8.  The block grid is cut into cells of cellW x cellH = shardCount blocks; every
    // shard owns exactly one block of every cell (which one rotates from cell to cell), so each shard's pixels are spread evenly over the
    // image whatever the image shows.
    uint32_t shardIndex = 0, shardCount = 1, cellW = 1, cellH = 1, cellsPerRow = 0, blockShift = 2;
};
struct RenderStats {
    uint64_t paths = 0, rays = 0, shadowRays = 0, launches = 0, bounces = 0, hostSyncs = 0;
    uint64_t shadowRaysTraced = 0;      // shadowRays minus the ones whose emitter sample cannot contribute (contribution exactly zero)
    uint64_t nodesVisited = 0, primsTested = 0, shadowNodesVisited = 0, shadowPrimsTested = 0, unsupportedLookups = 0, droppedSamples = 0;
    uint64_t fullTests = 0, shadowFullTests = 0;     // exact (FP64 cylinder / Wald triangle) tests that survived the fp32 pre-test
    double stageMs[6] = {0, 0, 0, 0, 0, 0};       // trace (closest + occlusion rays in one launch), shade, coherence sort (keys + radix sort), raygen, splat, film clear
    uint64_t stageLaunches[6] = {0, 0, 0, 0, 0, 0};
};

#define CP_CTR_RING 4
struct Wavefront {
    PathQueue q[2];
    ShadowQueue sq;
    float4 *hitPT = nullptr; uint32_t *hitPrim = nullptr;
    float4 *liAcc = nullptr;
    // Queue counters: a ring of CP_CTR_RING slots of 8 words (one per bounce) + an init slot written by k_raygen.  Slot of bounce b:
    // [0] paths that survive shade(b) = closest-hit rays of bounce b+1, [1] shadow rays emitted by shade(b), [2] work counter of the
    // trace launch of bounce b, [4] shadow rays counted but not traced (zero contribution).  Every kernel reads its input counts from
    // the previous slot ON THE DEVICE; the host only follows two bounces behind (pinned mirror `hCounters`, one event per slot) to size
    // the next launches and to notice the end of the wave -- it never waits for the bounce it has just launched.
    uint32_t *counters = nullptr, *hCounters = nullptr;
    cudaEvent_t slotEvent[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long *stats = nullptr;
    int *errFlag = nullptr;
    uint32_t *sortKeys[2] = {nullptr, nullptr}, *sortVals[2] = {nullptr, nullptr};
    void *sortTemp = nullptr; size_t sortTempBytes = 0;
    bool sortRays = true;
    uint32_t shardIndex = 0, shardCount = 1, shardBlockShift = 2;    // pixel shard of this context (WaveParams); block side = 8 << shift pixels
    bool fastMath = true;               // shade with the -DCP_FAST_MATH build of cp_shade.cu (cudapath_set_math_mode)
    // Bounces with more rays than this are sized from exact counters (one host wait each), smaller ones run ahead on the bounds of the bounce before.
    // The first bounce after the camera rays is always sized exactly: there the bound (every camera path) is 3-4x the queue (paths that hit
    // something) and sorting the slack costs 1.5 % of the step, against one host round trip per wave.  With the round-1 / early
    // round-2 default of 2^22 (six waits per wave) a host that answers late stalls the device once per bounce: 233-243 ms per step against
    // 214 ms of device time on two of five GPU boxes of the shared pool.
    uint32_t runAheadMax = 1u << 25;
    // Integrator::cancel() (include/mitsuba/render/integrator.h:76-84) may be called from another thread while render() blocks: the flag
    // is looked at once per bounce (after the host has read the queue counters); a cancelled render returns false with "render cancelled".
    std::atomic<int> cancelRequested{0}, inRender{0};
    // RenderJob progress (ProgressReporter in SamplingIntegrator::render, src/librender/integrator.cpp:95-138): called on the rendering
    // thread after every finished wave with the camera paths done so far and the total of this call
    void (*progress)(void *user, uint64_t done, uint64_t total) = nullptr; void *progressUser = nullptr;
    const uint32_t *coherence_order(const SceneDev &S, const PathQueue &q, const ShadowQueue &sq, const uint32_t *prevSlot, uint32_t ubClosest, uint32_t ubShadow, cudaStream_t stream);
    uint32_t capacity = 0;
    cudaStream_t allocStream = nullptr;
    bool reserve(uint32_t waveSize, cudaStream_t stream, std::string &err);
    void release();
    bool render(const SceneDev &S, uint32_t spp, uint64_t seed, uint32_t sampleBegin, uint32_t sampleEnd, float *d_film,
                uint32_t waveSize, bool collectStats, bool profileStages, cudaStream_t stream, RenderStats &rs, std::string &err);
    ~Wavefront() { release(); }
};

// shared by the stage kernels ----------------------------------------------------------------------------------------
// flags word: bits 0..15 depth, bit 16 first (camera) segment, bit 17 last BSDF sample was EDelta, bit 18 invalid (padding pixel)
// F_SCATTERED: some vertex of the path sampled a component other than ENull (`scattered` of path.cpp:125,213)
enum : uint32_t { F_FIRST = 1u << 16, F_DELTA = 1u << 17, F_INVALID = 1u << 18, F_SCATTERED = 1u << 19 };

#ifdef __CUDACC__
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
    samp = wp.sampleBegin + (uint32_t) s;
    if (wp.shardCount == 1u) {
        x = (tile % wp.tilesX) * 8u + (within & 7u);
        y = (tile / wp.tilesX) * 8u + (within >> 3);
    } else {      // owned block (one per cell) -> 8x8 tile inside it -> pixel
        const uint32_t bs = wp.blockShift, cell = tile >> (2u * bs), t = tile & ((1u << (2u * bs)) - 1u);
        const uint32_t cy = cell / wp.cellsPerRow, cx = cell - cy * wp.cellsPerRow;
        const uint32_t slot = (wp.shardIndex + cx + 3u * cy) % wp.shardCount;          // which block of the cell is ours
        const uint32_t bx = cx * wp.cellW + slot % wp.cellW, by = cy * wp.cellH + slot / wp.cellW;
        x = ((bx << bs) + (t & ((1u << bs) - 1u))) * 8u + (within & 7u);
        y = ((by << bs) + (t >> bs)) * 8u + (within >> 3);
    }
    return x < wp.filmW && y < wp.filmH;
}
#endif

// cp_shade.cu (built with -fmad=false)
void launch_raygen(const SceneDev &S, const WaveParams &wp, PathQueue q, float4 *liAcc, uint32_t n, uint32_t *initSlot, cudaStream_t stream);
// nPtr: device word holding the number of paths in `in` (<= nUpper, which only sizes the grid)
void launch_shade(const SceneDev &S, const WaveParams &wp, PathQueue in, const uint32_t *nPtr, uint32_t nUpper, const float4 *hitPT, const uint32_t *hitPrim, PathQueue out,
                  ShadowQueue sq, float4 *liAcc, uint32_t *counters, unsigned long long *unsupportedLookups, cudaStream_t stream);
void launch_shade_fast(const SceneDev &S, const WaveParams &wp, PathQueue in, const uint32_t *nPtr, uint32_t nUpper, const float4 *hitPT, const uint32_t *hitPrim, PathQueue out,
                       ShadowQueue sq, float4 *liAcc, uint32_t *counters, unsigned long long *unsupportedLookups, cudaStream_t stream);      // the -DCP_FAST_MATH build of cp_shade.cu
void launch_splat(const SceneDev &S, const WaveParams &wp, const float4 *liAcc, uint32_t n, float *film, unsigned long long *dropped, cudaStream_t stream);
bool splat_batch(const SceneDev &S, const float *d_pos, const float *d_rgb, const float *d_alpha, uint64_t n, float *d_film, cudaStream_t stream, std::string &err);

// cp_batch.cu -- parity hooks / stage micro-benchmarks on device-resident batches
bool bsdf_eval_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err, bool discrete = false, const float *d_uv = nullptr);
bool bsdf_sample_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_sample, const float *d_extra /* 4 per tuple or null */, float *d_wo, float *d_weight, float *d_pdf, int32_t *d_type, cudaStream_t s, std::string &err, const float *d_uv = nullptr);
bool bsdf_eval_world_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_frames, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err);
bool bsdf_eval_world_batch_fast(const SceneDev &S, int bsdf, uint64_t n, const float *d_frames, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err);
bool bsdf_eval_batch_fast(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err, bool discrete = false, const float *d_uv = nullptr);
bool bsdf_sample_batch_fast(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_sample, const float *d_extra, float *d_wo, float *d_weight, float *d_pdf, int32_t *d_type, cudaStream_t s, std::string &err, const float *d_uv = nullptr);
bool env_eval_batch_fast(const SceneDev &S, uint64_t n, const float *d_dir, float *d_rgb, float *d_pdf, cudaStream_t s, std::string &err);
bool env_sample_batch_fast(const SceneDev &S, uint64_t n, const float *d_ref, const float *d_sample, float *d_dir, float *d_value, float *d_pdfDist, cudaStream_t s, std::string &err);
bool intersect_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_mint, const float *d_maxt, int anyHit, bool stats,
                     int32_t *d_shape, uint32_t *d_prim, float *d_t, float *d_rec /*15 floats per ray or null*/, unsigned long long *d_stats, cudaStream_t s, std::string &err,
                     float *d_uv = nullptr /* its.uv + geometric normal, 5 floats per ray (needs d_rec) */);
void fill_records_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_t, const int32_t *d_shape, float *d_rec, float *d_uv, cudaStream_t s);
bool env_eval_batch(const SceneDev &S, uint64_t n, const float *d_dir, float *d_rgb, float *d_pdf, cudaStream_t s, std::string &err);
bool env_eval_filtered_batch(const SceneDev &S, uint64_t n, const float *d_dir, const float *d_rx, const float *d_ry, float *d_rgb, cudaStream_t s, std::string &err);
bool env_sample_batch(const SceneDev &S, uint64_t n, const float *d_ref, const float *d_sample, float *d_dir, float *d_value, float *d_pdfDist, cudaStream_t s, std::string &err);
bool read_bandwidth_probe(size_t bytes, int iters, cudaStream_t s, double &gbs, std::string &err);
bool camera_rays_batch(const SceneDev &S, uint64_t n, const float *d_pxy, float *d_o, float *d_d, float *d_minmax, cudaStream_t s, std::string &err);

} // namespace cp

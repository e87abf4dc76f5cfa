// cp_wavefront.h -- queues and host driver of the wavefront path tracer
#pragma once
#include <cuda_runtime.h>
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
};
struct RenderStats {
    uint64_t paths = 0, rays = 0, shadowRays = 0, launches = 0, bounces = 0;
    uint64_t nodesVisited = 0, primsTested = 0, unsupportedLookups = 0, droppedSamples = 0;
};

struct Wavefront {
    PathQueue q[2];
    ShadowQueue sq;
    float4 *hitPT = nullptr; uint32_t *hitPrim = nullptr;
    float4 *liAcc = nullptr;
    uint32_t *counters = nullptr, *hCounters = nullptr;
    unsigned long long *stats = nullptr;
    int *errFlag = nullptr;
    uint32_t capacity = 0;
    bool reserve(uint32_t waveSize, std::string &err);
    void release();
    bool render(const SceneDev &S, uint32_t spp, uint64_t seed, uint32_t sampleBegin, uint32_t sampleEnd, float *d_film,
                uint32_t waveSize, bool collectStats, cudaStream_t stream, RenderStats &rs, std::string &err);
    ~Wavefront() { release(); }
};

bool splat_batch(const SceneDev &S, const float *d_pos, const float *d_rgb, const float *d_alpha, uint64_t n, float *d_film, cudaStream_t stream, std::string &err);

// cp_batch.cu -- parity hooks / stage micro-benchmarks on device-resident batches
bool bsdf_eval_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err);
bool bsdf_sample_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_sample, float *d_wo, float *d_weight, float *d_pdf, int32_t *d_type, cudaStream_t s, std::string &err);
bool intersect_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_mint, const float *d_maxt, int anyHit, bool stats,
                     int32_t *d_shape, uint32_t *d_prim, float *d_t, float *d_rec /*15 floats per ray or null*/, unsigned long long *d_stats, cudaStream_t s, std::string &err);
bool env_eval_batch(const SceneDev &S, uint64_t n, const float *d_dir, float *d_rgb, float *d_pdf, cudaStream_t s, std::string &err);
bool env_sample_batch(const SceneDev &S, uint64_t n, const float *d_ref, const float *d_sample, float *d_dir, float *d_value, float *d_pdfDist, cudaStream_t s, std::string &err);
bool camera_rays_batch(const SceneDev &S, uint64_t n, const float *d_pxy, float *d_o, float *d_d, float *d_minmax, cudaStream_t s, std::string &err);

} // namespace cp

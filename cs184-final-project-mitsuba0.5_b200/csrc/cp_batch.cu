// cp_batch.cu -- per-stage batch kernels: parity hooks and the config-5 stage micro-benchmarks (sm_100a).
//
// (ray queries; the BSDF / emitter / sensor batches are in cp_batch_shade.cu, built with -fmad=false)
// These run exactly the device functions the wavefront uses (cp_bsdf.cuh, cp_traverse.cuh, cp_env.cuh,
// cp_camera.cuh) on flat device-resident batches, mirroring the reference interfaces
//   BSDF::eval / pdf / sample                 include/mitsuba/render/bsdf.h:369-441
//   Scene::rayIntersect (closest / shadow)    include/mitsuba/render/scene.h:187-189, src/librender/skdtree.cpp:112-142,207-226
//   Emitter::evalEnvironment / sampleDirect / pdfDirect   src/emitters/envmap.cpp:380-410,516-556
//   Sensor::sampleRayDifferential             src/sensors/perspective.cpp:271-298
// Input streams are read as coalesced fp32 arrays; one thread per tuple / ray.
#include "cp_host.h"
#include "cp_traverse.cuh"
#include "cp_wavefront.h"

namespace cp {

template <bool ANY, bool STATS>
__global__ void __launch_bounds__(128) k_intersect_batch(SceneDev S, uint64_t n, const float *__restrict__ o, const float *__restrict__ d,
                                                         const float *__restrict__ mint, const float *__restrict__ maxt,
                                                         int32_t *shape, uint32_t *prim, float *tOut, float *rec, unsigned long long *stats, int *errFlag) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 ro(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    RayHit h; uint32_t nv = 0, np = 0; int ovf = 0;
    const bool hit = traverse<ANY, STATS>(S, ro, rd, mint[i], maxt[i], h, nv, np, ovf);
    if (hit) {
        const float4 v1 = __ldg(S.vtx + h.gv);
        const uint32_t sh = vtx_shape(v1);
        shape[i] = (int32_t) sh; prim[i] = h.gv - S.shapes[sh].vertexOffset; tOut[i] = h.t;
        if (rec) {   // raw hit for the (un-fused) record kernel in cp_batch_shade.cu: stored hit point + global primitive
            float *p = rec + 15 * i;
            p[0] = h.p.x; p[1] = h.p.y; p[2] = h.p.z; p[3] = __uint_as_float(h.gv);
        }
    } else {
        shape[i] = -1; prim[i] = 0xffffffffu; tOut[i] = CP_INF;
        if (rec) for (int k = 0; k < 15; ++k) rec[15 * i + k] = 0.0f;
    }
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 0, (unsigned long long) nv); atomicAdd(stats + 1, (unsigned long long) np); }
}

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)
static inline unsigned grid_for(uint64_t n, int block) { return (unsigned) ((n + block - 1) / block); }

bool intersect_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_mint, const float *d_maxt, int anyHit, bool stats,
                     int32_t *d_shape, uint32_t *d_prim, float *d_t, float *d_rec, unsigned long long *d_stats, cudaStream_t s, std::string &err) {
    if (n == 0) return true;
    int *d_err = nullptr;
    CKB(cudaMalloc(&d_err, sizeof(int))); CKB(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    const unsigned g = grid_for(n, 128);
    if (anyHit) {
        if (stats) k_intersect_batch<true, true><<<g, 128, 0, s>>>(S, n, d_o, d_d, d_mint, d_maxt, d_shape, d_prim, d_t, d_rec, d_stats, d_err);
        else k_intersect_batch<true, false><<<g, 128, 0, s>>>(S, n, d_o, d_d, d_mint, d_maxt, d_shape, d_prim, d_t, d_rec, d_stats, d_err);
    } else {
        if (stats) k_intersect_batch<false, true><<<g, 128, 0, s>>>(S, n, d_o, d_d, d_mint, d_maxt, d_shape, d_prim, d_t, d_rec, d_stats, d_err);
        else k_intersect_batch<false, false><<<g, 128, 0, s>>>(S, n, d_o, d_d, d_mint, d_maxt, d_shape, d_prim, d_t, d_rec, d_stats, d_err);
    }
    if (d_rec && !anyHit) fill_records_batch(S, n, d_d, d_shape, d_rec, s);
    int herr = 0;
    CKB(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKB(cudaStreamSynchronize(s));
    cudaFree(d_err);
    CKB(cudaGetLastError());
    if (herr) { err = "BVH traversal stack overflow"; return false; }
    return true;
}
} // namespace cp

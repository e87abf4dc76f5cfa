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

// ray source / sink over flat fp32 arrays
struct BatchIO {
    const float *o, *d, *mint, *maxt; int32_t *shape; uint32_t *prim; float *tOut; float *rec;
    const float4 *vtx; const ShapeDev *shapes; const float4 *triAccel; const float4 *rects;
    CP_D bool load(uint32_t i, V3 &ro, V3 &rd, float &mn, float &mx, bool &, uint32_t &) const {
        ro = V3(o[3 * (size_t) i], o[3 * (size_t) i + 1], o[3 * (size_t) i + 2]); rd = V3(d[3 * (size_t) i], d[3 * (size_t) i + 1], d[3 * (size_t) i + 2]);
        mn = mint[i]; mx = maxt[i];
        return true;
    }
    CP_D void store(uint32_t i, bool, bool hit, const RayHit &h) const {
        if (hit) {
            if (h.gv & CP_RECT_FLAG) {  // rectangle: one primitive per shape
                shape[i] = (int32_t) __float_as_uint(__ldg(rects + CP_RECT_STRIDE * (size_t) (h.gv & CP_PRIM_MASK) + 3).w); prim[i] = 0u;
            } else if (h.gv & CP_TRI_FLAG) {   // triangle: shape / primitive index come from the TriAccel record (skdtree.h:296-299)
                const float4 C = __ldg(triAccel + 3 * (size_t) (h.gv & CP_PRIM_MASK) + 2);
                shape[i] = (int32_t) __float_as_uint(C.z); prim[i] = __float_as_uint(C.w);
            } else {
                const uint32_t sh = vtx_shape(__ldg(vtx + h.gv));
                shape[i] = (int32_t) sh; prim[i] = h.gv - shapes[sh].vertexOffset;
            }
            tOut[i] = h.t;
            if (rec) {   // raw hit for the (un-fused) record kernel in cp_batch_shade.cu: stored hit point + global primitive
                float *p = rec + 15 * (size_t) i;
                p[0] = h.p.x; p[1] = h.p.y; p[2] = h.p.z; p[3] = __uint_as_float(h.gv);
            }
        } else {
            shape[i] = -1; prim[i] = 0xffffffffu; tOut[i] = CP_INF;
            if (rec) for (int k = 0; k < 15; ++k) rec[15 * (size_t) i + k] = 0.0f;
        }
    }
};
template <bool ANY, bool STATS, bool MESH>
__global__ void __launch_bounds__(CP_TRACE_THREADS, CP_MIN_BLOCKS) k_intersect_batch(SceneDev S, BatchIO io, uint32_t n, uint32_t *rayCounter, unsigned long long *stats, int *errFlag) {
    TraceCounters tc[2] = {{0, 0, 0}, {0, 0, 0}}; int ovf = 0;
    trace_persistent<ANY ? TRACE_ANY : TRACE_CLOSEST, STATS, MESH>(S, io, n, rayCounter, tc, ovf);
    if (ovf) *errFlag = 1;
    if (STATS) { atomicAdd(stats + 0, tc[ANY ? 1 : 0].nodes); atomicAdd(stats + 1, tc[ANY ? 1 : 0].prims); atomicAdd(stats + 2, tc[ANY ? 1 : 0].fullTests); }
}

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)
static inline unsigned grid_for(uint64_t n, int block) { return (unsigned) ((n + block - 1) / block); }

bool intersect_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_mint, const float *d_maxt, int anyHit, bool stats,
                     int32_t *d_shape, uint32_t *d_prim, float *d_t, float *d_rec, unsigned long long *d_stats, cudaStream_t s, std::string &err, float *d_uv) {
    if (n == 0) return true;
    if (n > 0xfffffff0ull) { err = "ray batch too large (split it into chunks below 2^32 rays)"; return false; }
    int dev = 0, numSMs = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
    int *d_ctl = nullptr;   // [0] overflow flag, [1] persistent ray counter; from the caching allocator, returned on every path out of here
    CKB(dev_alloc(&d_ctl, 2 * sizeof(int)));
    struct CtlGuard { int *p; cudaStream_t s; ~CtlGuard() { cudaStreamSynchronize(s); dev_free(p); } } ctlGuard{d_ctl, s};
    CKB(cudaMemsetAsync(d_ctl, 0, 2 * sizeof(int), s));
    const unsigned need = (unsigned) ((n + 127) / 128), g = need < (unsigned) numSMs * 8u ? need : (unsigned) numSMs * 8u;
    BatchIO io{d_o, d_d, d_mint, d_maxt, d_shape, d_prim, d_t, anyHit ? nullptr : d_rec, S.vtx, S.shapes, S.mesh.triAccel, S.mesh.rects};
    uint32_t *ctr = (uint32_t *) (d_ctl + 1);
#define CP_LAUNCH_BATCH(AN, ST, ME) k_intersect_batch<AN, ST, ME><<<g, CP_TRACE_THREADS, 0, s>>>(S, io, (uint32_t) n, ctr, d_stats, d_ctl)
    const int variant = (anyHit ? 4 : 0) | (stats ? 2 : 0) | (S.mesh.triCount + S.mesh.rectCount > 0 ? 1 : 0);
    switch (variant) {
        case 0: CP_LAUNCH_BATCH(false, false, false); break; case 1: CP_LAUNCH_BATCH(false, false, true); break;
        case 2: CP_LAUNCH_BATCH(false, true, false); break;  case 3: CP_LAUNCH_BATCH(false, true, true); break;
        case 4: CP_LAUNCH_BATCH(true, false, false); break;  case 5: CP_LAUNCH_BATCH(true, false, true); break;
        case 6: CP_LAUNCH_BATCH(true, true, false); break;   default: CP_LAUNCH_BATCH(true, true, true); break;
    }
    if (d_rec && !anyHit) fill_records_batch(S, n, d_o, d_d, d_t, d_shape, d_rec, d_uv, s);
    int herr = 0;
    CKB(cudaMemcpyAsync(&herr, d_ctl, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKB(cudaStreamSynchronize(s));
    CKB(cudaGetLastError());
    if (herr) { err = "BVH traversal stack overflow"; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------ bandwidth probe
// Roofline denominators measured on the device the bench runs on: every thread streams 16-byte vectors of a resident buffer through
// L2 (ld.global.cg bypasses L1).  A buffer well below the L2 size measures L2 read bandwidth, a multi-GB one HBM read bandwidth.
__global__ void __launch_bounds__(256) k_read_probe(const float4 *__restrict__ buf, size_t n4, int iters, float *sink) {
    float acc = 0.0f;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}
bool read_bandwidth_probe(size_t bytes, int iters, cudaStream_t s, double &gbs, std::string &err) {
    if (bytes < (1u << 20) || iters < 1) { err = "probe needs at least 1 MiB and one iteration"; return false; }
    int dev = 0, numSMs = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
    float4 *buf = nullptr; float *sink = nullptr;
    CKB(dev_alloc(&buf, bytes)); CKB(dev_alloc(&sink, sizeof(float)));
    CKB(cudaMemsetAsync(buf, 0, bytes, s));
    const size_t n4 = bytes / sizeof(float4);
    cudaEvent_t e0, e1; CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
    k_read_probe<<<numSMs * 8, 256, 0, s>>>(buf, n4, 1, sink);              // warm-up: the buffer becomes L2-resident if it fits
    float best = 0.0f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, s);
        k_read_probe<<<numSMs * 8, 256, 0, s>>>(buf, n4, iters, sink);
        cudaEventRecord(e1, s);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) { err = std::string("probe kernel: ") + cudaGetErrorString(e); break; }
        const float g = (float) ((double) n4 * sizeof(float4) * iters / (ms * 1e-3) / 1e9);
        if (g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamSynchronize(s);
    dev_free(buf); dev_free(sink);
    gbs = best;
    return err.empty();
}
} // namespace cp

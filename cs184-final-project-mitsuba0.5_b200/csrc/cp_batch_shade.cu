// cp_batch_shade.cu -- per-stage batch kernels: parity hooks and the config-5 stage micro-benchmarks (sm_100a).
//
// These run exactly the device functions the wavefront uses (cp_bsdf.cuh, cp_traverse.cuh, cp_env.cuh,
// cp_camera.cuh) on flat device-resident batches, mirroring the reference interfaces
//   BSDF::eval / pdf / sample                 include/mitsuba/render/bsdf.h:369-441
//   Scene::rayIntersect (closest / shadow)    include/mitsuba/render/scene.h:187-189, src/librender/skdtree.cpp:112-142,207-226
//   Emitter::evalEnvironment / sampleDirect / pdfDirect   src/emitters/envmap.cpp:380-410,516-556
//   Sensor::sampleRayDifferential             src/sensors/perspective.cpp:271-298
// Input streams are read as coalesced fp32 arrays; one thread per tuple / ray.
//
// Compiled twice like cp_shade.cu: the strict build emits everything; the -DCP_FAST_MATH build emits the BSDF and emitter batches
// under *_fast names (the math mode of the context picks one).
#include "cp_host.h"
#include "cp_env.cuh"
#include "cp_camera.cuh"
#include "cp_wavefront.h"
#ifdef CP_FAST_MATH
#define k_bsdf_eval k_bsdf_eval_fast
#define k_bsdf_sample k_bsdf_sample_fast
#define k_env_eval k_env_eval_fast
#define k_env_sample k_env_sample_fast
#define k_bsdf_eval_world k_bsdf_eval_world_fast
#define bsdf_eval_world_batch bsdf_eval_world_batch_fast
#define bsdf_eval_batch bsdf_eval_batch_fast
#define bsdf_sample_batch bsdf_sample_batch_fast
#define env_eval_batch env_eval_batch_fast
#define env_sample_batch env_sample_batch_fast
#endif

namespace cp {

__global__ void __launch_bounds__(256) k_bsdf_eval(const BsdfDev *__restrict__ bsdfs, int bsdf, uint64_t n, const float *__restrict__ wi,
                                                   const float *__restrict__ wo, float *eval, float *pdf, bool discrete, const float *__restrict__ uv) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BsdfDev &b = bsdfs[bsdf];
    const V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), c(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    const V3 e = bsdf_eval(b, a, c, discrete, uv ? uv[2 * i] : 0.0f, uv ? uv[2 * i + 1] : 0.0f);
    eval[3 * i] = e.x; eval[3 * i + 1] = e.y; eval[3 * i + 2] = e.z;
    pdf[i] = bsdf_pdf(b, a, c, discrete);
}
// the same from WORLD-space directions and a shading frame per tuple (s, t, n): wi = frame.toLocal(wiWorld) as the integrator forms it
// (its.wi = its.toLocal(-ray.d), include/mitsuba/render/skdtree.h:426-427; bRec.wo = its.toLocal(d), src/integrators/path/path.cpp:186)
__global__ void __launch_bounds__(256) k_bsdf_eval_world(const BsdfDev *__restrict__ bsdfs, int bsdf, uint64_t n, const float *__restrict__ frames,
                                                         const float *__restrict__ wi, const float *__restrict__ wo, float *eval, float *pdf) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BsdfDev &b = bsdfs[bsdf];
    Frame f;
    f.s = V3(frames[9 * i], frames[9 * i + 1], frames[9 * i + 2]); f.t = V3(frames[9 * i + 3], frames[9 * i + 4], frames[9 * i + 5]); f.n = V3(frames[9 * i + 6], frames[9 * i + 7], frames[9 * i + 8]);
    const V3 a = f.toLocal(V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2])), c = f.toLocal(V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]));
    const V3 e = bsdf_eval(b, a, c, false);
    eval[3 * i] = e.x; eval[3 * i + 1] = e.y; eval[3 * i + 2] = e.z;
    pdf[i] = bsdf_pdf(b, a, c, false);
}
__global__ void __launch_bounds__(256) k_bsdf_sample(const BsdfDev *__restrict__ bsdfs, int bsdf, uint64_t n, const float *__restrict__ wi,
                                                     const float *__restrict__ sample, const float *__restrict__ extra, float *wo, float *weight, float *pdf, int32_t *type, const float *__restrict__ uv) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BsdfDev &b = bsdfs[bsdf];
    float4 ex = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (extra) ex = make_float4(extra[4 * i], extra[4 * i + 1], extra[4 * i + 2], extra[4 * i + 3]);
    const BsdfSampleOut r = bsdf_sample(b, V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), sample[2 * i], sample[2 * i + 1], ex, uv ? uv[2 * i] : 0.0f, uv ? uv[2 * i + 1] : 0.0f);
    wo[3 * i] = r.wo.x; wo[3 * i + 1] = r.wo.y; wo[3 * i + 2] = r.wo.z;
    weight[3 * i] = r.weight.x; weight[3 * i + 1] = r.weight.y; weight[3 * i + 2] = r.weight.z;
    pdf[i] = r.pdf; type[i] = r.type | (r.component << 8);
}

__global__ void k_env_eval(SceneDev S, uint64_t n, const float *__restrict__ dir, float *rgb, float *pdf) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    const V3 v = env_eval(S.env, d);
    rgb[3 * i] = v.x; rgb[3 * i + 1] = v.y; rgb[3 * i + 2] = v.z;
    pdf[i] = env_pdf_direct(S.env, d);
}
#ifndef CP_FAST_MATH
__global__ void k_env_eval_filtered(SceneDev S, uint64_t n, const float *__restrict__ dir, const float *__restrict__ rx, const float *__restrict__ ry, float *rgb) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 v = env_eval_filtered(S.env, V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), V3(rx[3 * i], rx[3 * i + 1], rx[3 * i + 2]), V3(ry[3 * i], ry[3 * i + 1], ry[3 * i + 2]), nullptr);
    rgb[3 * i] = v.x; rgb[3 * i + 1] = v.y; rgb[3 * i + 2] = v.z;
}
#endif
__global__ void k_env_sample(SceneDev S, uint64_t n, const float *__restrict__ ref, const float *__restrict__ sample, float *dir, float *value, float *pdfDist) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EnvSample r = env_sample_direct(S.env, V3(ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]), sample[2 * i], sample[2 * i + 1]);
    dir[3 * i] = r.d.x; dir[3 * i + 1] = r.d.y; dir[3 * i + 2] = r.d.z;
    value[3 * i] = r.value.x; value[3 * i + 1] = r.value.y; value[3 * i + 2] = r.value.z;
    pdfDist[2 * i] = r.pdf; pdfDist[2 * i + 1] = r.dist;
}
#ifndef CP_FAST_MATH
__global__ void k_camera_rays(SceneDev S, uint64_t n, const float *__restrict__ pxy, float *o, float *d, float *minmax) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CameraRay r = camera_ray(S.cam, pxy[2 * i], pxy[2 * i + 1], 1.0f);
    o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
    d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
    minmax[2 * i] = r.mint; minmax[2 * i + 1] = r.maxt;
}

// HairShape::fillIntersectionRecord + computeShadingFrame for the hits of k_intersect_batch (in place: rec[0..3] = p, gv)
// uvOut (optional): its.uv and its.geoFrame.n, 5 floats per ray
__global__ void k_fill_records(SceneDev S, uint64_t n, const float *__restrict__ o, const float *__restrict__ d, const float *__restrict__ t,
                               const int32_t *__restrict__ shape, float *rec, float *uvOut) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (shape[i] < 0) { if (uvOut) for (int k = 0; k < 5; ++k) uvOut[5 * i + k] = 0.0f; return; }
    float *p = rec + 15 * i;
    const uint32_t gv = __float_as_uint(p[3]);
    HitRecord r;
    if (gv & CP_RECT_FLAG) fill_intersection_rect(S.mesh, gv & CP_PRIM_MASK, p[0], p[1], t[i], V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), r);
    else if (gv & CP_TRI_FLAG) fill_intersection_mesh(S.mesh, S.shapes, gv & CP_PRIM_MASK, p[0], p[1], V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), r);
    else fill_intersection(__ldg(S.vtx + gv), __ldg(S.vtx + gv + 1), S.shapes[shape[i]].radius, V3(p[0], p[1], p[2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), r);
    p[0] = r.p.x; p[1] = r.p.y; p[2] = r.p.z; p[3] = r.sh.n.x; p[4] = r.sh.n.y; p[5] = r.sh.n.z;
    p[6] = r.sh.s.x; p[7] = r.sh.s.y; p[8] = r.sh.s.z; p[9] = r.sh.t.x; p[10] = r.sh.t.y; p[11] = r.sh.t.z;
    p[12] = r.wi.x; p[13] = r.wi.y; p[14] = r.wi.z;
    if (uvOut) { uvOut[5 * i] = r.u; uvOut[5 * i + 1] = r.v; uvOut[5 * i + 2] = r.geoN.x; uvOut[5 * i + 3] = r.geoN.y; uvOut[5 * i + 4] = r.geoN.z; }
}
void fill_records_batch(const SceneDev &S, uint64_t n, const float *d_o, const float *d_d, const float *d_t, const int32_t *d_shape, float *d_rec, float *d_uv, cudaStream_t s) {
    if (n) k_fill_records<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(S, n, d_o, d_d, d_t, d_shape, d_rec, d_uv);
}
#endif // !CP_FAST_MATH

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)
static inline unsigned grid_for(uint64_t n, int block) { return (unsigned) ((n + block - 1) / block); }

bool bsdf_eval_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err, bool discrete, const float *d_uv) {
    if (bsdf < 0 || bsdf >= S.bsdfCount) { err = "bsdf id out of range"; return false; }
    if (n == 0) return true;
    k_bsdf_eval<<<grid_for(n, 256), 256, 0, s>>>(S.bsdfs, bsdf, n, d_wi, d_wo, d_eval, d_pdf, discrete, d_uv);
    CKB(cudaGetLastError());
    return true;
}
bool bsdf_eval_world_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_frames, const float *d_wi, const float *d_wo, float *d_eval, float *d_pdf, cudaStream_t s, std::string &err) {
    if (bsdf < 0 || bsdf >= S.bsdfCount) { err = "bsdf id out of range"; return false; }
    if (n == 0) return true;
    k_bsdf_eval_world<<<grid_for(n, 256), 256, 0, s>>>(S.bsdfs, bsdf, n, d_frames, d_wi, d_wo, d_eval, d_pdf);
    CKB(cudaGetLastError());
    return true;
}
bool bsdf_sample_batch(const SceneDev &S, int bsdf, uint64_t n, const float *d_wi, const float *d_sample, const float *d_extra, float *d_wo, float *d_weight, float *d_pdf, int32_t *d_type, cudaStream_t s, std::string &err, const float *d_uv) {
    if (bsdf < 0 || bsdf >= S.bsdfCount) { err = "bsdf id out of range"; return false; }
    if (n == 0) return true;
    k_bsdf_sample<<<grid_for(n, 256), 256, 0, s>>>(S.bsdfs, bsdf, n, d_wi, d_sample, d_extra, d_wo, d_weight, d_pdf, d_type, d_uv);
    CKB(cudaGetLastError());
    return true;
}
bool env_eval_batch(const SceneDev &S, uint64_t n, const float *d_dir, float *d_rgb, float *d_pdf, cudaStream_t s, std::string &err) {
    if (!S.env.present) { err = "no environment map set"; return false; }
    if (n == 0) return true;
    k_env_eval<<<grid_for(n, 256), 256, 0, s>>>(S, n, d_dir, d_rgb, d_pdf);
    CKB(cudaGetLastError());
    return true;
}
#ifndef CP_FAST_MATH
bool env_eval_filtered_batch(const SceneDev &S, uint64_t n, const float *d_dir, const float *d_rx, const float *d_ry, float *d_rgb, cudaStream_t s, std::string &err) {
    if (!S.env.present) { err = "no environment map set"; return false; }
    if (n == 0) return true;
    k_env_eval_filtered<<<grid_for(n, 256), 256, 0, s>>>(S, n, d_dir, d_rx, d_ry, d_rgb);
    CKB(cudaGetLastError());
    return true;
}
#endif
bool env_sample_batch(const SceneDev &S, uint64_t n, const float *d_ref, const float *d_sample, float *d_dir, float *d_value, float *d_pdfDist, cudaStream_t s, std::string &err) {
    if (!S.env.present) { err = "no environment map set"; return false; }
    if (n == 0) return true;
    k_env_sample<<<grid_for(n, 256), 256, 0, s>>>(S, n, d_ref, d_sample, d_dir, d_value, d_pdfDist);
    CKB(cudaGetLastError());
    return true;
}
#ifndef CP_FAST_MATH
bool camera_rays_batch(const SceneDev &S, uint64_t n, const float *d_pxy, float *d_o, float *d_d, float *d_minmax, cudaStream_t s, std::string &err) {
    if (n == 0) return true;
    k_camera_rays<<<grid_for(n, 256), 256, 0, s>>>(S, n, d_pxy, d_o, d_d, d_minmax);
    CKB(cudaGetLastError());
    return true;
}
#endif // !CP_FAST_MATH

} // namespace cp

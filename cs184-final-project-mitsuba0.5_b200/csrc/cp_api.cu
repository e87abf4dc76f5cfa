// cp_api.cu -- implementation of the C ABI declared in include/cudapath.h
#include "../../include/cudapath.h"
#include "cp_host.h"
#include "cp_wavefront.h"
#include <cstring>
#include <cctype>
#include <cmath>
#include <algorithm>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <future>
#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only: libnccl.so.2 is bound at run time (cp_multi section), never at link time

using namespace cp;

static thread_local std::string g_lastError;
static int fail(const std::string &msg) { g_lastError = msg; return -1; }
#define CKA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

namespace {
struct EnvHost { std::vector<float> rgb; int w = 0, h = 0; float toWorld[16]; float scale = 1; bool present = false; };
struct CamHost { float toWorld[16]; float fov = 35, nearClip = 1e-2f, farClip = 1e4f; int w = 0, h = 0; bool present = false; };
struct BsdfHost { BsdfDev dev; MarschnerTables tables; float *rt = nullptr; };

// 4x4 helpers (row-major, fp32), arithmetic of the reference:
//   matrix product            include/mitsuba/core/matrix.h:743-757 (sum over k, starting from 0)
//   Matrix<4,4,float>::invert include/mitsuba/core/matrix.inl:138-193 (Gauss-Jordan with full pivoting, in place) -- what
//                             Transform(const Matrix4x4 &) runs on every matrix of a scene file and on Transform::perspective
void mat_mul(const float *a, const float *b, float *r) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0; for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j]; r[i * 4 + j] = s; }
}
bool mat_inv(const float *a, float *out) { return mat4_invert_f32(a, out); }      // cp_host_data.cpp (shared with the OBJ loader)
V3 h_xfm_point(const float *m, V3 p) {
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11], w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return V3(x, y, z);
    return V3(x, y, z) / w;
}
}

// The caller's current device is restored when an entry point returns (the library only borrows the thread's device binding).
struct DevGuard {
    int prev = -1; cudaError_t err;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); err = cudaSetDevice(dev); }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct cudapath_multi;
struct cudapath_ctx {
    int device = 0;
    // cudapath_create_multi: the context the caller holds is the one of devices[0]; `peers` are the contexts of the other devices.
    // Every scene-building call is repeated on them (so all devices hold the same flattened scene, ids included), cudapath_build
    // builds them concurrently and cudapath_render splits the sample range over all of them (cp_multi section below).
    std::vector<cudapath_ctx *> peers;
    cudapath_multi *multi = nullptr;
    cudaStream_t stream = nullptr;
    std::string dataDir;
    std::vector<BsdfHost> bsdfs;
    // per shape: raw caller arrays staged on the device (xyz fp32 triples + starts bytes); packed into `d_vtx` by build()
    // (meshes: xyz / optional normals fp32 triples + shape-local uint32 index triples)
    struct Staged { float *xyz = nullptr; uint8_t *starts = nullptr; uint32_t n = 0; float *nrm = nullptr; uint32_t *idx = nullptr; uint32_t nTris = 0; float *uv = nullptr; };
    std::vector<Staged> staged;
    uint32_t vtxTotal = 0, meshVtxTotal = 0, triTotal = 0;
    float4 *d_meshPos = nullptr, *d_meshNrm = nullptr, *d_triAccel = nullptr; uint32_t *d_meshIdx = nullptr; float2 *d_meshUV = nullptr;
    std::vector<float4> rects; float4 *d_rects = nullptr;
    int samplerKind = 0; uint64_t samplerScramble = 0; uint32_t sobolRows[2] = {0, 0};     // cudapath_set_sampler
    uint32_t *d_sobolM32 = nullptr; uint64_t *d_sobolVdc = nullptr, *d_sobolInv = nullptr;      // CP_RECT_STRIDE float4 per `rectangle` shape (cp_tri.cuh)
    std::vector<ShapeDev> shapes;
    EnvHost env; EnvTables envTables;
    CamHost cam;
    int filterType = 0; float filterParam = 0; int hasAlpha = 0;
    int filmHdr = 0; float filmGamma = -1.0f, filmExposure = 0.0f;   // how the film asks to be developed (ldrfilm.cpp:180-181); not used by the render
    IntegratorDev integ{-1, 5, 0, 0};
    // device
    float4 *d_vtx = nullptr; ShapeDev *d_shapes = nullptr; BsdfDev *d_bsdfs = nullptr;
    BVHDev bvh;
    SceneDev scene;
    bool built = false;
    Wavefront wf;
    uint32_t waveSize = 0; int collectStats = 0, profileStages = 0;     // 0 = sized from the free device memory at render time
    int maxSplit = getenv("CUDAPATH_MAX_SPLIT") ? std::max(1, std::min(64, atoi(getenv("CUDAPATH_MAX_SPLIT")))) : 16;    // cudapath_set_build_options; hair-curl: 286 / 299 / 311 / 322 / 328 Mpaths/s at 8 / 12 / 16 / 24 / 32 (build 19 ms per 32 M references)
    bool maxSplitExplicit = false;
    uint32_t shardIndex = 0, shardCount = 1;     // cudapath_set_pixel_shard
    float leafSplitCost = getenv("CUDAPATH_LEAF_SPLIT_COST") ? (float) atof(getenv("CUDAPATH_LEAF_SPLIT_COST")) : 1.0f;    // see k_collapse (cp_bvh.cu); < 0: leaves of up to CP_LEAF_MAX references, never opened
    int sortRays = getenv("CUDAPATH_NO_SORT") ? 0 : 1;
    // math mode of the shading stages (cudapath_set_math_mode): 1 = fast (default), 0 = strict; CUDAPATH_MATH=strict|fast sets the default
    int fastMath = (getenv("CUDAPATH_MATH") && std::string(getenv("CUDAPATH_MATH")) == "strict") ? 0 : 1;
    cudapath_stats stats{};
    float sceneAABB[6] = {0, 0, 0, 0, 0, 0};

    // Everything the context owns on the device comes from the caching allocator (cp_mem.cpp): a context that is created, built,
    // rendered and destroyed repeatedly never goes back to the driver for memory.  Renders may have run on a caller's stream
    // (cudapath_render_dev), so freeing first waits for the device.
    void dfree(const void *p) { dev_free(p); }
    void freeBuilt() {
        cudaDeviceSynchronize();
        dfree(d_vtx); dfree(d_shapes); dfree(d_bsdfs); dfree(bvh.nodes); dfree(bvh.prims); dfree(bvh.leafSeg);
        dfree(d_meshPos); dfree(d_meshNrm); dfree(d_triAccel); dfree(d_meshIdx); dfree(d_meshUV); dfree(d_rects);
        d_meshPos = d_meshNrm = d_triAccel = nullptr; d_meshIdx = nullptr; d_meshUV = nullptr; d_rects = nullptr;
        dfree(envTables.texels); dfree(envTables.cdfCols); dfree(envTables.cdfRows); dfree(envTables.rowWeights); dfree(envTables.mipTexels); dfree(envTables.mipInfo);
        d_vtx = nullptr; d_shapes = nullptr; d_bsdfs = nullptr; bvh = BVHDev(); envTables = EnvTables(); built = false;
    }
    ~cudapath_ctx() {
        DevGuard guard_(device);
        freeBuilt();                  // synchronises the device
        for (auto &st : staged) { dfree(st.xyz); dfree(st.starts); dfree(st.nrm); dfree(st.idx); dfree(st.uv); }
        dfree(d_sobolM32); dfree(d_sobolVdc); dfree(d_sobolInv);
        for (auto &b : bsdfs) { dfree(b.tables.tab); dfree(b.tables.cdf); dfree(b.tables.sums); dfree(b.tables.pdf); dfree(b.rt); }
        wf.release();
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    }
};

struct cudapath_hair_file { HairFileData data; };
struct cudapath_mesh_file { MeshFileData data; };

static int require_built(cudapath_ctx *ctx) {
    if (!ctx) return fail("null context");
    if (!ctx->built) return fail("cudapath_build() has not been called");
    return 0;
}
// Repeats a successful scene-building call on the peer contexts of a multi-device context; all must answer alike.
template <class F> static int fan_out(cudapath_ctx *ctx, int rc, F &&call) {
    if (rc < 0 || !ctx) return rc;
    for (cudapath_ctx *p : ctx->peers) {
        const int r = call(p);
        if (r < 0) return r;
        if (r != rc) return fail("the device contexts of a multi-GPU context diverged");
    }
    return rc;
}
#define CP_GUARD(ctx) DevGuard guard_((ctx)->device); if (guard_.err != cudaSuccess) return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(guard_.err))

// sin / cos of the three lobe shifts of the scale angle (fp32 shifts as the device forms them, fp64 functions): ma_lobe_angles
static void set_lobe_constants(BsdfDev &d) {
    const double off[3] = {-(double) (2.0f * d.scaleAngle), (double) d.scaleAngle, (double) (4.0f * d.scaleAngle)};
    for (int k = 0; k < 3; ++k) { d.lobeSin[k] = std::sin(off[k]); d.lobeCos[k] = std::cos(off[k]); }
}

extern "C" {

const char *cudapath_last_error(void) { return g_lastError.c_str(); }
// used by the host-side translation units (scene loader) to report through the same channel
int cudapath_set_error_message(const char *msg) { g_lastError = msg ? msg : ""; return -1; }

int cudapath_create(int cuda_device, cudapath_ctx **out) {
    if (!out) return fail("null output pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(std::string("no CUDA device available: ") + cudaGetErrorString(e));
    if (cuda_device < 0 || cuda_device >= count) return fail("CUDA device index out of range");
    DevGuard guard_(cuda_device); CKA(guard_.err);
    std::unique_ptr<cudapath_ctx> ctx(new cudapath_ctx());
    ctx->device = cuda_device;
    CKA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    *out = ctx.release();
    return 0;
}
static void multi_free(cudapath_multi *m);
void cudapath_destroy(cudapath_ctx *ctx) {
    if (!ctx) return;
    cudapath_multi *m = ctx->multi; ctx->multi = nullptr;
    std::vector<cudapath_ctx *> peers; peers.swap(ctx->peers);
    if (m) multi_free(m);            // device films and communicators first
    for (cudapath_ctx *p : peers) delete p;
    delete ctx;
}
int cudapath_trim_memory(int cuda_device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || cuda_device < 0 || cuda_device >= count) return fail("CUDA device index out of range");
    DevGuard guard_(cuda_device); CKA(guard_.err);
    dev_trim();
    return 0;
}

int cudapath_set_data_dir(cudapath_ctx *ctx, const char *path) { if (!ctx || !path) return fail("null argument"); ctx->dataDir = path; return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_data_dir(p, path); }); }

int cudapath_add_bsdf_kajiyakay(cudapath_ctx *ctx, const float d[3], const float s[3], float exponent) {
    if (!ctx || !d || !s) return fail("null argument");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    V3 diff(d[0], d[1], d[2]), spec(s[0], s[1], s[2]);
    // BSDF::ensureEnergyConservation (src/librender/bsdf.cpp:115-146)
    const float actualMax = maxc(spec + diff);
    if (actualMax > 1.0f) { const float scale = 0.99f * (1.0f / actualMax); spec = spec * scale; diff = diff * scale; }
    b.dev.kind = 0; b.dev.diffuse = diff; b.dev.specular = spec; b.dev.exponent = exponent;
    const float dAvg = luminance(diff), sAvg = luminance(spec);
    b.dev.specW = sAvg / (dAvg + sAvg);                    // kajiyakay.cpp:96-98
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_kajiyakay(p, d, s, exponent); });
}

int cudapath_add_bsdf_marschner(cudapath_ctx *ctx, float int_ior, float ext_ior, const float d[3], const float s[3], float alpha, int distribution, int nonlinear) {
    if (!ctx || !d || !s) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    if (ctx->dataDir.empty()) return fail("marschner needs data/microfacet/*.dat: call cudapath_set_data_dir() first");
    CP_GUARD(ctx);
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 1;
    b.dev.eta = int_ior / ext_ior;
    b.dev.invEta2 = 1.0f / (b.dev.eta * b.dev.eta);
    b.dev.alpha = std::max(alpha, 1e-4f);                  // microfacet.h:131
    b.dev.nonlinear = nonlinear ? 1 : 0;
    const float betaR = 0.1f, betaTT = betaR * 0.5f, betaTRT = betaR * 2.0f;   // marschner_diffuse.cpp:152-155 (hard-coded)
    b.dev.vR = betaR * betaR; b.dev.vTT = betaTT * betaTT; b.dev.vTRT = betaTRT * betaTRT;
    b.dev.scaleAngle = -0.1f;
    set_lobe_constants(b.dev);
    V3 spec(s[0], s[1], s[2]);
    { const float mx = maxc(spec); if (mx > 1.0f) spec = spec * (0.99f * (1.0f / mx)); }   // bsdf.cpp:88-113
    b.dev.diffuse = V3(d[0], d[1], d[2]); b.dev.specular = spec;
    const float dAvg = luminance(b.dev.diffuse), sAvg = luminance(spec);
    b.dev.specW = sAvg / (dAvg + sAvg);                    // marschner_diffuse.cpp:214-216
    std::vector<float> T; float Fdr; std::string err;
    if (!rough_transmittance_slice(ctx->dataDir, distribution, b.dev.eta, b.dev.alpha, T, Fdr, err)) return fail(err);
    b.dev.Fdr = Fdr; b.dev.rtSize = (int) T.size();
    CKA(dev_alloc(&b.rt, T.size() * 4));
    CKA(cudaMemcpyAsync(b.rt, T.data(), T.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    float pts[140], wts[140];
    gauss_legendre_140(pts, wts);
    const float sigmaA[3] = {0.5f, 0.5f, 0.5f};           // marschner_diffuse.cpp:125 (hard-coded)
    if (!build_marschner_tables(b.dev.eta, betaR, sigmaA, pts, wts, ctx->stream, b.tables, err)) return fail(err);
    b.dev.tab = b.tables.tab; b.dev.cdf = b.tables.cdf; b.dev.sums = b.tables.sums; b.dev.pdfs = b.tables.pdf; b.dev.rt = b.rt;
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_marschner(p, int_ior, ext_ior, d, s, alpha, distribution, nonlinear); });
}

int cudapath_add_bsdf_roughplastic(cudapath_ctx *ctx, float int_ior, float ext_ior, const float d[3], const float s[3], float alpha, int distribution,
                                   int sample_visible, int nonlinear) {
    if (!ctx || !d || !s) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    if (distribution < 0 || distribution > 2) return fail("Specified an invalid distribution, must be \"beckmann\", \"ggx\", or \"phong\"/\"as\"!");
    if (ctx->dataDir.empty()) return fail("roughplastic needs data/microfacet/*.dat: call cudapath_set_data_dir() first");
    CP_GUARD(ctx);
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 4;
    b.dev.eta = int_ior / ext_ior;
    b.dev.invEta2 = 1.0f / (b.dev.eta * b.dev.eta);
    b.dev.alpha = std::max(alpha, 1e-4f);                  // microfacet.h:131
    b.dev.distr = distribution; b.dev.sampleVisible = sample_visible ? 1 : 0;
    if (distribution == 2) { b.dev.sampleVisible = 0; b.dev.exponent = std::max(2.0f / (b.dev.alpha * b.dev.alpha) - 2.0f, 0.0f); }   // :140-143, :677-680
    b.dev.nonlinear = nonlinear ? 1 : 0;
    V3 spec(s[0], s[1], s[2]), diff(d[0], d[1], d[2]);
    { const float mx = maxc(spec); if (mx > 1.0f) spec = spec * (0.99f * (1.0f / mx)); }   // roughplastic.cpp:260-263 -> bsdf.cpp:88-113
    { const float mx = maxc(diff); if (mx > 1.0f) diff = diff * (0.99f * (1.0f / mx)); }
    b.dev.diffuse = diff; b.dev.specular = spec;
    const float dAvg = luminance(diff), sAvg = luminance(spec);
    b.dev.specW = sAvg / (dAvg + sAvg);                    // roughplastic.cpp:276-278
    std::vector<float> T; float Fdr; std::string err;
    if (!rough_transmittance_slice(ctx->dataDir, distribution, b.dev.eta, b.dev.alpha, T, Fdr, err)) return fail(err);
    b.dev.Fdr = Fdr; b.dev.rtSize = (int) T.size();
    CKA(dev_alloc(&b.rt, T.size() * 4));
    CKA(cudaMemcpyAsync(b.rt, T.data(), T.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    b.dev.rt = b.rt;
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_roughplastic(p, int_ior, ext_ior, d, s, alpha, distribution, sample_visible, nonlinear); });
}

int cudapath_add_bsdf_marschner_fixed(cudapath_ctx *ctx, float int_ior, float ext_ior) {
    const float sigmaA[3] = {0.22f, 0.22f, 0.22f};                               // marschner.cpp:122, :130-137 (hard-coded); eval keeps TRT only (:333-334)
    return cudapath_add_bsdf_marschner_full(ctx, int_ior, ext_ior, sigmaA, 0.1f, -0.1f, 4);
}
int cudapath_add_bsdf_marschner_full(cudapath_ctx *ctx, float int_ior, float ext_ior, const float sigmaA[3], float betaR, float scale_angle_rad, int lobe_mask) {
    if (!ctx || !sigmaA) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    if (!(betaR > 0)) return fail("marschner: the longitudinal roughness betaR must be positive");
    if (lobe_mask < 1 || lobe_mask > 7) return fail("marschner: lobe_mask selects at least one of R (1), TT (2), TRT (4)");
    CP_GUARD(ctx);
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 3;
    b.dev.eta = int_ior / ext_ior;
    const float betaTT = betaR * 0.5f, betaTRT = betaR * 2.0f;                   // marschner.cpp:133-134
    b.dev.vR = betaR * betaR; b.dev.vTT = betaTT * betaTT; b.dev.vTRT = betaTRT * betaTRT;
    b.dev.scaleAngle = scale_angle_rad;
    b.dev.lobeMask = lobe_mask;
    set_lobe_constants(b.dev);
    b.dev.diffuse = V3(0.0f); b.dev.specular = V3(1.0f);
    float pts[140], wts[140];
    gauss_legendre_140(pts, wts);
    std::string err;
    if (!build_marschner_tables(b.dev.eta, betaR, sigmaA, pts, wts, ctx->stream, b.tables, err)) return fail(err);
    b.dev.tab = b.tables.tab; b.dev.cdf = b.tables.cdf; b.dev.sums = b.tables.sums; b.dev.pdfs = b.tables.pdf;
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_marschner_full(p, int_ior, ext_ior, sigmaA, betaR, scale_angle_rad, lobe_mask); });
}

static V3 ensure_energy_conservation(const float v[3]) {      // BSDF::ensureEnergyConservation for a constant texture, bsdf.cpp:88-113
    V3 c(v[0], v[1], v[2]);
    const float mx = maxc(c);
    if (mx > 1.0f) c = c * (0.99f * (1.0f / mx));
    return c;
}

int cudapath_add_bsdf_thindielectric(cudapath_ctx *ctx, float int_ior, float ext_ior, const float r[3], const float t[3]) {
    if (!ctx || !r || !t) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 5;
    b.dev.eta = int_ior / ext_ior;                         // thindielectric.cpp:84
    b.dev.specular = ensure_energy_conservation(r);        // :112-115
    b.dev.diffuse = ensure_energy_conservation(t);         // the transmittance rides in the `diffuse` slot
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_thindielectric(p, int_ior, ext_ior, r, t); });
}

int cudapath_add_bsdf_marschnerdielectric(cudapath_ctx *ctx, float int_ior, float ext_ior, const float d[3], const float r[3], const float t[3], float exponent) {
    if (!ctx || !d || !r || !t) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 6;
    b.dev.eta = int_ior / ext_ior;                         // marschnerdielectric.cpp:156
    b.dev.diffuse = V3(d[0], d[1], d[2]);
    b.dev.specular = ensure_energy_conservation(r);        // :191-194
    b.dev.specT = ensure_energy_conservation(t);
    b.dev.exponent = exponent;                             // only read by the dead cone term of eval()
    const float dAvg = luminance(b.dev.diffuse), sAvg = luminance(b.dev.specular), tAvg = luminance(b.dev.specT);
    b.dev.specW = (sAvg + tAvg) / (dAvg + sAvg + tAvg);    // :211-214
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_marschnerdielectric(p, int_ior, ext_ior, d, r, t, exponent); });
}

int cudapath_add_bsdf_diffuse(cudapath_ctx *ctx, const float reflectance[3], int two_sided) {
    if (!ctx || !reflectance) return fail("null argument");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    V3 r(reflectance[0], reflectance[1], reflectance[2]);
    { const float mx = maxc(r); if (mx > 1.0f) r = r * (0.99f * (1.0f / mx)); }           // ensureEnergyConservation, bsdf.cpp:88-113
    b.dev.kind = 2; b.dev.twoSided = two_sided ? 1 : 0; b.dev.diffuse = r; b.dev.specular = V3(0.0f);
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_diffuse(p, reflectance, two_sided); });
}

// `plastic` (src/bsdfs/plastic.cpp:140-217): constant reflectances; a checkerboard can replace the diffuse one (cudapath_bsdf_set_checkerboard)
int cudapath_add_bsdf_plastic(cudapath_ctx *ctx, float int_ior, float ext_ior, const float d[3], const float s[3], int nonlinear) {
    if (!ctx || !d || !s) return fail("null argument");
    if (int_ior < 0 || ext_ior < 0) return fail("The interior and exterior indices of refraction must be positive!");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 7;
    b.dev.eta = int_ior / ext_ior;                               // :152
    b.dev.specular = ensure_energy_conservation(s);              // :188-191
    b.dev.diffuse = ensure_energy_conservation(d);
    b.dev.color1 = b.dev.diffuse;
    b.dev.nonlinear = nonlinear ? 1 : 0;
    b.dev.Fdr = fresnel_diffuse_reflectance(1 / b.dev.eta);      // m_fdrInt, :194 (m_fdrExt only feeds getDiffuseReflectance)
    const float dAvg = luminance(b.dev.diffuse), sAvg = luminance(b.dev.specular);
    b.dev.specW = sAvg / (dAvg + sAvg);                          // :199-202
    b.dev.invEta2 = 1 / (b.dev.eta * b.dev.eta);
    b.dev.uvScale[0] = b.dev.uvScale[1] = 1.0f;
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_plastic(p, int_ior, ext_ior, d, s, nonlinear); });
}

// `mirror`, the fork's own plugin (src/bsdfs/mirror.cpp:187-221)
int cudapath_add_bsdf_mirror(cudapath_ctx *ctx, const float s[3]) {
    if (!ctx || !s) return fail("null argument");
    BsdfHost b; std::memset((void *) &b.dev, 0, sizeof(b.dev));
    b.dev.kind = 8;
    b.dev.specular = ensure_energy_conservation(s);
    ctx->bsdfs.push_back(b); ctx->built = false;
    return fan_out(ctx, (int) ctx->bsdfs.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_bsdf_mirror(p, s); });
}

// <texture type="checkerboard"> as the (diffuse) reflectance of a `diffuse` or `plastic` BSDF (src/textures/checkerboard.cpp:49-52,
// src/librender/texture.cpp:81-95); the BSDF's configure() runs again on it: energy conservation over both colours (bsdf.cpp:88-113),
// sampling weights from the average colour (plastic.cpp:199-202)
int cudapath_bsdf_set_checkerboard(cudapath_ctx *ctx, int bsdf_id, const float color0[3], const float color1[3], float uoffset, float voffset, float uscale, float vscale) {
    if (!ctx || !color0 || !color1) return fail("null argument");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size()) return fail("unknown bsdf id");
    BsdfDev &b = ctx->bsdfs[bsdf_id].dev;
    if (b.kind != 2 && b.kind != 7) return fail("textured reflectance: only `diffuse` and `plastic` carry a texture on this path");
    V3 c0(color0[0], color0[1], color0[2]), c1(color1[0], color1[1], color1[2]);
    const float actualMax = std::max(maxc(c0), maxc(c1));
    if (actualMax > 1.0f) { const float sc = 0.99f * (1.0f / actualMax); c0 = c0 * sc; c1 = c1 * sc; }
    b.texKind = 1; b.diffuse = c0; b.color1 = c1;
    b.uvOffset[0] = uoffset; b.uvOffset[1] = voffset; b.uvScale[0] = uscale; b.uvScale[1] = vscale;
    if (b.kind == 7) { const float dAvg = luminance((c0 + c1) * 0.5f), sAvg = luminance(b.specular); b.specW = sAvg / (dAvg + sAvg); }
    ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_bsdf_set_checkerboard(p, bsdf_id, color0, color1, uoffset, voffset, uscale, vscale); });
}

// <bsdf type="twosided"> around a diffuse / roughplastic / plastic BSDF, the same nested BRDF on both sides (src/bsdfs/twosided.cpp:84-110)
int cudapath_bsdf_set_twosided(cudapath_ctx *ctx, int bsdf_id) {
    if (!ctx) return fail("null argument");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size()) return fail("unknown bsdf id");
    BsdfDev &b = ctx->bsdfs[bsdf_id].dev;
    if (b.kind != 2 && b.kind != 4 && b.kind != 7 && b.kind != 8) return fail("Only materials without a transmission component can be nested!");
    b.twoSided = 1; ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_bsdf_set_twosided(p, bsdf_id); });
}

int cudapath_write_exr(const char *filename, const float *rgb, int w, int h, int half_float) {
    if (!filename || !rgb) return fail("null argument");
    std::string err;
    if (!write_exr_file(filename, rgb, w, h, half_float != 0, err)) return fail(err);
    return 0;
}
int cudapath_float_to_half(const float *in, uint64_t n, uint16_t *out) { if (!in || !out) return fail("null argument"); float_to_half_array(in, (size_t) n, out); return 0; }
int cudapath_random_floats(uint64_t seed, uint64_t n, float *out) { if (!out) return fail("null argument"); mitsuba_random_floats(seed, (size_t) n, out); return 0; }
int cudapath_fresnel_diffuse_reflectance(float eta, float *out) { if (!out) return fail("null argument"); *out = fresnel_diffuse_reflectance(eta); return 0; }

// `rectangle` (src/shapes/rectangle.cpp:81-125): the square [-1,1]^2 x {0} under toWorld, one analytic primitive.  Host arithmetic in plain
// fp32 without FMA, like the reference: Transform(Matrix4x4) inverts by Gauss-Jordan, flipNormals multiplies by scale(1, 1, -1).
int cudapath_add_rectangle(cudapath_ctx *ctx, const float to_world[16], int flip_normals, int bsdf_id) {
    if (!ctx || !to_world) return fail("null argument");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size()) return fail("rectangle references an unknown bsdf id");
    if (ctx->shapes.size() >= (1u << 23)) return fail("too many shapes");
    float o2w[16], w2o[16];
    std::memcpy(o2w, to_world, 64);
    if (!mat_inv(o2w, w2o)) return fail("rectangle: singular toWorld transform");
    if (flip_normals) {
        float sc[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1}, scInv[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1.0f / -1.0f, 0, 0, 0, 0, 1}, a[16], b[16];
        mat_mul(o2w, sc, a); mat_mul(scInv, w2o, b);                // Transform::operator* : (A.m B.m, B.inv A.inv), transform.cpp:28-31
        std::memcpy(o2w, a, 64); std::memcpy(w2o, b, 64);
    }
    auto xv = [&](const float *m, V3 v) { return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); };
    const V3 dpdu = xv(o2w, V3(2, 0, 0)), dpdv = xv(o2w, V3(0, 2, 0));
    const V3 nl(0, 0, 1);                                          // Transform::operator()(Normal): transpose of the inverse
    const V3 nw(w2o[0] * nl.x + w2o[4] * nl.y + w2o[8] * nl.z, w2o[1] * nl.x + w2o[5] * nl.y + w2o[9] * nl.z, w2o[2] * nl.x + w2o[6] * nl.y + w2o[10] * nl.z);
    const V3 n = normalize(nw);
    if (std::abs(dot(normalize(dpdu), normalize(dpdv))) > kEpsilon) return fail("Error: 'toWorld' transformation contains shear!");
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    const V3 corners[4] = {V3(-1, -1, 0), V3(1, -1, 0), V3(1, 1, 0), V3(-1, 1, 0)};
    for (const V3 &c : corners) { const V3 q = h_xfm_point(o2w, c); bmin[0] = std::min(bmin[0], q.x); bmin[1] = std::min(bmin[1], q.y); bmin[2] = std::min(bmin[2], q.z);
                                  bmax[0] = std::max(bmax[0], q.x); bmax[1] = std::max(bmax[1], q.y); bmax[2] = std::max(bmax[2], q.z); }
    const uint32_t shapeIndex = (uint32_t) ctx->shapes.size();
    float4 rec[CP_RECT_STRIDE];
    for (int r = 0; r < 3; ++r) rec[r] = make_float4(w2o[4 * r], w2o[4 * r + 1], w2o[4 * r + 2], w2o[4 * r + 3]);
    uint32_t si = shapeIndex; float sif; std::memcpy(&sif, &si, 4);
    rec[3] = make_float4(dpdu.x, dpdu.y, dpdu.z, sif); rec[4] = make_float4(n.x, n.y, n.z, 0.0f); rec[5] = make_float4(dpdv.x, dpdv.y, dpdv.z, 0.0f);
    rec[6] = make_float4(bmin[0], bmin[1], bmin[2], 0.0f); rec[7] = make_float4(bmax[0], bmax[1], bmax[2], 0.0f);
    ShapeDev sd; std::memset(&sd, 0, sizeof(sd));
    sd.kind = 2; sd.bsdf = bsdf_id; sd.triOffset = (uint32_t) (ctx->rects.size() / CP_RECT_STRIDE); sd.triCount = 1;
    ctx->rects.insert(ctx->rects.end(), rec, rec + CP_RECT_STRIDE);
    ctx->staged.push_back(cudapath_ctx::Staged());
    ctx->shapes.push_back(sd); ctx->built = false;
    return fan_out(ctx, (int) ctx->shapes.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_rectangle(p, to_world, flip_normals, bsdf_id); });
}

int cudapath_add_mesh(cudapath_ctx *ctx, const float *xyz, const float *normals, uint32_t n_vertices, const uint32_t *indices, uint32_t n_triangles, int bsdf_id) {
    return cudapath_add_mesh_uv(ctx, xyz, normals, nullptr, n_vertices, indices, n_triangles, bsdf_id);
}
int cudapath_add_mesh_uv(cudapath_ctx *ctx, const float *xyz, const float *normals, const float *uvs, uint32_t n_vertices, const uint32_t *indices, uint32_t n_triangles, int bsdf_id) {
    if (!ctx || !xyz || !indices) return fail("null argument");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size()) return fail("mesh references an unknown bsdf id");
    if (n_vertices == 0 || n_triangles == 0) return fail("mesh needs at least one vertex and one triangle");
    if (ctx->shapes.size() >= (1u << 23)) return fail("too many shapes");
    if ((uint64_t) ctx->triTotal + n_triangles >= (1ull << 28) || (uint64_t) ctx->meshVtxTotal + n_vertices >= 0xfffffff0ull) return fail("too many triangles");
    for (size_t i = 0; i < 3 * (size_t) n_triangles; ++i) if (indices[i] >= n_vertices) return fail("mesh index out of range");
    CP_GUARD(ctx);
    ShapeDev sd; std::memset(&sd, 0, sizeof(sd));
    sd.kind = 1; sd.bsdf = bsdf_id; sd.vertexOffset = ctx->meshVtxTotal; sd.vertexCount = n_vertices; sd.triOffset = ctx->triTotal; sd.triCount = n_triangles;
    sd.hasNormals = normals ? 1 : 0; sd.hasUV = uvs ? 1 : 0;
    cudapath_ctx::Staged st; st.n = n_vertices; st.nTris = n_triangles;
    if (uvs) {
        CKA(dev_alloc(&st.uv, sizeof(float) * 2 * (size_t) n_vertices));
        CKA(cudaMemcpyAsync(st.uv, uvs, sizeof(float) * 2 * (size_t) n_vertices, cudaMemcpyHostToDevice, ctx->stream));
    }
    CKA(dev_alloc(&st.xyz, sizeof(float) * 3 * (size_t) n_vertices));
    CKA(cudaMemcpyAsync(st.xyz, xyz, sizeof(float) * 3 * (size_t) n_vertices, cudaMemcpyHostToDevice, ctx->stream));
    if (normals) {
        CKA(dev_alloc(&st.nrm, sizeof(float) * 3 * (size_t) n_vertices));
        CKA(cudaMemcpyAsync(st.nrm, normals, sizeof(float) * 3 * (size_t) n_vertices, cudaMemcpyHostToDevice, ctx->stream));
    }
    CKA(dev_alloc(&st.idx, sizeof(uint32_t) * 3 * (size_t) n_triangles));
    CKA(cudaMemcpyAsync(st.idx, indices, sizeof(uint32_t) * 3 * (size_t) n_triangles, cudaMemcpyHostToDevice, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    ctx->staged.push_back(st);
    ctx->meshVtxTotal += n_vertices; ctx->triTotal += n_triangles;
    ctx->shapes.push_back(sd); ctx->built = false;
    return fan_out(ctx, (int) ctx->shapes.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_mesh_uv(p, xyz, normals, uvs, n_vertices, indices, n_triangles, bsdf_id); });
}

int cudapath_mesh_file_load(const char *filename, const float to_world[16], int face_normals, int flip_normals, cudapath_mesh_file **out) {
    if (!filename || !to_world || !out) return fail("null argument");
    std::unique_ptr<cudapath_mesh_file> m(new cudapath_mesh_file());
    std::string err;
    if (!load_obj_file(filename, to_world, face_normals != 0, flip_normals != 0, true, m->data, err)) return fail(err);
    *out = m.release();
    return 0;
}
uint32_t cudapath_mesh_file_vertex_count(const cudapath_mesh_file *m) { return m ? (uint32_t) (m->data.xyz.size() / 3) : 0; }
uint32_t cudapath_mesh_file_triangle_count(const cudapath_mesh_file *m) { return m ? (uint32_t) (m->data.indices.size() / 3) : 0; }
int cudapath_mesh_file_has_normals(const cudapath_mesh_file *m) { return m && !m->data.normals.empty() ? 1 : 0; }
int cudapath_mesh_file_has_texcoords(const cudapath_mesh_file *m) { return m && !m->data.uvs.empty() ? 1 : 0; }
void cudapath_mesh_file_copy_texcoords(const cudapath_mesh_file *m, float *uvs) { if (m && uvs && !m->data.uvs.empty()) std::memcpy(uvs, m->data.uvs.data(), m->data.uvs.size() * 4); }
void cudapath_mesh_file_copy(const cudapath_mesh_file *m, float *xyz, float *normals, uint32_t *indices) {
    if (!m) return;
    if (xyz) std::memcpy(xyz, m->data.xyz.data(), m->data.xyz.size() * 4);
    if (normals && !m->data.normals.empty()) std::memcpy(normals, m->data.normals.data(), m->data.normals.size() * 4);
    if (indices) std::memcpy(indices, m->data.indices.data(), m->data.indices.size() * 4);
}
void cudapath_mesh_file_free(cudapath_mesh_file *m) { delete m; }

int cudapath_add_mesh_file(cudapath_ctx *ctx, const char *filename, const float to_world[16], int face_normals, int flip_normals, int bsdf_id) {
    cudapath_mesh_file *m = nullptr;
    if (cudapath_mesh_file_load(filename, to_world, face_normals, flip_normals, &m) != 0) return -1;
    int r = cudapath_add_mesh_uv(ctx, m->data.xyz.data(), m->data.normals.empty() ? nullptr : m->data.normals.data(), m->data.uvs.empty() ? nullptr : m->data.uvs.data(),
                                 (uint32_t) (m->data.xyz.size() / 3), m->data.indices.data(), (uint32_t) (m->data.indices.size() / 3), bsdf_id);
    cudapath_mesh_file_free(m);
    return r;
}

int cudapath_add_hair(cudapath_ctx *ctx, const float *xyz, const uint8_t *starts, uint32_t n, float radius, int bsdf_id) {
    if (!ctx || !xyz || !starts) return fail("null argument");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size()) return fail("hair shape references an unknown bsdf id");
    if (n < 2) return fail("hair shape needs at least two vertices");
    if (!starts[0]) return fail("the first hair vertex must start a fiber");
    if (ctx->shapes.size() >= (1u << 23)) return fail("too many shapes");
    if ((uint64_t) ctx->vtxTotal + n + 1 >= 0xfffffff0ull) return fail("too many hair vertices");
    CP_GUARD(ctx);
    ShapeDev sd; std::memset(&sd, 0, sizeof(sd));
    sd.radius = radius; sd.bsdf = bsdf_id; sd.vertexOffset = ctx->vtxTotal; sd.vertexCount = n;
    // The caller's arrays go straight to the device (a true DMA when they are page-locked); the float4 vertex stream with
    // the start-of-fiber / shape bits and the sentinel of hair.cpp:782 is assembled by a kernel in cudapath_build().
    cudapath_ctx::Staged st; st.n = n;
    CKA(dev_alloc(&st.xyz, sizeof(float) * 3 * (size_t) n));
    CKA(dev_alloc(&st.starts, (size_t) n));
    CKA(cudaMemcpyAsync(st.xyz, xyz, sizeof(float) * 3 * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
    CKA(cudaMemcpyAsync(st.starts, starts, (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));              // the caller may release its buffers when this returns
    ctx->staged.push_back(st);
    ctx->vtxTotal += n + 1;
    ctx->shapes.push_back(sd); ctx->built = false;
    return fan_out(ctx, (int) ctx->shapes.size() - 1, [&](cudapath_ctx *p) { return cudapath_add_hair(p, xyz, starts, n, radius, bsdf_id); });
}

int cudapath_hair_file_load(const char *filename, float radius, float angle_threshold_deg, float reduction, const float to_world[16], cudapath_hair_file **out) {
    if (!filename || !to_world || !out) return fail("null argument");
    std::unique_ptr<cudapath_hair_file> h(new cudapath_hair_file());
    std::string err;
    if (!load_hair_file(filename, radius, angle_threshold_deg, reduction, to_world, h->data, err)) return fail(err);
    *out = h.release();
    return 0;
}
uint32_t cudapath_hair_file_vertex_count(const cudapath_hair_file *h) { return h ? (uint32_t) h->data.startsFiber.size() : 0; }
float cudapath_hair_file_radius(const cudapath_hair_file *h) { return h ? h->data.radius : 0.0f; }
void cudapath_hair_file_copy(const cudapath_hair_file *h, float *xyz, uint8_t *starts) {
    if (!h) return;
    if (xyz) std::memcpy(xyz, h->data.xyz.data(), h->data.xyz.size() * 4);
    if (starts) std::memcpy(starts, h->data.startsFiber.data(), h->data.startsFiber.size());
}
void cudapath_hair_file_free(cudapath_hair_file *h) { delete h; }

int cudapath_add_hair_file(cudapath_ctx *ctx, const char *filename, float radius, float angle_threshold_deg, float reduction, const float to_world[16], int bsdf_id) {
    cudapath_hair_file *h = nullptr;
    if (cudapath_hair_file_load(filename, radius, angle_threshold_deg, reduction, to_world, &h) != 0) return -1;
    int r = cudapath_add_hair(ctx, h->data.xyz.data(), h->data.startsFiber.data(), (uint32_t) h->data.startsFiber.size(), h->data.radius, bsdf_id);
    cudapath_hair_file_free(h);
    return r;
}

int cudapath_set_envmap(cudapath_ctx *ctx, const float *rgb, int w, int h, const float to_world[16], float scale) {
    if (!ctx || !rgb || !to_world) return fail("null argument");
    if (w <= 0 || h <= 0 || std::max(w, h) > 0xFFFF) return fail("Environment maps images must be smaller than 65536 pixels in width and height");
    ctx->env.rgb.assign(rgb, rgb + (size_t) 3 * w * h);
    ctx->env.w = w; ctx->env.h = h; ctx->env.scale = scale; std::memcpy(ctx->env.toWorld, to_world, 64);
    ctx->env.present = true; ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_envmap(p, rgb, w, h, to_world, scale); });
}

int cudapath_load_rgbe(const char *filename, float *out_rgb, int *out_width, int *out_height) {
    if (!filename || !out_width || !out_height) return fail("null argument");
    std::vector<float> rgb; int w = 0, h = 0; std::string err;
    if (!load_rgbe_file(filename, rgb, w, h, err)) return fail(err);
    *out_width = w; *out_height = h;
    if (out_rgb) std::memcpy(out_rgb, rgb.data(), rgb.size() * 4);
    return 0;
}
int cudapath_set_envmap_file(cudapath_ctx *ctx, const char *filename, const float to_world[16], float scale) {
    if (!ctx || !filename || !to_world) return fail("null argument");
    std::string name = filename;
    const size_t dot = name.rfind('.');
    std::string ext = dot == std::string::npos ? "" : name.substr(dot + 1);
    for (auto &c : ext) c = (char) std::tolower((unsigned char) c);
    if (ext != "hdr" && ext != "rgbe") return fail("envmap: only Radiance RGBE files (.hdr, .rgbe) are read here; other formats go through cudapath_set_envmap");
    std::vector<float> rgb; int w = 0, h = 0; std::string err;
    if (!load_rgbe_file(name, rgb, w, h, err)) return fail(err);
    return cudapath_set_envmap(ctx, rgb.data(), w, h, to_world, scale);
}

int cudapath_bake_sunsky(const char *data_dir, float turbidity, const float albedo[3], const float sun_direction[3], float sky_scale,
                         float sun_scale, float sun_radius_scale, int resolution, float *out_rgb) {
    if (!data_dir || !albedo || !sun_direction || !out_rgb) return fail("null argument");
    SunSkyParams p; p.turbidity = turbidity; p.skyScale = sky_scale; p.sunScale = sun_scale; p.sunRadiusScale = sun_radius_scale; p.resolution = resolution;
    for (int i = 0; i < 3; ++i) { p.albedo[i] = albedo[i]; p.sunDirection[i] = sun_direction[i]; }
    std::vector<float> rgb; int w, h; std::string err;
    if (!bake_sunsky(data_dir, p, rgb, w, h, err)) return fail(err);
    std::memcpy(out_rgb, rgb.data(), rgb.size() * 4);
    return 0;
}
int cudapath_sun_radiance(const char *data_dir, float turbidity, const float sun_direction[3], float out_rgb[3]) {
    if (!data_dir || !sun_direction || !out_rgb) return fail("null argument");
    SunSkyParams p; p.turbidity = turbidity; p.resolution = 2;
    for (int i = 0; i < 3; ++i) p.sunDirection[i] = sun_direction[i];
    std::vector<float> rgb; int w, h; std::string err;
    if (!bake_sunsky(data_dir, p, rgb, w, h, err, out_rgb)) return fail(err);
    return 0;
}
int cudapath_env_pyramid_level(const float *rgb, int width, int height, int level, int *out_width, int *out_height, float *out_rgb) {
    if (!rgb || width <= 0 || height <= 0) return fail("null argument");
    std::vector<EnvMipLevel> levels; float lut[64];
    build_env_pyramid(rgb, width, height, levels, lut);
    if (level < 0 || level >= (int) levels.size()) return fail("MIP level out of range");
    if (out_width) *out_width = levels[level].w;
    if (out_height) *out_height = levels[level].h;
    if (out_rgb) std::memcpy(out_rgb, levels[level].rgb.data(), levels[level].rgb.size() * 4);
    return (int) levels.size();
}
int cudapath_set_sunsky(cudapath_ctx *ctx, float turbidity, const float albedo[3], const float sun_direction[3], float sky_scale,
                        float sun_scale, float sun_radius_scale, int resolution) {
    if (!ctx) return fail("null context");
    if (ctx->dataDir.empty()) return fail("sunsky needs the sky-model data: call cudapath_set_data_dir() first");
    if (resolution <= 1) return fail("invalid sunsky resolution");
    std::vector<float> rgb((size_t) 3 * resolution * (resolution / 2));
    if (cudapath_bake_sunsky(ctx->dataDir.c_str(), turbidity, albedo, sun_direction, sky_scale, sun_scale, sun_radius_scale, resolution, rgb.data()) != 0) return -1;
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    return cudapath_set_envmap(ctx, rgb.data(), resolution, resolution / 2, ident, 1.0f);
}

int cudapath_set_camera_perspective(cudapath_ctx *ctx, const float to_world[16], float fov, float nearClip, float farClip, int w, int h) {
    if (!ctx || !to_world) return fail("null argument");
    if (nearClip <= 0) return fail("The 'nearClip' parameter must be greater than zero!");
    if (nearClip >= farClip) return fail("The 'nearClip' parameter must be smaller than 'farClip'.");
    if (fov <= 0 || fov >= 180) return fail("The horizontal field of view must be in the interval (0, 180)!");
    if (w <= 0 || h <= 0) return fail("invalid film size");
    std::memcpy(ctx->cam.toWorld, to_world, 64);
    ctx->cam.fov = fov; ctx->cam.nearClip = nearClip; ctx->cam.farClip = farClip; ctx->cam.w = w; ctx->cam.h = h; ctx->cam.present = true;
    ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_camera_perspective(p, to_world, fov, nearClip, farClip, w, h); });
}
int cudapath_set_film(cudapath_ctx *ctx, int filter, float param, int has_alpha) {
    if (!ctx) return fail("null context");
    if (filter < 0 || filter > 2) return fail("unsupported reconstruction filter");
    ctx->filterType = filter; ctx->filterParam = param; ctx->hasAlpha = has_alpha ? 1 : 0; ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_film(p, filter, param, has_alpha); });
}
int cudapath_set_integrator(cudapath_ctx *ctx, int max_depth, int rr_depth, int strict_normals, int hide_emitters) {
    if (!ctx) return fail("null context");
    if (max_depth <= 0 && max_depth != -1) return fail("'maxDepth' must be set to -1 (infinite) or a value greater than zero!");  // integrator.cpp:208-210
    if (rr_depth <= 0) return fail("'rrDepth' must be set to a value greater than zero!");
    ctx->integ.maxDepth = max_depth; ctx->integ.rrDepth = rr_depth; ctx->integ.strictNormals = strict_normals ? 1 : 0; ctx->integ.hideEmitters = hide_emitters ? 1 : 0;
    if (ctx->built) ctx->scene.integ = ctx->integ;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_integrator(p, max_depth, rr_depth, strict_normals, hide_emitters); });
}
int cudapath_set_sampler(cudapath_ctx *ctx, int kind, uint64_t scramble) {
    if (!ctx) return fail("null context");
    if (kind == 2) {                       // "what the scene file asks for": resolved by cudapath_load_scene_xml when it meets the <sampler> element
        ctx->samplerKind = 2; ctx->built = false;
        return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_sampler(p, kind, scramble); });
    }
    if (kind != 0 && kind != 1) return fail("sampler: 0 = counter-based Philox stream, 1 = the reference's sobol sampler, 2 = the sampler of the scene file loaded next");
    if (kind == 1 && !ctx->d_sobolM32) {
        if (ctx->dataDir.empty()) return fail("the sobol sampler needs sobol.bin: call cudapath_set_data_dir() first");
        CP_GUARD(ctx);
        const std::string path = ctx->dataDir + "/sobol.bin";
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return fail("cannot open \"" + path + "\" (tools/mirror_refdata.py writes it)");
        uint32_t h[5] = {0, 0, 0, 0, 0};
        bool ok = fread(h, 4, 5, f) == 5 && h[0] == 0x4c424f53u && h[1] == CP_SOBOL_DIMS && h[2] == CP_SOBOL_SIZE && h[3] >= 1 && h[3] <= 64 && h[4] >= 1 && h[4] <= 64;
        std::vector<uint32_t> m32; std::vector<uint64_t> vdc, inv;
        if (ok) { m32.resize((size_t) CP_SOBOL_DIMS * CP_SOBOL_SIZE); vdc.resize((size_t) h[3] * CP_SOBOL_SIZE); inv.resize((size_t) h[4] * CP_SOBOL_SIZE);
                  ok = fread(m32.data(), 4, m32.size(), f) == m32.size() && fread(vdc.data(), 8, vdc.size(), f) == vdc.size() && fread(inv.data(), 8, inv.size(), f) == inv.size(); }
        fclose(f);
        if (!ok) return fail("malformed \"" + path + "\"");
        CKA(dev_alloc(&ctx->d_sobolM32, m32.size() * 4)); CKA(dev_alloc(&ctx->d_sobolVdc, vdc.size() * 8)); CKA(dev_alloc(&ctx->d_sobolInv, inv.size() * 8));
        CKA(cudaMemcpyAsync(ctx->d_sobolM32, m32.data(), m32.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CKA(cudaMemcpyAsync(ctx->d_sobolVdc, vdc.data(), vdc.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CKA(cudaMemcpyAsync(ctx->d_sobolInv, inv.data(), inv.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CKA(cudaStreamSynchronize(ctx->stream));
        ctx->sobolRows[0] = h[3]; ctx->sobolRows[1] = h[4];
    }
    ctx->samplerKind = kind; ctx->samplerScramble = scramble; ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_sampler(p, kind, scramble); });
}
int cudapath_get_sampler(cudapath_ctx *ctx) { return ctx ? ctx->samplerKind : -1; }

int cudapath_set_options(cudapath_ctx *ctx, uint32_t wave_size, int collect_stats, int profile_stages) {
    if (!ctx) return fail("null context");
    if (wave_size) ctx->waveSize = std::max(wave_size, 1024u);
    ctx->collectStats = collect_stats; ctx->profileStages = profile_stages;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_options(p, wave_size, collect_stats, profile_stages); });
}

static int build_one(cudapath_ctx *ctx) {
    if (!ctx) return fail("null context");
    if (ctx->shapes.empty()) return fail("the scene contains no shapes");
    if (!ctx->cam.present) return fail("no sensor set");
    CP_GUARD(ctx);
    const bool trace = getenv("CUDAPATH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tb0 = now();
    ctx->freeBuilt();
    std::string err;
    // The Lanczos MIP pyramid of the environment map is host work (3 ms for 512x256): it runs on a worker thread while this one drives the device build.
    std::vector<EnvMipLevel> pyrLevels; EnvMipInfo pyrInfo; std::memset(&pyrInfo, 0, sizeof(pyrInfo));
    std::future<void> pyramidJob;
    struct JoinGuard { std::future<void> &f; ~JoinGuard() { if (f.valid()) f.wait(); } } pyramidGuard{pyramidJob};      // error returns must not leave the worker running
    if (ctx->env.present) pyramidJob = std::async(std::launch::async, [&]() { build_env_pyramid(ctx->env.rgb.data(), ctx->env.w, ctx->env.h, pyrLevels, pyrInfo.lut); });
    cudaEvent_t e0, e1; CKA(cudaEventCreate(&e0)); CKA(cudaEventCreate(&e1));
    CKA(cudaEventRecord(e0, ctx->stream));
    // geometry
    CKA(dev_alloc(&ctx->d_vtx, sizeof(float4) * ((size_t) ctx->vtxTotal + 4)));
    CKA(cudaMemsetAsync(ctx->d_vtx, 0, sizeof(float4) * ((size_t) ctx->vtxTotal + 4), ctx->stream));
    MeshDev mesh; std::memset(&mesh, 0, sizeof(mesh));
    if (!ctx->rects.empty()) {
        CKA(dev_alloc(&ctx->d_rects, sizeof(float4) * ctx->rects.size()));
        CKA(cudaMemcpyAsync(ctx->d_rects, ctx->rects.data(), sizeof(float4) * ctx->rects.size(), cudaMemcpyHostToDevice, ctx->stream));
        mesh.rects = ctx->d_rects; mesh.rectCount = (uint32_t) (ctx->rects.size() / CP_RECT_STRIDE);
    }
    if (ctx->triTotal) {
        CKA(dev_alloc(&ctx->d_meshUV, sizeof(float2) * (size_t) ctx->meshVtxTotal));
        CKA(cudaMemsetAsync(ctx->d_meshUV, 0, sizeof(float2) * (size_t) ctx->meshVtxTotal, ctx->stream));
        mesh.uv = ctx->d_meshUV;
        CKA(dev_alloc(&ctx->d_meshPos, sizeof(float4) * (size_t) ctx->meshVtxTotal)); CKA(dev_alloc(&ctx->d_meshNrm, sizeof(float4) * (size_t) ctx->meshVtxTotal));
        CKA(dev_alloc(&ctx->d_meshIdx, sizeof(uint32_t) * 3 * (size_t) ctx->triTotal)); CKA(dev_alloc(&ctx->d_triAccel, sizeof(float4) * 3 * (size_t) ctx->triTotal));
        mesh.triAccel = ctx->d_triAccel; mesh.pos = ctx->d_meshPos; mesh.nrm = ctx->d_meshNrm; mesh.idx = ctx->d_meshIdx;
        mesh.triCount = ctx->triTotal; mesh.vertCount = ctx->meshVtxTotal;
    }
    int hairShapes = 0;
    for (size_t i = 0; i < ctx->staged.size(); ++i) {
        const cudapath_ctx::Staged &st = ctx->staged[i];
        const ShapeDev &sd = ctx->shapes[i];
        if (sd.kind == 2) continue;                 // rectangle: its record is already in d_rects
        if (sd.kind == 1) {
            pack_mesh(st.xyz, st.nrm, st.n, ctx->d_meshPos + sd.vertexOffset, ctx->d_meshNrm + sd.vertexOffset, ctx->stream);
            if (st.uv) CKA(cudaMemcpyAsync(ctx->d_meshUV + sd.vertexOffset, st.uv, sizeof(float2) * (size_t) st.n, cudaMemcpyDeviceToDevice, ctx->stream));
            build_tri_accel(st.idx, st.nTris, sd.vertexOffset, (uint32_t) i, ctx->d_meshPos, ctx->d_meshIdx + 3 * (size_t) sd.triOffset,
                            ctx->d_triAccel + 3 * (size_t) sd.triOffset, ctx->stream);
        } else {
            pack_vertices(st.xyz, st.starts, st.n, (uint32_t) i, ctx->d_vtx + sd.vertexOffset, ctx->stream);
            hairShapes++;
        }
    }
    CKA(dev_alloc(&ctx->d_shapes, sizeof(ShapeDev) * ctx->shapes.size()));
    CKA(cudaMemcpyAsync(ctx->d_shapes, ctx->shapes.data(), sizeof(ShapeDev) * ctx->shapes.size(), cudaMemcpyHostToDevice, ctx->stream));
    BuildInfo info;
    const double tb1 = now();
    if (!build_bvh(ctx->d_vtx, ctx->vtxTotal, ctx->d_shapes, (int) ctx->shapes.size(), mesh, ctx->maxSplit, ctx->leafSplitCost, ctx->stream, ctx->bvh, info, err)) return fail(err);
    CKA(cudaMemcpyAsync(ctx->shapes.data(), ctx->d_shapes, sizeof(ShapeDev) * ctx->shapes.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    const double tb2 = now();
    // KDTreeBase::buildInternal enlarges the box of a finished kd-tree by MTS_KD_AABB_EPSILON = 1e-3, relative to its extent plus
    // absolute (include/mitsuba/render/gkdtree.h:50,1213-1220; the second line already sees the lowered minimum).  getAABB() returns
    // that box: HairKDTree clips rays against it (hair.cpp:205) and hands it to the scene-level tree (hair.cpp:944-946), whose own
    // box -- the union over its primitives, enlarged once more -- is what ShapeKDTree::rayIntersect clips against (skdtree.cpp:124).
    // Plain fp32 host arithmetic (no FMA), like the reference.
    auto enlarge = [](float *mn, float *mx) {
        const float eps = 1e-3f;
        for (int k = 0; k < 3; ++k) { const volatile float e = (mx[k] - mn[k]) * eps; mn[k] = mn[k] - (e + eps); }
        for (int k = 0; k < 3; ++k) { const volatile float e = (mx[k] - mn[k]) * eps; mx[k] = mx[k] + (e + eps); }
    };
    for (auto &sh : ctx->shapes) if (sh.kind == 0 && sh.bmin[0] <= sh.bmax[0]) enlarge(sh.bmin, sh.bmax);
    CKA(cudaMemcpyAsync(ctx->d_shapes, ctx->shapes.data(), sizeof(ShapeDev) * ctx->shapes.size(), cudaMemcpyHostToDevice, ctx->stream));
    // bsdfs
    std::vector<BsdfDev> devs; for (auto &b : ctx->bsdfs) devs.push_back(b.dev);
    CKA(dev_alloc(&ctx->d_bsdfs, sizeof(BsdfDev) * std::max<size_t>(devs.size(), 1)));
    CKA(cudaMemcpyAsync(ctx->d_bsdfs, devs.data(), sizeof(BsdfDev) * devs.size(), cudaMemcpyHostToDevice, ctx->stream));

    SceneDev &S = ctx->scene;
    std::memset(&S, 0, sizeof(S));
    S.vtx = ctx->d_vtx; S.vtxCount = ctx->vtxTotal; S.shapes = ctx->d_shapes; S.shapeCount = (int) ctx->shapes.size();
    S.bsdfs = ctx->d_bsdfs; S.bsdfCount = (int) devs.size(); S.bvh = ctx->bvh; S.integ = ctx->integ;
    S.sobol.kind = 0;
    if (ctx->samplerKind == 2) return fail("cudapath_set_sampler(2) waits for a scene file: no <sampler> element has been loaded");
    if (ctx->samplerKind == 1) {        // SobolSampler::setFilmResolution(crop size, bucketed = true), sobol.cpp:146-156
        uint32_t r = (uint32_t) std::max(ctx->cam.w, ctx->cam.h);
        r--; r |= r >> 1; r |= r >> 2; r |= r >> 4; r |= r >> 8; r |= r >> 16; r++;
        uint32_t lg = 0; while ((1u << (lg + 1)) <= r) ++lg;
        if (lg > 1 && (lg > ctx->sobolRows[0] || lg > ctx->sobolRows[1])) return fail("sobol: film resolution outside the enumeration tables");
        uint64_t scr = ctx->samplerScramble;
        if (scr) {                       // sampleTEA(lo, hi), qmc.h:146-156 (sobol.cpp:96-103)
            uint32_t v0 = (uint32_t) scr, v1 = (uint32_t) (scr >> 32), sum = 0;
            for (int i = 0; i < 4; ++i) { sum += 0x9e3779b9; v0 += ((v1 << 4) + 0xA341316C) ^ (v1 + sum) ^ ((v1 >> 5) + 0xC8013EA4); v1 += ((v0 << 4) + 0xAD90777D) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7E95761E); }
            scr = ((uint64_t) v1 << 32) + v0;
        }
        S.sobol.kind = 1; S.sobol.m32 = ctx->d_sobolM32; S.sobol.vdc = ctx->d_sobolVdc; S.sobol.inv = ctx->d_sobolInv;
        S.sobol.logRes = lg; S.sobol.res = (float) r; S.sobol.scramble = (uint32_t) scr; S.sobol.err = nullptr;
    }
    S.mesh = mesh; S.clipPerShape = (hairShapes > 1 || ctx->triTotal > 0 || !ctx->rects.empty()) ? 1 : 0;
    for (int k = 0; k < 3; ++k) { S.sceneMin[k] = INFINITY; S.sceneMax[k] = -INFINITY; }
    for (auto &sh : ctx->shapes) for (int k = 0; k < 3; ++k) { S.sceneMin[k] = std::min(S.sceneMin[k], sh.bmin[k]); S.sceneMax[k] = std::max(S.sceneMax[k], sh.bmax[k]); }
    enlarge(S.sceneMin, S.sceneMax);
    for (int k = 0; k < 3; ++k) { ctx->sceneAABB[k] = S.sceneMin[k]; ctx->sceneAABB[3 + k] = S.sceneMax[k]; }

    // camera (perspective.cpp:126-160; Transform::perspective transform.cpp:99-121)
    {
        CameraDev &C = S.cam;
        const CamHost &H = ctx->cam;
        // fp32 Transform algebra of the reference: every factory carries its inverse (transform.cpp:33-63, perspective inverts numerically
        // :99-123), a product carries (A.m * B.m, B.inv * A.inv) (transform.cpp:28-31); sampleToCamera is the inverse member of
        //   scale(1/relSize) * translate(-relOffset) * scale(-0.5, -0.5 aspect, 1) * translate(-1, -1/aspect, 0) * perspective(...)
        // evaluated left to right (the first two factors are exact identities without a crop window)
        const float aspect = H.w / (float) H.h;                                        // sensor.cpp:101-102
        const float recip = 1.0f / (H.farClip - H.nearClip);
        const float cot = 1.0f / std::tan((H.fov / 2.0f) * (kPi / 180.0f));
        const float P[16] = {cot, 0, 0, 0, 0, cot, 0, 0, 0, 0, H.farClip * recip, -H.nearClip * H.farClip * recip, 0, 0, 1, 0};
        float Pinv[16];
        if (!mat_inv(P, Pinv)) return fail("singular camera projection");
        const float sx = -0.5f, sy = -0.5f * aspect, tx = -1.0f, ty = -1.0f / aspect;
        const float ScInv[16] = {1.0f / sx, 0, 0, 0, 0, 1.0f / sy, 0, 0, 0, 0, 1.0f / 1.0f, 0, 0, 0, 0, 1};
        const float TrInv[16] = {1, 0, 0, -tx, 0, 1, 0, -ty, 0, 0, 1, -0.0f, 0, 0, 0, 1};
        float TS[16], inv[16];
        mat_mul(TrInv, ScInv, TS); mat_mul(Pinv, TS, inv);
        for (int i = 0; i < 16; ++i) { C.s2c[i] = inv[i]; C.toWorld[i] = H.toWorld[i]; }
        C.invResX = (float) 1 / (float) H.w; C.invResY = (float) 1 / (float) H.h; C.nearClip = H.nearClip; C.farClip = H.farClip; C.filmW = H.w; C.filmH = H.h;
        const V3 p0 = h_xfm_point(C.s2c, V3(0.0f)), px = h_xfm_point(C.s2c, V3(C.invResX, 0, 0)), py = h_xfm_point(C.s2c, V3(0, C.invResY, 0));
        const V3 dx = px - p0, dy = py - p0;
        C.dx[0] = dx.x; C.dx[1] = dx.y; C.dx[2] = dx.z; C.dy[0] = dy.x; C.dy[1] = dy.y; C.dy[2] = dy.z;
    }
    // film filter (rfilter.cpp:37-55; tent.cpp:42-44, box.cpp, gaussian.cpp)
    {
        FilmDev &F = S.film;
        float radius = 1.0f, stddev = ctx->filterParam > 0 ? ctx->filterParam : 0.5f;
        if (ctx->filterType == 0 && ctx->filterParam > 0) radius = ctx->filterParam;
        if (ctx->filterType == 1) radius = 0.5f + 1e-5f;   // box.cpp:38: props.getFloat("radius", 0.5f) + 1e-5f
        if (ctx->filterType == 2) radius = 4 * stddev;
        auto ev = [&](float x) -> float {
            if (ctx->filterType == 0) return std::max(0.0f, 1.0f - std::fabs(x / radius));
            if (ctx->filterType == 1) return std::fabs(x) <= radius ? 1.0f : 0.0f;
            const float a = -1.0f / (2.0f * stddev * stddev);
            return std::max(0.0f, cr_exp(a * x * x) - cr_exp(a * radius * radius));
        };
        float sum = 0.0f;
        for (int i = 0; i < 31; ++i) { F.filterValues[i] = ev((radius * i) / 31); sum += F.filterValues[i]; }
        F.filterValues[31] = 0.0f;
        sum *= 2 * radius / 31;
        const float norm = 1.0f / sum;
        for (int i = 0; i < 31; ++i) F.filterValues[i] *= norm;
        F.filterRadius = radius; F.filterScale = 31 / radius; F.hasAlpha = ctx->hasAlpha;
        if (radius > 7.5f) return fail("reconstruction filter radius too large");
    }
    // environment map
    S.env.present = 0;
    if (ctx->env.present) {
        float *d_rgb = nullptr;
        const size_t bytes = ctx->env.rgb.size() * 4;
        CKA(dev_alloc(&d_rgb, bytes));
        CKA(cudaMemcpyAsync(d_rgb, ctx->env.rgb.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
        const bool ok = build_env_tables(d_rgb, ctx->env.w, ctx->env.h, ctx->stream, ctx->envTables, err);
        dev_free(d_rgb);
        if (!ok) return fail(err);
        {   // Lanczos MIP pyramid + EWA weights for camera rays that leave the scene (MIPMap::eval, envmap.cpp:391-407)
            pyramidJob.get();            // host work that ran beside the BVH build (see the top of this function)
            std::vector<EnvMipLevel> &levels = pyrLevels; EnvMipInfo &info = pyrInfo;
            if ((int) levels.size() > CP_ENV_MAX_LEVELS) return fail("environment map has too many MIP levels");
            info.levels = (int) levels.size();
            size_t total = 0;
            for (size_t l = 0; l < levels.size(); ++l) {
                info.w[l] = levels[l].w; info.h[l] = levels[l].h;
                info.ratioX[l] = (float) levels[l].w / (float) ctx->env.w; info.ratioY[l] = (float) levels[l].h / (float) ctx->env.h;     // m_sizeRatio, mipmap.h:264-266
                if (l >= 1) { info.offset[l] = (uint32_t) total; total += (size_t) levels[l].w * levels[l].h; }
            }
            std::vector<float> upper; upper.reserve(3 * total);
            for (size_t l = 1; l < levels.size(); ++l) upper.insert(upper.end(), levels[l].rgb.begin(), levels[l].rgb.end());
            float *d_upper = nullptr;
            CKA(dev_alloc(&d_upper, std::max<size_t>(upper.size(), 1) * 4));
            CKA(dev_alloc(&ctx->envTables.mipTexels, std::max<size_t>(total, 1) * sizeof(float4)));
            CKA(dev_alloc(&ctx->envTables.mipInfo, sizeof(EnvMipInfo)));
            if (total) CKA(cudaMemcpyAsync(d_upper, upper.data(), upper.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            quantize_texels(d_upper, (int) total, ctx->envTables.mipTexels, ctx->stream);
            CKA(cudaMemcpyAsync(ctx->envTables.mipInfo, &info, sizeof(info), cudaMemcpyHostToDevice, ctx->stream));
            CKA(cudaStreamSynchronize(ctx->stream));
            dev_free(d_upper);
        }
        EnvDev &E = S.env;
        E.w = ctx->env.w; E.h = ctx->env.h; E.texels = ctx->envTables.texels; E.cdfCols = ctx->envTables.cdfCols; E.cdfRows = ctx->envTables.cdfRows;
        E.rowWeights = ctx->envTables.rowWeights; E.normalization = ctx->envTables.normalization; E.scale = ctx->env.scale;
        E.mipTexels = ctx->envTables.mipTexels; E.mip = ctx->envTables.mipInfo;
        E.pixelSizeX = 2 * kPi / E.w; E.pixelSizeY = kPi / E.h;
        float tw[16], ti[16];
        for (int i = 0; i < 16; ++i) tw[i] = ctx->env.toWorld[i];
        if (!mat_inv(tw, ti)) return fail("singular environment map transform");      // Transform(const Matrix4x4 &), transform.h:50-55
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { E.toWorld[r * 3 + c] = tw[r * 4 + c]; E.toLocal[r * 3 + c] = ti[r * 4 + c]; }
        // scene bounds for the emitter: kd-tree AABB + sensor position (scene.cpp:387-413), bounding sphere x1.5 (envmap.cpp:331-341)
        float mn[3], mx[3];
        const V3 camPos = h_xfm_point(ctx->cam.toWorld, V3(0.0f));
        const float cp3[3] = {camPos.x, camPos.y, camPos.z};
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(S.sceneMin[k], cp3[k]); mx[k] = std::max(S.sceneMax[k], cp3[k]); }
        float c[3]; for (int k = 0; k < 3; ++k) c[k] = (mx[k] + mn[k]) * 0.5f;
        const V3 diff(c[0] - mx[0], c[1] - mx[1], c[2] - mx[2]);
        E.bsCenter[0] = c[0]; E.bsCenter[1] = c[1]; E.bsCenter[2] = c[2];
        E.bsRadius = std::max(kEpsilon, length(diff) * 1.5f);
        E.present = 1;
    }
    CKA(cudaEventRecord(e1, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    float ms = 0; CKA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (trace) fprintf(stderr, "[cudapath] build: pack %.4f s, bvh %.4f s, tables+rest %.4f s (device %.4f s)\n", tb1 - tb0, tb2 - tb1, now() - tb2, ms * 1e-3);
    ctx->stats.build_ms = ms; ctx->stats.segments = info.segments; ctx->stats.triangles = info.triangles; ctx->stats.bvh_nodes = info.nodes; ctx->stats.bvh_references = info.references;
    ctx->built = true;
    return 0;
}

int cudapath_render_dev(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *film_dev, void *stream) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!film_dev) return fail("null film buffer");
    if (spp == 0 || sample_end > spp || sample_begin > sample_end) return fail("invalid sample range");
    cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;
    cudaEvent_t e0, e1; CKA(cudaEventCreate(&e0)); CKA(cudaEventCreate(&e1));
    CKA(cudaEventRecord(e0, st));
    RenderStats rs; std::string err;
    ctx->wf.sortRays = ctx->sortRays != 0; ctx->wf.fastMath = ctx->fastMath != 0;
    ctx->wf.shardIndex = ctx->shardIndex; ctx->wf.shardCount = ctx->shardCount;
    if (const char *e = getenv("CUDAPATH_SHARD_BLOCK")) { const int px = atoi(e); uint32_t sh = 0; while (sh < 6u && (16 << sh) <= px) ++sh; ctx->wf.shardBlockShift = sh; }     // pixels per block side: 8, 16, 32 (default) ... 512
    if (const char *e = getenv("CUDAPATH_RUNAHEAD_MAX")) ctx->wf.runAheadMax = (uint32_t) strtoul(e, nullptr, 0);
    ctx->wf.cancelRequested.store(0); ctx->wf.inRender.store(1);
    struct RenderScope { Wavefront &w; cudaEvent_t a, b; ~RenderScope() { w.inRender.store(0); w.cancelRequested.store(0); cudaEventDestroy(a); cudaEventDestroy(b); } } scope_{ctx->wf, e0, e1};
    // Deep bounces leave only a few live paths per wave, so few, large waves keep the GPU full for longer (hair-curl at 64 spp: one 2^26 wave is
    // 17 % faster than four 2^24 waves).  A path costs ~228 B of queue space; a wave is up to 2^26 paths (15 GB of a 180 GB B200; Wavefront::render
    // never reserves more slots than the job has paths).  The size is NOT derived from cudaMemGetInfo: on the shared GPU hosts that query takes
    // 0.6 ms most of the time and 25-80 ms now and then (it was inside every render until the end of round 2: the cause of the steps that took 10-45 ms
    // longer, profiles/r2_step_time_spread_g25.log).  Instead the reservation is tried and halved while the device says it is out of memory.
    uint32_t waveSize = ctx->waveSize;
    const bool autoSize = waveSize == 0;
    if (autoSize) waveSize = 1u << 26;
    for (;;) {
        const bool ok = ctx->wf.render(ctx->scene, spp, seed, sample_begin, sample_end, film_dev, waveSize, ctx->collectStats != 0, ctx->profileStages != 0, st, rs, err);
        if (ok) break;
        if (!autoSize || waveSize <= (1u << 20) || err.find("out of memory") == std::string::npos || ctx->wf.capacity != 0) return fail(err);
        cudaGetLastError();                  // the failed allocation is not sticky; nothing was launched (reserve() comes first and leaves capacity at 0)
        ctx->wf.release();
        waveSize >>= 1; rs = RenderStats(); err.clear();
    }
    CKA(cudaEventRecord(e1, st));
    CKA(cudaEventSynchronize(e1));
    float ms = 0; CKA(cudaEventElapsedTime(&ms, e0, e1));
    cudapath_stats &s = ctx->stats;
    s.paths = rs.paths; s.rays = rs.rays; s.shadow_rays = rs.shadowRays; s.kernel_launches = rs.launches; s.bounces = rs.bounces;
    s.nodes_visited = rs.nodesVisited; s.prims_tested = rs.primsTested; s.shadow_nodes_visited = rs.shadowNodesVisited; s.shadow_prims_tested = rs.shadowPrimsTested;
    s.trace_ms = rs.stageMs[0]; s.shade_ms = rs.stageMs[1]; s.sort_ms = rs.stageMs[2]; s.raygen_ms = rs.stageMs[3]; s.splat_ms = rs.stageMs[4];
    s.trace_launches = rs.stageLaunches[0]; s.shade_launches = rs.stageLaunches[1]; s.sort_launches = rs.stageLaunches[2]; s.host_waits = rs.hostSyncs; s.unsupported_filtered_lookups = rs.unsupportedLookups; s.dropped_samples = rs.droppedSamples;
    s.full_tests = rs.fullTests; s.shadow_full_tests = rs.shadowFullTests; s.shadow_rays_traced = rs.shadowRaysTraced;
    s.render_ms = ms;
    return 0;
}

int cudapath_cancel(cudapath_ctx *ctx) {
    if (!ctx) return fail("null context");
    // the only entry point that may be called while another thread is inside a render; like Integrator::cancel() it acts on the job
    // that is running -- a request that finds no render in progress is dropped
    if (ctx->wf.inRender.load()) ctx->wf.cancelRequested.store(1);
    for (cudapath_ctx *p : ctx->peers) if (p->wf.inRender.load()) p->wf.cancelRequested.store(1);
    return 0;
}

int cudapath_set_progress_callback(cudapath_ctx *ctx, void (*callback)(void *user, uint64_t paths_done, uint64_t paths_total), void *user) {
    if (!ctx) return fail("null context");
    ctx->wf.progress = callback; ctx->wf.progressUser = user;
    return 0;
}

static int render_multi(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *out_film);
int cudapath_render(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *out_film) {
    if (ctx && !ctx->peers.empty()) return render_multi(ctx, spp, seed, sample_begin, sample_end, out_film);
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!out_film) return fail("null film buffer");
    const size_t bytes = sizeof(float) * 5 * (size_t) ctx->cam.w * ctx->cam.h;
    const bool trace = getenv("CUDAPATH_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    float *d_film = nullptr;
    CKA(dev_alloc(&d_film, bytes));
    CKA(cudaMemsetAsync(d_film, 0, bytes, ctx->stream));
    const double t1 = now();
    int r = cudapath_render_dev(ctx, spp, seed, sample_begin, sample_end, d_film, ctx->stream);
    const double t2 = now();
    if (r == 0) {
        cudaError_t e = cudaMemcpyAsync(out_film, d_film, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { dev_free(d_film); return fail(std::string("film copy: ") + cudaGetErrorString(e)); }
    }
    const double t3 = now();
    dev_free(d_film);
    if (trace) fprintf(stderr, "[cudapath] render: film alloc %.3f s, render_dev %.3f s (device %.3f s), read-back %.3f s, free %.3f s\n", t1 - t0, t2 - t1,
                       ctx->stats.render_ms * 1e-3, t3 - t2, now() - t3);
    return r;
}

int cudapath_develop(const float *film, int w, int h, float *out_rgb) {
    if (!film || !out_rgb) return fail("null argument");
    for (size_t i = 0; i < (size_t) w * h; ++i) {
        const float wt = film[5 * i + 4];
        const float inv = wt != 0 ? 1.0f / wt : 0.0f;      // fmtconv.cpp: invWeight = weight != 0 ? 1/weight : 0
        for (int k = 0; k < 3; ++k) out_rgb[3 * i + k] = film[5 * i + k] * inv;
    }
    return 0;
}

// LDRFilm::develop with the default `gamma` tonemapper (src/films/ldrfilm.cpp:300-321): Bitmap::convert(ERGB, EUInt8, gamma, 2^exposure)
// = per pixel value/weight (fmtconv.cpp:984-995), times the multiplier, gamma curve (applyGamma :1104-1111: sRGB when gamma == -1,
// else pow(v, 1/gamma)), then round-to-nearest and clamp to [0, 255] (convertScalar :1157-1159).
int cudapath_develop_ldr(const float *film, int w, int h, float gamma, float exposure, uint8_t *out_rgb8) {
    if (!film || !out_rgb8) return fail("null argument");
    if (gamma == 0) return fail("gamma must not be zero");
    const float multiplier = cr_pow(2.0f, exposure);
    const float invDestGamma = 1.0f / gamma;
    for (size_t i = 0; i < (size_t) w * h; ++i) {
        const float wt = film[5 * i + 4];
        const float inv = wt != 0 ? 1.0f / wt : wt;
        for (int k = 0; k < 3; ++k) {
            float value = film[5 * i + k] * inv;
            value *= multiplier;
            if (invDestGamma != 1) {
                if (invDestGamma == -1) value = (value <= 0.0031308f) ? (12.92f * value) : (1.055f * cr_pow(value, (float) (1.0 / 2.4)) - 0.055f);
                else value = cr_pow(value, invDestGamma);
            }
            const float scaled = value * 255.0f + 0.5f;
            const float lo = (0.0f < scaled) ? scaled : 0.0f;               // std::max((Float) 0, x): NaN -> 0
            out_rgb8[3 * i + k] = (uint8_t) ((lo < 255.0f) ? lo : 255.0f);  // std::min(255, .)
        }
    }
    return 0;
}

int cudapath_set_pixel_shard(cudapath_ctx *ctx, uint32_t shard_index, uint32_t shard_count) {
    if (!ctx) return fail("null context");
    if (shard_count < 1 || shard_count > 64 || shard_index >= shard_count) return fail("invalid pixel shard");
    ctx->shardIndex = shard_index; ctx->shardCount = shard_count;
    return 0;
}

int cudapath_set_math_mode(cudapath_ctx *ctx, int strict) {
    if (!ctx) return fail("null context");
    ctx->fastMath = strict ? 0 : 1;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_math_mode(p, strict); });
}
int cudapath_get_math_mode(cudapath_ctx *ctx) { return ctx ? (ctx->fastMath ? 0 : 1) : -1; }

int cudapath_set_build_options(cudapath_ctx *ctx, int max_split) {
    if (!ctx) return fail("null context");
    if (max_split < 1 || max_split > 64) return fail("max_split must be in [1, 64]");
    ctx->maxSplit = max_split; ctx->maxSplitExplicit = true; ctx->built = false;
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_build_options(p, max_split); });
}
int cudapath_set_job_size_hint(cudapath_ctx *ctx, uint64_t paths_per_device) {
    if (!ctx) return fail("null context");
    if (!ctx->maxSplitExplicit && !getenv("CUDAPATH_MAX_SPLIT")) {
        const int want = paths_per_device >= (1ull << 25) ? 16 : 8;
        if (want != ctx->maxSplit) { ctx->maxSplit = want; ctx->built = false; }
    }
    return fan_out(ctx, 0, [&](cudapath_ctx *p) { return cudapath_set_job_size_hint(p, paths_per_device); });
}
int cudapath_get_stats(cudapath_ctx *ctx, cudapath_stats *out) { if (!ctx || !out) return fail("null argument"); *out = ctx->stats; return 0; }
int cudapath_film_size(cudapath_ctx *ctx, int *w, int *h) {
    if (!ctx || !w || !h) return fail("null argument");
    if (!ctx->cam.present) return fail("no sensor set");
    *w = ctx->cam.w; *h = ctx->cam.h;
    return 0;
}
int cudapath_set_film_output(cudapath_ctx *ctx, int hdr, float gamma, float exposure) {
    if (!ctx) return fail("null context");
    if (gamma == 0) return fail("gamma must not be zero");
    ctx->filmHdr = hdr ? 1 : 0; ctx->filmGamma = gamma; ctx->filmExposure = exposure;
    return 0;
}
int cudapath_get_film_output(cudapath_ctx *ctx, int *hdr, float *gamma, float *exposure) {
    if (!ctx || !hdr || !gamma || !exposure) return fail("null argument");
    *hdr = ctx->filmHdr; *gamma = ctx->filmGamma; *exposure = ctx->filmExposure;
    return 0;
}
int cudapath_scene_bounds(cudapath_ctx *ctx, float aabb[6], float bs[4]) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    std::memcpy(aabb, ctx->sceneAABB, 24);
    for (int k = 0; k < 3; ++k) bs[k] = ctx->scene.env.bsCenter[k];
    bs[3] = ctx->scene.env.bsRadius;
    return 0;
}

} // extern "C"


// ------------------------------------------------------------------------------------------ multi-GPU contexts
// Replaces the tile scheduler of the reference for this path -- BlockedRenderProcess handing 32x32 tiles to the LocalWorkers that
// `mitsuba -p N` starts, and the mutex-protected Film::put of every finished tile (src/librender/renderproc.cpp:117-182,
// src/mitsuba/mitsuba.cpp:280-329).  Here every GPU holds the whole flattened scene and renders a contiguous range of SAMPLE INDICES
// of every pixel into a private full-size film (perfectly balanced wherever the hair is, no tile border to exchange; the RNG is
// keyed by pixel / sample / vertex, so the image does not depend on the device count up to fp32 summation order); the films are
// summed onto devices[0] with ONE ncclReduce over NVLink.  libnccl.so.2 is bound with dlopen when the first multi-GPU context is
// created: a process that already carries an NCCL (torch) shares it, a single-GPU process never needs it.
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = std::string("multi-GPU rendering needs NCCL: ") + dlerror(); return false; }
        auto sym = [&](const char *name) { void *p = dlsym(handle, name); if (!p) error = std::string("libnccl lacks ") + name; return p; };
        CommInitAll = (decltype(CommInitAll)) sym("ncclCommInitAll"); CommDestroy = (decltype(CommDestroy)) sym("ncclCommDestroy");
        Reduce = (decltype(Reduce)) sym("ncclReduce"); GroupStart = (decltype(GroupStart)) sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd)) sym("ncclGroupEnd"); GetErrorString = (decltype(GetErrorString)) sym("ncclGetErrorString");
        return CommInitAll && CommDestroy && Reduce && GroupStart && GroupEnd && GetErrorString;
    }
};
NcclApi g_nccl;
}

struct cudapath_multi {
    std::vector<cudapath_ctx *> all;        // devices[0] (the context the caller holds) first
    std::vector<ncclComm_t> comms;
    std::vector<float *> films;             // per-device film, kept between renders
    size_t filmBytes = 0;
    double reduceMs = 0;
};

static void multi_free(cudapath_multi *m) {
    if (!m) return;
    for (size_t i = 0; i < m->films.size(); ++i) if (m->films[i]) { DevGuard g(m->all[i]->device); cudaDeviceSynchronize(); dev_free(m->films[i]); }
    for (ncclComm_t c : m->comms) if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    delete m;
}

// contiguous, balanced split of [begin, end) over n parts (the first `rem` parts get one index more)
static void split_range(uint32_t begin, uint32_t end, int n, int i, uint32_t &b, uint32_t &e) {
    const uint32_t total = end - begin, base = total / (uint32_t) n, rem = total % (uint32_t) n;
    b = begin + (uint32_t) i * base + std::min<uint32_t>((uint32_t) i, rem);
    e = b + base + ((uint32_t) i < rem ? 1u : 0u);
}

static int render_multi(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *out_film) {
    cudapath_multi *m = ctx->multi;
    if (!m) return fail("not a multi-GPU context");
    if (!out_film) return fail("null film buffer");
    if (spp == 0 || sample_end > spp || sample_begin > sample_end) return fail("invalid sample range");
    const int n = (int) m->all.size();
    for (cudapath_ctx *c : m->all) if (!c->built) return fail("cudapath_build() has not been called");
    const size_t count = 5 * (size_t) ctx->cam.w * ctx->cam.h, bytes = sizeof(float) * count;
    if (m->filmBytes != bytes) {
        for (int i = 0; i < n; ++i) { DevGuard g(m->all[i]->device); cudaDeviceSynchronize(); if (m->films[i]) dev_free(m->films[i]); m->films[i] = nullptr; CKA(dev_alloc(&m->films[i], bytes)); }
        m->filmBytes = bytes;
    }
    // Which axis of the pixel-sample space is split: pixel blocks when the image has enough of them to balance the devices (every
    // device then keeps all sample indices of its pixels, i.e. the ray density of the single-device render), sample ranges otherwise.
    const uint32_t blocks = (uint32_t) ((ctx->cam.w + 31) / 32) * (uint32_t) ((ctx->cam.h + 31) / 32);
    const bool byPixels = blocks >= 16u * (uint32_t) n && !getenv("CUDAPATH_SHARD_SAMPLES");
    // one host thread per device: zero the private film, render this device's share
    std::vector<int> rc(n, 0); std::vector<std::string> msg(n);
    auto work = [&](int i) {
        cudapath_ctx *c = m->all[i];
        DevGuard g(c->device);
        uint32_t b = sample_begin, e = sample_end;
        if (byPixels) { c->shardIndex = (uint32_t) i; c->shardCount = (uint32_t) n; }
        else { c->shardIndex = 0; c->shardCount = 1; split_range(sample_begin, sample_end, n, i, b, e); }
        if (cudaMemsetAsync(m->films[i], 0, bytes, c->stream) != cudaSuccess) { rc[i] = -1; msg[i] = "film clear failed"; return; }
        c->stats.paths = c->stats.rays = c->stats.shadow_rays = c->stats.shadow_rays_traced = c->stats.kernel_launches = c->stats.bounces = 0; c->stats.render_ms = 0;
        if (e > b) { rc[i] = cudapath_render_dev(c, spp, seed, b, e, m->films[i], c->stream); if (rc[i] < 0) msg[i] = cudapath_last_error(); }
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for (auto &t : threads) t.join();
    for (int i = 0; i < n; ++i) if (rc[i] < 0) return fail(msg[i]);
    // ONE collective: sum of the private films onto devices[0], in place on the root
    cudaEvent_t e0, e1;
    {
        DevGuard g(ctx->device);
        CKA(cudaEventCreate(&e0)); CKA(cudaEventCreate(&e1));
        CKA(cudaEventRecord(e0, ctx->stream));
    }
    ncclResult_t nr = g_nccl.GroupStart();
    for (int i = 0; i < n && nr == ncclSuccess; ++i) {
        DevGuard g(m->all[i]->device);
        nr = g_nccl.Reduce(m->films[i], m->films[0], count, ncclFloat, ncclSum, 0, m->comms[i], m->all[i]->stream);
    }
    if (nr == ncclSuccess) nr = g_nccl.GroupEnd(); else g_nccl.GroupEnd();
    if (nr != ncclSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return fail(std::string("ncclReduce: ") + g_nccl.GetErrorString(nr)); }
    for (int i = 1; i < n; ++i) { DevGuard g(m->all[i]->device); CKA(cudaStreamSynchronize(m->all[i]->stream)); }
    {
        DevGuard g(ctx->device);
        CKA(cudaEventRecord(e1, ctx->stream));
        CKA(cudaMemcpyAsync(out_film, m->films[0], bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CKA(cudaStreamSynchronize(ctx->stream));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); m->reduceMs = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    // statistics of the job: counters add up, the render time is that of the slowest device
    cudapath_stats &s = ctx->stats;
    for (int i = 1; i < n; ++i) {
        const cudapath_stats &p = m->all[i]->stats;
        s.paths += p.paths; s.rays += p.rays; s.shadow_rays += p.shadow_rays; s.shadow_rays_traced += p.shadow_rays_traced; s.kernel_launches += p.kernel_launches;
        s.bounces += p.bounces; s.unsupported_filtered_lookups += p.unsupported_filtered_lookups; s.dropped_samples += p.dropped_samples;
        s.nodes_visited += p.nodes_visited; s.prims_tested += p.prims_tested; s.shadow_nodes_visited += p.shadow_nodes_visited; s.shadow_prims_tested += p.shadow_prims_tested;
        s.full_tests += p.full_tests; s.shadow_full_tests += p.shadow_full_tests; s.host_waits += p.host_waits;
        s.render_ms = std::max(s.render_ms, p.render_ms);
    }
    return 0;
}

extern "C" {

int cudapath_create_multi(const int *cuda_devices, int n_devices, cudapath_ctx **out) {
    if (!cuda_devices || !out) return fail("null argument");
    if (n_devices < 1) return fail("a multi-GPU context needs at least one device");
    for (int i = 0; i < n_devices; ++i) for (int j = 0; j < i; ++j) if (cuda_devices[i] == cuda_devices[j]) return fail("a device is listed twice");
    if (n_devices == 1) return cudapath_create(cuda_devices[0], out);
    if (!g_nccl.load()) return fail(g_nccl.error);
    std::unique_ptr<cudapath_multi, void (*)(cudapath_multi *)> m(new cudapath_multi(), multi_free);
    for (int i = 0; i < n_devices; ++i) {
        cudapath_ctx *c = nullptr;
        if (cudapath_create(cuda_devices[i], &c) != 0) { for (cudapath_ctx *p : m->all) delete p; m->all.clear(); return -1; }
        m->all.push_back(c);
    }
    m->films.assign(n_devices, nullptr);
    m->comms.assign(n_devices, nullptr);
    const ncclResult_t nr = g_nccl.CommInitAll(m->comms.data(), n_devices, cuda_devices);
    if (nr != ncclSuccess) { for (cudapath_ctx *p : m->all) delete p; m->all.clear(); return fail(std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nr)); }
    cudapath_ctx *root = m->all[0];
    root->peers.assign(m->all.begin() + 1, m->all.end());
    root->multi = m.release();
    *out = root;
    return 0;
}

int cudapath_device_count(cudapath_ctx *ctx) { return ctx ? 1 + (int) ctx->peers.size() : 0; }

int cudapath_visible_devices(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int cudapath_build(cudapath_ctx *ctx) {
    if (!ctx) return fail("null context");
    if (ctx->peers.empty()) return build_one(ctx);
    // every device builds its own copy of the BVH and the tables, concurrently
    const int n = 1 + (int) ctx->peers.size();
    std::vector<int> rc(n, 0); std::vector<std::string> msg(n);
    auto work = [&](int i) { cudapath_ctx *c = i == 0 ? ctx : ctx->peers[i - 1]; rc[i] = build_one(c); if (rc[i] < 0) msg[i] = cudapath_last_error(); };
    std::vector<std::thread> threads;
    for (int i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for (auto &t : threads) t.join();
    for (int i = 0; i < n; ++i) if (rc[i] < 0) return fail(msg[i]);
    return 0;
}

int cudapath_measure_read_bandwidth(cudapath_ctx *ctx, size_t bytes, int iterations, double *out_gb_per_s) {
    if (!ctx || !out_gb_per_s) return fail("null argument");
    CP_GUARD(ctx);
    std::string err;
    if (!read_bandwidth_probe(bytes, iterations, ctx->stream, *out_gb_per_s, err)) return fail(err);
    return 0;
}

double cudapath_last_reduce_ms(cudapath_ctx *ctx) { return ctx && ctx->multi ? ctx->multi->reduceMs : 0.0; }

} // extern "C"

// ------------------------------------------------------------------------------------------ host-buffer hooks
namespace {
struct DevBuf {
    void *p = nullptr; size_t bytes = 0;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t b) { bytes = b; return cudaMalloc(&p, b ? b : 1); }
    cudaError_t upload(const void *src, size_t b, cudaStream_t s) { cudaError_t e = alloc(b); if (e != cudaSuccess) return e; return cudaMemcpyAsync(p, src, b, cudaMemcpyHostToDevice, s); }
    cudaError_t download(void *dst, cudaStream_t s) const { return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, s); }
    template <typename T> T *as() { return (T *) p; }
};
}

extern "C" {

int cudapath_bsdf_eval_batch(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf a, b, e, p; std::string err;
    CKA(a.upload(wi, n * 12, ctx->stream)); CKA(b.upload(wo, n * 12, ctx->stream)); CKA(e.alloc(n * 12)); CKA(p.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_eval_batch_fast : bsdf_eval_batch)(ctx->scene, bsdf_id, n, a.as<float>(), b.as<float>(), e.as<float>(), p.as<float>(), ctx->stream, err, false, nullptr)) return fail(err);
    CKA(e.download(out_eval, ctx->stream)); CKA(p.download(out_pdf, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_bsdf_eval_batch_discrete(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf a, b, e, p; std::string err;
    CKA(a.upload(wi, n * 12, ctx->stream)); CKA(b.upload(wo, n * 12, ctx->stream)); CKA(e.alloc(n * 12)); CKA(p.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_eval_batch_fast : bsdf_eval_batch)(ctx->scene, bsdf_id, n, a.as<float>(), b.as<float>(), e.as<float>(), p.as<float>(), ctx->stream, err, true, nullptr)) return fail(err);
    CKA(e.download(out_eval, ctx->stream)); CKA(p.download(out_pdf, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// ... with texture coordinates per tuple (its.uv): the textured kinds (`diffuse` / `plastic` with a checkerboard); discrete != 0: EDiscrete measure
int cudapath_bsdf_eval_batch_uv(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, const float *uv, int discrete, float *out_eval, float *out_pdf) {
    if (require_built(ctx)) return -1;
    if (!wi || !wo || !uv || !out_eval || !out_pdf) return fail("null argument");
    CP_GUARD(ctx);
    DevBuf a, b, t, e, p; std::string err;
    CKA(a.upload(wi, n * 12, ctx->stream)); CKA(b.upload(wo, n * 12, ctx->stream)); CKA(t.upload(uv, n * 8, ctx->stream)); CKA(e.alloc(n * 12)); CKA(p.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_eval_batch_fast : bsdf_eval_batch)(ctx->scene, bsdf_id, n, a.as<float>(), b.as<float>(), e.as<float>(), p.as<float>(), ctx->stream, err, discrete != 0, t.as<float>())) return fail(err);
    CKA(e.download(out_eval, ctx->stream)); CKA(p.download(out_pdf, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_bsdf_sample_batch_uv(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, const float *uv, float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type) {
    if (require_built(ctx)) return -1;
    if (!wi || !sample || !uv || !out_wo || !out_weight || !out_pdf || !out_type) return fail("null argument");
    CP_GUARD(ctx);
    DevBuf a, s, x, wo, wt, p, t; std::string err;
    CKA(a.upload(wi, n * 12, ctx->stream)); CKA(s.upload(sample, n * 8, ctx->stream)); CKA(x.upload(uv, n * 8, ctx->stream));
    CKA(wo.alloc(n * 12)); CKA(wt.alloc(n * 12)); CKA(p.alloc(n * 4)); CKA(t.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_sample_batch_fast : bsdf_sample_batch)(ctx->scene, bsdf_id, n, a.as<float>(), s.as<float>(), nullptr, wo.as<float>(), wt.as<float>(), p.as<float>(), t.as<int32_t>(), ctx->stream, err, x.as<float>())) return fail(err);
    CKA(wo.download(out_wo, ctx->stream)); CKA(wt.download(out_weight, ctx->stream)); CKA(p.download(out_pdf, ctx->stream)); CKA(t.download(out_type, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_bsdf_eval_batch_world(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *frames, const float *wi_world, const float *wo_world, float *out_eval, float *out_pdf) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!frames || !wi_world || !wo_world || !out_eval || !out_pdf) return fail("null argument");
    DevBuf f, a, b, e, p; std::string err;
    CKA(f.upload(frames, n * 36, ctx->stream)); CKA(a.upload(wi_world, n * 12, ctx->stream)); CKA(b.upload(wo_world, n * 12, ctx->stream)); CKA(e.alloc(n * 12)); CKA(p.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_eval_world_batch_fast : bsdf_eval_world_batch)(ctx->scene, bsdf_id, n, f.as<float>(), a.as<float>(), b.as<float>(), e.as<float>(), p.as<float>(), ctx->stream, err)) return fail(err);
    CKA(e.download(out_eval, ctx->stream)); CKA(p.download(out_pdf, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_bsdf_sample_batch(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type) {
    return cudapath_bsdf_sample_batch_ex(ctx, bsdf_id, n, wi, sample, nullptr, out_wo, out_weight, out_pdf, out_type);
}
int cudapath_bsdf_sample_batch_ex(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, const float *extra, float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf a, s, x, wo, wt, p, t; std::string err;
    CKA(a.upload(wi, n * 12, ctx->stream)); CKA(s.upload(sample, n * 8, ctx->stream));
    if (extra) CKA(x.upload(extra, n * 16, ctx->stream));
    CKA(wo.alloc(n * 12)); CKA(wt.alloc(n * 12)); CKA(p.alloc(n * 4)); CKA(t.alloc(n * 4));
    if (!(ctx->fastMath ? bsdf_sample_batch_fast : bsdf_sample_batch)(ctx->scene, bsdf_id, n, a.as<float>(), s.as<float>(), extra ? x.as<float>() : nullptr, wo.as<float>(), wt.as<float>(), p.as<float>(), t.as<int32_t>(), ctx->stream, err, nullptr)) return fail(err);
    CKA(wo.download(out_wo, ctx->stream)); CKA(wt.download(out_weight, ctx->stream)); CKA(p.download(out_pdf, ctx->stream)); CKA(t.download(out_type, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_intersect_batch(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                             int any_hit, int32_t *out_shape, uint32_t *out_prim, float *out_t, float *out_record) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf o, d, mn, mx, sh, pr, t, rec; std::string err;
    CKA(o.upload(origin, n * 12, ctx->stream)); CKA(d.upload(direction, n * 12, ctx->stream)); CKA(mn.upload(mint, n * 4, ctx->stream)); CKA(mx.upload(maxt, n * 4, ctx->stream));
    CKA(sh.alloc(n * 4)); CKA(pr.alloc(n * 4)); CKA(t.alloc(n * 4));
    if (out_record) CKA(rec.alloc(n * 60));
    if (!intersect_batch(ctx->scene, n, o.as<float>(), d.as<float>(), mn.as<float>(), mx.as<float>(), any_hit, false, sh.as<int32_t>(), pr.as<uint32_t>(), t.as<float>(),
                         out_record ? rec.as<float>() : nullptr, nullptr, ctx->stream, err)) return fail(err);
    CKA(sh.download(out_shape, ctx->stream)); CKA(pr.download(out_prim, ctx->stream)); CKA(t.download(out_t, ctx->stream));
    if (out_record) CKA(rec.download(out_record, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// closest hit with the full record, its.uv and the geometric normal (out_uv_geo_n: 5 floats per ray)
int cudapath_intersect_batch_uv(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                                int32_t *out_shape, uint32_t *out_prim, float *out_t, float *out_record, float *out_uv_geo_n) {
    if (require_built(ctx)) return -1;
    if (!origin || !direction || !mint || !maxt || !out_shape || !out_prim || !out_t || !out_record || !out_uv_geo_n) return fail("null argument");
    CP_GUARD(ctx);
    DevBuf o, d, mn, mx, sh, pr, t, rec, uv; std::string err;
    CKA(o.upload(origin, n * 12, ctx->stream)); CKA(d.upload(direction, n * 12, ctx->stream)); CKA(mn.upload(mint, n * 4, ctx->stream)); CKA(mx.upload(maxt, n * 4, ctx->stream));
    CKA(sh.alloc(n * 4)); CKA(pr.alloc(n * 4)); CKA(t.alloc(n * 4)); CKA(rec.alloc(n * 60)); CKA(uv.alloc(n * 20));
    if (!intersect_batch(ctx->scene, n, o.as<float>(), d.as<float>(), mn.as<float>(), mx.as<float>(), 0, false, sh.as<int32_t>(), pr.as<uint32_t>(), t.as<float>(),
                         rec.as<float>(), nullptr, ctx->stream, err, uv.as<float>())) return fail(err);
    CKA(sh.download(out_shape, ctx->stream)); CKA(pr.download(out_prim, ctx->stream)); CKA(t.download(out_t, ctx->stream));
    CKA(rec.download(out_record, ctx->stream)); CKA(uv.download(out_uv_geo_n, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_env_eval_batch(cudapath_ctx *ctx, uint64_t n, const float *direction, float *out_rgb, float *out_pdf) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf d, c, p; std::string err;
    CKA(d.upload(direction, n * 12, ctx->stream)); CKA(c.alloc(n * 12)); CKA(p.alloc(n * 4));
    if (!(ctx->fastMath ? env_eval_batch_fast : env_eval_batch)(ctx->scene, n, d.as<float>(), c.as<float>(), p.as<float>(), ctx->stream, err)) return fail(err);
    CKA(c.download(out_rgb, ctx->stream)); CKA(p.download(out_pdf, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_env_eval_filtered_batch(cudapath_ctx *ctx, uint64_t n, const float *direction, const float *rx_direction, const float *ry_direction, float *out_rgb) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!direction || !rx_direction || !ry_direction || !out_rgb) return fail("null argument");
    DevBuf d, rx, ry, c; std::string err;
    CKA(d.upload(direction, n * 12, ctx->stream)); CKA(rx.upload(rx_direction, n * 12, ctx->stream)); CKA(ry.upload(ry_direction, n * 12, ctx->stream)); CKA(c.alloc(n * 12));
    if (!env_eval_filtered_batch(ctx->scene, n, d.as<float>(), rx.as<float>(), ry.as<float>(), c.as<float>(), ctx->stream, err)) return fail(err);
    CKA(c.download(out_rgb, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_env_mip_level(cudapath_ctx *ctx, int level, int *out_width, int *out_height, float *out_rgb) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!ctx->scene.env.present) return fail("no environment map set");
    EnvMipInfo info;
    CKA(cudaMemcpy(&info, ctx->envTables.mipInfo, sizeof(info), cudaMemcpyDeviceToHost));
    if (level < 0 || level >= info.levels) return fail("MIP level out of range");
    if (out_width) *out_width = info.w[level];
    if (out_height) *out_height = info.h[level];
    if (out_rgb) {
        const size_t n = (size_t) info.w[level] * info.h[level];
        std::vector<float4> t(n);
        CKA(cudaMemcpy(t.data(), level == 0 ? ctx->envTables.texels : ctx->envTables.mipTexels + info.offset[level], n * sizeof(float4), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) { out_rgb[3 * i] = t[i].x; out_rgb[3 * i + 1] = t[i].y; out_rgb[3 * i + 2] = t[i].z; }
    }
    return info.levels;
}
int cudapath_env_sample_batch(cudapath_ctx *ctx, uint64_t n, const float *ref_point, const float *sample, float *out_direction, float *out_value, float *out_pdf_dist) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf r, s, d, v, p; std::string err;
    CKA(r.upload(ref_point, n * 12, ctx->stream)); CKA(s.upload(sample, n * 8, ctx->stream)); CKA(d.alloc(n * 12)); CKA(v.alloc(n * 12)); CKA(p.alloc(n * 8));
    if (!(ctx->fastMath ? env_sample_batch_fast : env_sample_batch)(ctx->scene, n, r.as<float>(), s.as<float>(), d.as<float>(), v.as<float>(), p.as<float>(), ctx->stream, err)) return fail(err);
    CKA(d.download(out_direction, ctx->stream)); CKA(v.download(out_value, ctx->stream)); CKA(p.download(out_pdf_dist, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_camera_rays_batch(cudapath_ctx *ctx, uint64_t n, const float *pixel_sample, float *out_origin, float *out_direction, float *out_mint_maxt) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf p, o, d, m; std::string err;
    CKA(p.upload(pixel_sample, n * 8, ctx->stream)); CKA(o.alloc(n * 12)); CKA(d.alloc(n * 12)); CKA(m.alloc(n * 8));
    if (!camera_rays_batch(ctx->scene, n, p.as<float>(), o.as<float>(), d.as<float>(), m.as<float>(), ctx->stream, err)) return fail(err);
    CKA(o.download(out_origin, ctx->stream)); CKA(d.download(out_direction, ctx->stream)); CKA(m.download(out_mint_maxt, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_splat_batch(cudapath_ctx *ctx, uint64_t n, const float *position, const float *rgb, const float *alpha, float *out_film) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    DevBuf p, c, a, f; std::string err;
    const size_t fb = sizeof(float) * 5 * (size_t) ctx->cam.w * ctx->cam.h;
    CKA(p.upload(position, n * 8, ctx->stream)); CKA(c.upload(rgb, n * 12, ctx->stream)); CKA(a.upload(alpha, n * 4, ctx->stream));
    CKA(f.alloc(fb)); CKA(cudaMemsetAsync(f.p, 0, fb, ctx->stream));
    if (!splat_batch(ctx->scene, p.as<float>(), c.as<float>(), a.as<float>(), n, f.as<float>(), ctx->stream, err)) return fail(err);
    CKA(f.download(out_film, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cudapath_marschner_tables(cudapath_ctx *ctx, int bsdf_id, float *out_tables, float *out_pdfs, float *out_cdfs, float *out_sums, float *out_rt, float *out_consts) {
    if (!ctx) return fail("null context");
    if (bsdf_id < 0 || bsdf_id >= (int) ctx->bsdfs.size() || (ctx->bsdfs[bsdf_id].dev.kind != 1 && ctx->bsdfs[bsdf_id].dev.kind != 3)) return fail("not a marschner bsdf");
    CP_GUARD(ctx);
    const BsdfHost &b = ctx->bsdfs[bsdf_id];
    std::vector<float4> t(3 * 4096);
    CKA(cudaMemcpyAsync(t.data(), b.tables.tab, sizeof(float4) * t.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaMemcpyAsync(out_pdfs, b.tables.pdf, 4 * 3 * 4096, cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaMemcpyAsync(out_cdfs, b.tables.cdf, 4 * 3 * 64 * 65, cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaMemcpyAsync(out_sums, b.tables.sums, 4 * 3 * 64, cudaMemcpyDeviceToHost, ctx->stream));
    if (b.dev.rtSize) CKA(cudaMemcpyAsync(out_rt, b.rt, 4 * (size_t) b.dev.rtSize, cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < t.size(); ++i) { out_tables[3 * i] = t[i].x; out_tables[3 * i + 1] = t[i].y; out_tables[3 * i + 2] = t[i].z; }
    out_consts[0] = b.dev.Fdr; out_consts[1] = b.dev.specW; out_consts[2] = b.dev.eta; out_consts[3] = (float) b.dev.rtSize;
    return 0;
}
int cudapath_env_tables(cudapath_ctx *ctx, float *out_cdf_rows, float *out_cdf_cols, float *out_row_weights, float *out_normalization) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    if (!ctx->scene.env.present) return fail("no environment map set");
    const EnvDev &E = ctx->scene.env;
    CKA(cudaMemcpyAsync(out_cdf_rows, E.cdfRows, 4 * (size_t) (E.h + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaMemcpyAsync(out_cdf_cols, E.cdfCols, 4 * (size_t) (E.w + 1) * E.h, cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaMemcpyAsync(out_row_weights, E.rowWeights, 4 * (size_t) E.h, cudaMemcpyDeviceToHost, ctx->stream));
    CKA(cudaStreamSynchronize(ctx->stream));
    *out_normalization = E.normalization;
    return 0;
}
int cudapath_filter_table(cudapath_ctx *ctx, float out32[32]) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    std::memcpy(out32, ctx->scene.film.filterValues, 128);
    return 0;
}

// ------------------------------------------------------------------------------------------ device-resident variants
int cudapath_bsdf_eval_batch_dev(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf, void *stream) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    std::string err;
    if (!(ctx->fastMath ? bsdf_eval_batch_fast : bsdf_eval_batch)(ctx->scene, bsdf_id, n, wi, wo, out_eval, out_pdf, stream ? (cudaStream_t) stream : ctx->stream, err, false, nullptr)) return fail(err);
    return 0;
}
int cudapath_bsdf_sample_batch_dev(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type, void *stream) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    std::string err;
    if (!(ctx->fastMath ? bsdf_sample_batch_fast : bsdf_sample_batch)(ctx->scene, bsdf_id, n, wi, sample, nullptr, out_wo, out_weight, out_pdf, out_type, stream ? (cudaStream_t) stream : ctx->stream, err, nullptr)) return fail(err);
    return 0;
}
int cudapath_intersect_batch_dev(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                                 int any_hit, int32_t *out_shape, uint32_t *out_prim, float *out_t, unsigned long long *out_stats, void *stream) {
    if (require_built(ctx)) return -1;
    CP_GUARD(ctx);
    std::string err;
    if (!intersect_batch(ctx->scene, n, origin, direction, mint, maxt, any_hit, out_stats != nullptr, out_shape, out_prim, out_t, nullptr, out_stats,
                         stream ? (cudaStream_t) stream : ctx->stream, err)) return fail(err);
    return 0;
}

} // extern "C"

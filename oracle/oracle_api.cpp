// oracle/oracle_api.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points of the CPU oracle (a restatement of the reference's hair path-tracing hot path,
// see o_*.h for the file:line citations).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product never does.
//
// Parity status: the reference has no test, golden vector or fixture for hair, Marschner or Kajiya-Kay and
// cannot be built as a whole here, so the oracle is pinned against the reference's own source run as it is (oracle/Makefile `ref`,
// outputs in oracle/_ref, tests in tests/test_oracle_cpu.py): all eight BSDF plugin files and the path integrator (path.cpp) compiled
// unmodified against interface scaffolding (libref_bsdf.so, libref_path.so); the hair cylinder / miter / bounds / intersection-record
// code, the two-level ray query, the hair file loader, TriAccel, the AABB slab test, the envmap emitter, the perspective sensor with the fp32 Transform / Matrix algebra, ImageBlock::put with
// the reconstruction filters and the libcore helpers they
// call executed from text cut out of the reference at build time (libref_geom.so); GaussLegendre, InterpolatedDistribution1D and the
// Hosek-Wilkie sky model compiled directly (libref_pieces.so).  NOT pinned that way, restatement only: the sunsky bake around the sky
// model, the OBJ / RGBE readers and the film develop ("Parity status").
#include "o_math.h"
#include "o_hair.h"
#include "o_bsdf.h"
#include "o_env.h"
#include "o_render.h"
#include <thread>
#include <mutex>
#include <dlfcn.h>
#include <cstring>

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception &e) { g_err = e.what(); return -1; } catch (...) { g_err = "unknown error"; return -1; }

// Batch entry points split their index range over the host's cores (each tuple / ray is independent; results are position-wise).
template <class F> static void parallel_for(uint64_t n, F &&body) {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("ORC_THREADS")) nt = (unsigned) std::max(1, atoi(e));
    if (nt < 2 || n < 4096) { body((uint64_t) 0, n); return; }
    nt = (unsigned) std::min<uint64_t>(nt, n / 2048);
    std::vector<std::thread> th;
    std::exception_ptr error; std::mutex m;
    const uint64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([&, b, e]() { try { body(b, e); } catch (...) { std::lock_guard<std::mutex> g(m); error = std::current_exception(); } });
    }
    for (auto &t : th) t.join();
    if (error) std::rethrow_exception(error);
}

extern "C" {

// what accelerates the ray queries of this build (reported with the CPU baseline)
const char *orc_accel_description() { return orc::Geometry::accelDescription(); }
const char *orc_last_error() { return g_err.c_str(); }

void *orc_scene_create() { return new Scene(); }
void orc_scene_destroy(void *s) { delete (Scene *) s; }

int orc_add_bsdf_kajiyakay(void *sp, const float *diffuse, const float *specular, float exponent) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 0;
    b.kk.configure(V3(diffuse[0], diffuse[1], diffuse[2]), V3(specular[0], specular[1], specular[2]), exponent);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

int orc_add_bsdf_marschner(void *sp, float intIOR, float extIOR, const float *diffuse, const float *specRefl,
                           float alpha, int distribution, int nonlinear, const char *dataDir) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 1;
    b.ma = std::make_shared<Marschner>();
    b.ma->configure(intIOR, extIOR, V3(diffuse[0], diffuse[1], diffuse[2]), V3(specRefl[0], specRefl[1], specRefl[2]),
                    alpha, distribution, nonlinear != 0, dataDir);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// the unbuilt `Marschner` of src/bsdfs/marschner.cpp ("fixed" mode)
int orc_add_bsdf_marschner_fixed(void *sp, float intIOR, float extIOR) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 3;
    b.mf = std::make_shared<MarschnerFixed>();
    b.mf->configure(intIOR, extIOR);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// the same class with all three lobes in eval() and its constants taken from the scene (lobeMask: bit 0 R, 1 TT, 2 TRT)
int orc_add_bsdf_marschner_full(void *sp, float intIOR, float extIOR, const float *sigmaA, float betaR, float scaleAngleRad, int lobeMask) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 3;
    b.mf = std::make_shared<MarschnerFixed>();
    b.mf->configure(intIOR, extIOR, V3(sigmaA[0], sigmaA[1], sigmaA[2]), betaR, scaleAngleRad, lobeMask);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// `roughplastic` plugin (src/bsdfs/roughplastic.cpp); distribution: 0 beckmann, 1 ggx, 2 phong
int orc_add_bsdf_roughplastic(void *sp, float intIOR, float extIOR, const float *diffuse, const float *specular, float alpha, int distribution,
                              int sampleVisible, int nonlinear, const char *dataDir) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 4;
    b.rp = std::make_shared<RoughPlastic>();
    b.rp->configure(intIOR, extIOR, V3(diffuse[0], diffuse[1], diffuse[2]), V3(specular[0], specular[1], specular[2]), alpha, distribution,
                    sampleVisible != 0, nonlinear != 0, dataDir);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// `thindielectric` plugin (src/bsdfs/thindielectric.cpp)
int orc_add_bsdf_thindielectric(void *sp, float intIOR, float extIOR, const float *specR, const float *specT) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 5;
    b.td.configure(intIOR, extIOR, V3(specR[0], specR[1], specR[2]), V3(specT[0], specT[1], specT[2]));
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// `marschnerdielectric` plugin (src/bsdfs/marschnerdielectric.cpp)
int orc_add_bsdf_marschnerdielectric(void *sp, float intIOR, float extIOR, const float *diffuse, const float *specR, const float *specT, float exponent) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 6;
    b.md.configure(intIOR, extIOR, V3(diffuse[0], diffuse[1], diffuse[2]), V3(specR[0], specR[1], specR[2]), V3(specT[0], specT[1], specT[2]), exponent);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// `diffuse` plugin with a constant reflectance (src/bsdfs/diffuse.cpp), optionally inside `twosided`
int orc_add_bsdf_diffuse(void *sp, const float *reflectance, int twoSided) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 2;
    b.df.configure(V3(reflectance[0], reflectance[1], reflectance[2]), twoSided != 0);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}

// `plastic` plugin (src/bsdfs/plastic.cpp) with constant reflectances; a checkerboard can replace the diffuse one afterwards
int orc_add_bsdf_plastic(void *sp, float intIOR, float extIOR, const float *diffuse, const float *specular, int nonlinear) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 7;
    Texture2D d; d.setConstant(V3(diffuse[0], diffuse[1], diffuse[2]));
    b.pl.configure(intIOR, extIOR, d, V3(specular[0], specular[1], specular[2]), nonlinear != 0);
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}
// `mirror` plugin of the fork (src/bsdfs/mirror.cpp)
int orc_add_bsdf_mirror(void *sp, const float *specular) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny b; b.kind = 8;
    b.mi.configure(V3(specular[0], specular[1], specular[2]));
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
    ORC_CATCH
}
// <texture type="checkerboard"> as the (diffuse) reflectance of a `diffuse` or `plastic` BSDF: the BSDF is configured again with it
int orc_bsdf_set_checkerboard(void *sp, int bsdf, const float *color0, const float *color1, float uoffset, float voffset, float uscale, float vscale) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny &b = s->bsdfs.at(bsdf);
    Texture2D t; t.setCheckerboard(V3(color0[0], color0[1], color0[2]), V3(color1[0], color1[1], color1[2]), uoffset, voffset, uscale, vscale);
    if (b.kind == 2) b.df.configure(t, b.df.twoSided);
    else if (b.kind == 7) { const SmoothPlastic old = b.pl; b.pl.diffuse = t; b.pl.diffuse.ensureEnergyConservation();
                            b.pl.specularSamplingWeight = luminance(old.specR) / (luminance(b.pl.diffuse.getAverage()) + luminance(old.specR)); }
    else throw std::runtime_error("textured reflectance: only `diffuse` and `plastic` carry a texture on this path");
    return 0;
    ORC_CATCH
}
// <bsdf type="twosided"> around a roughplastic / plastic / diffuse BSDF (the same nested BRDF on both sides)
int orc_bsdf_set_twosided(void *sp, int bsdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    BSDFAny &b = s->bsdfs.at(bsdf);
    if (b.kind == 2) b.df.twoSided = true;
    else if (b.kind == 4 || b.kind == 7 || b.kind == 8) b.twoSided = true;
    else throw std::runtime_error("twosided: only materials without a transmission component can be nested");
    return 0;
    ORC_CATCH
}
int orc_plastic_constants(void *sp, int bsdf, float *out4) {
    ORC_TRY
    const BSDFAny &b = ((Scene *) sp)->bsdfs.at(bsdf);
    if (b.kind != 7) throw std::runtime_error("not a plastic BSDF");
    out4[0] = b.pl.fdrInt; out4[1] = b.pl.fdrExt; out4[2] = b.pl.specularSamplingWeight; out4[3] = b.pl.invEta2;
    return 0;
    ORC_CATCH
}
float orc_fresnel_diffuse_reflectance(float eta) { return fresnelDiffuseReflectance(eta); }
// eval / pdf / sample with texture coordinates (its.uv); measure: 0 solid angle, 1 discrete
int orc_bsdf_eval_batch_uv(void *sp, int bsdf, uint64_t n, const float *wi, const float *wo, const float *uv, int discrete, float *outEval, float *outPdf) {
    ORC_TRY
    const BSDFAny &b = ((Scene *) sp)->bsdfs.at(bsdf);
    for (uint64_t i = 0; i < n; ++i) {
        V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), c(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        V3 e = b.eval(a, c, discrete != 0, uv[2 * i], uv[2 * i + 1]);
        outEval[3 * i] = e.x; outEval[3 * i + 1] = e.y; outEval[3 * i + 2] = e.z;
        outPdf[i] = b.pdf(a, c, discrete != 0);
    }
    return 0;
    ORC_CATCH
}
int orc_bsdf_sample_batch_uv(void *sp, int bsdf, uint64_t n, const float *wi, const float *sample, const float *uv, float *outWo, float *outWeight,
                             float *outPdf, int32_t *outType) {
    ORC_TRY
    const BSDFAny &b = ((Scene *) sp)->bsdfs.at(bsdf);
    const float zero[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
        V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        BSDFSample r = b.sample(a, sample[2 * i], sample[2 * i + 1], zero, uv[2 * i], uv[2 * i + 1]);
        outWo[3 * i] = r.wo.x; outWo[3 * i + 1] = r.wo.y; outWo[3 * i + 2] = r.wo.z;
        outWeight[3 * i] = r.weight.x; outWeight[3 * i + 1] = r.weight.y; outWeight[3 * i + 2] = r.weight.z;
        outPdf[i] = r.pdf; outType[i] = r.sampledType | (r.sampledComponent << 8);
    }
    return 0;
    ORC_CATCH
}
// `rectangle` shape (src/shapes/rectangle.cpp): toWorld as one row-major 4x4 matrix, inverted like Transform(const Matrix4x4 &) does
int orc_add_rectangle(void *sp, const float *toWorld16, int flipNormals, int bsdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    HairShape h;
    const M44 tw = M44::fromRowMajor(toWorld16); M44 inv;
    if (!invert(tw, inv)) throw std::runtime_error("rectangle: singular toWorld transform");
    h.mesh.isRect = true; h.mesh.rect.configure(tw, inv, flipNormals != 0);
    h.bsdf = bsdf;
    h.finalizeMesh((uint32_t) s->geo.shapes.size());
    s->geo.shapes.push_back(std::move(h));
    return (int) s->geo.shapes.size() - 1;
    ORC_CATCH
}
// texture coordinates and geometric normal of the closest hit (its.uv, its.geoFrame.n): 5 floats per ray
int orc_intersect_uv_batch(void *sp, uint64_t n, const float *o, const float *d, const float *mint, const float *maxt, float *out5) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i]);
        Intersection its;
        const bool hit = s->geo.rayIntersect(r, its);
        float *q = out5 + 5 * i;
        q[0] = hit ? its.u : 0; q[1] = hit ? its.v : 0; q[2] = hit ? its.geoFrame.n.x : 0; q[3] = hit ? its.geoFrame.n.y : 0; q[4] = hit ? its.geoFrame.n.z : 0;
    }
    return 0;
    ORC_CATCH
}

// Triangle mesh as TriMesh exposes it after configure(): positions, optional vertex normals (null = face normals), indices
int orc_add_mesh_uv(void *sp, const float *xyz, const float *normals, const float *uvs, uint32_t nVerts, const uint32_t *indices, uint32_t nTris, int bsdf);
int orc_add_mesh(void *sp, const float *xyz, const float *normals, uint32_t nVerts, const uint32_t *indices, uint32_t nTris, int bsdf) {
    return orc_add_mesh_uv(sp, xyz, normals, nullptr, nVerts, indices, nTris, bsdf);
}
// ... with optional per-vertex texture coordinates (TriMesh::getVertexTexcoords)
int orc_add_mesh_uv(void *sp, const float *xyz, const float *normals, const float *uvs, uint32_t nVerts, const uint32_t *indices, uint32_t nTris, int bsdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    HairShape h;
    if (uvs) h.mesh.uv.assign(uvs, uvs + 2 * (size_t) nVerts);
    h.mesh.pos.resize(nVerts);
    for (uint32_t i = 0; i < nVerts; ++i) h.mesh.pos[i] = V3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (normals) { h.mesh.nrm.resize(nVerts); for (uint32_t i = 0; i < nVerts; ++i) h.mesh.nrm[i] = V3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]); }
    h.mesh.idx.assign(indices, indices + 3 * (size_t) nTris);
    for (uint32_t i : h.mesh.idx) if (i >= nVerts) throw std::runtime_error("mesh index out of range");
    h.bsdf = bsdf;
    h.finalizeMesh((uint32_t) s->geo.shapes.size());
    s->geo.shapes.push_back(std::move(h));
    return (int) s->geo.shapes.size() - 1;
    ORC_CATCH
}

int orc_add_hair(void *sp, const float *xyz, const uint8_t *startsFiber, uint32_t n, float radius, int bsdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    HairShape h;
    h.verts.resize(n);
    for (uint32_t i = 0; i < n; ++i) h.verts[i] = V3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    h.startsFiber.assign(startsFiber, startsFiber + n);
    h.startsFiber.push_back(1);
    h.radius = radius; h.bsdf = bsdf;
    h.finalize();
    s->geo.shapes.push_back(std::move(h));
    return (int) s->geo.shapes.size() - 1;
    ORC_CATCH
}

// Radiance RGBE file -> top-down RGB fp32; sizes always, pixels when out is not null
int orc_load_rgbe(const char *path, float *out, int *w, int *h) {
    ORC_TRY
    RGBEImage img = readRGBE(path);
    *w = img.w; *h = img.h;
    if (out) std::memcpy(out, img.rgb.data(), img.rgb.size() * sizeof(float));
    return 0;
    ORC_CATCH
}

// Two-phase hair file load: returns a handle, then sizes / copies.
void *orc_hair_file_load_reduced(const char *path, float radius, float angleThresholdDeg, const float *toWorld16, float reduction) {
    try {
        HairShape *h = new HairShape();
        loadHairFile(path, radius, angleThresholdDeg, M44::fromRowMajor(toWorld16), *h, reduction);
        return h;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *orc_hair_file_load(const char *path, float radius, float angleThresholdDeg, const float *toWorld16) { return orc_hair_file_load_reduced(path, radius, angleThresholdDeg, toWorld16, 0.0f); }
// Random(seed).nextFloat(), n times
void orc_random_floats(uint64_t seed, uint64_t n, float *out) { MitsubaRandom r(seed); for (uint64_t i = 0; i < n; ++i) out[i] = r.nextFloat(); }
uint32_t orc_hair_file_vertex_count(void *h) { return (uint32_t) ((HairShape *) h)->verts.size(); }
uint32_t orc_hair_file_segment_count(void *h) { return (uint32_t) ((HairShape *) h)->segIndex.size(); }
float orc_hair_file_radius(void *h) { return ((HairShape *) h)->radius; }
void orc_hair_file_copy(void *hp, float *xyz, uint8_t *startsFiber) {
    HairShape *h = (HairShape *) hp;
    for (size_t i = 0; i < h->verts.size(); ++i) { xyz[3 * i] = h->verts[i].x; xyz[3 * i + 1] = h->verts[i].y; xyz[3 * i + 2] = h->verts[i].z; startsFiber[i] = h->startsFiber[i]; }
}
void orc_hair_file_free(void *h) { delete (HairShape *) h; }

int orc_set_envmap(void *sp, const float *rgb, int w, int h, const float *toWorld16, float scale) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    s->env.init(rgb, w, h, M44::fromRowMajor(toWorld16), scale);
    s->hasEnv = true;
    return 0;
    ORC_CATCH
}
int orc_set_camera(void *sp, const float *toWorld16, float fovX, float nearClip, float farClip, int w, int h) {
    Scene *s = (Scene *) sp;
    s->cam.toWorld = M44::fromRowMajor(toWorld16); s->cam.xfov = fovX; s->cam.nearClip = nearClip; s->cam.farClip = farClip;
    s->cam.filmW = w; s->cam.filmH = h;
    return 0;
}
int orc_set_film(void *sp, int filterType, float param, int hasAlpha) {
    Scene *s = (Scene *) sp;
    s->filter.type = filterType;
    if (filterType == 0 && param > 0) s->filter.radius = param;
    if (filterType == 2 && param > 0) s->filter.stddev = param;
    s->filmHasAlpha = hasAlpha != 0;
    return 0;
}
int orc_set_integrator(void *sp, int maxDepth, int rrDepth, int strictNormals, int hideEmitters) {
    Scene *s = (Scene *) sp;
    s->maxDepth = maxDepth; s->rrDepth = rrDepth; s->strictNormals = strictNormals != 0; s->hideEmitters = hideEmitters != 0;
    return 0;
}
int orc_finalize(void *sp) {
    ORC_TRY
    ((Scene *) sp)->finalize();
    return 0;
    ORC_CATCH
}

int orc_scene_bounds(void *sp, float *aabb6, float *bsphere4) {
    Scene *s = (Scene *) sp;
    aabb6[0] = s->geo.aabb.mn.x; aabb6[1] = s->geo.aabb.mn.y; aabb6[2] = s->geo.aabb.mn.z;
    aabb6[3] = s->geo.aabb.mx.x; aabb6[4] = s->geo.aabb.mx.y; aabb6[5] = s->geo.aabb.mx.z;
    bsphere4[0] = s->env.sceneBSphere.center.x; bsphere4[1] = s->env.sceneBSphere.center.y; bsphere4[2] = s->env.sceneBSphere.center.z; bsphere4[3] = s->env.sceneBSphere.radius;
    return 0;
}

int orc_segment_bounds(void *sp, int shape, float *out6n) {
    Scene *s = (Scene *) sp;
    const HairShape &h = s->geo.shapes[shape];
    for (size_t i = 0; i < h.segIndex.size(); ++i) {
        AABB b = h.segmentAABB(h.segIndex[i]);
        float *o = out6n + 6 * i;
        o[0] = b.mn.x; o[1] = b.mn.y; o[2] = b.mn.z; o[3] = b.mx.x; o[4] = b.mx.y; o[5] = b.mx.z;
    }
    return 0;
}

int orc_bsdf_eval_batch(void *sp, int bsdf, uint64_t n, const float *wi, const float *wo, float *outEval, float *outPdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    const BSDFAny &b = s->bsdfs.at(bsdf);
    parallel_for(n, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), c(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
            V3 e = b.eval(a, c);
            outEval[3 * i] = e.x; outEval[3 * i + 1] = e.y; outEval[3 * i + 2] = e.z;
            outPdf[i] = b.pdf(a, c);
        }
    });
    return 0;
    ORC_CATCH
}
// the same in the discrete measure (EDiscrete, common.h:56-67): non-zero only for the delta components of `thindielectric`
int orc_bsdf_eval_batch_discrete(void *sp, int bsdf, uint64_t n, const float *wi, const float *wo, float *outEval, float *outPdf) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    const BSDFAny &b = s->bsdfs.at(bsdf);
    for (uint64_t i = 0; i < n; ++i) {
        V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), c(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        V3 e = b.eval(a, c, true);
        outEval[3 * i] = e.x; outEval[3 * i + 1] = e.y; outEval[3 * i + 2] = e.z;
        outPdf[i] = b.pdf(a, c, true);
    }
    return 0;
    ORC_CATCH
}
// `extra` (optional, 4 per tuple): the additional sampler draws of the fixed Marschner
int orc_bsdf_sample_batch(void *sp, int bsdf, uint64_t n, const float *wi, const float *sample, const float *extra, float *outWo, float *outWeight,
                          float *outPdf, int32_t *outType) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    const BSDFAny &b = s->bsdfs.at(bsdf);
    const float zero[4] = {0, 0, 0, 0};
    parallel_for(n, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            V3 a(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
            BSDFSample r = b.sample(a, sample[2 * i], sample[2 * i + 1], extra ? extra + 4 * i : zero);
            outWo[3 * i] = r.wo.x; outWo[3 * i + 1] = r.wo.y; outWo[3 * i + 2] = r.wo.z;
            outWeight[3 * i] = r.weight.x; outWeight[3 * i + 1] = r.weight.y; outWeight[3 * i + 2] = r.weight.z;
            outPdf[i] = r.pdf; outType[i] = r.sampledType | (r.sampledComponent << 8);
        }
    });
    return 0;
    ORC_CATCH
}

// Marschner precomputed tables: out = 3 lobes x 64 x 64 x RGB; cdfs = 3 x 64 x 65; sums = 3 x 64; rt = 100 T samples + Fdr const
int orc_marschner_tables(void *sp, int bsdf, float *outTables, float *outPdfs, float *outCdfs, float *outSums, float *outRT, float *outConsts) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    const BSDFAny &b = s->bsdfs.at(bsdf);
    if (b.kind != 1 && b.kind != 3) throw std::runtime_error("not a marschner bsdf");
    const Marschner *m = b.kind == 1 ? b.ma.get() : &b.mf->base;
    const Azimuthal *lobes[3] = {&m->nR, &m->nTT, &m->nTRT};
    for (int l = 0; l < 3; ++l) {
        for (int i = 0; i < 64 * 64; ++i) { outTables[(l * 4096 + i) * 3] = lobes[l]->table[i].x; outTables[(l * 4096 + i) * 3 + 1] = lobes[l]->table[i].y; outTables[(l * 4096 + i) * 3 + 2] = lobes[l]->table[i].z; }
        std::memcpy(outPdfs + l * 4096, lobes[l]->sampler.pdfs.data(), 4096 * 4);
        std::memcpy(outCdfs + l * 64 * 65, lobes[l]->sampler.cdfs.data(), 64 * 65 * 4);
        std::memcpy(outSums + l * 64, lobes[l]->sampler.sums.data(), 64 * 4);
    }
    if (b.kind == 3) { outConsts[0] = 0; outConsts[1] = 0; outConsts[2] = m->eta; outConsts[3] = 0; return 0; }   // no rough-transmittance data
    std::memcpy(outRT, b.ma->extRT.trans.data(), b.ma->extRT.thetaSamples * 4);
    outConsts[0] = 1 - b.ma->intRT.evalDiffuse(b.ma->alpha); // Fdr
    outConsts[1] = b.ma->specularSamplingWeight;
    outConsts[2] = b.ma->eta;
    outConsts[3] = (float) b.ma->extRT.thetaSamples;
    return 0;
    ORC_CATCH
}

// mode: 0 = BVH closest, 1 = BVH any-hit, 2 = brute closest, 3 = brute any-hit
int orc_intersect_batch(void *sp, uint64_t n, const float *o, const float *d, const float *mint, const float *maxt, int mode,
                        int32_t *outShape, uint32_t *outPrim, float *outT) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    parallel_for(n, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            Ray r(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i]);
            Hit h; bool hit;
            if (mode < 2) hit = s->geo.intersectBVH(r, mode == 1, h);
            else hit = s->geo.intersectBrute(r, mode == 3, h);
            outShape[i] = hit ? h.shape : -1; outPrim[i] = hit ? h.iv : 0xffffffffu; outT[i] = hit ? h.t : kInf;
        }
    });
    return 0;
    ORC_CATCH
}
// All segments hit by the ray inside [mint,maxt] with their t (tie analysis for the bit-exact prim-id test)
int orc_intersect_candidates(void *sp, const float *o, const float *d, float mint, float maxt, int maxOut, int32_t *outShape, uint32_t *outPrim, float *outT) {
    Scene *s = (Scene *) sp;
    Ray r(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), mint, maxt);
    int cnt = 0;
    float smin, smax;
    if (!s->geo.sceneInterval(r, false, smin, smax)) return 0;
    for (size_t si = 0; si < s->geo.shapes.size(); ++si) {
        const HairShape &h = s->geo.shapes[si];
        float a, b;
        if (h.isMesh) {
            for (uint32_t j = 0; j < h.mesh.triCount(); ++j) {
                float t, u, v;
                if (h.mesh.intersectPrim(j, r.o, r.d, smin, smax, u, v, t) && cnt < maxOut) { outShape[cnt] = (int) si; outPrim[cnt] = j; outT[cnt] = t; cnt++; }
            }
            continue;
        }
        if (!s->geo.shapeInterval(h, r, smin, smax, a, b)) continue;
        for (uint32_t iv : h.segIndex) {
            float t; V3 p;
            if (h.intersect(r, iv, a, b, t, p) && cnt < maxOut) { outShape[cnt] = (int) si; outPrim[cnt] = iv; outT[cnt] = t; cnt++; }
        }
    }
    return cnt;
}
// Raw per-segment operations (pinning against the reference text itself, tests/test_oracle_cpu.py): HairKDTree::intersect for
// (ray, segment, interval) tuples, and HairShape::fillIntersectionRecord for (segment, stored hit point) pairs: out = p n s t
int orc_segment_intersect_batch(void *sp, int shape, uint64_t n, const float *o, const float *d, const uint32_t *iv, const float *mint, const float *maxt,
                                int32_t *outHit, float *outT, float *outP) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    const HairShape &h = s->geo.shapes.at(shape);
    for (uint64_t i = 0; i < n; ++i) {
        Ray r(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i]);
        float t = 0; V3 p(0.0f);
        const bool hit = h.intersect(r, iv[i], mint[i], maxt[i], t, p);
        outHit[i] = hit ? 1 : 0; outT[i] = hit ? t : 0.0f;
        outP[3 * i] = hit ? p.x : 0; outP[3 * i + 1] = hit ? p.y : 0; outP[3 * i + 2] = hit ? p.z : 0;
    }
    return 0;
    ORC_CATCH
}
int orc_segment_record_batch(void *sp, int shape, uint64_t n, const uint32_t *iv, const float *p, float *out12) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        Hit h; h.shape = shape; h.iv = iv[i]; h.t = 1.0f; h.p = V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        Ray r(V3(0.0f), V3(0.0f, 0.0f, 1.0f), 0.0f, kInf);
        Intersection its;
        s->geo.fillIntersection(r, h, its);
        const V3 vs[4] = {its.p, its.geoFrame.n, its.geoFrame.s, its.geoFrame.t};
        for (int k = 0; k < 4; ++k) { out12[12 * i + 3 * k] = vs[k].x; out12[12 * i + 3 * k + 1] = vs[k].y; out12[12 * i + 3 * k + 2] = vs[k].z; }
    }
    return 0;
    ORC_CATCH
}
// TriAccel::load + rayIntersect and AABB::rayIntersect on flat tuples (pinning against the reference text)
int orc_triaccel_batch(uint64_t n, const float *A, const float *B, const float *C, const float *o, const float *d, const float *mint, const float *maxt,
                       float *outAccel, int32_t *outHit, float *outTUV) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) {
        TriAccel acc;
        acc.load(V3(A[3 * i], A[3 * i + 1], A[3 * i + 2]), V3(B[3 * i], B[3 * i + 1], B[3 * i + 2]), V3(C[3 * i], C[3 * i + 1], C[3 * i + 2]));
        const float vals[10] = {(float) acc.k, acc.n_u, acc.n_v, acc.n_d, acc.a_u, acc.a_v, acc.b_nu, acc.b_nv, acc.c_nu, acc.c_nv};
        for (int k = 0; k < 10; ++k) outAccel[10 * i + k] = acc.k == 3 && k > 0 ? 0.0f : vals[k];
        float u = 0, v = 0, t = 0;
        const bool hit = acc.rayIntersect(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i], u, v, t);
        outHit[i] = hit ? 1 : 0; outTUV[3 * i] = hit ? t : 0; outTUV[3 * i + 1] = hit ? u : 0; outTUV[3 * i + 2] = hit ? v : 0;
    }
    return 0;
    ORC_CATCH
}
int orc_aabb_ray_batch(uint64_t n, const float *bmin, const float *bmax, const float *o, const float *d, int32_t *outHit, float *outNearFar) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) {
        AABB b; b.mn = V3(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2]); b.mx = V3(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2]);
        Ray r(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), 0.0f, kInf);
        float nearT = 0, farT = 0;
        const bool hit = b.rayIntersect(r.o, r.d, r.dRcp, nearT, farT);
        outHit[i] = hit ? 1 : 0; outNearFar[2 * i] = hit ? nearT : 0; outNearFar[2 * i + 1] = hit ? farT : 0;
    }
    return 0;
    ORC_CATCH
}
// closest hit + intersection record: outRec = p(3) n(3) s(3) t(3) wi(3) per ray
int orc_intersect_full_batch(void *sp, uint64_t n, const float *o, const float *d, const float *mint, const float *maxt,
                             int32_t *outShape, uint32_t *outPrim, float *outT, float *outRec) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i]);
        Intersection its;
        bool hit = s->geo.rayIntersect(r, its);
        outShape[i] = hit ? its.shape : -1; outPrim[i] = hit ? its.iv : 0xffffffffu; outT[i] = hit ? its.t : kInf;
        float *rec = outRec + 15 * i;
        V3 vs[5] = {its.p, its.shFrame.n, its.shFrame.s, its.shFrame.t, its.wi};
        for (int k = 0; k < 5; ++k) { rec[3 * k] = hit ? vs[k].x : 0; rec[3 * k + 1] = hit ? vs[k].y : 0; rec[3 * k + 2] = hit ? vs[k].z : 0; }
    }
    return 0;
    ORC_CATCH
}

int orc_camera_rays(void *sp, uint64_t n, const float *pxy, float *outO, float *outD, float *outMinMax) {
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r; V3 rx, ry;
        s->cam.sampleRayDifferential(pxy[2 * i], pxy[2 * i + 1], r, rx, ry);
        outO[3 * i] = r.o.x; outO[3 * i + 1] = r.o.y; outO[3 * i + 2] = r.o.z;
        outD[3 * i] = r.d.x; outD[3 * i + 1] = r.d.y; outD[3 * i + 2] = r.d.z;
        outMinMax[2 * i] = r.mint; outMinMax[2 * i + 1] = r.maxt;
    }
    return 0;
}

int orc_camera_differentials(void *sp, float *out6) {
    Scene *s = (Scene *) sp;
    out6[0] = s->cam.dx.x; out6[1] = s->cam.dx.y; out6[2] = s->cam.dx.z; out6[3] = s->cam.dy.x; out6[4] = s->cam.dy.y; out6[5] = s->cam.dy.z;
    return 0;
}
int orc_matrix_invert(const float *m16, float *out16) {
    M44 a = M44::fromRowMajor(m16), b;
    const bool ok = invert(a, b);
    std::memcpy(out16, b.m, 64);
    return ok ? 1 : 0;
}

// Environment map hooks: eval (no differentials), sampleDirect from `ref`, pdfDirect
int orc_env_eval_batch(void *sp, uint64_t n, const float *d, float *outRGB, float *outPdf) {
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        V3 dir(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        V3 v = s->env.evalEnvironment(dir);
        outRGB[3 * i] = v.x; outRGB[3 * i + 1] = v.y; outRGB[3 * i + 2] = v.z;
        outPdf[i] = s->env.pdfDirect(dir);
    }
    return 0;
}
// evalEnvironment of rays with differentials (EWA lookups in the MIP pyramid)
int orc_env_eval_filtered_batch(void *sp, uint64_t n, const float *d, const float *rx, const float *ry, float *outRGB) {
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        V3 v = s->env.evalEnvironment(V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), true, V3(rx[3 * i], rx[3 * i + 1], rx[3 * i + 2]), V3(ry[3 * i], ry[3 * i + 1], ry[3 * i + 2]));
        outRGB[3 * i] = v.x; outRGB[3 * i + 1] = v.y; outRGB[3 * i + 2] = v.z;
    }
    return 0;
}
// one level of the MIP pyramid (half-quantised texels as fp32); returns the number of levels
int orc_env_mip_level(void *sp, int level, int *outW, int *outH, float *outRGB) {
    Scene *s = (Scene *) sp;
    const int n = (int) s->env.levels.size();
    if (level < 0 || level >= n) return -1;
    const int w = s->env.levels[level].w, h = s->env.levels[level].h;
    if (outW) *outW = w;
    if (outH) *outH = h;
    if (outRGB) for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) { V3 t = s->env.evalTexel(level, x, y); float *o = outRGB + 3 * ((size_t) y * w + x); o[0] = t.x; o[1] = t.y; o[2] = t.z; }
    return n;
}
int orc_env_sample_batch(void *sp, uint64_t n, const float *ref, const float *sample, float *outD, float *outValue, float *outPdfDist) {
    Scene *s = (Scene *) sp;
    for (uint64_t i = 0; i < n; ++i) {
        EnvMap::DirectSample r = s->env.sampleDirect(V3(ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]), sample[2 * i], sample[2 * i + 1]);
        outD[3 * i] = r.d.x; outD[3 * i + 1] = r.d.y; outD[3 * i + 2] = r.d.z;
        outValue[3 * i] = r.value.x; outValue[3 * i + 1] = r.value.y; outValue[3 * i + 2] = r.value.z;
        outPdfDist[2 * i] = r.pdf; outPdfDist[2 * i + 1] = r.dist;
    }
    return 0;
}
int orc_env_tables(void *sp, float *outCdfRows, float *outCdfCols, float *outRowWeights, float *outNormalization, uint16_t *outTexels) {
    Scene *s = (Scene *) sp;
    std::memcpy(outCdfRows, s->env.cdfRows.data(), s->env.cdfRows.size() * 4);
    std::memcpy(outCdfCols, s->env.cdfCols.data(), s->env.cdfCols.size() * 4);
    std::memcpy(outRowWeights, s->env.rowWeights.data(), s->env.rowWeights.size() * 4);
    *outNormalization = s->env.normalization;
    if (outTexels) std::memcpy(outTexels, s->env.texels.data(), s->env.texels.size() * 2);
    return 0;
}

int orc_filter_table(void *sp, float *out32) { Scene *s = (Scene *) sp; std::memcpy(out32, s->filter.values, 32 * 4); return 0; }

// Render sample range [sBegin,sEnd) of spp into outFilm (5 x w x h, accumulated sums; not normalised)
// sampler: 0 Philox counters (default), 1 the reference's `sobol` sampler with `scramble`; the tables come from <dataDir>/sobol.bin
int orc_set_sampler(void *sp, int kind, uint64_t scramble, const char *dataDir) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    if (kind != 0 && kind != 1) throw std::runtime_error("sampler: 0 = philox, 1 = sobol");
    if (kind == 1 && !s->sobolTables) { s->sobolTables = std::make_shared<SobolTables>(); s->sobolTables->load(std::string(dataDir) + "/sobol.bin"); }
    s->samplerKind = kind; s->sobolScramble = scramble;
    return 0;
    ORC_CATCH
}
// raw sampler sequence for one pixel: `pattern` holds 1 / 2 per request (next1D / next2D), per sample index; out gets the numbers in order
int orc_sobol_sequence(void *sp, int px, int py, uint32_t firstSample, uint32_t nSamples, int nReq, const int32_t *pattern, float *out) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    if (!s->sobolTables) throw std::runtime_error("orc_set_sampler(1) first");
    SobolSampler sb = s->makeSobol();
    sb.generate(px, py);
    for (uint32_t k = 0; k < firstSample; ++k) sb.advance();
    for (uint32_t j = 0; j < nSamples; ++j) {
        for (int r = 0; r < nReq; ++r) { if (pattern[r] == 1) *out++ = sb.next1D(); else { float a, b; sb.next2D(a, b); *out++ = a; *out++ = b; } }
        sb.advance();
    }
    return 0;
    ORC_CATCH
}

int orc_render(void *sp, uint32_t spp, uint64_t seed, uint32_t sBegin, uint32_t sEnd, int nThreads, float *outFilm, uint64_t *outStats) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    s->seed = seed;
    s->stats.rays = 0; s->stats.shadowRays = 0; s->stats.paths = 0; s->stats.pathLength = 0; s->stats.dropped = 0;
    Film film; film.init(s->cam.filmW, s->cam.filmH);
    s->render(film, spp, sBegin, sEnd, nThreads);
    std::memcpy(outFilm, film.data.data(), film.data.size() * 4);
    if (outStats) { outStats[0] = s->stats.rays; outStats[1] = s->stats.shadowRays; outStats[2] = s->stats.paths; outStats[3] = s->stats.pathLength; outStats[4] = s->stats.dropped; outStats[5] = (uint64_t) s->env.unsupportedFiltered; }
    return 0;
    ORC_CATCH
}

// Per-sample radiance for path-by-path comparison: pixels (x,y), sample index -> Li, sample position, depth
int orc_render_samples(void *sp, uint64_t n, const uint32_t *xy, const uint32_t *samp, uint32_t spp, uint64_t seed, float *outLi, float *outPos) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    s->seed = seed;
    Film film; film.init(s->cam.filmW, s->cam.filmH);
    for (uint64_t i = 0; i < n; ++i) {
        V3 L; float pos[2];
        s->renderSample(film, xy[2 * i], xy[2 * i + 1], samp[i], spp, &L, pos);
        outLi[3 * i] = L.x; outLi[3 * i + 1] = L.y; outLi[3 * i + 2] = L.z;
        outPos[2 * i] = pos[0]; outPos[2 * i + 1] = pos[1];
    }
    return 0;
    ORC_CATCH
}

// ---------------------------------------------------------------------------------------------------------------------------------
// The reference's own MIPathTracer::Li (src/integrators/path/path.cpp compiled unmodified into oracle/_ref/libref_path.so) run on THIS
// scene: geometry, BSDFs, emitter and random numbers are supplied through the callback table of ref_shim/ref_path_callbacks.h, so the
// outcome isolates the integrator logic.  Same camera rays and Philox counters as renderSample(); compared with orc_render_samples().
#include "ref_shim/ref_path_callbacks.h"
namespace {
struct LiBridge { Scene *s; uint32_t pix, samp, k0, k1; SobolSampler *sob = nullptr; };
int br_rayIntersect(void *u, const float o[3], const float d[3], float mint, float maxt, RefPathIts *out) {
    LiBridge *b = (LiBridge *) u;
    Ray r(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), mint, maxt);
    Intersection its;
    b->s->geo.rayIntersect(r, its);
    out->valid = its.valid ? 1 : 0;
    if (!its.valid) return 0;
    out->t = its.t;
    const V3 vs[6] = {its.p, its.geoFrame.n, its.shFrame.s, its.shFrame.t, its.shFrame.n, its.wi};
    float *dst[6] = {out->p, out->geoN, out->shS, out->shT, out->shN, out->wi};
    for (int k = 0; k < 6; ++k) { dst[k][0] = vs[k].x; dst[k][1] = vs[k].y; dst[k][2] = vs[k].z; }
    out->bsdf = b->s->geo.shapes[its.shape].bsdf;
    return 1;
}
void br_sampleEmitterDirect(void *u, const float ref[3], float sx, float sy, float value[3], float d[3], float *pdf) {
    LiBridge *b = (LiBridge *) u;
    value[0] = value[1] = value[2] = 0; d[0] = d[1] = d[2] = 0; *pdf = 0;
    if (!b->s->hasEnv) return;
    const V3 p(ref[0], ref[1], ref[2]);
    EnvMap::DirectSample ds = b->s->env.sampleDirect(p, sx, sy);          // emitter->sampleDirect, then the visibility test of scene.cpp:838-849
    d[0] = ds.d.x; d[1] = ds.d.y; d[2] = ds.d.z; *pdf = ds.pdf;
    if (ds.pdf != 0) {
        Ray shadow(p, ds.d, kEpsilon, ds.dist * (1 - kShadowEpsilon));
        if (!b->s->geo.rayOccluded(shadow)) { value[0] = ds.value.x; value[1] = ds.value.y; value[2] = ds.value.z; }
    }
}
void br_evalEnvironment(void *u, const float d[3], int hasDiff, const float rx[3], const float ry[3], float value[3]) {
    LiBridge *b = (LiBridge *) u;
    const V3 v = b->s->env.evalEnvironment(V3(d[0], d[1], d[2]), hasDiff != 0, V3(rx[0], rx[1], rx[2]), V3(ry[0], ry[1], ry[2]));
    value[0] = v.x; value[1] = v.y; value[2] = v.z;
}
int br_hasEnvironment(void *u) { return ((LiBridge *) u)->s->hasEnv ? 1 : 0; }
int br_fillDirect(void *u, const float o[3], const float d[3]) { return ((LiBridge *) u)->s->env.fillDirectSamplingRecord(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2])) ? 1 : 0; }
float br_pdfEmitterDirect(void *u, const float d[3]) { return ((LiBridge *) u)->s->env.pdfDirect(V3(d[0], d[1], d[2])); }
unsigned br_bsdfType(void *u, int id) {
    const BSDFAny &m = ((LiBridge *) u)->s->bsdfs[id];
    return m.kind == 5 ? (ENull | EDeltaReflection) : m.kind == 6 ? (ENull | EDeltaReflection | EDiffuseReflection) : (EDiffuseReflection | EGlossyReflection);
}
void br_bsdfEval(void *u, int id, const float wi[3], const float wo[3], float out[3]) {
    const V3 v = ((LiBridge *) u)->s->bsdfs[id].eval(V3(wi[0], wi[1], wi[2]), V3(wo[0], wo[1], wo[2]));
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
float br_bsdfPdf(void *u, int id, const float wi[3], const float wo[3]) { return ((LiBridge *) u)->s->bsdfs[id].pdf(V3(wi[0], wi[1], wi[2]), V3(wo[0], wo[1], wo[2])); }
void br_bsdfSample(void *u, int id, int depth, const float wi[3], float sx, float sy, float wo[3], float weight[3], float *pdf, unsigned *type, float *eta) {
    LiBridge *b = (LiBridge *) u;
    const BSDFAny &m = b->s->bsdfs[id];
    float extra[4] = {0, 0, 0, 0};
    if (m.drawsExtra()) { Philox4 ue = philox4x32_10(b->pix, b->samp, (uint32_t) depth, 2, b->k0, b->k1); for (int k = 0; k < 4; ++k) extra[k] = u32_to_unit(ue.v[k]);
                          if (b->sob) { b->sob->next2D(extra[0], extra[1]); b->sob->next2D(extra[2], extra[3]); } }
    const BSDFSample bs = m.sample(V3(wi[0], wi[1], wi[2]), sx, sy, extra);
    wo[0] = bs.wo.x; wo[1] = bs.wo.y; wo[2] = bs.wo.z; weight[0] = bs.weight.x; weight[1] = bs.weight.y; weight[2] = bs.weight.z;
    *pdf = bs.pdf; *type = (unsigned) bs.sampledType; *eta = bs.eta;
}
void br_next2D(void *u, int depth, int which, float out[2]) {
    LiBridge *b = (LiBridge *) u;
    if (b->sob) { b->sob->next2D(out[0], out[1]); return; }        // the stateful sampler: whatever order path.cpp asks in
    const Philox4 v = philox4x32_10(b->pix, b->samp, (uint32_t) depth, 0, b->k0, b->k1);
    out[0] = u32_to_unit(v.v[which ? 2 : 0]); out[1] = u32_to_unit(v.v[which ? 3 : 1]);
}
float br_next1D(void *u, int depth) { LiBridge *b = (LiBridge *) u; if (b->sob) return b->sob->next1D(); return u32_to_unit(philox4x32_10(b->pix, b->samp, (uint32_t) depth, 1, b->k0, b->k1).v[0]); }
}

int orc_render_samples_ref_li(void *sp, const char *refPathLib, uint64_t n, const uint32_t *xy, const uint32_t *samp, uint32_t spp, uint64_t seed,
                              float *outLi, float *outAlpha, int32_t *outDepth) {
    ORC_TRY
    Scene *s = (Scene *) sp;
    s->seed = seed;
    void *lib = dlopen(refPathLib, RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("cannot load ") + refPathLib + ": " + dlerror());
    typedef int (*LiFn)(const RefPathCallbacks *, int, int, int, int, int, const float *, const float *, float, float, const float *, const float *, float *, float *);
    LiFn li = (LiFn) dlsym(lib, "ref_path_li");
    if (!li) throw std::runtime_error("ref_path_li not found");
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    for (uint64_t i = 0; i < n; ++i) {
        LiBridge b{s, xy[2 * i + 1] * (uint32_t) s->cam.filmW + xy[2 * i], samp[i], k0, k1};
        RefPathCallbacks cb{&b, br_rayIntersect, br_sampleEmitterDirect, br_evalEnvironment, br_hasEnvironment, br_fillDirect, br_pdfEmitterDirect,
                            br_bsdfType, br_bsdfEval, br_bsdfPdf, br_bsdfSample, br_next2D, br_next1D};
        // the camera ray exactly as Scene::renderSample() makes it (integrator.cpp:140-188)
        const Philox4 u = philox4x32_10(b.pix, b.samp, 0, 0, k0, k1);
        float px = (float) xy[2 * i] + u32_to_unit(u.v[0]), py = (float) xy[2 * i + 1] + u32_to_unit(u.v[1]);
        SobolSampler sob;
        if (s->samplerKind == 1) {       // the reference's sampler object, asked by path.cpp itself from here on
            sob = s->makeSobol(); sob.px = (int) xy[2 * i]; sob.py = (int) xy[2 * i + 1]; sob.setSampleIndex(samp[i]);
            float a0, a1; sob.next2D(a0, a1);
            px = (float) (int) xy[2 * i] + a0; py = (float) (int) xy[2 * i + 1] + a1;
            b.sob = &sob;
        }
        Ray ray; V3 rx, ry;
        s->cam.sampleRayDifferential(px, py, ray, rx, ry);
        const float ds = 1.0f / std::sqrt((float) spp);
        rx = ray.d + (rx - ray.d) * ds; ry = ray.d + (ry - ray.d) * ds;
        float L[3], alpha = 0;
        const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z}, rxa[3] = {rx.x, rx.y, rx.z}, rya[3] = {ry.x, ry.y, ry.z};
        outDepth[i] = li(&cb, s->maxDepth, s->rrDepth, s->strictNormals ? 1 : 0, s->hideEmitters ? 1 : 0, s->filmHasAlpha ? 1 : 0, o, d, ray.mint, ray.maxt, rxa, rya, L, &alpha);
        outLi[3 * i] = L[0]; outLi[3 * i + 1] = L[1]; outLi[3 * i + 2] = L[2]; outAlpha[i] = alpha;
    }
    return 0;
    ORC_CATCH
}

// Film splat of explicit samples (F1 parity hook): pos(2), value(3), alpha per sample -> film
int orc_splat_batch(void *sp, uint64_t n, const float *pos, const float *rgb, const float *alpha, float *outFilm) {
    Scene *s = (Scene *) sp;
    Film film; film.init(s->cam.filmW, s->cam.filmH);
    for (uint64_t i = 0; i < n; ++i) film.put(s->filter, pos[2 * i], pos[2 * i + 1], V3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), alpha[i]);
    std::memcpy(outFilm, film.data.data(), film.data.size() * 4);
    return 0;
}

// sunsky bake through the compiled reference pieces (oracle/_ref/libref_pieces.so)
// computeSunRadiance (sunmodel.h:316-371) of the oracle alone: elevation angle from the zenith, linear RGB out
int orc_sun_radiance(const char *refLib, float theta, float turbidity, float *outRGB) {
    ORC_TRY
    void *h = dlopen(refLib, RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("oracle/_ref not built: ") + dlerror());
    RefPieces ref;
    auto cie = (int (*)(const float **, const float **, const float **, const float **)) dlsym(h, "ref_cie_tables");
    if (!cie) throw std::runtime_error("oracle/_ref library lacks expected symbols");
    ref.cie_n = cie(&ref.cie_wl, &ref.cie_x, &ref.cie_y, &ref.cie_z);
    const V3 r = computeSunRadiance(theta, turbidity, ref);
    outRGB[0] = r.x; outRGB[1] = r.y; outRGB[2] = r.z;
    return 0;
    ORC_CATCH
}
int orc_bake_sunsky(const char *refLib, float turbidity, float albedo, const float *sunDir, float skyScale, float sunScale,
                    float sunRadiusScale, int resolution, float *outRGB) {
    ORC_TRY
    void *h = dlopen(refLib, RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("oracle/_ref not built: ") + dlerror());
    RefPieces ref;
    ref.sky_alloc = (void *(*)(double, double, double)) dlsym(h, "ref_sky_alloc");
    ref.sky_radiance = (double (*)(void *, double, double, int)) dlsym(h, "ref_sky_radiance");
    ref.sky_free = (void (*)(void *)) dlsym(h, "ref_sky_free");
    auto cie = (int (*)(const float **, const float **, const float **, const float **)) dlsym(h, "ref_cie_tables");
    if (!ref.sky_alloc || !ref.sky_radiance || !ref.sky_free || !cie) throw std::runtime_error("oracle/_ref library lacks expected symbols");
    ref.cie_n = cie(&ref.cie_wl, &ref.cie_x, &ref.cie_y, &ref.cie_z);
    SunSkyParams P; P.turbidity = turbidity; P.albedo = albedo; P.sunDirection = V3(sunDir[0], sunDir[1], sunDir[2]);
    P.skyScale = skyScale; P.sunScale = sunScale; P.sunRadiusScale = sunRadiusScale; P.resolution = resolution;
    std::vector<float> rgb; int W, H;
    bakeSunSky(P, ref, rgb, W, H);
    std::memcpy(outRGB, rgb.data(), rgb.size() * 4);
    return 0;
    ORC_CATCH
}

// Small exported helpers used by the pinning tests
void orc_gauss_legendre_140(float *points, float *weights) {
    GaussLegendre<140> g; std::memcpy(points, g.points, 140 * 4); std::memcpy(weights, g.weights, 140 * 4);
}
void orc_interp_dist(const float *weights, int size, int num, float *outPdfs, float *outCdfs, float *outSums) {
    InterpolatedDistribution1D d; d.init(std::vector<float>(weights, weights + size * num), size, num);
    std::memcpy(outPdfs, d.pdfs.data(), d.pdfs.size() * 4); std::memcpy(outCdfs, d.cdfs.data(), d.cdfs.size() * 4); std::memcpy(outSums, d.sums.data(), d.sums.size() * 4);
}
void orc_interp_dist_warp(const float *weights, int size, int num, int n, const float *distribution, const float *u, float *outU, int *outX) {
    InterpolatedDistribution1D d; d.init(std::vector<float>(weights, weights + size * num), size, num);
    for (int i = 0; i < n; ++i) { float uu = u[i]; int x; d.warp(distribution[i], uu, x); outU[i] = uu; outX[i] = x; }
}
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out4) {
    Philox4 p = philox4x32_10(c0, c1, c2, c3, k0, k1); std::memcpy(out4, p.v, 16);
}
void orc_half_roundtrip(const float *in, int n, uint16_t *outHalf, float *outFloat) {
    for (int i = 0; i < n; ++i) { outHalf[i] = float_to_half(in[i]); outFloat[i] = half_to_float(outHalf[i]); }
}

} // extern "C"

#ifdef ORC_FAST_STATS
extern "C" void orc_fast_stats(uint64_t *out5) {
    out5[0] = orc::Geometry::stRays.exchange(0); out5[1] = orc::Geometry::stNodes.exchange(0); out5[2] = orc::Geometry::stPre.exchange(0);
    out5[3] = orc::Geometry::stExact.exchange(0); out5[4] = orc::Geometry::stLeafPop.exchange(0);
}
// tree metrics: out[0] wide nodes, out[1] refs, out[2] sum area(inner child boxes)/area(root), out[3] sum area(ref boxes)/area(root), out[4] max depth, out[5] mean ref depth, out[6] mean fill
extern "C" void orc_fast_tree_metrics(void *sp, double *out) {
    using namespace orc;
    Scene *s = (Scene *) sp; const auto &W = s->geo.wide;
    auto area = [&](const Geometry::Wide &w, int k) { double ex = w.hi[0][k] - w.lo[0][k], ey = w.hi[1][k] - w.lo[1][k], ez = w.hi[2][k] - w.lo[2][k]; return ex * ey + ey * ez + ez * ex; };
    double rootA = 0; { double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30}; for (int k = 0; k < 8; ++k) if (W[0].child[k] != 0xffffffffu) for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], (double) W[0].lo[c][k]); hi[c] = std::max(hi[c], (double) W[0].hi[c][k]); }
      double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2]; rootA = ex * ey + ey * ez + ez * ex; }
    double sumInner = 0, sumRef = 0, depthSum = 0, fill = 0; uint64_t refs = 0; int maxDepth = 0;
    std::vector<std::pair<uint32_t, int>> st; st.push_back({0u, 1});
    while (!st.empty()) { auto [i, dpt] = st.back(); st.pop_back(); maxDepth = std::max(maxDepth, dpt);
        for (int k = 0; k < 8; ++k) { uint32_t c = W[i].child[k]; if (c == 0xffffffffu) continue; fill += 1;
            if (c & 0x80000000u) { sumRef += area(W[i], k) / rootA; refs++; depthSum += dpt; } else { sumInner += area(W[i], k) / rootA; st.push_back({c, dpt + 1}); } } }
    out[0] = (double) W.size(); out[1] = (double) refs; out[2] = sumInner; out[3] = sumRef; out[4] = maxDepth; out[5] = depthSum / refs; out[6] = fill / W.size();
}
#endif

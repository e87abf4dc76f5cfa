// cp_shade.cu -- shading stages of the wavefront path tracer (sm_100a), compiled with -fmad=false.
//
// Replaces (reference file:line):
//   SamplingIntegrator::renderBlock      src/librender/integrator.cpp:140-188   (jitter, sensor ray, put)
//   PerspectiveCameraImpl::sampleRayDifferential  src/sensors/perspective.cpp:271-298
//   MIPathTracer::Li                     src/integrators/path/path.cpp:119-300  (NEE + MIS + BSDF sampling + RR)
//   Scene::sampleEmitterDirect           src/librender/scene.cpp:828-853        (shadow ray [Epsilon, dist(1-ShadowEpsilon)])
//   ImageBlock::put                      include/mitsuba/render/imageblock.h:144-186 (discretised filter splat)
// No fused multiply-adds here: the reference's x86 build has none, and the Marschner lobes amplify last-bit differences.
//
// This file is compiled twice (csrc/Makefile): as is -- the strict build, every elementary function correctly rounded, bit-identical
// to the oracle -- and with -DCP_FAST_MATH, where the calls that are not amplified use the fp32 CUDA functions (cp_common.cuh) and
// only the shading kernel is emitted, under the name launch_shade_fast.  A context picks one at run time (cudapath_set_math_mode).
#include "cp_host.h"
#include "cp_env.cuh"
#include "cp_camera.cuh"
#include "cp_wavefront.h"
#ifdef CP_FAST_MATH
#define k_shade k_shade_fast
#define launch_shade launch_shade_fast
#endif

namespace cp {

// samplePos of renderBlock (integrator.cpp:171): pixel + the sampler's first 2-D number.  Philox: counter stream 0 of vertex 0; sobol: dimensions
// 0 and 1, scaled to the pixel the enumerated index belongs to (sobol.cpp:233-239)
__device__ __forceinline__ void pixel_sample(const SceneDev &S, const WaveParams &wp, uint32_t x, uint32_t y, uint32_t samp, float &px, float &py) {
    if (S.sobol.kind == 1) {
        const uint64_t idx = sobol_index(S.sobol, samp, x, y);
        float a = sobol_sample(S.sobol, idx, 0u), b = sobol_sample(S.sobol, idx, 1u);
        if (idx != (uint64_t) samp) { a = a * S.sobol.res - (float) (int) x; b = b * S.sobol.res - (float) (int) y; }
        px = (float) (int) x + a; py = (float) (int) y + b;
        return;
    }
    const Philox4 u = philox4x32_10(y * wp.filmW + x, samp, 0u, 0u, wp.seedLo, wp.seedHi);
    px = (float) x + u32_to_unit(u.v[0]); py = (float) y + u32_to_unit(u.v[1]);
}
#define CP_DIM_SHIFT 20       // flags word, bits 20..31: the next Sobol dimension of the path (sampler-faithful mode)

#ifndef CP_FAST_MATH

__global__ void __launch_bounds__(256) k_raygen(SceneDev S, WaveParams wp, PathQueue q, float4 *liAcc, uint32_t n, uint32_t *initSlot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) { initSlot[0] = n; initSlot[1] = 0u; }      // queue lengths the first trace launch reads (closest-hit rays, shadow rays)
    uint32_t x, y, samp;
    const bool valid = path_to_pixel(wp, wp.waveBase + i, x, y, samp);
    liAcc[i] = make_float4(0, 0, 0, 1.0f);
    if (!valid) {
        q.ro[i] = make_float4(0, 0, 0, 1.0f); q.rd[i] = make_float4(0, 0, 1, 0.0f);
        q.thr[i] = make_float4(0, 0, 0, 0); q.id[i] = make_uint2(i, F_INVALID | F_FIRST | 1u);
        return;
    }
    float px, py;
    pixel_sample(S, wp, x, y, samp, px, py);
    const CameraRay r = camera_ray(S.cam, px, py, wp.diffScale);
    q.ro[i] = make_float4(r.o.x, r.o.y, r.o.z, r.mint);
    q.rd[i] = make_float4(r.d.x, r.d.y, r.d.z, r.maxt);
    q.thr[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    q.id[i] = make_uint2(i, F_FIRST | 1u | (2u << CP_DIM_SHIFT));                 // depth starts at 1 (integrator.h:218-224); two sampler dimensions are spent
}

#endif // !CP_FAST_MATH

__device__ __forceinline__ float mi_weight(float pdfA, float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); } // path.cpp:296-300

#ifndef CP_SHADE_MIN_BLOCKS
#ifdef CP_FAST_MATH
#define CP_SHADE_MIN_BLOCKS 5      // 96 registers, 64 B of spills: 280 / 280 vs 274 / 257 Mpaths/s at 4 CTAs (120 registers) and 277 / 273 at 6, two interleaved rounds (round 2, g12)
#else
#define CP_SHADE_MIN_BLOCKS 4      // the strict mode (fp64-evaluated elementary functions) needs its 122 registers
#endif
#endif
__global__ void __launch_bounds__(128, CP_SHADE_MIN_BLOCKS) k_shade(SceneDev S, WaveParams wp, PathQueue in, const uint32_t *__restrict__ nPtr, const float4 *__restrict__ hitPT,
                                               const uint32_t *__restrict__ hitPrim, PathQueue out, ShadowQueue sq, float4 *liAcc,
                                               uint32_t *counters, unsigned long long *unsupportedLookups) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *nPtr;                   // queue length left by the previous bounce; the grid covers the host's upper bound
    if (blockIdx.x * blockDim.x >= n) return;
    bool survive = false, wantShadow = false, countOnlyShadow = false;
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
            const bool sobol = S.sobol.kind == 1;
            uint32_t dim = flags >> CP_DIM_SHIFT; bool dimOverflow = false;
            const uint64_t sobolIdx = sobol ? sobol_index(S.sobol, samp, x, y) : 0ull;

            if (gv == 0xffffffffu) {
                // ---- the ray escaped
                float4 acc = liAcc[pathId];
                if (first) {
                    acc.w = S.film.hasAlpha ? 0.0f : 1.0f;                      // records.inl:117-135
                    if (S.env.present && !S.integ.hideEmitters) {               // path.cpp:136-143
                        float spx, spy;
                        pixel_sample(S, wp, x, y, samp, spx, spy);
                        const CameraRay cr = camera_ray(S.cam, spx, spy, wp.diffScale);
                        V3 L = thr * env_eval_filtered(S.env, rayD, cr.rx, cr.ry, unsupportedLookups);
                        acc.x += L.x; acc.y += L.y; acc.z += L.z;
                    }
                } else if (S.env.present && !(S.integ.hideEmitters && !(flags & F_SCATTERED))) {   // path.cpp:233-264 (:237-238 hides the emitter from unscattered paths)
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
                    float rr;
                    if (sobol) { rr = sobol_next1D(S.sobol, sobolIdx, dim, dimOverflow); if (dimOverflow) { *S.sobol.err = 2; break; } }
                    else rr = u32_to_unit(philox4x32_10(pix, samp, (uint32_t) depth, 1u, wp.seedLo, wp.seedHi).v[0]);
                    if (rr >= qv) break;
                    thr = thr / qv;
                }
                depth++;
                if (!(depth <= S.integ.maxDepth || S.integ.maxDepth < 0)) break;  // loop condition path.cpp:135
            }
            // ---- intersection record
            const float4 hp = hitPT[i];
            HitRecord rec;
            uint32_t shapeIdx;
            if (gv & CP_RECT_FLAG) shapeIdx = fill_intersection_rect(S.mesh, gv & CP_PRIM_MASK, hp.x, hp.y, hp.w, rayO, rayD, rec);
            else if (gv & CP_TRI_FLAG) shapeIdx = fill_intersection_mesh(S.mesh, S.shapes, gv & CP_PRIM_MASK, hp.x, hp.y, rayD, rec);
            else {
                const float4 v1 = __ldg(S.vtx + gv), v2 = __ldg(S.vtx + gv + 1);
                shapeIdx = vtx_shape(v1);
                fill_intersection(v1, v2, S.shapes[shapeIdx].radius, V3(hp.x, hp.y, hp.z), rayD, rec);
            }
            const ShapeDev &shape = S.shapes[shapeIdx];
            if ((depth >= S.integ.maxDepth && S.integ.maxDepth > 0) ||
                (S.integ.strictNormals && dot(rayD, rec.geoN) * rec.wi.z >= 0)) break;   // path.cpp:156-165
            const BsdfDev &bsdf = S.bsdfs[shape.bsdf];
            // the random numbers of this vertex: Philox counter stream 0 = (emitter.x, emitter.y, bsdf.x, bsdf.y); sobol: drawn in the order path.cpp asks
            float e0, e1, b0, b1;
            if (!sobol) {
                const Philox4 u = philox4x32_10(pix, samp, (uint32_t) depth, 0u, wp.seedLo, wp.seedHi);
                e0 = u32_to_unit(u.v[0]); e1 = u32_to_unit(u.v[1]); b0 = u32_to_unit(u.v[2]); b1 = u32_to_unit(u.v[3]);
            } else {
                e0 = e1 = 0.0f;
                if (bsdf_has_smooth(bsdf)) sobol_next2D(S.sobol, sobolIdx, dim, e0, e1, dimOverflow);     // path.cpp:179: drawn whether or not the scene has an emitter
                sobol_next2D(S.sobol, sobolIdx, dim, b0, b1, dimOverflow);                                                // path.cpp:210
                if (dimOverflow) { *S.sobol.err = 2; break; }
            }
            // ---- emitter sampling (path.cpp:174-200)
            if (S.env.present && bsdf_has_smooth(bsdf)) {       // path.cpp:174-175: only BSDFs with a smooth component
                const EnvSample es = env_sample_direct(S.env, rec.p, e0, e1);
                if (es.pdf != 0) {
                    // The shadow ray is always traced when pdf != 0 (scene.cpp:838-845); its contribution may be zero.
                    V3 contrib(0.0f);
                    if (!isZero(es.value) && !bsdf_eval_is_zero(bsdf)) {
                        const V3 wo = rec.sh.toLocal(es.d);
                        const V3 bsdfVal = bsdf_eval(bsdf, rec.wi, wo, false, rec.u, rec.v);
                        if (!isZero(bsdfVal) && (!S.integ.strictNormals || dot(rec.geoN, es.d) * wo.z > 0)) {
                            const float bsdfPdf = bsdf_pdf(bsdf, rec.wi, wo);
                            contrib = thr * es.value * bsdfVal * mi_weight(es.pdf, bsdfPdf);
                        }
                    }
                    // A shadow ray whose sample cannot contribute (BSDF value zero, wrong side) cannot change the image: it is counted
                    // like the reference counts it, but not traced.
                    if (isZero(contrib)) { countOnlyShadow = true; }
                    wantShadow = !countOnlyShadow;
                    so = make_float4(rec.p.x, rec.p.y, rec.p.z, kEpsilon);
                    sd = make_float4(es.d.x, es.d.y, es.d.z, es.dist * (1 - kShadowEpsilon));
                    sc = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(pathId));
                }
            }
            // ---- BSDF sampling (path.cpp:207-226)
            float4 extra = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (bsdf_draws_extra(bsdf)) {          // counter stream 2 of this vertex (0: emitter + BSDF sample, 1: roulette)
                if (sobol) {                       // bRec.sampler->next2D() twice inside sample() (marschner.cpp:473-474)
                    sobol_next2D(S.sobol, sobolIdx, dim, extra.x, extra.y, dimOverflow); sobol_next2D(S.sobol, sobolIdx, dim, extra.z, extra.w, dimOverflow);
                    if (dimOverflow) { *S.sobol.err = 2; break; }
                } else {
                    const Philox4 ue = philox4x32_10(pix, samp, (uint32_t) depth, 2u, wp.seedLo, wp.seedHi);
                    extra = make_float4(u32_to_unit(ue.v[0]), u32_to_unit(ue.v[1]), u32_to_unit(ue.v[2]), u32_to_unit(ue.v[3]));
                }
            }
            const BsdfSampleOut bs = bsdf_sample(bsdf, rec.wi, b0, b1, extra, rec.u, rec.v);
            if (isZero(bs.weight)) break;
            const V3 wo = rec.sh.toWorld(bs.wo);
            if (S.integ.strictNormals && dot(rec.geoN, wo) * bs.wo.z <= 0) break;
            thr = thr * bs.weight;
            nro = make_float4(rec.p.x, rec.p.y, rec.p.z, kEpsilon);
            nrd = make_float4(wo.x, wo.y, wo.z, CP_INF);
            nthr = make_float4(thr.x, thr.y, thr.z, bs.pdf);
            nid = make_uint2(pathId, (uint32_t) depth | ((bs.type & EDelta) ? F_DELTA : 0u) | ((bs.type != ENull || (flags & F_SCATTERED)) ? F_SCATTERED : 0u) | (dim << CP_DIM_SHIFT));
            survive = true;
        } while (false);
    }
    const uint32_t oi = warp_append(counters + 0, survive);
    if (survive) { out.ro[oi] = nro; out.rd[oi] = nrd; out.thr[oi] = nthr; out.id[oi] = nid; }
    const uint32_t si = warp_append(counters + 1, wantShadow);
    if (wantShadow) { sq.o[si] = so; sq.d[si] = sd; sq.c[si] = sc; }
    warp_append(counters + 4, countOnlyShadow);
}

#ifndef CP_FAST_MATH
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
    float px, py;
    pixel_sample(S, wp, x, y, samp, px, py);            // the sample position is recomputed, not stored
    const float4 acc = liAcc[i];
    film_put(S.film, film, (int) wp.filmW, (int) wp.filmH, px, py, V3(acc.x, acc.y, acc.z), acc.w, dropped);
}

// Parity hook for F1: splat explicit samples
__global__ void k_splat_batch(SceneDev S, const float *__restrict__ pos, const float *__restrict__ rgb, const float *__restrict__ alpha,
                              uint64_t n, float *film, int W, int H) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    film_put(S.film, film, W, H, pos[2 * i], pos[2 * i + 1], V3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), alpha[i], nullptr);
}


void launch_raygen(const SceneDev &S, const WaveParams &wp, PathQueue q, float4 *liAcc, uint32_t n, uint32_t *initSlot, cudaStream_t stream) {
    k_raygen<<<(n + 255) / 256, 256, 0, stream>>>(S, wp, q, liAcc, n, initSlot);
}
#endif // !CP_FAST_MATH
void launch_shade(const SceneDev &S, const WaveParams &wp, PathQueue in, const uint32_t *nPtr, uint32_t nUpper, const float4 *hitPT, const uint32_t *hitPrim, PathQueue out,
                  ShadowQueue sq, float4 *liAcc, uint32_t *counters, unsigned long long *unsupportedLookups, cudaStream_t stream) {
    if (nUpper == 0) return;
    k_shade<<<(nUpper + 127) / 128, 128, 0, stream>>>(S, wp, in, nPtr, hitPT, hitPrim, out, sq, liAcc, counters, unsupportedLookups);
}
#ifndef CP_FAST_MATH
void launch_splat(const SceneDev &S, const WaveParams &wp, const float4 *liAcc, uint32_t n, float *film, unsigned long long *dropped, cudaStream_t stream) {
    k_splat<<<(n + 255) / 256, 256, 0, stream>>>(S, wp, liAcc, n, film, dropped);
}

#define CKW(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)
bool splat_batch(const SceneDev &S, const float *d_pos, const float *d_rgb, const float *d_alpha, uint64_t n, float *d_film, cudaStream_t stream, std::string &err) {
    if (n == 0) return true;
    k_splat_batch<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(S, d_pos, d_rgb, d_alpha, n, d_film, S.cam.filmW, S.cam.filmH);
    CKW(cudaStreamSynchronize(stream));
    CKW(cudaGetLastError());
    return true;
}
#endif // !CP_FAST_MATH

} // namespace cp

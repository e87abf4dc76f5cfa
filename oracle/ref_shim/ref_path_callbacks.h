// oracle/ref_shim/ref_path_callbacks.h -- TEST INFRASTRUCTURE ONLY.
// Plain-C callback table through which the reference's path integrator (src/integrators/path/path.cpp, compiled unmodified into
// oracle/_ref/libref_path.so) reaches the scene components.  The oracle fills it with ITS components (oracle_api.cpp), so that the
// reference's Li() and the oracle's Li() run on identical geometry, BSDFs, emitter and random numbers: what is compared is the
// integrator logic itself (path.cpp:119-300).
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    int valid; float t; float p[3]; float geoN[3]; float shS[3], shT[3], shN[3]; float wi[3]; int bsdf;
} RefPathIts;
typedef struct {
    void *user;
    /* Scene::rayIntersect(ray, its) */
    int (*rayIntersect)(void *user, const float o[3], const float d[3], float mint, float maxt, RefPathIts *its);
    /* Scene::sampleEmitterDirect(dRec, sample) incl. the shadow ray (scene.cpp:828-853): value (0 when occluded / no sample), direction, pdf */
    void (*sampleEmitterDirect)(void *user, const float ref[3], float sx, float sy, float value[3], float d[3], float *pdf);
    /* Scene::evalEnvironment(ray): camera rays carry differentials */
    void (*evalEnvironment)(void *user, const float d[3], int hasDifferentials, const float rxDir[3], const float ryDir[3], float value[3]);
    int (*hasEnvironment)(void *user);
    int (*fillDirectSamplingRecord)(void *user, const float o[3], const float d[3]);
    float (*pdfEmitterDirect)(void *user, const float d[3]);
    /* BSDF of an intersection */
    unsigned (*bsdfType)(void *user, int bsdf);
    void (*bsdfEval)(void *user, int bsdf, const float wi[3], const float wo[3], float out[3]);
    float (*bsdfPdf)(void *user, int bsdf, const float wi[3], const float wo[3]);
    void (*bsdfSample)(void *user, int bsdf, int depth, const float wi[3], float sx, float sy, float wo[3], float weight[3], float *pdf, unsigned *sampledType, float *eta);
    /* RadianceQueryRecord::nextSample2D / nextSample1D: which = 0 emitter sample, 1 BSDF sample (2-D); the 1-D draw is the roulette of `depth` */
    void (*next2D)(void *user, int depth, int which, float out[2]);
    float (*next1D)(void *user, int depth);
} RefPathCallbacks;
#ifdef __cplusplus
}
#endif

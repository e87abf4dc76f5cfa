// cp_scene.cuh -- device-resident scene description shared by all kernels (sm_100a).
#pragma once
#include "cp_common.cuh"
#include "cp_hair.cuh"
#include "cp_tri.cuh"
#include "cp_bsdf.cuh"

namespace cp {

// 4-wide BVH node, 128 B = one L2 line, read as 8 x LDG.128.
//   lo[0..2] = child min x/y/z (4 lanes each), hi[0..2] = child max x/y/z
//   child[i] >= 0  : index of an inner node
//   child[i] <  0  : leaf, ~child = (firstPrim << 3) | (primCount - 1), primCount in [1,8]
//   unused slots have an inverted box (+inf, -inf) and are never entered.
struct __align__(128) BVH4Node {
    float4 lo[3];
    float4 hi[3];
    int4 child;
    int4 pad;
};
static_assert(sizeof(BVH4Node) == 128, "BVH4Node must be one 128-byte line");

struct BVHDev {
    const BVH4Node *nodes;
    const uint32_t *prims;      // sorted reference list: global first-vertex index gv of each reference's segment
    const float4 *leafSeg;      // per reference, in leaf order: (p1.xyz, flags) (p2.xyz, gv bits) -- what the fp32 pre-test reads,
                                // so a leaf visit is one dependent load (32 contiguous bytes per reference) instead of index -> vertex.
                                // Triangle references carry flags bit 3 and CP_TRI_FLAG | triangle index in the id slot.
    uint32_t nodeCount, primCount;
};

// MIP levels above level 0 (TMIPMap::m_pyramid, m_sizeRatio, m_weightLut: include/mitsuba/render/mipmap.h:245-302)
#define CP_ENV_MAX_LEVELS 18
struct EnvMipInfo {
    int levels;                                  // including level 0
    int w[CP_ENV_MAX_LEVELS], h[CP_ENV_MAX_LEVELS];
    uint32_t offset[CP_ENV_MAX_LEVELS];          // level l >= 1 starts at mipTexels + offset[l]
    float ratioX[CP_ENV_MAX_LEVELS], ratioY[CP_ENV_MAX_LEVELS];
    float lut[64];                               // EWA Gaussian weights
};

struct EnvDev {
    int w, h;
    const float4 *texels;       // half-quantised RGB stored as exact fp32 (envmap.cpp:102-103)
    const float *cdfCols;       // (w+1) x h
    const float *cdfRows;       // h+1
    const float *rowWeights;    // h
    const float4 *mipTexels;    // levels 1.. of the Lanczos pyramid, half-quantised like level 0 (camera rays that miss: EWA lookups)
    const EnvMipInfo *mip;
    float normalization, scale;
    float pixelSizeX, pixelSizeY;
    float toWorld[9], toLocal[9]; // 3x3 linear parts (directions only)
    float bsCenter[3], bsRadius;  // scene bounding sphere x1.5 (envmap.cpp:331-341)
    int present;
};

struct CameraDev {
    float s2c[16];              // sampleToCamera (row-major 4x4)
    float toWorld[16];
    float dx[3], dy[3];         // near-plane position differentials (perspective.cpp:156-159)
    float nearClip, farClip;
    float invResX, invResY;
    int filmW, filmH;
};

struct FilmDev {
    float filterValues[32];     // discretised reconstruction filter (rfilter.cpp:37-55)
    float filterRadius, filterScale;
    int hasAlpha;
};

struct IntegratorDev { int maxDepth, rrDepth, strictNormals, hideEmitters; };

// Sampler-faithful mode: the reference's `sobol` sampler (src/samplers/sobol.cpp over src/samplers/sobolseq.h) instead of the counter-based
// Philox stream.  kind 0 = Philox (default), 1 = sobol.  The direction numbers (src/samplers/sobolseq.cpp: 1024 dimensions x 52 32-bit columns,
// and the enumeration matrices with their inverses per log2 resolution) live in HBM (234 KB, L2-resident).
#define CP_SOBOL_DIMS 1024u
#define CP_SOBOL_SIZE 52u
struct SobolDev {
    int kind;
    const uint32_t *m32; const uint64_t *vdc, *inv;
    uint32_t logRes, scramble;      // log2 of the film resolution rounded up to a power of two (sobol.cpp:146-156); low word of the TEA-hashed scramble
    float res;
    int *err;                       // set to 2 when a path asks for a dimension beyond the table (the plugin raises an error there, sobol.cpp:222-224)
};
// sobolseq.h:59-74 (sampleSingle)
CP_D float sobol_sample(const SobolDev &Q, uint64_t index, uint32_t dimension) {
    uint32_t result = Q.scramble;
    for (uint32_t i = dimension * CP_SOBOL_SIZE; index; index >>= 1, ++i) if (index & 1ull) result ^= __ldg(Q.m32 + i);
    return fminf(__uint2float_rn(result) * (1.0f / 4294967296.0f), 0.999999940395355225f);
}
// sobolseq.h:104-133 (look_up, SINGLE_PRECISION): index of sample `frame` of pixel (px, py) in the global sequence
CP_D uint64_t sobol_look_up(const SobolDev &Q, uint32_t frame, uint32_t px, uint32_t py) {
    const uint32_t m = Q.logRes, m2 = m << 1;
    uint64_t index = (uint64_t) frame << m2, delta = 0;
    for (uint32_t c = 0; frame; frame >>= 1, ++c) if (frame & 1u) delta ^= __ldg(Q.vdc + (size_t) (m - 1) * CP_SOBOL_SIZE + c);
    const uint64_t scr = (uint64_t) (Q.scramble >> (32u - m));
    uint64_t b = ((((uint64_t) px ^ scr) << m) | ((uint64_t) py ^ scr)) ^ delta;
    for (uint32_t c = 0; b; b >>= 1, ++c) if (b & 1ull) index ^= __ldg(Q.inv + (size_t) (m - 1) * CP_SOBOL_SIZE + c);
    return index;
}
// SobolSampler::setSampleIndex (sobol.cpp:206-217): the enumerated index when the film spans more than one pixel per axis
CP_D uint64_t sobol_index(const SobolDev &Q, uint32_t samp, uint32_t px, uint32_t py) { return Q.logRes > 1u ? sobol_look_up(Q, samp, px, py) : (uint64_t) samp; }
// next1D / next2D (sobol.cpp:219-245) on an explicit dimension counter; no sample arrays are requested on this path, so the reserved range is
// [5, 5) -- whose tests still move a 2-D request that would begin at dimension 4 to dimension 5
CP_D float sobol_next1D(const SobolDev &Q, uint64_t index, uint32_t &dim, bool &overflow) {
    if (dim >= CP_SOBOL_DIMS) { overflow = true; return 0.0f; }
    return sobol_sample(Q, index, dim++);
}
CP_D void sobol_next2D(const SobolDev &Q, uint64_t index, uint32_t &dim, float &a, float &b, bool &overflow) {
    if (dim + 1u >= 5u && dim < 5u) dim = 5u;
    if (dim + 1u >= CP_SOBOL_DIMS) { overflow = true; a = b = 0.0f; return; }
    a = sobol_sample(Q, index, dim++); b = sobol_sample(Q, index, dim++);
}

struct SceneDev {
    const float4 *vtx;          // see cp_hair.cuh
    uint32_t vtxCount;
    const ShapeDev *shapes;
    int shapeCount;
    const BsdfDev *bsdfs;
    int bsdfCount;
    BVHDev bvh;
    MeshDev mesh;               // triangle meshes (cp_tri.cuh); triCount == 0 for pure hair scenes
    int clipPerShape;           // hair primitives see the ray interval clipped to their own shape's bounds (hair.cpp:205-209):
                                // needed as soon as the scene bounds differ from a hair shape's bounds (several shapes, or meshes)
    float sceneMin[3], sceneMax[3];   // ShapeKDTree::m_aabb
    EnvDev env;
    CameraDev cam;
    FilmDev film;
    IntegratorDev integ;
    SobolDev sobol;
};

} // namespace cp

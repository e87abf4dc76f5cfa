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
};

} // namespace cp

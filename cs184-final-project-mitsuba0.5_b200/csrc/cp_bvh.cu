// cp_bvh.cu -- device-side BVH construction over hair segments and triangles (sm_100a).
//
// Replaces HairKDTree / ShapeKDTree construction (src/shapes/hair.cpp:108-159,
// include/mitsuba/render/gkdtree.h:958-2400, src/librender/skdtree.cpp:68-110) for this path.  The SAH
// kd-trees are NOT reproduced; the only build products the ray query's results depend on are kept
// bit-compatible with the reference: the segment list (hair.cpp:117-124), the per-shape bounds
// (union of getAABB(index), hair.cpp:368-397, gkdtree.h:997-1002) and the scene bounds.
//
// Pipeline (all on the device): segment compaction -> bounds + centroid box -> 48-bit Morton keys on a cubic grid ->
// radix sort (cub::DeviceRadixSort, build-time only) -> Karras 2012 binary radix tree -> bottom-up refit
// -> collapse into 128-byte 4-wide nodes with leaves of up to CP_LEAF_MAX consecutive sorted segments.
#include "cp_scene.cuh"
#include "cp_host.h"
#include <cub/cub.cuh>
#include <cstdlib>
#include <vector>

namespace cp {

#ifndef CP_LEAF_MAX
#define CP_LEAF_MAX 4      // references per leaf (<= 8: 3 bits of the leaf encoding); measured best of 2/3/4/6/8
#endif

__device__ __forceinline__ void atomicMinFloat(float *addr, float v) {
    if (v >= 0) atomicMin((int *) addr, __float_as_int(v)); else atomicMax((unsigned int *) addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxFloat(float *addr, float v) {
    if (v >= 0) atomicMax((int *) addr, __float_as_int(v)); else atomicMin((unsigned int *) addr, __float_as_uint(v));
}

// (x, y, z, bits) vertex stream of one shape + its sentinel (see cp_hair.cuh).  bits: 1 = starts a fiber, 2 = the miter joint
// at this vertex bends by at most 120 degrees (or the vertex is a fiber end), i.e. the miter plane overshoots the
// perpendicular end cut by at most r*tan(60 deg) < 2r -- lets the fp32 pre-test bound the segment's axial extent.
__global__ void k_pack_vertices(const float *__restrict__ xyz, const uint8_t *__restrict__ starts, uint32_t n, uint32_t shape, float4 *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const bool st = (i == n) || starts[i];
    uint32_t bits = (shape << 8) | (st ? 1u : 0u);
    if (i < n) {
        const V3 p(xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2]);
        const bool hasPrev = !st && i > 0, hasNext = (i + 1 < n) && !starts[i + 1];
        bool mild = true;
        if (hasPrev && hasNext) {
            const V3 a = normalize(p - V3(xyz[3 * (size_t) (i - 1)], xyz[3 * (size_t) (i - 1) + 1], xyz[3 * (size_t) (i - 1) + 2]));
            const V3 b = normalize(V3(xyz[3 * (size_t) (i + 1)], xyz[3 * (size_t) (i + 1) + 1], xyz[3 * (size_t) (i + 1) + 2]) - p);
            mild = dot(a, b) >= -0.45f;      // cos(120 deg) = -0.5, with slack for fp32 rounding
        }
        if (mild) bits |= 2u;
        out[i] = make_float4(p.x, p.y, p.z, __uint_as_float(bits));
    } else out[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(bits | 2u));
}
void pack_vertices(const float *d_xyz, const uint8_t *d_starts, uint32_t n, uint32_t shape, float4 *d_out, cudaStream_t stream) {
    k_pack_vertices<<<(n + 1 + 255) / 256, 256, 0, stream>>>(d_xyz, d_starts, n, shape, d_out);
}

struct IsSegmentStart {
    const float4 *vtx; uint32_t n;
    __device__ bool operator()(uint32_t i) const { return i + 1 < n && !(__float_as_uint(vtx[i + 1].w) & 1u); }
};

// Number of BVH references a segment is split into.  Hair segments are long and thin, so the box of a diagonal segment
// is almost empty; cutting the segment into pieces along its axis (each piece keeps the parent's primitive id) shrinks the
// summed box volume by ~1/k^2.  This only affects which boxes lead to a primitive test, never the test itself.
__device__ __forceinline__ int split_count(const float4 &v1, const float4 &v2, float radius, int maxSplit) {
    const V3 a = vtx_pos(v2) - vtx_pos(v1);
    const float ext = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fabsf(a.z));
    int k = (int) (ext / (8.0f * radius));
    return max(1, min(k, maxSplit));
}
__global__ void k_split_counts(const float4 *__restrict__ vtx, const uint32_t *__restrict__ segs, uint32_t nSeg, const ShapeDev *__restrict__ shapes,
                               int maxSplit, uint32_t *counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSeg) return;
    const uint32_t gv = segs[i];
    const float4 v1 = vtx[gv], v2 = vtx[gv + 1];
    counts[i] = (uint32_t) split_count(v1, v2, shapes[vtx_shape(v1)].radius, maxSplit);
}

// per segment: reference-tight bounds -> per-shape union (atomics); per reference: conservative BVH box; centroid box
__global__ void k_segment_bounds(const float4 *__restrict__ vtx, const uint32_t *__restrict__ segs, uint32_t nSeg, const uint32_t *__restrict__ refOffset,
                                 int maxSplit, ShapeDev *shapes, float *leafBox /*6*nRef*/, uint32_t *refPrim, float *centroidBox /*6*/) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < nSeg;
    if (!live) i = nSeg - 1;                              // keep the whole warp alive for the reductions below (results of the clone are discarded)
    const uint32_t gv = segs[i];
    const float4 v1 = vtx[gv], v2 = vtx[gv + 1], v0 = vtx[gv > 0 ? gv - 1 : 0], v3 = vtx[gv + 2];
    const uint32_t shapeIdx = vtx_shape(v1);
    ShapeDev &sd = shapes[shapeIdx];
    const float radius = sd.radius;
    float minA[3], maxA[3], minB[3], maxB[3];
    segment_end_boxes(v0, v1, v2, v3, radius, minA, maxA, minB, maxB);
    {   // per-shape union: min / max are exact, so reducing inside the warp first (when all lanes touch the same shape, the usual
        // case) gives the same bounds as one atomic per segment at 1/32 of the traffic to the same six words
        const bool uniform = __match_any_sync(0xffffffffu, shapeIdx) == 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = fminf(minA[k], minB[k]), hi = fmaxf(maxA[k], maxB[k]);
            if (uniform) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
                if ((threadIdx.x & 31u) == 0u) { atomicMinFloat(&sd.bmin[k], lo); atomicMaxFloat(&sd.bmax[k], hi); }
            } else { atomicMinFloat(&sd.bmin[k], lo); atomicMaxFloat(&sd.bmax[k], hi); }
        }
    }
    float cmin[3] = {CP_INF, CP_INF, CP_INF}, cmax[3] = {-CP_INF, -CP_INF, -CP_INF};      // centroid box of this thread's references
    const int nPieces = live ? split_count(v1, v2, radius, maxSplit) : 0;
    const uint32_t base = refOffset[i];
    const V3 p1 = vtx_pos(v1), a = vtx_pos(v2) - p1;
    const float invLen = 1.0f / length(a);
    // half-extent of a perpendicular cross-section disc along each coordinate axis
    const float ex = radius * sqrtf(fmaxf(0.0f, 1.0f - (a.x * invLen) * (a.x * invLen)));
    const float ey = radius * sqrtf(fmaxf(0.0f, 1.0f - (a.y * invLen) * (a.y * invLen)));
    const float ez = radius * sqrtf(fmaxf(0.0f, 1.0f - (a.z * invLen) * (a.z * invLen)));
    for (int j = 0; j < nPieces; ++j) {
        const float t0 = (float) j / nPieces, t1 = (float) (j + 1) / nPieces;
        const V3 q0 = p1 + a * t0, q1 = p1 + a * t1;
        float bmin[3] = {fminf(q0.x, q1.x) - ex, fminf(q0.y, q1.y) - ey, fminf(q0.z, q1.z) - ez};
        float bmax[3] = {fmaxf(q0.x, q1.x) + ex, fmaxf(q0.y, q1.y) + ey, fmaxf(q0.z, q1.z) + ez};
        if (j == 0) for (int k = 0; k < 3; ++k) { bmin[k] = fminf(bmin[k], minA[k]); bmax[k] = fmaxf(bmax[k], maxA[k]); }             // miter-cut end ellipses
        if (j == nPieces - 1) for (int k = 0; k < 3; ++k) { bmin[k] = fminf(bmin[k], minB[k]); bmax[k] = fmaxf(bmax[k], maxB[k]); }
        // The reference's bound uses radius*(1-Epsilon); widen so that every point the FP64 test can accept lies strictly
        // inside the box (radius*Epsilon for the shrink + rounding slack of the fp32 box arithmetic).
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pad = radius * 4e-4f + 4e-7f * fmaxf(fabsf(bmin[k]), fabsf(bmax[k]));
            bmin[k] -= pad; bmax[k] += pad;
            leafBox[6 * (size_t) (base + j) + k] = bmin[k]; leafBox[6 * (size_t) (base + j) + 3 + k] = bmax[k];
            const float c = 0.5f * (bmin[k] + bmax[k]);
            cmin[k] = fminf(cmin[k], c); cmax[k] = fmaxf(cmax[k], c);
        }
        refPrim[base + j] = gv;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo = cmin[k], hi = cmax[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
        if ((threadIdx.x & 31u) == 0u && lo <= hi) { atomicMinFloat(&centroidBox[k], lo); atomicMaxFloat(&centroidBox[3 + k], hi); }
    }
}

// per triangle: exact vertex box -> per-shape union (Triangle::getAABB, triangle.h:40-45); one padded BVH reference
__global__ void k_tri_bounds(MeshDev mesh, uint32_t refBase, ShapeDev *shapes, float *leafBox, uint32_t *refPrim, float *centroidBox) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mesh.triCount) return;
    const float4 p0 = mesh.pos[mesh.idx[3 * (size_t) j]], p1 = mesh.pos[mesh.idx[3 * (size_t) j + 1]], p2 = mesh.pos[mesh.idx[3 * (size_t) j + 2]];
    float bmin[3] = {fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z)};
    float bmax[3] = {fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z)};
    ShapeDev &sd = shapes[__float_as_uint(mesh.triAccel[3 * (size_t) j + 2].z)];
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMinFloat(&sd.bmin[k], bmin[k]); atomicMaxFloat(&sd.bmax[k], bmax[k]); }
    // A box only decides which tests run: pad it so that hits on the rim of a (possibly flat) box survive the fp32 slab test.
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float pad = 1e-5f * (bmax[k] - bmin[k]) + 2e-6f * fmaxf(fabsf(bmin[k]), fabsf(bmax[k])) + 1e-30f;
        bmin[k] -= pad; bmax[k] += pad;
        leafBox[6 * (size_t) (refBase + j) + k] = bmin[k]; leafBox[6 * (size_t) (refBase + j) + 3 + k] = bmax[k];
        const float c = 0.5f * (bmin[k] + bmax[k]);
        atomicMinFloat(&centroidBox[k], c); atomicMaxFloat(&centroidBox[3 + k], c);
    }
    refPrim[refBase + j] = CP_TRI_FLAG | j;
}

// per rectangle: the shape box (the four transformed corners, computed on the host like Rectangle::getAABB) and one padded BVH reference
__global__ void k_rect_bounds(MeshDev mesh, uint32_t refBase, ShapeDev *shapes, float *leafBox, uint32_t *refPrim, float *centroidBox) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mesh.rectCount) return;
    const float4 lo = mesh.rects[CP_RECT_STRIDE * (size_t) j + 6], hi = mesh.rects[CP_RECT_STRIDE * (size_t) j + 7];
    float bmin[3] = {lo.x, lo.y, lo.z}, bmax[3] = {hi.x, hi.y, hi.z};
    ShapeDev &sd = shapes[__float_as_uint(mesh.rects[CP_RECT_STRIDE * (size_t) j + 3].w)];
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMinFloat(&sd.bmin[k], bmin[k]); atomicMaxFloat(&sd.bmax[k], bmax[k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float pad = 1e-5f * (bmax[k] - bmin[k]) + 2e-6f * fmaxf(fabsf(bmin[k]), fabsf(bmax[k])) + 1e-30f;
        bmin[k] -= pad; bmax[k] += pad;
        leafBox[6 * (size_t) (refBase + j) + k] = bmin[k]; leafBox[6 * (size_t) (refBase + j) + 3 + k] = bmax[k];
        const float c = 0.5f * (bmin[k] + bmax[k]);
        atomicMinFloat(&centroidBox[k], c); atomicMaxFloat(&centroidBox[3 + k], c);
    }
    refPrim[refBase + j] = CP_TRI_FLAG | CP_RECT_FLAG | j;
}

// Grid bits per axis of the Morton key.  16 bits (cells of 1.5e-5 of the scene's largest extent) order hair references as well as 21 do and the
// radix sort of the 48-bit keys takes six 8-bit passes instead of eight.
#ifndef CP_MORTON_BITS
#define CP_MORTON_BITS 16
#endif
__device__ __forceinline__ uint64_t expandBits21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
// cubic: all three axes are quantised with the LARGEST extent of the centroid box, so that the grid cells are cubes whatever the
// aspect of the scene (a flat scene otherwise gets cells that are thin along its short axis, i.e. splits along it come too early)
__global__ void k_morton(const float *__restrict__ leafBox, uint32_t nSeg, const float *__restrict__ centroidBox, uint64_t *keys, uint32_t *ids, int cubic) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSeg) return;
    uint64_t code = 0;
    const float extMax = fmaxf(fmaxf(centroidBox[3] - centroidBox[0], centroidBox[4] - centroidBox[1]), centroidBox[5] - centroidBox[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float c = 0.5f * (leafBox[6 * (size_t) i + k] + leafBox[6 * (size_t) i + 3 + k]);
        float ext = cubic ? extMax : centroidBox[3 + k] - centroidBox[k];
        float f = ext > 0 ? (c - centroidBox[k]) / ext : 0.0f;
        uint64_t q = (uint64_t) fminf(fmaxf(f * (float) (1u << CP_MORTON_BITS), 0.0f), (float) ((1u << CP_MORTON_BITS) - 1u));
        code |= expandBits21(q) << (2 - k);
    }
    keys[i] = code; ids[i] = i;
}

// Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees"
__device__ __forceinline__ int delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long) (a ^ b));
}
// child encoding in the binary tree: >= 0 inner node, < 0 -> ~leaf index (sorted position)
__global__ void k_radix_tree(const uint64_t *__restrict__ keys, int n, int2 *children, int *parentInner, int *parentLeaf, int2 *ranges) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(lo, hi);
    if (left >= 0) parentInner[left] = i; else parentLeaf[~left] = i;
    if (right >= 0) parentInner[right] = i; else parentLeaf[~right] = i;
    if (i == 0) parentInner[0] = -1;
}

__global__ void k_gather_leaf_boxes(const float *__restrict__ leafBox, const uint32_t *__restrict__ ids, const uint32_t *__restrict__ refPrim,
                                    uint32_t n, float *sortedBox, uint32_t *sortedPrims) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = ids[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) sortedBox[6 * (size_t) i + k] = leafBox[6 * (size_t) src + k];
    sortedPrims[i] = refPrim[src];
}

// leaf records: flags = vertex bits of the first vertex (bit 0 starts fiber, bit 1 mild joint, bits 8.. shape) | bit 2 = mild joint at the second vertex
__global__ void k_leaf_records(const float4 *__restrict__ vtx, const uint32_t *__restrict__ sortedPrims, uint32_t n, float4 *leafSeg) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t gv = sortedPrims[i];
    if (gv & CP_TRI_FLAG) {   // triangle reference: bit 3 of the flags, id slot = CP_TRI_FLAG | triangle index (tested from MeshDev::triAccel)
        leafSeg[2 * (size_t) i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(8u));
        leafSeg[2 * (size_t) i + 1] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(gv));
        return;
    }
    const float4 v1 = vtx[gv], v2 = vtx[gv + 1];
    const uint32_t flags = __float_as_uint(v1.w) | ((__float_as_uint(v2.w) & 2u) ? 4u : 0u);
    leafSeg[2 * (size_t) i] = make_float4(v1.x, v1.y, v1.z, __uint_as_float(flags));
    leafSeg[2 * (size_t) i + 1] = make_float4(v2.x, v2.y, v2.z, __uint_as_float(gv));
}

__device__ __forceinline__ float boxArea(const float *b) {
    float ex = b[3] - b[0], ey = b[4] - b[1], ez = b[5] - b[2];
    return ex * ey + ey * ez + ez * ex;
}

// Bottom-up boxes of the binary tree.  The thread that completes a node holds its box and both child boxes, so it also decides what the collapse
// needs to know about the node -- whether a wide node may OPEN it -- and leaves the answer in bit 31 of ranges[node].y:
// a binary subtree with more than CP_LEAF_MAX references is always opened; a smaller one may become ONE leaf (one box, every reference pre-tested on
// entry) or stay a subtree (a box per child).  Surface-area heuristic, one level deep: it is opened when
//     splitCost * A(node) + A(left) n_left + A(right) n_right  <  A(node) n,
// i.e. when the children's boxes are so much smaller than their union that testing them first saves pre-tests (short fibers in open space); bundles of
// long, diagonal, pre-split fibers, whose pieces' boxes overlap anyway, stay leaves.  splitCost < 0: never.
#define CP_OPEN_FLAG 0x80000000u
__global__ void k_refit(const int2 *__restrict__ children, const int *__restrict__ parentInner, const int *__restrict__ parentLeaf,
                        const float *__restrict__ sortedBox, int n, float *innerBox, int *flags, int2 *ranges, float splitCost) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parentLeaf[i];
    while (node >= 0) {
        const int2 ch = children[node], r = ranges[node];  // read-only inputs of this node: in flight while the arrival counter is updated
        if (atomicAdd(&flags[node], 1) == 0) return;      // first arrival waits for the sibling
        __threadfence();
        float a[6], c[6], b[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            a[k] = ch.x >= 0 ? __ldcg(&innerBox[6 * (size_t) ch.x + k]) : sortedBox[6 * (size_t) (~ch.x) + k];
            c[k] = ch.y >= 0 ? __ldcg(&innerBox[6 * (size_t) ch.y + k]) : sortedBox[6 * (size_t) (~ch.y) + k];
            b[k] = k < 3 ? fminf(a[k], c[k]) : fmaxf(a[k], c[k]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcg(&innerBox[6 * (size_t) node + k], b[k]);
        {
            const int gamma = ch.x >= 0 ? ch.x : ~ch.x;             // the left child ends at the split position (k_radix_tree)
            const int cnt = r.y - r.x + 1, nL = gamma - r.x + 1, nR = r.y - gamma;
            bool open = cnt > CP_LEAF_MAX;
            if (!open && splitCost >= 0.0f) { const float aN = boxArea(b); open = splitCost * aN + boxArea(a) * nL + boxArea(c) * nR < aN * cnt; }
            if (open) ranges[node].y = (int) ((uint32_t) r.y | CP_OPEN_FLAG);
        }
        __threadfence();
        node = parentInner[node];
    }
}

struct CollapseItem { int bin; int wide; };

// One thread per wide node to emit.  A binary subtree that k_refit did not mark as open becomes a leaf reference; the candidates of a wide node are
// opened largest box first.  The child wide nodes of a warp's items are allocated with ONE atomic per counter (the two counters advance in lockstep).
__global__ void k_collapse(const CollapseItem *__restrict__ in, int nIn, CollapseItem *out, int *outCount, int *wideCount,
                           const int2 *__restrict__ children, const int2 *__restrict__ ranges,
                           const float *__restrict__ innerBox, const float *__restrict__ sortedBox, BVH4Node *nodes, int maxWide, int *err) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < nIn;                              // every lane stays for the warp scan below
    CollapseItem it{0, 0};
    int cand[4]; int2 rng[4]; float area[4]; int nc = 0;      // area < 0: not to be opened (a reference, or a subtree that stays one leaf)
    auto add = [&](int c) {
        cand[nc] = c; rng[nc] = make_int2(0, 0); area[nc] = -1.0f;
        if (c >= 0) {
            rng[nc] = ranges[c];
            if ((uint32_t) rng[nc].y & CP_OPEN_FLAG) area[nc] = boxArea(innerBox + 6 * (size_t) c);
        }
        ++nc;
    };
    int mine = 0;
    if (live) {
        it = in[idx];
        { const int2 ch = children[it.bin]; add(ch.x); add(ch.y); }
        while (nc < 4) {
            int best = -1; float bestArea = -1.0f;
            for (int k = 0; k < nc; ++k) if (area[k] > bestArea) { bestArea = area[k]; best = k; }
            if (best < 0) break;
            const int2 c2 = children[cand[best]];
            const int keep = nc; nc = best; add(c2.x); nc = keep; add(c2.y);
        }
        for (int k = 0; k < nc; ++k) if (area[k] >= 0.0f) ++mine;
    }
    __syncwarp();
    int wBase = 0, oBase = 0;
    {
        const unsigned lane = threadIdx.x & 31u;
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (unsigned) off) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 31u && total > 0) { wBase = atomicAdd(wideCount, total); oBase = atomicAdd(outCount, total); }
        wBase = __shfl_sync(0xffffffffu, wBase, 31) + incl - mine; oBase = __shfl_sync(0xffffffffu, oBase, 31) + incl - mine;
    }
    if (!live) return;
    float lo[3][4], hi[3][4]; int ref[4];
    for (int k = 0; k < 4; ++k) {
        if (k >= nc) {
            for (int a = 0; a < 3; ++a) { lo[a][k] = CP_INF; hi[a][k] = -CP_INF; }
            ref[k] = (int) 0x80000000; // empty slot (skipped by the traversal)
            continue;
        }
        int c = cand[k];
        const float *b = c >= 0 ? innerBox + 6 * (size_t) c : sortedBox + 6 * (size_t) (~c);
        for (int a = 0; a < 3; ++a) { lo[a][k] = b[a]; hi[a][k] = b[3 + a]; }
        if (c < 0) ref[k] = ~(int) ((((uint32_t) ~c) << 3) | 0u);
        else if (area[k] < 0.0f) {
            const int2 r = rng[k];
            ref[k] = ~(int) ((((uint32_t) r.x) << 3) | (uint32_t) (r.y - r.x));
        } else {
            const int w = wBase++, o = oBase++;
            if (w >= maxWide) { *err = 1; ref[k] = (int) 0x80000000; continue; }
            ref[k] = w;
            out[o].bin = c; out[o].wide = w;
        }
    }
    BVH4Node nd;
    for (int a = 0; a < 3; ++a) {
        nd.lo[a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        nd.hi[a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    nd.child = make_int4(ref[0], ref[1], ref[2], ref[3]);
    nd.pad = make_int4(0, 0, 0, 0);
    nodes[it.wide] = nd;
}

__global__ void k_single_leaf_root(const float *__restrict__ sortedBox, int n, BVH4Node *nodes) {
    // n <= CP_LEAF_MAX segments: one inner node with one leaf child covering all of them
    float b[6] = {CP_INF, CP_INF, CP_INF, -CP_INF, -CP_INF, -CP_INF};
    for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) b[k] = k < 3 ? fminf(b[k], sortedBox[6 * i + k]) : fmaxf(b[k], sortedBox[6 * i + k]);
    BVH4Node nd;
    for (int a = 0; a < 3; ++a) { nd.lo[a] = make_float4(b[a], CP_INF, CP_INF, CP_INF); nd.hi[a] = make_float4(b[3 + a], -CP_INF, -CP_INF, -CP_INF); }
    int leaf = ~(int) ((0u << 3) | (uint32_t) (n - 1));
    nd.child = make_int4(leaf, (int) 0x80000000, (int) 0x80000000, (int) 0x80000000);
    nd.pad = make_int4(0, 0, 0, 0);
    nodes[0] = nd;
}

__global__ void k_init_shape_bounds(ShapeDev *shapes, int n, float *centroidBox) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int k = 0; k < 3; ++k) { shapes[i].bmin[k] = CP_INF; shapes[i].bmax[k] = -CP_INF; }
    if (i == 0) for (int k = 0; k < 3; ++k) { centroidBox[k] = CP_INF; centroidBox[3 + k] = -CP_INF; }
}

namespace {
struct Scratch {   // scratch allocations from the caching allocator (cp_mem.cpp): returned to its free list, not to the driver
    std::vector<void *> ptrs; cudaStream_t stream;
    explicit Scratch(cudaStream_t s) : stream(s) {}
    ~Scratch() { cudaStreamSynchronize(stream); for (void *p : ptrs) if (p) dev_free(p); }
    template <typename T> cudaError_t alloc(T **p, size_t bytes) {
        cudaError_t e = dev_alloc((void **) p, bytes);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void release(void *p) { for (auto &q : ptrs) if (q == p) q = nullptr; }
};
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// Builds the BVH for the vertex array already resident on the device.  On success the caller owns
// out.nodes / out.prims (cudaFree).  `shapes` is updated in place with the per-shape bounds.
bool build_bvh(const float4 *d_vtx, uint32_t vtxCount, ShapeDev *d_shapes, int shapeCount, const MeshDev &mesh, int maxSplit, float leafSplitCost, cudaStream_t stream,
               BVHDev &out, BuildInfo &info, std::string &err) {
    out = BVHDev(); info = BuildInfo();
    Scratch S(stream);
    uint32_t *d_segs = nullptr, *d_ids = nullptr, *d_idsSorted = nullptr, *d_prims = nullptr;
    uint64_t *d_keys = nullptr, *d_keysSorted = nullptr;
    float *d_leafBox = nullptr, *d_sortedBox = nullptr, *d_innerBox = nullptr, *d_cbox = nullptr;
    int *d_num = nullptr, *d_parentInner = nullptr, *d_parentLeaf = nullptr, *d_flags = nullptr, *d_counters = nullptr;
    int2 *d_children = nullptr, *d_ranges = nullptr;
    CollapseItem *d_q0 = nullptr, *d_q1 = nullptr;
    BVH4Node *d_wide = nullptr, *d_final = nullptr;
    void *d_temp = nullptr; size_t need = 0;
    uint32_t nSeg = 0; int wideCount = 0;
    const int B = 256;
    cub::CountingInputIterator<uint32_t> counting(0);

    CK(S.alloc(&d_segs, sizeof(uint32_t) * (size_t) (vtxCount + 1)));
    CK(S.alloc(&d_num, sizeof(int) * 4));
    CK(S.alloc(&d_cbox, sizeof(float) * 6));
    k_init_shape_bounds<<<(shapeCount + B - 1) / B + 1, B, 0, stream>>>(d_shapes, shapeCount, d_cbox);
    {
        IsSegmentStart pred{d_vtx, vtxCount};
        CK(cub::DeviceSelect::If(nullptr, need, counting, d_segs, d_num, (int) vtxCount, pred, stream));
        CK(S.alloc(&d_temp, need));
        CK(cub::DeviceSelect::If(d_temp, need, counting, d_segs, d_num, (int) vtxCount, pred, stream));
        int h = 0;
        CK(cudaMemcpyAsync(&h, d_num, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        nSeg = (uint32_t) h;
    }
    info.segments = nSeg; info.triangles = mesh.triCount;
    if (nSeg == 0 && mesh.triCount == 0 && mesh.rectCount == 0) { err = "scene contains no hair segments, triangles or rectangles"; return false; }
    if (nSeg >= (1u << 28)) { err = "too many segments for the leaf encoding"; return false; }

    // references: every segment contributes split_count() boxes
    uint32_t *d_counts = nullptr, *d_offsets = nullptr, *d_refPrim = nullptr;
    uint32_t nRef = 0;
    {
        CK(S.alloc(&d_counts, sizeof(uint32_t) * (size_t) (nSeg + 1))); CK(S.alloc(&d_offsets, sizeof(uint32_t) * (size_t) (nSeg + 1)));
        CK(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * (size_t) (nSeg + 1), stream));
        if (nSeg) k_split_counts<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_vtx, d_segs, nSeg, d_shapes, maxSplit, d_counts);
        void *d_temp3 = nullptr;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, d_counts, d_offsets, (int) nSeg + 1, stream));
        CK(S.alloc(&d_temp3, need));
        CK(cub::DeviceScan::ExclusiveSum(d_temp3, need, d_counts, d_offsets, (int) nSeg + 1, stream));
        CK(cudaMemcpyAsync(&nRef, d_offsets + nSeg, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    const uint32_t nHairRef = nRef;
    nRef += mesh.triCount + mesh.rectCount;     // one reference per triangle / rectangle, after the hair references
    info.references = nRef;
    if ((uint64_t) nHairRef + mesh.triCount + mesh.rectCount >= (1u << 28)) { err = "too many BVH references for the leaf encoding"; return false; }
    CK(S.alloc(&d_leafBox, sizeof(float) * 6 * (size_t) nRef));
    CK(S.alloc(&d_refPrim, sizeof(uint32_t) * (size_t) nRef));
    if (nSeg) k_segment_bounds<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_vtx, d_segs, nSeg, d_offsets, maxSplit, d_shapes, d_leafBox, d_refPrim, d_cbox);
    if (mesh.triCount) k_tri_bounds<<<(mesh.triCount + B - 1) / B, B, 0, stream>>>(mesh, nHairRef, d_shapes, d_leafBox, d_refPrim, d_cbox);
    if (mesh.rectCount) k_rect_bounds<<<(mesh.rectCount + B - 1) / B, B, 0, stream>>>(mesh, nHairRef + mesh.triCount, d_shapes, d_leafBox, d_refPrim, d_cbox);
    nSeg = nRef;   // from here on the builder works on references
    CK(S.alloc(&d_keys, sizeof(uint64_t) * (size_t) nSeg)); CK(S.alloc(&d_keysSorted, sizeof(uint64_t) * (size_t) nSeg));
    CK(S.alloc(&d_ids, sizeof(uint32_t) * (size_t) nSeg)); CK(S.alloc(&d_idsSorted, sizeof(uint32_t) * (size_t) nSeg));
    k_morton<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_leafBox, nSeg, d_cbox, d_keys, d_ids, getenv("CUDAPATH_MORTON_CUBE") ? atoi(getenv("CUDAPATH_MORTON_CUBE")) : 1);
    {
        void *d_temp2 = nullptr;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_keys, d_keysSorted, d_ids, d_idsSorted, (int) nSeg, 0, 3 * CP_MORTON_BITS, stream));
        CK(S.alloc(&d_temp2, need));
        CK(cub::DeviceRadixSort::SortPairs(d_temp2, need, d_keys, d_keysSorted, d_ids, d_idsSorted, (int) nSeg, 0, 3 * CP_MORTON_BITS, stream));
    }
    CK(S.alloc(&d_sortedBox, sizeof(float) * 6 * (size_t) nSeg));
    CK(S.alloc(&d_prims, sizeof(uint32_t) * (size_t) nSeg));
    k_gather_leaf_boxes<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_leafBox, d_idsSorted, d_refPrim, nSeg, d_sortedBox, d_prims);
    float4 *d_leafSeg = nullptr;
    CK(S.alloc(&d_leafSeg, sizeof(float4) * 2 * (size_t) nSeg));
    k_leaf_records<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_vtx, d_prims, nSeg, d_leafSeg);

    if (nSeg <= CP_LEAF_MAX) {
        CK(S.alloc(&d_final, sizeof(BVH4Node)));
        k_single_leaf_root<<<1, 1, 0, stream>>>(d_sortedBox, (int) nSeg, d_final);
        wideCount = 1;
    } else {
        const int nInner = (int) nSeg - 1;
        CK(S.alloc(&d_children, sizeof(int2) * (size_t) nInner)); CK(S.alloc(&d_ranges, sizeof(int2) * (size_t) nInner));
        CK(S.alloc(&d_parentInner, sizeof(int) * (size_t) nInner)); CK(S.alloc(&d_parentLeaf, sizeof(int) * (size_t) nSeg));
        CK(S.alloc(&d_flags, sizeof(int) * (size_t) nInner)); CK(cudaMemsetAsync(d_flags, 0, sizeof(int) * (size_t) nInner, stream));
        CK(S.alloc(&d_innerBox, sizeof(float) * 6 * (size_t) nInner));
        k_radix_tree<<<(nInner + B - 1) / B, B, 0, stream>>>(d_keysSorted, (int) nSeg, d_children, d_parentInner, d_parentLeaf, d_ranges);
        k_refit<<<(nSeg + B - 1) / B, B, 0, stream>>>(d_children, d_parentInner, d_parentLeaf, d_sortedBox, (int) nSeg, d_innerBox, d_flags, d_ranges, leafSplitCost);
        // Collapse level by level.  A wide node is created only for a binary inner node (the one it absorbs), and distinct wide
        // nodes absorb distinct binary nodes, so nInner bounds the wide-node count.  In practice half of that is never reached
        // (hair scenes: 0.24 to 0.38 nodes per reference, growing with the pre-split cap), so the scratch array starts there
        // and the collapse is repeated with the full bound in the unlikely case that it overflows.
        CK(S.alloc(&d_q0, sizeof(CollapseItem) * (size_t) nInner)); CK(S.alloc(&d_q1, sizeof(CollapseItem) * (size_t) nInner));
        CK(S.alloc(&d_counters, sizeof(int) * 4));
        int levels = 0, wideCapacity = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            int capacity = attempt == 0 ? (int) std::min<long long>(nInner, (long long) nSeg / 2 + 1024) : nInner;
            if (attempt == 0 && getenv("CUDAPATH_TEST_COLLAPSE_CAP")) capacity = std::max(1, std::min(nInner, atoi(getenv("CUDAPATH_TEST_COLLAPSE_CAP"))));   // tests: force the retry
            if (d_wide) { S.release(d_wide); cudaStreamSynchronize(stream); dev_free(d_wide); d_wide = nullptr; }     // the first attempt's array
            CK(S.alloc(&d_wide, sizeof(BVH4Node) * (size_t) capacity)); wideCapacity = capacity;
            CollapseItem root{0, 0};
            int init[4] = {0, 1, 0, 0}; // [0]=next-level count, [1]=wide count, [2]=error
            CK(cudaMemcpyAsync(d_q0, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(d_counters, init, sizeof(init), cudaMemcpyHostToDevice, stream));
            int nIn = 1; bool overflow = false;
            levels = 0;
            while (nIn > 0) {
                k_collapse<<<(nIn + 127) / 128, 128, 0, stream>>>(d_q0, nIn, d_q1, d_counters, d_counters + 1, d_children, d_ranges,
                                                                 d_innerBox, d_sortedBox, d_wide, capacity, d_counters + 2);
                int h[3];
                CK(cudaMemcpyAsync(h, d_counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                if (h[2]) { overflow = true; break; }
                nIn = h[0]; wideCount = h[1];
                CK(cudaMemsetAsync(d_counters, 0, sizeof(int), stream));
                std::swap(d_q0, d_q1);
                if (++levels > 4096) { err = "BVH collapse did not terminate"; return false; }
            }
            if (!overflow) break;
            if (capacity == nInner) { err = "BVH collapse overflow"; return false; }
        }
        info.levels = levels;
        if ((size_t) wideCount * 2 >= (size_t) wideCapacity) { d_final = d_wide; }      // the scratch array is mostly full: it becomes the node array as it is
        else {
            CK(S.alloc(&d_final, sizeof(BVH4Node) * (size_t) wideCount));
            CK(cudaMemcpyAsync(d_final, d_wide, sizeof(BVH4Node) * (size_t) wideCount, cudaMemcpyDeviceToDevice, stream));
        }
    }
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    out.nodes = d_final; out.prims = d_prims; out.leafSeg = d_leafSeg; out.nodeCount = (uint32_t) wideCount; out.primCount = nSeg;
    info.nodes = (uint32_t) wideCount;
    S.release(d_final); S.release(d_prims); S.release(d_leafSeg);
    return true;
}

} // namespace cp

// cp_mesh.cu -- device-side set-up of triangle meshes (sm_100a), compiled with -fmad=false.
//
// Replaces the TriAccel precomputation of ShapeKDTree::build (src/librender/skdtree.cpp:78-110, TriAccel::load
// include/mitsuba/render/triaccel.h:61-97) and flattens the caller's TriMesh arrays (positions / optional vertex normals /
// indices, include/mitsuba/render/trimesh.h) into the scene-wide float4 streams described in cp_tri.cuh.
// No FMA contraction here: the records decide hits bit for bit against the reference's un-fused x86 arithmetic.
#include "cp_host.h"

namespace cp {

__global__ void k_pack_mesh(const float *__restrict__ xyz, const float *__restrict__ nrm, uint32_t nVerts, float4 *pos, float4 *outNrm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVerts) return;
    pos[i] = make_float4(xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2], 0.0f);
    outNrm[i] = nrm ? make_float4(nrm[3 * (size_t) i], nrm[3 * (size_t) i + 1], nrm[3 * (size_t) i + 2], 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// one thread per triangle: global indices + the 48-byte TriAccel record
__global__ void k_tri_accel(const uint32_t *__restrict__ localIdx, uint32_t nTris, uint32_t vertexOffset, uint32_t shapeIndex,
                            const float4 *__restrict__ pos /* scene-wide */, uint32_t *outIdx, float4 *outAccel) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nTris) return;
    const uint32_t i0 = localIdx[3 * (size_t) j] + vertexOffset, i1 = localIdx[3 * (size_t) j + 1] + vertexOffset, i2 = localIdx[3 * (size_t) j + 2] + vertexOffset;
    outIdx[3 * (size_t) j] = i0; outIdx[3 * (size_t) j + 1] = i1; outIdx[3 * (size_t) j + 2] = i2;
    const float4 qa = pos[i0], qb = pos[i1], qc = pos[i2];
    const V3 A(qa.x, qa.y, qa.z), B(qb.x, qb.y, qb.z), C(qc.x, qc.y, qc.z);
    const V3 b = C - A, c = B - A, N = cross(c, b);
    int k = 0;
    for (int a = 0; a < 3; ++a) if (fabsf(comp(N, a)) > fabsf(comp(N, k))) k = a;
    const int u = (k + 1) % 3, v = (k + 2) % 3;                  // waldModulo
    const float n_k = comp(N, k), denom = comp(b, u) * comp(c, v) - comp(b, v) * comp(c, u);
    float4 ra, rb, rc;
    if (denom == 0) {                                            // degenerate: k = 3, never hit (triaccel.h:77-80)
        ra = make_float4(__uint_as_float(3u), 0.0f, 0.0f, 0.0f); rb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        rc = make_float4(0.0f, 0.0f, __uint_as_float(shapeIndex), __uint_as_float(j));
    } else {
        ra = make_float4(__uint_as_float((uint32_t) k), comp(N, u) / n_k, comp(N, v) / n_k, dot(A, N) / n_k);
        rb = make_float4(comp(A, u), comp(A, v), comp(b, u) / denom, -comp(b, v) / denom);
        rc = make_float4(comp(c, v) / denom, -comp(c, u) / denom, __uint_as_float(shapeIndex), __uint_as_float(j));
    }
    outAccel[3 * (size_t) j] = ra; outAccel[3 * (size_t) j + 1] = rb; outAccel[3 * (size_t) j + 2] = rc;
}

void pack_mesh(const float *d_xyz, const float *d_nrm, uint32_t nVerts, float4 *d_pos, float4 *d_outNrm, cudaStream_t stream) {
    if (nVerts) k_pack_mesh<<<(nVerts + 255) / 256, 256, 0, stream>>>(d_xyz, d_nrm, nVerts, d_pos, d_outNrm);
}
void build_tri_accel(const uint32_t *d_localIdx, uint32_t nTris, uint32_t vertexOffset, uint32_t shapeIndex, const float4 *d_pos,
                     uint32_t *d_outIdx, float4 *d_outAccel, cudaStream_t stream) {
    if (nTris) k_tri_accel<<<(nTris + 255) / 256, 256, 0, stream>>>(d_localIdx, nTris, vertexOffset, shapeIndex, d_pos, d_outIdx, d_outAccel);
}

} // namespace cp

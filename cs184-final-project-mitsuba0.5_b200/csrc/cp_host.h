// cp_host.h -- host-side declarations shared by the translation units of libcudapath.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "cp_scene.cuh"

namespace cp {

// cp_mem.cpp -- caching device allocator (size-keyed free list over cudaMalloc).  dev_free(): the caller guarantees that no device
// work touching the block is still pending.
cudaError_t dev_alloc(void **p, size_t bytes);
template <typename T> inline cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc((void **) p, bytes); }
void dev_free(const void *p);
size_t dev_cached_bytes();
void dev_trim();

struct BuildInfo { uint32_t segments = 0, triangles = 0, references = 0, nodes = 0; int levels = 0; };

// cp_bvh.cu
bool build_bvh(const float4 *d_vtx, uint32_t vtxCount, ShapeDev *d_shapes, int shapeCount, const MeshDev &mesh, int maxSplit, float leafSplitCost, cudaStream_t stream,
               BVHDev &out, BuildInfo &info, std::string &err);

// cp_mesh.cu
void pack_mesh(const float *d_xyz, const float *d_nrm, uint32_t nVerts, float4 *d_pos, float4 *d_outNrm, cudaStream_t stream);
void build_tri_accel(const uint32_t *d_localIdx, uint32_t nTris, uint32_t vertexOffset, uint32_t shapeIndex, const float4 *d_pos,
                     uint32_t *d_outIdx, float4 *d_outAccel, cudaStream_t stream);

void pack_vertices(const float *d_xyz, const uint8_t *d_starts, uint32_t n, uint32_t shape, float4 *d_out, cudaStream_t stream);

// cp_tables.cu -- device-side precomputation of the Marschner azimuthal tables and the envmap CDFs
struct MarschnerTables { float4 *tab = nullptr; float *cdf = nullptr; float *sums = nullptr; float *pdf = nullptr; };
bool build_marschner_tables(float eta, float betaR, const float sigmaA[3], const float *glPoints140, const float *glWeights140,
                            cudaStream_t stream, MarschnerTables &out, std::string &err);
struct EnvTables { float4 *texels = nullptr; float *cdfCols = nullptr; float *cdfRows = nullptr; float *rowWeights = nullptr; float normalization = 0;
                   float4 *mipTexels = nullptr; EnvMipInfo *mipInfo = nullptr; };
void quantize_texels(const float *d_rgb, int n, float4 *d_texels, cudaStream_t stream);      // fp32 RGB -> half-quantised float4 (cp_tables.cu)
bool build_env_tables(const float *d_rgb, int w, int h, cudaStream_t stream, EnvTables &out, std::string &err);

// cp_host_data.cpp -- host-side set-up that stays on the CPU (file parsing, tiny tables)
void gauss_legendre_140(float *points, float *weights);
// Loads <dataDir>/microfacet/<beckmann|ggx|phong>.dat and reduces it to the 1-D slice used by the Marschner
// diffuse term: out T(|cos|^(1/4)) samples for (eta, alpha) and the internal diffuse reflectance Fdr.
bool rough_transmittance_slice(const std::string &dataDir, int distribution, float eta, float alpha,
                               std::vector<float> &outT, float &outFdr, std::string &err);

struct HairFileData { std::vector<float> xyz; std::vector<uint8_t> startsFiber; float radius = 0; };
bool load_hair_file(const std::string &path, float radius, float angleThresholdDeg, float reduction, const float toWorld[16],
                    HairFileData &out, std::string &err);

struct MeshFileData { std::vector<float> xyz, normals /* empty = face normals */, uvs /* 2 per vertex; empty = the file has no texture coordinates */; std::vector<uint32_t> indices; };
// Radiance RGBE (.hdr) image -> top-down RGB fp32 (Bitmap::readRGBE, src/libcore/bitmap.cpp:3590-3678)
bool load_rgbe_file(const std::string &path, std::vector<float> &rgb, int &w, int &h, std::string &err);
bool mat4_invert_f32(const float *a, float *out);     // Matrix<4,4,float>::invert of the reference (matrix.inl:138-193), row-major
// fresnelDiffuseReflectance(eta, fast = false): adaptive Gauss-Lobatto quadrature of F(sqrt(xi), eta) over [0, 1] (util.cpp:807-862, quad.cpp:287-420)
float fresnel_diffuse_reflectance(float eta);
// Random(seed).nextFloat() of the reference, n times (SFMT-19937: src/libcore/random.cpp); the hair loader's `reduction` draws from Random()
void mitsuba_random_floats(uint64_t seed, size_t n, float *out);
// hdrfilm's OpenEXR output (scan lines, uncompressed, half or float RGB)
bool write_exr_file(const std::string &path, const float *rgb, int w, int h, bool half, std::string &err);
void float_to_half_array(const float *in, size_t n, uint16_t *out);
bool load_obj_file(const std::string &path, const float toWorld[16], bool faceNormals, bool flipNormals, bool flipTexCoords, MeshFileData &out, std::string &err);

// cp_host_mip.cpp -- Lanczos MIP pyramid of the environment map + EWA weight table (mipmap.h:180-302)
struct EnvMipLevel { int w = 0, h = 0; std::vector<float> rgb; };
void build_env_pyramid(const float *rgb, int w, int h, std::vector<EnvMipLevel> &levels, float lut[64]);

// cp_host_sunsky.cpp
struct SunSkyParams { float turbidity = 3, albedo[3] = {0.2f, 0.2f, 0.2f}, sunDirection[3] = {0, 1, 0}, skyScale = 1, sunScale = 1, sunRadiusScale = 1, stretch = 1; int resolution = 512; };
bool bake_sunsky(const std::string &dataDir, const SunSkyParams &p, std::vector<float> &rgb, int &w, int &h, std::string &err, float *sunRadianceOut = nullptr);

} // namespace cp

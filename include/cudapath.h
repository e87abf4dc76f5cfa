/* cudapath.h -- C ABI of the B200-native hair path-tracing hot path.
 *
 * This is the boundary a librender-based `cudapath` integrator plugin binds: the plugin
 * flattens the loaded Mitsuba scene into plain arrays and calls these entry points; nothing C++ crosses it.
 * All pointers are caller-owned HOST memory unless the name ends in `_dev`.  Every function returns 0 (or a
 * non-negative id) on success and a negative value on failure; cudapath_last_error() then returns a message
 * (the reference reports errors by throwing from Log(EError), src/libcore/logger.cpp:100,147).
 * One context owns one CUDA device (cudapath_create) or several of one box (cudapath_create_multi); a context is
 * thread-compatible (use it from one thread at a time).  Entry points restore the caller's current CUDA device before returning.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference tree).
 */
#ifndef CUDAPATH_H
#define CUDAPATH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cudapath_ctx cudapath_ctx;

/* ---- context ---------------------------------------------------------------------------------------------- */
/* Replaces Scheduler/LocalWorker set-up for this path (src/mitsuba/mitsuba.cpp:280-329): one context per GPU. */
int cudapath_create(int cuda_device, cudapath_ctx **out);
/* `mitsuba -p N` (src/mitsuba/mitsuba.cpp:218-222,280-282: N LocalWorkers) and the tile scheduler that feeds them
 * (BlockedRenderProcess, src/librender/renderproc.cpp:117-182) for this path: ONE context spanning n_devices GPUs of this box.
 * The returned context is used exactly like a single-device one: every scene-building call is repeated on all devices,
 * cudapath_build builds them concurrently (one host thread per device), and cudapath_render splits the sample range
 * [sample_begin, sample_end) into n_devices contiguous parts, renders each on its own device into a private full-size film and
 * sums the films onto cuda_devices[0] with a single ncclReduce over NVLink before the read-back.  The image does not depend on
 * n_devices up to fp32 summation order (the RNG is keyed by pixel, sample index and path vertex).  cudapath_render_dev, the
 * parity hooks and the statistics of a build refer to cuda_devices[0]; cudapath_get_stats after a render holds the job's totals.
 * n_devices == 1 is cudapath_create.  libnccl.so.2 is bound at run time, only by this call. */
int cudapath_create_multi(const int *cuda_devices, int n_devices, cudapath_ctx **out);
/* Number of devices a context spans (1 for cudapath_create) / CUDA devices visible to the process (0 when there is no driver). */
int cudapath_device_count(cudapath_ctx *ctx);
int cudapath_visible_devices(void);
/* Device time of the film reduce of the last multi-GPU cudapath_render, in ms (0 for a single device). */
double cudapath_last_reduce_ms(cudapath_ctx *ctx);
void cudapath_destroy(cudapath_ctx *ctx);
/* Device memory freed by contexts is parked in a per-device free list and reused by the next build / render of similar size (a
 * repeated job then never waits for the driver).  This returns the parked blocks of `cuda_device` to the driver. */
int cudapath_trim_memory(int cuda_device);
const char *cudapath_last_error(void);
/* Directory holding a Mitsuba `data/` tree (microfacet/{beckmann,ggx,phong}.dat, and for the sunsky helper
 * sunsky/hosek_rgb.f64 + cie1931.f32).  Replaces FileResolver look-ups (src/bsdfs/rtrans.h:95-97). */
int cudapath_set_data_dir(cudapath_ctx *ctx, const char *path);

/* ---- BSDF plugins ------------------------------------------------------------------------------------------ */
/* `kajiyakay` plugin: KajiyaKay(props)+configure(), src/bsdfs/kajiyakay.cpp:60-107. Returns the bsdf id. */
int cudapath_add_bsdf_kajiyakay(cudapath_ctx *ctx, const float diffuse_reflectance[3], const float specular_reflectance[3], float exponent);
/* `marschner` plugin as built (class MarschnerDiffuse): ctor+configure(), src/bsdfs/marschner_diffuse.cpp:113-160,193-247;
 * builds the azimuthal tables (:751-847) on the device.  distribution: 0 beckmann, 1 ggx, 2 phong
 * (src/bsdfs/microfacet.h:99-135).  Returns the bsdf id. */
int cudapath_add_bsdf_marschner(cudapath_ctx *ctx, float int_ior, float ext_ior, const float diffuse_reflectance[3],
                                const float specular_reflectance[3], float alpha, int distribution, int nonlinear);

/* `roughplastic` plugin (the BSDF of the default models/{straight-hair,curly-hair,furball}/scene.xml): RoughPlastic ctor+configure(),
 * src/bsdfs/roughplastic.cpp:186-304, with MicrofacetDistribution(props) of src/bsdfs/microfacet.h:99-146.  distribution: 0 beckmann,
 * 1 ggx, 2 phong; sample_visible as the `sampleVisible` property (default true; ignored for phong).  Reference defaults: int_ior 1.49
 * (polypropylene), ext_ior 1.000277 (air), alpha 0.1, specular 1, diffuse 0.5.  Returns the bsdf id. */
int cudapath_add_bsdf_roughplastic(cudapath_ctx *ctx, float int_ior, float ext_ior, const float diffuse_reflectance[3], const float specular_reflectance[3],
                                   float alpha, int distribution, int sample_visible, int nonlinear);
/* The fork's second Marschner class, src/bsdfs/marschner.cpp (NOT part of its build; SURVEY M7, "fixed" mode): ctor :110-138 with the
 * hard-coded sigmaA = 0.22 / beta = 0.1 / scale angle -0.1, eval with the TRT lobe only (:309-341), the real pdf (:347-407) and a
 * sample() that draws two additional 2-D numbers from the sampler (:421-535).  Reference defaults: int_ior 1.55 (amber), ext_ior
 * 1.000277 (air).  Returns the bsdf id. */
int cudapath_add_bsdf_marschner_fixed(cudapath_ctx *ctx, float int_ior, float ext_ior);
/* The same class as a physically meaningful, scene-driven mode: what the constructor hard-codes comes from the caller --
 * sigma_a (absorption, :122), beta_r (longitudinal roughness of the R lobe; TT = beta_r / 2, TRT = 2 beta_r, :131-134) and the scale
 * angle in radians (:135) -- and lobe_mask says which lobes eval() keeps (bit 0 R, bit 1 TT, bit 2 TRT; 7 = all three, i.e. the file without
 * the two lines :333-334 that zero R and TT; 4 = cudapath_add_bsdf_marschner_fixed).  pdf (:347-407) and sample (:421-535) always cover the
 * three lobes, so with lobe_mask = 7 the sampling density matches the function it samples.  Returns the bsdf id. */
int cudapath_add_bsdf_marschner_full(cudapath_ctx *ctx, float int_ior, float ext_ior, const float sigma_a[3], float beta_r, float scale_angle_rad,
                                     int lobe_mask);
/* `diffuse` plugin with a constant reflectance: SmoothDiffuse ctor+configure(), src/bsdfs/diffuse.cpp:70-103; two_sided != 0
 * wraps it in the `twosided` adapter (src/bsdfs/twosided.cpp:50-181) with the same BRDF on both sides.  For triangle meshes
 * that accompany the fibers (a scalp or head under the hair).  Returns the bsdf id. */
int cudapath_add_bsdf_diffuse(cudapath_ctx *ctx, const float reflectance[3], int two_sided);
/* `thindielectric` plugin (models/straight-hair/scene_thindielectric.xml): ThinDielectric ctor+configure(),
 * src/bsdfs/thindielectric.cpp:73-125.  Reference defaults: int_ior 1.5046 (bk7), ext_ior 1.000277 (air), both colours 1.
 * Two discrete components (EDeltaReflection, ENull): the path tracer does no emitter sampling at such a vertex.  Returns the bsdf id. */
int cudapath_add_bsdf_thindielectric(cudapath_ctx *ctx, float int_ior, float ext_ior, const float specular_reflectance[3], const float specular_transmittance[3]);
/* `marschnerdielectric` plugin, the fork's third hair BSDF (models/straight-hair/scene_dielectric*.xml): ctor+configure(),
 * src/bsdfs/marschnerdielectric.cpp:128-221.  Reference defaults: int_ior 1.501 (benzene), ext_ior 1.000277 (air), diffuse 0.5,
 * specular reflectance / transmittance 0.1, exponent 30.  Reproduced as committed (eval() identically zero, see cp_bsdf.cuh).
 * Returns the bsdf id. */
int cudapath_add_bsdf_marschnerdielectric(cudapath_ctx *ctx, float int_ior, float ext_ior, const float diffuse_reflectance[3],
                                          const float specular_reflectance[3], const float specular_transmittance[3], float exponent);
/* `plastic` plugin (models/teapot/scene.xml:31-38): SmoothPlastic ctor+configure(), src/bsdfs/plastic.cpp:140-217 -- a delta reflection over a
 * diffuse base, fdrInt / fdrExt from fresnelDiffuseReflectance (src/libcore/util.cpp:807-862).  Reference defaults: int_ior 1.49
 * (polypropylene), ext_ior 1.000277 (air), specular 1, diffuse 0.5.  Returns the bsdf id. */
int cudapath_add_bsdf_plastic(cudapath_ctx *ctx, float int_ior, float ext_ior, const float diffuse_reflectance[3], const float specular_reflectance[3], int nonlinear);
/* `mirror`, a plugin the fork adds (src/bsdfs/mirror.cpp:187-276; models/teapot/mirror_scene.xml:32-36): one delta reflection on the front side, reproduced as
 * committed -- eval() is zero in both measures, pdf() is 1 in the discrete one.  Reference default: specularReflectance 1.  Returns the bsdf id. */
int cudapath_add_bsdf_mirror(cudapath_ctx *ctx, const float specular_reflectance[3]);
/* <texture type="checkerboard"> (src/textures/checkerboard.cpp:49-72 behind Texture2D, src/librender/texture.cpp:81-121) as the `reflectance` of a
 * `diffuse` or the `diffuseReflectance` of a `plastic` BSDF (models/teapot/scene.xml:43-52); the BSDF's configure() runs again on it.
 * Reference defaults: color0 0.4, color1 0.2, offsets 0, scales 1. */
int cudapath_bsdf_set_checkerboard(cudapath_ctx *ctx, int bsdf_id, const float color0[3], const float color1[3], float uoffset, float voffset,
                                   float uscale, float vscale);
/* <bsdf type="twosided"> around an existing `diffuse`, `roughplastic`, `plastic` or `mirror` BSDF with the same nested BRDF on both sides:
 * TwoSidedBRDF::configure / eval / pdf / sample, src/bsdfs/twosided.cpp:84-181. */
int cudapath_bsdf_set_twosided(cudapath_ctx *ctx, int bsdf_id);
/* fresnelDiffuseReflectance(eta, fast = false), src/libcore/util.cpp:814-862 (host only: adaptive Gauss-Lobatto, src/libcore/quad.cpp:287-420). */
int cudapath_fresnel_diffuse_reflectance(float eta, float *out);
/* Random(seed).nextFloat(), n times: Mitsuba's SFMT-19937 generator (src/libcore/random.cpp:396-405,551-553,630-639; default seed 5489,
 * include/mitsuba/core/random.h:113) -- the stream the hair loader's `reduction` parameter draws from (src/shapes/hair.cpp:629,672-673).  Host only. */
int cudapath_random_floats(uint64_t seed, uint64_t n, float *out);

/* ---- shapes ------------------------------------------------------------------------------------------------ */
/* Triangle mesh as ShapeKDTree sees it: TriMesh::getVertexPositions() / getVertexNormals() (NULL = face normals) /
 * getTriangles() after configure() (include/mitsuba/render/trimesh.h:60-140).  Triangles join the hair segments in the device
 * BVH and are tested with Wald's projection test (TriAccel, include/mitsuba/render/triaccel.h:61-158).  Returns the shape id. */
int cudapath_add_mesh(cudapath_ctx *ctx, const float *xyz, const float *normals, uint32_t n_vertices, const uint32_t *indices,
                      uint32_t n_triangles, int bsdf_id);
/* ... with TriMesh::getVertexTexcoords() (2 floats per vertex, NULL = none: its.uv is then the barycentric pair, include/mitsuba/render/skdtree.h:399-406). */
int cudapath_add_mesh_uv(cudapath_ctx *ctx, const float *xyz, const float *normals, const float *uvs, uint32_t n_vertices, const uint32_t *indices,
                         uint32_t n_triangles, int bsdf_id);
/* `rectangle` shape (models/teapot/scene.xml:56-62): Rectangle ctor+configure() / getAABB / rayIntersect / fillIntersectionRecord,
 * src/shapes/rectangle.cpp:81-171 -- the square [-1,1]^2 x {0} under to_world (row-major 4x4; inverted like Transform(const Matrix4x4 &),
 * include/mitsuba/core/transform.h:50-55), one analytic primitive of the top-level tree with uv = (x+1, y+1)/2.  A sheared to_world fails
 * with the reference's message.  Returns the shape id. */
int cudapath_add_rectangle(cudapath_ctx *ctx, const float to_world[16], int flip_normals, int bsdf_id);
/* `obj` shape: WavefrontOBJ(props) with all faces collapsed into one mesh (src/shapes/obj.cpp:186-349, createMesh :608-700) followed
 * by TriMesh::computeNormals (src/librender/trimesh.cpp:608-672).  Materials of the file are ignored.  Returns the shape id. */
int cudapath_add_mesh_file(cudapath_ctx *ctx, const char *filename, const float to_world[16], int face_normals, int flip_normals, int bsdf_id);
/* Loader only (no context, no GPU): the flattened arrays a host would pass to cudapath_add_mesh. */
typedef struct cudapath_mesh_file cudapath_mesh_file;
int cudapath_mesh_file_load(const char *filename, const float to_world[16], int face_normals, int flip_normals, cudapath_mesh_file **out);
uint32_t cudapath_mesh_file_vertex_count(const cudapath_mesh_file *m);
int cudapath_mesh_file_has_texcoords(const cudapath_mesh_file *m);                 /* `vt` records referenced by a face (src/shapes/obj.cpp:633-636) */
void cudapath_mesh_file_copy_texcoords(const cudapath_mesh_file *m, float *uvs);  /* 2 floats per vertex */
uint32_t cudapath_mesh_file_triangle_count(const cudapath_mesh_file *m);
int cudapath_mesh_file_has_normals(const cudapath_mesh_file *m);
void cudapath_mesh_file_copy(const cudapath_mesh_file *m, float *xyz, float *normals, uint32_t *indices);
void cudapath_mesh_file_free(cudapath_mesh_file *m);

/* `hair` shape from already-loaded fibers: HairShape::getVertices()/getStartFiber() (src/shapes/hair.h:51-57) and the
 * world-space radius of HairKDTree (src/shapes/hair.cpp:108-124).  starts_fiber has n_vertices entries (the sentinel
 * of hair.cpp:782 is added internally).  Returns the shape id. */
int cudapath_add_hair(cudapath_ctx *ctx, const float *xyz, const uint8_t *starts_fiber, uint32_t n_vertices, float radius, int bsdf_id);
/* `hair` shape from a .mitshair file (binary "BINARY_HAIR" or ASCII): HairShape::HairShape(props), src/shapes/hair.cpp:609-785.
 * Applies to_world (row-major 4x4), the radius scaling and the angle-threshold vertex merge.  Returns the shape id. */
int cudapath_add_hair_file(cudapath_ctx *ctx, const char *filename, float radius, float angle_threshold_deg, float reduction,
                           const float to_world[16], int bsdf_id);
/* Loader only (no context): returns a handle to query/copy, for hosts that want the flattened arrays. */
typedef struct cudapath_hair_file cudapath_hair_file;
int cudapath_hair_file_load(const char *filename, float radius, float angle_threshold_deg, float reduction, const float to_world[16], cudapath_hair_file **out);
uint32_t cudapath_hair_file_vertex_count(const cudapath_hair_file *h);
float cudapath_hair_file_radius(const cudapath_hair_file *h);
void cudapath_hair_file_copy(const cudapath_hair_file *h, float *xyz, uint8_t *starts_fiber);
void cudapath_hair_file_free(cudapath_hair_file *h);

/* ---- emitter ----------------------------------------------------------------------------------------------- */
/* `envmap` emitter from a lat-long RGB fp32 bitmap (what SunSkyEmitter hands to its nested envmap, src/emitters/sunsky.cpp:218-229):
 * EnvironmentMap ctor + configure(), src/emitters/envmap.cpp:100-190,260-329 (half quantisation, CDFs built on the device). */
int cudapath_set_envmap(cudapath_ctx *ctx, const float *rgb, int width, int height, const float to_world[16], float scale);
/* `envmap` emitter from a Radiance RGBE file (`<string name="filename" value="textures/envmap.hdr"/>`, models/teapot/scene.xml):
 * Bitmap::readRGBE, src/libcore/bitmap.cpp:3590-3678 (flat and run-length encoded scanlines), then as cudapath_set_envmap.  The
 * `gamma` override and the .mip cache file of the plugin are not part of this path. */
int cudapath_set_envmap_file(cudapath_ctx *ctx, const char *filename, const float to_world[16], float scale);
/* The reader alone (host only): width/height always, pixels (top-down RGB fp32, 3*w*h floats) when out_rgb is not NULL. */
int cudapath_load_rgbe(const char *filename, float *out_rgb, int *out_width, int *out_height);
/* Host only (no context, no GPU): level `level` of the MIP pyramid the envmap emitter builds from a lat-long fp32 bitmap -- TMIPMap's
 * progressive downsampling with the 2-lobed Lanczos filter, ERepeat along u / EClamp along v, clamped to [0, inf)
 * (include/mitsuba/render/mipmap.h:180-271, src/libcore/bitmap.cpp:2230-2328, include/mitsuba/core/rfilter.h:107-460,
 * src/rfilters/lanczos.cpp:42-55) -- as fp32 BEFORE the half quantisation of the stored texels.  Returns the number of levels. */
int cudapath_env_pyramid_level(const float *rgb, int width, int height, int level, int *out_width, int *out_height, float *out_rgb);
/* `sunsky` emitter: SunSkyEmitter ctor, src/emitters/sunsky.cpp:100-236 (bakes the map on the host, then cudapath_set_envmap). */
int cudapath_set_sunsky(cudapath_ctx *ctx, float turbidity, const float albedo[3], const float sun_direction[3], float sky_scale,
                        float sun_scale, float sun_radius_scale, int resolution);
/* Bake only: writes resolution x resolution/2 x 3 floats. */
int cudapath_bake_sunsky(const char *data_dir, float turbidity, const float albedo[3], const float sun_direction[3], float sky_scale,
                         float sun_scale, float sun_radius_scale, int resolution, float *out_rgb);
/* Radiance of the sun disc in linear RGB for a sun direction and turbidity: computeSunRadiance, src/emitters/sunsky/sunmodel.h:260-371, with the
 * spectrum -> RGB conversion SunSkyEmitter applies (sunsky.cpp:170-171); the value the bake splats.  Host only. */
int cudapath_sun_radiance(const char *data_dir, float turbidity, const float sun_direction[3], float out_rgb[3]);

/* ---- sensor / film / integrator ---------------------------------------------------------------------------- */
/* `perspective` sensor: ProjectiveCamera/PerspectiveCamera props + configure(), src/librender/sensor.cpp:156-160,225-300,
 * src/sensors/perspective.cpp:126-160.  fov is along x. */
int cudapath_set_camera_perspective(cudapath_ctx *ctx, const float to_world[16], float fov_x_deg, float near_clip, float far_clip,
                                    int film_width, int film_height);
/* film + reconstruction filter: ReconstructionFilter::configure(), src/libcore/rfilter.cpp:37-55.
 * filter: 0 tent (radius 1), 1 box, 2 gaussian (param = stddev, default 0.5). has_alpha: film pixelFormat carries alpha. */
int cudapath_set_film(cudapath_ctx *ctx, int filter, float param, int has_alpha);
/* `path` integrator params: MonteCarloIntegrator(props), src/librender/integrator.cpp:190-225. */
int cudapath_set_integrator(cudapath_ctx *ctx, int max_depth, int rr_depth, int strict_normals, int hide_emitters);

/* ---- build + render ---------------------------------------------------------------------------------------- */
/* Scene::initialize(), src/librender/scene.cpp:322-413: uploads the geometry, builds the BVH on the device (replacing
 * ShapeKDTree::build / HairKDTree), computes shape/scene bounds and the environment bounding sphere. */
int cudapath_build(cudapath_ctx *ctx);
/* SamplingIntegrator::render -> renderBlock -> Li -> ImageBlock::put for sample indices [sample_begin, sample_end) of `spp`
 * (src/librender/integrator.cpp:95-188, src/integrators/path/path.cpp:119-294).  Writes the ACCUMULATED film
 * (width*height*5 floats: sum w*R, w*G, w*B, w*alpha, w) to host memory.  Films of disjoint sample ranges add up. */
int cudapath_render(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *out_film);
/* Same, accumulating into a caller-provided DEVICE buffer on a caller-provided CUDA stream (cudaStream_t as void*, may be 0);
 * the buffer must be zeroed by the caller before the first range.  Used with torch/NCCL for the multi-GPU film reduce. */
int cudapath_render_dev(cudapath_ctx *ctx, uint32_t spp, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float *film_dev, void *stream);
/* Pixel-space sharding (the image tiles of BlockedRenderProcess, src/librender/renderproc.cpp:117-182; its default tile is the same
 * 32x32 pixels, src/mitsuba/mitsuba.cpp:144): after this call cudapath_render / cudapath_render_dev only trace the paths of the pixel
 * blocks owned by shard `shard_index` of `shard_count` (every group of shard_count neighbouring blocks gives one block to each shard,
 * so each shard's pixels are spread evenly over the image); all other pixels stay zero except for the
 * reconstruction filter's one-pixel border, so the films of the `shard_count` shards ADD UP to the full image exactly like the films of
 * disjoint sample ranges do.  Keeping all sample indices of a pixel on one device keeps the ray density per region of the scene --
 * and with it the cache hit rates of the traversal -- at the level of the single-device render, which sample-range sharding does
 * not.  (1, 1)-style reset: shard_index 0 of shard_count 1. */
int cudapath_set_pixel_shard(cudapath_ctx *ctx, uint32_t shard_index, uint32_t shard_count);
/* Integrator::cancel() (include/mitsuba/render/integrator.h:76-84, SamplingIntegrator::cancel src/librender/integrator.cpp:90-93; reached
 * from RenderJob::cancel, include/mitsuba/render/renderjob.h:81, on another thread while render() blocks and then returns false).  The ONLY entry point that may be called concurrently with a render of the same
 * context.  The running (or next) cudapath_render / cudapath_render_dev returns -1 with the message "render cancelled" within one
 * bounce; the film then holds the finished waves only.  The request is consumed by that render. */
int cudapath_cancel(cudapath_ctx *ctx);
/* Progress of a render (ProgressReporter of SamplingIntegrator::render, src/librender/integrator.cpp:95-138): `callback` runs on the
 * rendering thread after every finished wave with the path slots done and the total of the call (pixels are padded to 8x8 tiles).
 * NULL removes it. */
int cudapath_set_progress_callback(cudapath_ctx *ctx, void (*callback)(void *user, uint64_t paths_done, uint64_t paths_total), void *user);
/* Film::develop normalisation, src/libcore/fmtconv.cpp:955-1056: rgb = sum / weight (0 where weight == 0). */
int cudapath_develop(const float *film, int width, int height, float *out_rgb);
/* HDRFilm::develop with its default fileFormat "openexr" (src/films/hdrfilm.cpp:213-246, 580-640): writes developed linear RGB (cudapath_develop) as a
 * single-part scan-line OpenEXR file, channels B / G / R, uncompressed; half_float != 0: componentFormat float16 (the plugin's default, round to
 * nearest even), else float32.  Host only. */
int cudapath_write_exr(const char *filename, const float *rgb, int w, int h, int half_float);
int cudapath_float_to_half(const float *in, uint64_t n, uint16_t *out);
/* LDRFilm::develop with the `gamma` tonemapper, src/films/ldrfilm.cpp:300-321 -> Bitmap::convert(ERGB, EUInt8, gamma, 2^exposure)
 * (src/libcore/fmtconv.cpp:984-995,1104-1111,1137-1160): width*height*3 bytes.  gamma = -1 selects the sRGB curve (the ldrfilm
 * default); the hair scene files use 2.2.  Banner, Reinhard tonemapping and the PNG/JPEG encoders are not part of this path. */
int cudapath_develop_ldr(const float *film, int width, int height, float gamma, float exposure, uint8_t *out_rgb8);
/* Sampler.  kind 0 (default): the counter-based Philox4x32-10 stream keyed by (pixel, sample index, path vertex, seed) that replaces the sampler
 * plugins (include/mitsuba/render/sampler.h:105-117) -- any <sampler> of a scene file selects it.  kind 1, the sampler-faithful mode: the
 * reference's `sobol` plugin reproduced number for number -- SobolSampler (src/samplers/sobol.cpp:90-245: scramble through sampleTEA,
 * setFilmResolution(size, bucketed), generate / advance / setSampleIndex, next1D / next2D with the skipped dimension 4) over Gruenschloss'
 * enumeration of the (0,2)-sequence per pixel (src/samplers/sobolseq.h:59-133), with the direction numbers of src/samplers/sobolseq.cpp read from
 * <data dir>/sobol.bin; the path tracer draws in the order renderBlock and MIPathTracer::Li do (integrator.cpp:171, path.cpp:176,210,284,
 * marschner.cpp:473-474).  A path that needs more than 1024 dimensions fails the render with the plugin's message.  `independent`
 * (src/samplers/independent.cpp) cannot be reproduced on a wavefront: every worker of the reference consumes ONE sequential SFMT stream whose
 * position at a pixel depends on the lengths of all paths the worker traced before (and on which blocks the scheduler handed it). */
int cudapath_set_sampler(cudapath_ctx *ctx, int kind, uint64_t scramble);
int cudapath_get_sampler(cudapath_ctx *ctx);
/* Tunables: wave size in paths (0 = default), collect traversal statistics (slower, counting kernels),
 * profile_stages (CUDA events around every stage launch). */
int cudapath_set_options(cudapath_ctx *ctx, uint32_t wave_size, int collect_stats, int profile_stages);

/* Math mode of the shading stages (BSDF eval / pdf / sample, emitter lookups).  strict = 1: every elementary function is correctly
 * rounded fp32 (evaluated in fp64) -- bit-identical to a faithful CPU evaluation, the mode the path-replay parity tests run in.
 * strict = 0 (default; the environment variable CUDAPATH_MATH=strict changes the default of new contexts): only the amplified chain
 * of the Marschner lobes (asin of the incident inclination, sin / cos of the three shifted lobe angles that M() multiplies by 1/v)
 * keeps its exact bits; every other call uses the 1-2 ulp fp32 CUDA functions.  BSDF values then stay within ~1e-6 relative of the
 * strict mode (BASELINE.json asks for 1e-4 against the reference's CPU build, whose own libm is only defined to an ulp).
 * Ray queries, the FP64 cylinder test, camera rays and the film are the same in both modes.  get: 1 strict, 0 fast. */
int cudapath_set_math_mode(cudapath_ctx *ctx, int strict);
int cudapath_get_math_mode(cudapath_ctx *ctx);

/* BVH build tunable: a long thin segment is referenced by up to max_split boxes cut along its axis (default 16; 1 = off). */
int cudapath_set_build_options(cudapath_ctx *ctx, int max_split);
/* Build effort from the size of the job: paths_per_device = width x height x samples per pixel / devices the scene is going to be
 * rendered with.  The finer pre-split (16) costs 13 ms more build per 32 M references and saves 8 % of the render: it is chosen at
 * 2^25 paths per device and above, the cap of 8 below (a job split over many devices).  Like the reference's kd-tree build
 * (src/shapes/hair.cpp:108-159, one tree whatever the job) this never changes a result.  Call before cudapath_build(); an explicit
 * cudapath_set_build_options() or CUDAPATH_MAX_SPLIT wins.  cudapath_load_scene_xml() calls it with the film size and sampleCount of the file. */
int cudapath_set_job_size_hint(cudapath_ctx *ctx, uint64_t paths_per_device);

typedef struct cudapath_stats {
    uint64_t paths, rays, shadow_rays;          /* same definitions as the reference's "Normal rays traced"/"Shadow rays traced" (src/librender/skdtree.cpp:46-47) */
    uint64_t kernel_launches, bounces;
    uint64_t nodes_visited, prims_tested;       /* closest-hit rays; only with collect_stats */
    uint64_t shadow_nodes_visited, shadow_prims_tested;
    uint64_t unsupported_filtered_lookups, dropped_samples;
    uint64_t segments, bvh_nodes, bvh_references, triangles;
    double build_ms, render_ms;                 /* device time of the last build / render (CUDA events) */
    /* per-stage device time of the last render, summed over launches (CUDA events on the launching stream; only with
     * profile_stages) and the number of launches of each stage.  trace = BVH traversal (the closest-hit rays of a bounce and the
     * shadow rays of the bounce before share one launch), sort = coherence keys + radix sort of those rays */
    double trace_ms, shade_ms, sort_ms, raygen_ms, splat_ms;
    uint64_t trace_launches, shade_launches, sort_launches;
    uint64_t shadow_rays_traced;                 /* shadow_rays minus those whose emitter sample has an exactly zero contribution (not traced here) */
    uint64_t full_tests, shadow_full_tests;      /* exact primitive tests (FP64 cylinder / Wald triangle) after the fp32 pre-test; only with collect_stats */
    uint64_t host_waits;                         /* times the host had to wait for the device inside the last render's bounce loops */
} cudapath_stats;
int cudapath_get_stats(cudapath_ctx *ctx, cudapath_stats *out);
int cudapath_scene_bounds(cudapath_ctx *ctx, float aabb_min_max[6], float bsphere_center_radius[4]);
/* Film::getSize() (include/mitsuba/render/film.h:49-92) */
int cudapath_film_size(cudapath_ctx *ctx, int *width, int *height);
/* How the scene's film wants to be developed: `hdrfilm` (linear float output) or `ldrfilm` with its `gamma` (-1 = sRGB, the plugin's
 * default) and `exposure` (src/films/ldrfilm.cpp:180-181).  Set by the scene loader, read by whoever writes the image
 * (cudapath_develop / cudapath_develop_ldr); the render itself does not depend on it. */
int cudapath_set_film_output(cudapath_ctx *ctx, int hdr, float gamma, float exposure);
int cudapath_get_film_output(cudapath_ctx *ctx, int *out_hdr, float *out_gamma, float *out_exposure);

/* ---- parity hooks (host buffers; same device functions as the render path) ---------------------------------- */
/* BSDF::eval + BSDF::pdf (include/mitsuba/render/bsdf.h:369-441), wi/wo local, measure = ESolidAngle, typeMask = EAll */
int cudapath_bsdf_eval_batch(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf);
/* The same with measure = EDiscrete (include/mitsuba/render/common.h:56-67): non-zero only for discrete components (`thindielectric`) */
int cudapath_bsdf_eval_batch_discrete(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf);
/* BSDF::eval + BSDF::pdf from WORLD-space directions and one shading frame per tuple (frames: s, t, n = 9 floats): the directions go
 * through Frame::toLocal (include/mitsuba/core/frame.h:55-85) on the device, as Intersection::toLocal does for its.wi / bRec.wo
 * (include/mitsuba/render/skdtree.h:426-427, src/integrators/path/path.cpp:186).  The "random world frame" batch of config 5. */
int cudapath_bsdf_eval_batch_world(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *frames, const float *wi_world, const float *wo_world,
                                   float *out_eval, float *out_pdf);
/* BSDF::sample: out_type = sampledType | sampledComponent << 8 */
int cudapath_bsdf_sample_batch(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, float *out_wo,
                               float *out_weight, float *out_pdf, int32_t *out_type);
/* Same with the additional sampler draws of BSDFs that pull more than the 2-D sample they are handed (`extra`: 4 floats per
 * tuple = xiN.x, xiN.y, xiM.x, xiM.y of the fixed Marschner; may be NULL). */
int cudapath_bsdf_sample_batch_ex(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, const float *extra,
                                  float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type);
/* Same with texture coordinates per tuple (its.uv, 2 floats): what Texture2D::eval hands the textured BSDFs (src/librender/texture.cpp:112-121);
 * discrete != 0 selects the EDiscrete measure (the delta reflection of `plastic`, src/bsdfs/plastic.cpp:246-313). */
int cudapath_bsdf_eval_batch_uv(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, const float *uv, int discrete,
                                float *out_eval, float *out_pdf);
int cudapath_bsdf_sample_batch_uv(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, const float *uv,
                                  float *out_wo, float *out_weight, float *out_pdf, int32_t *out_type);
/* Scene::rayIntersect (any_hit = 0) / shadow-ray query (any_hit = 1).  out_prim = shape-local first-vertex index iv
 * (the reference's primitive id, src/shapes/hair.cpp:151-155) or, for a mesh, the triangle index within the mesh
 * (TriAccel::primIndex); out_record (optional) = p, n, s, t, wi (15 floats per ray) as filled by
 * HairShape::fillIntersectionRecord (src/shapes/hair.cpp:825-862) / the mesh branch of include/mitsuba/render/skdtree.h:346-427. */
int cudapath_intersect_batch(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                             int any_hit, int32_t *out_shape, uint32_t *out_prim, float *out_t, float *out_record);
/* Closest hit with the record plus its.uv and its.geoFrame.n (out_uv_geo_n: 5 floats per ray; zero on a miss): the mesh branch of
 * include/mitsuba/render/skdtree.h:346-427 (uv :399-406) and Rectangle::fillIntersectionRecord (src/shapes/rectangle.cpp:158-171). */
int cudapath_intersect_batch_uv(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                                int32_t *out_shape, uint32_t *out_prim, float *out_t, float *out_record, float *out_uv_geo_n);
int cudapath_env_eval_batch(cudapath_ctx *ctx, uint64_t n, const float *direction, float *out_rgb, float *out_pdf);
/* Emitter::evalEnvironment for rays WITH differentials (camera rays that leave the scene): MIPMap::eval with the EWA filter over the
 * 2-lobed-Lanczos MIP pyramid, maxAnisotropy 10 (src/emitters/envmap.cpp:150-181,391-407; include/mitsuba/render/mipmap.h:629-836).
 * rx_direction / ry_direction are the offset ray directions of RayDifferential. */
int cudapath_env_eval_filtered_batch(cudapath_ctx *ctx, uint64_t n, const float *direction, const float *rx_direction, const float *ry_direction, float *out_rgb);
/* One level of the environment map's MIP pyramid as stored (half-quantised texels, width*height*3 floats; out_rgb may be NULL to query
 * the size).  Returns the number of levels. */
int cudapath_env_mip_level(cudapath_ctx *ctx, int level, int *out_width, int *out_height, float *out_rgb);
int cudapath_env_sample_batch(cudapath_ctx *ctx, uint64_t n, const float *ref_point, const float *sample, float *out_direction,
                              float *out_value, float *out_pdf_dist);
int cudapath_camera_rays_batch(cudapath_ctx *ctx, uint64_t n, const float *pixel_sample, float *out_origin, float *out_direction, float *out_mint_maxt);
int cudapath_splat_batch(cudapath_ctx *ctx, uint64_t n, const float *position, const float *rgb, const float *alpha, float *out_film);
/* Precomputed tables, for table-level parity: 3x64x64x3 azimuthal values, 3x64x64 pdfs, 3x64x65 cdfs, 3x64 sums,
 * rt_size rough-transmittance samples, consts = {Fdr, specularSamplingWeight, eta, rt_size}. */
int cudapath_marschner_tables(cudapath_ctx *ctx, int bsdf_id, float *out_tables, float *out_pdfs, float *out_cdfs, float *out_sums,
                              float *out_rt, float *out_consts);
int cudapath_env_tables(cudapath_ctx *ctx, float *out_cdf_rows, float *out_cdf_cols, float *out_row_weights, float *out_normalization);
int cudapath_filter_table(cudapath_ctx *ctx, float out32[32]);

/* ---- device-resident variants for stage benchmarks (all pointers are DEVICE memory; asynchronous on `stream`) ---- */
int cudapath_bsdf_eval_batch_dev(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *wo, float *out_eval, float *out_pdf, void *stream);
int cudapath_bsdf_sample_batch_dev(cudapath_ctx *ctx, int bsdf_id, uint64_t n, const float *wi, const float *sample, float *out_wo,
                                   float *out_weight, float *out_pdf, int32_t *out_type, void *stream);
/* out_stats (optional, device, 3 x uint64): nodes visited, primitives pre-tested, exact tests (enables the counting variant of the kernel) */
int cudapath_intersect_batch_dev(cudapath_ctx *ctx, uint64_t n, const float *origin, const float *direction, const float *mint, const float *maxt,
                                 int any_hit, int32_t *out_shape, uint32_t *out_prim, float *out_t, unsigned long long *out_stats, void *stream);

/* Roofline denominators for the stage reports, measured on the context's device: read bandwidth of a resident buffer of `bytes`
 * streamed `iterations` times with 16-byte loads that bypass L1 (a buffer well below the L2 size gives the L2 bandwidth, a
 * multi-GB one the HBM read bandwidth).  Best of three timed passes, CUDA events. */
int cudapath_measure_read_bandwidth(cudapath_ctx *ctx, size_t bytes, int iterations, double *out_gb_per_s);

/* ---- scene files -------------------------------------------------------------------------------------------- */
/* SceneHandler (src/librender/scenehandler.cpp:70-250) for the subset of tags the hair scenes use: integrator `path`,
 * sensor `perspective` (+ film, rfilter, sampler sampleCount), bsdf `kajiyakay` / `marschner` / `diffuse` / `twosided`, shape `hair` / `obj`, emitter
 * `sunsky` / `envmap`-from-memory.  `defines` is a ';'-separated list of name=value pairs replacing $name in the file
 * (mitsuba -D, src/mitsuba/mitsuba.cpp:168).  Hair files that are missing are reported as an error.  On success the
 * scene is loaded into ctx (cudapath_build still has to be called) and *out_spp receives the sampler's sampleCount. */
int cudapath_load_scene_xml(cudapath_ctx *ctx, const char *filename, const char *defines, uint32_t *out_spp);
/* Dry run of the same loader without a context or a GPU: parses the file, checks every plugin and parameter it uses and writes one
 * line per object it would create into `report` (geometry files that are missing are noted, not treated as errors).  Returns 0 when
 * the whole file is inside the supported path; otherwise -1 and cudapath_last_error() names the first unsupported element. */
int cudapath_validate_scene_xml(const char *filename, const char *defines, char *report, size_t report_size);

#ifdef __cplusplus
}
#endif
#endif /* CUDAPATH_H */

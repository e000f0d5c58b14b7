/*
 * ne_b200.h — C ABI of the B200-native path-tracing backend for NarvalEngine.
 *
 * This is the drop-in boundary. The reference has no plugin/FFI mechanism; its seams are
 * two C++ types, and every entry point below names the reference interface it replaces:
 *
 *   Seam 2 (renderer)   OfflineEngine(Camera, SceneSettings, Scene*)            src/core/OfflineEngine.h:23-47
 *                       OfflineEngine::renderTile / postProcessing / pixels     src/core/OfflineEngine.cpp:39-76
 *   Seam 1 (integrator) Integrator::Li(Ray, Scene*)                             src/integrators/Integrator.h:13-25
 *   Scene feed          SceneReader::loadScene -> Scene, Camera, SceneSettings  src/io/SceneReader.cpp:10-675
 *                       ResourceManager::loadVDBasTexture / loadVolasTexture    src/core/ResourceManager.cpp:165-286
 *
 * Plain C: pointers + sizes only, no C++/torch types. The caller keeps ownership of every pointer in the
 * descriptors; the library copies what it needs into HBM during ne_b200_scene_upload(). One context drives one
 * GPU (one process per GPU; see INTEGRATION.md for the 8-GPU sample-index split). A context may be driven by
 * one host thread at a time. All functions return NE_B200_OK (0) or a negative ne_b200_status and never abort
 * (the reference LOG(FATAL)s in its loaders, SceneReader.cpp:24-41); the message is in ne_b200_last_error().
 *
 * There is NO CPU fallback: every entry point that computes runs CUDA kernels (sm_100a) and fails with
 * NE_B200_ERR_CUDA when no device is present.
 */
#ifndef NE_B200_H
#define NE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NE_B200_API_VERSION 1

typedef enum ne_b200_status {
	NE_B200_OK = 0,
	NE_B200_ERR_INVALID = -1,     /* bad argument / malformed descriptor */
	NE_B200_ERR_CUDA = -2,        /* CUDA runtime error or no device */
	NE_B200_ERR_NOMEM = -3,
	NE_B200_ERR_STATE = -4,       /* call order (render before upload, ...) */
	NE_B200_ERR_UNSUPPORTED = -5  /* valid reference feature this build does not cover yet */
} ne_b200_status;

/* ------------------------------------------------------------------------------------------------------------
 * Scene description (what SceneReader builds, in POD form).
 * ---------------------------------------------------------------------------------------------------------- */

/* TextureLayout subset the CPU sampler understands: src/materials/Texture.cpp:37-87 */
typedef enum ne_b200_tex_format {
	NE_B200_TEX_R32F = 0, NE_B200_TEX_RG32F = 1, NE_B200_TEX_RGB32F = 2, NE_B200_TEX_RGBA32F = 3, NE_B200_TEX_RGBA8 = 4
} ne_b200_tex_format;

/* Texture::wrapTextureCoordinates, src/materials/Texture.cpp:88-109 (clamp is also the default) */
typedef enum ne_b200_wrap { NE_B200_WRAP_CLAMP = 0, NE_B200_WRAP_MIRROR = 1 } ne_b200_wrap;

/* 2-D material texture (Texture, src/materials/Texture.h:77-98). Row-major, index = width*y + x. */
typedef struct ne_b200_texture {
	int32_t width, height;
	int32_t format;              /* ne_b200_tex_format */
	int32_t wrap_u, wrap_v;      /* ne_b200_wrap */
	const void* texels;
} ne_b200_texture;

/*
 * Density grid of a GridMedia (src/materials/GridMedia.h:16-92), i.e. the Texture produced by
 * ResourceManager::loadVolasTexture (.vol, ResourceManager.cpp:222-286) or loadVDBasTexture (.vdb, :165-220).
 * Texture space: index = width*height*z + width*y + x (Math.h:104-106).
 * Give EITHER `dense` (width*height*depth floats) OR `n_leaves` 8x8x8 leaf bricks exactly as an OpenVDB
 * FloatGrid leaf iterator yields them: leaf_origin = 3 ints per leaf (multiples of 8, texture space),
 * leaf_values = 512 floats per leaf with texture-x fastest: value(x,y,z) = leaf_values[512*l + 64*z + 8*y + x].
 * (For a .vdb the reference swaps axes, Q29: texture x = vdb z, texture z = vdb x, so an OpenVDB leaf buffer,
 * whose linear offset is (x<<6)|(y<<3)|z, is already in this order. INTEGRATION.md shows the loop.)
 * Voxels not covered by any leaf are 0 (inactive background).
 */
typedef struct ne_b200_volume {
	int32_t width, height, depth;
	const float* dense;
	int32_t n_leaves;
	const int32_t* leaf_origin;
	const float* leaf_values;
} ne_b200_volume;

/* Material kinds created by SceneReader::processMaterial, src/io/SceneReader.cpp:67-222 */
typedef enum ne_b200_material_type {
	NE_B200_MAT_MICROFACET = 0,   /* GlossyBSDF(GGX, Schlick 0.04) :76-140 */
	NE_B200_MAT_EMITTER = 1,      /* DiffuseLight :141-155 */
	NE_B200_MAT_VOLUME = 2,       /* GridMedia (volume >= 0) or HomogeneousMedia (volume < 0) + VolumeBSDF :187-219 */
	NE_B200_MAT_DIRECTIONAL = 3,  /* DirectionalLight :156-168: `li` = le (JSON albedo), `direction` = normalize(-position) */
	NE_B200_MAT_INFINITE = 4      /* InfiniteAreaLight :169-186: `env_tex` = the lat-long map; carried by a sphere primitive */
} ne_b200_material_type;

typedef enum ne_b200_phase { NE_B200_PHASE_ISOTROPIC = 0, NE_B200_PHASE_HG = 1 } ne_b200_phase;

typedef struct ne_b200_material {
	int32_t type;                                  /* ne_b200_material_type */
	/* microfacet: texture indices into scene.textures, -1 = absent (Material::sampleMaterial returns (0,0,0,1)) */
	int32_t albedo_tex, roughness_tex, metallic_tex, normal_tex;
	int32_t has_normal_flag;                       /* Material::hasTexture(NORMAL): true only if NORMAL was the LAST texture added (Q24, Material.h:38-44) */
	/* emitter: radiance li (DiffuseLight::li). directional: le, direction. */
	float li[3];
	float direction[3];
	/* volume */
	float scattering[3], absorption[3];
	float density_multiplier;                      /* JSON "density" */
	int32_t phase;                                 /* ne_b200_phase */
	float g;
	int32_t volume;                                /* index into scene.volumes, or -1 */
	int32_t env_tex;                               /* infinite area light map */
} ne_b200_material;

/* Primitive kinds created by SceneReader::processPrimitives, src/io/SceneReader.cpp:224-648 */
typedef enum ne_b200_primitive_type {
	NE_B200_PRIM_RECTANGLE = 0,   /* unit square z=0, corners (-.5,-.5,0),(.5,.5,0), normal (0,0,-1), uv (0,0)-(1,1) :385-513 */
	NE_B200_PRIM_SPHERE = 1,      /* centre (0,0,0) OCS + radius; transform = translate(position) only :332-384 */
	NE_B200_PRIM_POINT = 2,       /* vertex = position AND transform contains position (Q25) :268-331 */
	NE_B200_PRIM_VOLUME = 3,      /* proxy AABB [-0.5,0.5]^3 whose material has a medium :515-645 */
	NE_B200_PRIM_MESH = 4         /* obj/gltf: triangles + BVH, one material for all triangles :245-267, Model.cpp:159-370 */
} ne_b200_primitive_type;

typedef struct ne_b200_primitive {
	int32_t type;                 /* ne_b200_primitive_type */
	int32_t material;             /* index into scene.materials; -1 = none (mesh without material: paths end there) */
	float to_world[16];           /* InstancedModel::transformToWCS, column-major as glm::mat4 */
	float to_object[16];          /* InstancedModel::invTransformToWCS (= glm::inverse(to_world)); see ne_b200_make_transform */
	float radius;                 /* sphere */
	float point[3];               /* point: the vertex (JSON position) */
	int32_t collision;            /* InstancedModel::isCollisionEnabled (spheres only in the JSON) */
	/* mesh (type MESH): triangle soup in OCS. positions: 3 floats per vertex; uvs: 2 floats per vertex or NULL;
	 * indices: 3 uint32 per triangle, in the order the importer lists the faces. */
	int32_t n_vertices, n_triangles;
	const float* positions;
	const float* uvs;
	const uint32_t* indices;
} ne_b200_primitive;

typedef struct ne_b200_scene_desc {
	int32_t n_textures;   const ne_b200_texture* textures;
	int32_t n_volumes;    const ne_b200_volume* volumes;
	int32_t n_materials;  const ne_b200_material* materials;
	/* In JSON order. Like SceneReader, primitives whose material is an emitter go to Scene::lights and the rest
	 * to Scene::instancedModels, each list keeping this order (Scene.cpp:30-56 folds over both in list order). */
	int32_t n_primitives; const ne_b200_primitive* primitives;
	/* SceneEditor::sortAndGroup (SceneEditor.cpp:2039-2054): move models with a medium to the end of
	 * instancedModels. The editor does this in init() only; 0 = keep the order given. */
	int32_t sort_and_group;
} ne_b200_scene_desc;

/* The vectors Camera::getRayPassingThrough reads (src/core/Camera.cpp:140-144). */
typedef struct ne_b200_camera {
	float position[3], lower_left[3], horizontal[3], vertical[3], side[3], up[3];
	float lens_radius;
} ne_b200_camera;

/* ------------------------------------------------------------------------------------------------------------
 * Host-side helpers restating glm / Camera arithmetic (pure host code, no GPU needed).
 * ---------------------------------------------------------------------------------------------------------- */

/* getTransform(T, R_deg, S) = translate * eulerAngleXYZ(radians) * scale (src/utils/Math.h:848-859) and its
 * glm::inverse (InstancedModel.cpp:11-22). Column-major. */
int ne_b200_make_transform(const float position[3], const float rotation_deg[3], const float scale[3],
                           float to_world[16], float to_object[16]);

/* Camera::Camera(lookFrom, lookAt, up, vfov, aspect, aperture, focus), src/core/Camera.cpp:7-26.
 * SceneReader passes up=(0,1,0), aperture=1e-4 and focus=3 for autoFocus (Q26, SceneReader.cpp:660-668). */
int ne_b200_camera_make(const float look_from[3], const float look_at[3], const float up[3], float vfov_deg,
                        float aspect, float aperture, float focus_distance, ne_b200_camera* out);

/* Host-side scene builders, exposed so a front end (and the CPU test-suite) can inspect exactly what
 * ne_b200_scene_upload will put in HBM. Pure host code. Two-call pattern: pass NULL output arrays to get the sizes.
 *
 * Brick-sparse grid built from a ne_b200_volume (replaces the dense Texture that ResourceManager::loadVolasTexture /
 * loadVDBasTexture hand to GridMedia, ResourceManager.cpp:165-286): dims = {bx, by, bz, n_slots}; table[bz][by][bx] =
 * slot or -1; inv_majorant[bz][by][bx] = 1 / max voxel over the brick's trilinear support [8b-1, 8b+8]^3 (0 = empty);
 * pool = n_slots records of 9x9x9 voxels [8b, 8b+8]^3 (apron layout, x fastest); *max_density =
 * GridMedia::calculateMaxDensity (GridMedia.h:16-21). */
int ne_b200_host_build_bricks(const ne_b200_volume* volume, int32_t dims[4], int32_t* table, float* inv_majorant,
                              float* pool, float* max_density);
/* Binned-SAH 2-wide BVH over a triangle soup (replaces BVH::init, src/primitives/BVH.cpp:6-106): counts =
 * {n_nodes, n_triangle_slots}; nodes = n_nodes records of 16 x 4 bytes {lo0[3], hi0[3], lo1[3], hi1[3], child0,
 * child1, 0, 0} (child >= 0: node index; child < 0: leaf, ~child = (first_slot << 3) | count);
 * triangles = n_triangle_slots records of 12 floats {v0.xyz, bits(original triangle index), v1.xyz, 0, v2.xyz, 0}. */
int ne_b200_host_build_bvh(const float* positions, int32_t n_vertices, const uint32_t* indices, int32_t n_triangles,
                           int32_t counts[2], void* nodes, float* triangles);

/* ------------------------------------------------------------------------------------------------------------
 * Scene front end and framebuffer consumers (pure host code): what SceneReader / ResourceManager / saveImage do.
 * ---------------------------------------------------------------------------------------------------------- */

/* SceneSettings as SceneReader::processCameraAndRenderer fills it (src/io/SceneReader.cpp:669-674, core/Settings.h). */
typedef struct ne_b200_render_settings {
	int32_t width, height;        /* renderer.resolution */
	int32_t spp, bounces;
	int32_t hdr;                  /* stored, unused by the offline path (OfflineEngine.cpp:39-52) */
} ne_b200_render_settings;

/* A scene file loaded into a ne_b200_scene_desc: replaces SceneReader::loadScene (src/io/SceneReader.cpp:10-53) with
 * processMaterial :67-222, processPrimitives :224-648 (obj meshes through this library's own OBJ reader: one vertex
 * per face corner, fan triangulation, V flipped, like assimp with Triangulate|FlipUVs, ResourceManager.cpp:59),
 * processCameraAndRenderer :650-675, ResourceManager::loadVolasTexture (.vol, ResourceManager.cpp:222-286) and
 * loadTexture (PNG -> RGBA8, mirror wrap, :288-315). `resources_dir` is the reference's RESOURCES_DIR: asset paths in
 * the JSON are relative to it. The object owns every array its descriptor points to.
 * `gltf` primitives: glTF 2.0 .gltf (external or base64 buffers) and .glb, triangle primitives of the node tree in
 * depth-first order without node transforms (what Model::processNode keeps, Model.cpp:323-335).
 * Not covered: .vdb (needs OpenVDB: pass leaf bricks through ne_b200_volume) -> error. */
typedef struct ne_b200_scene_file ne_b200_scene_file;
int ne_b200_scene_file_load(const char* json_path, const char* resources_dir, ne_b200_scene_file** out);
int ne_b200_scene_file_parse(const char* json_text, const char* resources_dir, ne_b200_scene_file** out);
const ne_b200_scene_desc* ne_b200_scene_file_desc(const ne_b200_scene_file* file);
/* Camera(position, lookAt, (0,1,0), vfov, W/H, 1e-4, focus): SceneReader ignores the JSON's `up` and `aperture`, and
 * autoFocus yields 3 (Q26, SceneReader.cpp:660-668). */
int ne_b200_scene_file_camera(const ne_b200_scene_file* file, ne_b200_camera* out);
int ne_b200_scene_file_settings(const ne_b200_scene_file* file, ne_b200_render_settings* out);
void ne_b200_scene_file_free(ne_b200_scene_file* file);

/* `.vol` density grids, the text format of ResourceManager::loadVolasTexture (ResourceManager.cpp:222-286): "W H D \n",
 * one discarded line, then densities x fastest, every token followed by one space. Two-call pattern: voxels = NULL
 * returns dims only. The writer emits exactly what that parser captures. */
int ne_b200_vol_read(const char* path, int32_t dims[3], float* voxels);
int ne_b200_vol_write(const char* path, const int32_t dims[3], const float* voxels);

/* stbi_load(path, ..., STBI_rgb_alpha) for PNG files (ResourceManager::loadTexture, ResourceManager.cpp:288-315):
 * dims = {width, height}; rgba = NULL returns dims only. */
int ne_b200_image_read_png(const char* path, int32_t dims[2], uint8_t* rgba);

/* saveImage(pixels, W, H, RGB32F, PNG | EXR, path), src/materials/Texture.h:44-75: PNG = clamp, (uint8_t)(v*255),
 * 3 channels; EXR = three float32 channels (tinyexr SaveEXR(..., 3, 0, ...)), here written uncompressed.
 * rgb = width*height*3 floats, row-major, row 0 first (the layout of OfflineEngine::pixels). */
int ne_b200_image_write_png(const char* path, int width, int height, const float* rgb);
int ne_b200_image_write_exr(const char* path, int width, int height, const float* rgb);
/* OfflineEngine::coreLoop's output.ppm (OfflineEngine.cpp:82,119-138): P6, maxval 65535, big-endian 16-bit samples,
 * pixels written last to first. rgb = the tone-mapped frame. */
int ne_b200_image_write_ppm(const char* path, int width, int height, const float* rgb);

/* ------------------------------------------------------------------------------------------------------------
 * Context, scene, render (replaces OfflineEngine, src/core/OfflineEngine.cpp).
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct ne_b200_ctx ne_b200_ctx;

const char* ne_b200_last_error(void);           /* thread-local, never NULL */
int ne_b200_device_count(void);                 /* number of CUDA devices visible, 0 if none/driver missing */

int ne_b200_create(int cuda_device, ne_b200_ctx** out);
void ne_b200_destroy(ne_b200_ctx* ctx);

/* Run the context's kernels and copies on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the
 * legacy default stream) instead of the private stream made by ne_b200_create, so that a host that already owns a
 * stream (an NCCL communicator, a GL interop queue, a profiler's event pair) orders against the render without
 * extra synchronisation. The stream must outlive the context or be replaced before it is destroyed. */
int ne_b200_set_stream(ne_b200_ctx* ctx, void* cuda_stream);

/* Flatten + copy the scene into HBM: SoA primitive/instance tables in reference fold order, SoA triangles +
 * binned-SAH BVH, brick-sparse density grids with per-brick majorants, emitter table. Replaces the Scene*
 * handed to OfflineEngine's ctor. May be called again to replace the scene. */
int ne_b200_scene_upload(ne_b200_ctx* ctx, const ne_b200_scene_desc* scene);

int ne_b200_camera_set(ne_b200_ctx* ctx, const ne_b200_camera* camera);

/* ne_b200_render flags */
#define NE_B200_RENDER_GLOBAL_MAJORANT 1u   /* track with the reference's single global majorant (GridMedia.cpp:56,82) instead of per-brick majorants */
#define NE_B200_RENDER_MEGAKERNEL      2u   /* one thread per path, no queues (debug / A-B check of the wavefront) */

/*
 * Render samples [spp_begin, spp_end) of every pixel of a width x height frame and ADD their radiance into the
 * context's fp32 linear accumulation buffer (OfflineEngine::renderTile's two hot loops, OfflineEngine.cpp:61-69,
 * for the whole frame). Philox4x32-10 keyed (seed, pixel, sample): any sample range can be rendered on any
 * GPU in any order and the sum is the same up to fp32 addition order. ASYNCHRONOUS: the whole render is enqueued on the
 * context's stream as one CUDA graph (its loop condition is evaluated on the device) and the call returns at once, so one
 * host thread can keep several GPUs busy; ne_b200_wait() joins and reports the outcome. (The debug megakernel flag and
 * NE_B200_HOST_LOOP=1 are synchronous.)
 * The buffer is (re)allocated and zeroed when width/height change or after ne_b200_clear().
 */
int ne_b200_render(ne_b200_ctx* ctx, int width, int height, int spp_begin, int spp_end, int bounces,
                   uint64_t seed, uint32_t flags);
int ne_b200_wait(ne_b200_ctx* ctx);
int ne_b200_clear(ne_b200_ctx* ctx);

/* Device pointer of the linear accumulation buffer (width*height*3 floats, Σ radiance, row-major W*y+x, y=0
 * top row) so the caller can ncclReduce it across ranks, and the sample count it holds. */
int ne_b200_accum_buffer(ne_b200_ctx* ctx, void** device_ptr, size_t* n_floats, int* samples_accumulated);
int ne_b200_set_samples_accumulated(ne_b200_ctx* ctx, int samples);

/* Progressive / resumable rendering: the accumulation buffer is a plain sum over sample indices, so
 * a frame can be checkpointed and continued later - in another context, process or GPU - by rendering further sample
 * ranges on top of it. download: sums[W*H*3] (host) and the number of samples they hold; upload: (re)allocates a
 * width x height buffer, fills it with `sums` and sets the sample count. Philox keys make [0,a) + [a,b) = [0,b). */
int ne_b200_accum_download(ne_b200_ctx* ctx, float* sums, int* samples_accumulated);
int ne_b200_accum_upload(ne_b200_ctx* ctx, int width, int height, const float* sums, int samples_accumulated);

/* rgb[W*H*3] = accumulation / samples (what `color / float(spp)` is at OfflineEngine.cpp:70). Host pointers. */
int ne_b200_read_linear(ne_b200_ctx* ctx, float* rgb);
/* rgb[W*H*3] = OfflineEngine::postProcessing(mean) = clamp(pow(1-exp(-0.5c), 1/2.2), 0, 1) (OfflineEngine.cpp:39-52):
 * the contents of OfflineEngine::pixels (glm::vec3[W*H]). Host pointer. */
int ne_b200_read_tonemapped(ne_b200_ctx* ctx, float* rgb);

/* The single call a renderer front end makes per frame: camera from host, render all spp, resolve, copy the
 * tone-mapped frame (and optionally the linear one, may be NULL) to host memory. Synchronous. */
int ne_b200_render_frame(ne_b200_ctx* ctx, const ne_b200_camera* camera, int width, int height, int spp, int bounces,
                         uint64_t seed, uint32_t flags, float* pixels_tonemapped, float* pixels_linear);

/* ------------------------------------------------------------------------------------------------------------
 * Several GPUs of one box behind one handle, in ONE process (SURVEY 8b `ne_b200_create(gpu_ids, n_gpus, ...)`, 8e).
 * NarvalEngine is a single process (SceneEditor::startOffEngine, src/SceneEditor.cpp:590-604): these calls give it the
 * sample-index partition and the exchange step without a launcher or a collective of its own. GPU g of G holds a scene
 * replica and renders samples [g*spp/G, (g+1)*spp/G) of every pixel (one worker thread per device issues its work; the
 * renders run concurrently); ONE kernel on the first device then reads every peer's accumulation buffer over NVLink
 * (peer access; staged copies where unavailable), sums them in rank order, divides by spp and tone-maps.
 * The image is the single-GPU image up to fp32 addition order. gpu_ids may name a device more than once (each entry gets
 * its own context; used by the tests on one-GPU boxes).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct ne_b200_multi ne_b200_multi;
int ne_b200_create_multi(const int* gpu_ids, int n_gpus, ne_b200_multi** out);
void ne_b200_multi_destroy(ne_b200_multi* m);
int ne_b200_multi_count(const ne_b200_multi* m);
/* rank's own context (counters, stream, checkpointing); owned by the multi handle */
ne_b200_ctx* ne_b200_multi_ctx(ne_b200_multi* m, int rank);
/* 1 when the first device reads rank's accumulation buffer in place (peer access or same device), 0 when it is copied first */
int ne_b200_multi_peer_access(const ne_b200_multi* m, int rank);
/* the scene replicated on every GPU (uploads run concurrently) */
int ne_b200_multi_scene_upload(ne_b200_multi* m, const ne_b200_scene_desc* scene);
/* every GPU renders its share of the frame's spp samples per pixel. Asynchronous on every device. */
int ne_b200_multi_render(ne_b200_multi* m, const ne_b200_camera* camera, int width, int height, int spp, int bounces,
                         uint64_t seed, uint32_t flags);
/* joins the renders; fused reduce + resolve on the first device; frames to HOST memory (either pointer may be NULL).
 * Rank 0's accumulation buffer holds the whole frame's sums afterwards (checkpoint with ne_b200_accum_download). */
int ne_b200_multi_resolve(ne_b200_multi* m, float* pixels_tonemapped, float* pixels_linear);
/* ne_b200_multi_render + ne_b200_multi_resolve: the multi-GPU ne_b200_render_frame. Synchronous. */
int ne_b200_multi_render_frame(ne_b200_multi* m, const ne_b200_camera* camera, int width, int height, int spp, int bounces,
                               uint64_t seed, uint32_t flags, float* pixels_tonemapped, float* pixels_linear);

/* Progressive rendering with a stopping rule ("adaptive stopping"): samples are rendered in batches of
 * spp_batch; even and odd batches accumulate separately, and after every pair the GPU estimates the frame's rel-MSE
 * (SURVEY 8d's image metric, eps 1e-4) from the difference of the two half-estimates. Rendering stops when the estimate
 * is <= target_rel_mse with at least spp_min samples in, or at spp_max. The accumulation buffer then holds all samples
 * (checkpointable / resumable as usual). Host pointers may be NULL. Synchronous. */
typedef struct ne_b200_adaptive_result {
	int32_t spp_rendered;
	float rel_mse_estimate;       /* of the frame returned, against the converged image */
	int32_t converged;            /* 1: stopped by the target; 0: stopped by spp_max */
} ne_b200_adaptive_result;
int ne_b200_render_adaptive(ne_b200_ctx* ctx, const ne_b200_camera* camera, int width, int height, int spp_min, int spp_max,
                            int spp_batch, float target_rel_mse, int bounces, uint64_t seed, uint32_t flags,
                            float* pixels_tonemapped, float* pixels_linear, ne_b200_adaptive_result* result);

/* Work counters of everything rendered since the last ne_b200_counters_reset (device counters; they are what
 * bench.py's roofline uses). Byte sizes of the records are exported so the check is reproducible. */
typedef struct ne_b200_counters {
	uint64_t paths;               /* camera paths started */
	uint64_t extend_rays;         /* Scene::intersectScene equivalents with tMin=1e-11 (Li) */
	uint64_t shadow_rays;         /* intersectScene equivalents from visibilityTr / intersectTr (tMin=1e-3) */
	uint64_t delta_steps;         /* GridMedia::sample tracking iterations (density lookups) */
	uint64_t ratio_steps;         /* GridMedia::Tr tracking iterations (density lookups) */
	uint64_t brick_visits;        /* macro-DDA brick entries */
	uint64_t bvh_nodes;           /* BVH node visits */
	uint64_t tri_tests;           /* ray/triangle tests */
	uint64_t prim_tests;          /* analytic primitive / instance tests */
	uint64_t scatter_events;      /* real collisions */
	uint64_t surface_events;      /* surface shading events */
	uint64_t wavefront_iterations;
	uint64_t kernel_launches;     /* CUDA kernels launched by the library */
	double ms_render;             /* device time of ne_b200_render calls (CUDA events around each on the context's stream). The stage
	                                 accounts below are stamps of the GPU's %globaltimer taken by the kernels of the render graph
	                                 (lane 0's clock when two lanes overlap) or, with NE_B200_HOST_LOOP=1, CUDA events */
	double ms_volume_kernel;      /* device time of the volume tracking kernels (delta + ratio tracking) */
	double ms_extend_kernel;      /* ray casting: extend, shadow, transmittance-search */
	double ms_shade_kernel;       /* scatter + surface shading */
	double ms_upload;             /* host wall time of the last ne_b200_scene_upload */
	uint32_t bytes_per_tracking_step;  /* 32 B cell + 4 B brick-table entry + 4 B majorant */
	uint32_t bytes_per_bvh_node;
	uint32_t bytes_per_triangle;
	uint32_t bytes_per_path_record;    /* sizeof the wavefront's path + hit record (one L2 line) */
	double ms_other_kernel;       /* camera-ray generation (with the first intersectScene) + queue bookkeeping */
} ne_b200_counters;
int ne_b200_get_counters(ne_b200_ctx* ctx, ne_b200_counters* out);
int ne_b200_counters_reset(ne_b200_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------------
 * Test hooks: single device functions on arrays of inputs, optionally driven by an explicit uniform TAPE
 * (the sequence narvalengine::random() would return) so results can be compared with the reference
 * draw for draw. All pointers are HOST pointers; each call launches kernels on the context's GPU.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct ne_b200_hit {      /* RayIntersection, src/primitives/Ray.h:8-15 */
	float hit_point[3], normal[3], uv[2];
	float t_near, t_far;
	int32_t hit;                  /* bool result of Scene::intersectScene */
	int32_t instance;             /* index in fold order: instancedModels..., then lights... ; -1 = none */
	int32_t is_light;
	int32_t primitive;            /* triangle index inside a mesh (as given in `indices`), else 0 */
} ne_b200_hit;

/* Scene::intersectScene(ray, hit, t_min, t_max), src/core/Scene.cpp:30-56 */
int ne_b200_test_intersect(ne_b200_ctx* ctx, int n, const float* origins, const float* directions,
                           float t_min, float t_max, ne_b200_hit* out);

/* Camera::getRayPassingThrough(x,y) with 2 tape uniforms per ray (theta, r), src/core/Camera.cpp:140-144 */
int ne_b200_test_camera_rays(ne_b200_ctx* ctx, int n, const float* xy, const float* tape, float* origins, float* directions);

/* BSDF::eval / BSDF::pdf / (with tape) BSDF::sample at a hit on `instance` (fold order) with the given uv and
 * normal, src/core/BSDF.h:100-142. eval: 3 floats, pdf: 1 float, sampled: 3 floats per item (tape: 2 uniforms each). */
int ne_b200_test_bsdf(ne_b200_ctx* ctx, int n, int instance, const float* incoming, const float* scattered,
                      const float* normals, const float* uvs, const float* tape, float* eval, float* pdf, float* sampled);

/* GridMedia::Tr(ray, isect) and GridMedia::sample(ray, scattered, isect) on `instance` with a tape of
 * `tape_stride` uniforms per item; *_used returns uniforms consumed. src/materials/GridMedia.cpp:45-100.
 * Always global-majorant tracking (tape-exact). */
int ne_b200_test_grid_tr(ne_b200_ctx* ctx, int n, int instance, const float* origins, const float* directions,
                         const float* t_near, const float* t_far, const float* tape, int tape_stride,
                         float* tr, int32_t* used);
int ne_b200_test_grid_sample(ne_b200_ctx* ctx, int n, int instance, const float* origins, const float* directions,
                             const float* t_near, const float* t_far, const float* tape, int tape_stride,
                             float* transmittance, float* scattered_o, float* scattered_d, int32_t* used);

/* VolumetricPathIntegrator::Li(ray, scene) for n rays, each with its own tape (tape_stride uniforms), one GPU
 * thread per ray, reference draw order. src/integrators/VolumetricPathIntegrator.cpp:176-301.
 * This is Seam 1 (Integrator::Li). radiance: 3 floats per ray; used: uniforms consumed (or -1 if the tape ran out). */
int ne_b200_test_li_tape(ne_b200_ctx* ctx, int n, const float* origins, const float* directions, int bounces,
                         const float* tape, int tape_stride, float* radiance, int32_t* used);

/* Same with Philox (seed, path index i, sample 0): the megakernel estimator used for wavefront A/B checks. */
int ne_b200_test_li_philox(ne_b200_ctx* ctx, int n, const float* origins, const float* directions, int bounces,
                           uint64_t seed, uint32_t flags, float* radiance);

/* uniformSampleOneLight(incoming, isect, scene) (VolumetricPathIntegrator.cpp:159-174, incl. estimateDirect
 * :74-157) at explicit hits: incoming ray dir, hit record (point, normal, uv, instance). Tape-driven. */
int ne_b200_test_sample_one_light(ne_b200_ctx* ctx, int n, const float* incoming_dirs, const ne_b200_hit* hits,
                                  const float* tape, int tape_stride, float* radiance, int32_t* used);

/* Trilinear density GridMedia::interpolatedDensity at OCS points of `instance` (GridMedia.cpp:25-43) and the
 * grid's invMaxDensity (GridMedia.cpp:12). */
int ne_b200_test_density(ne_b200_ctx* ctx, int n, int instance, const float* ocs_points, float* density, float* inv_max_density);

/* The brick-sparse grid of scene volume `volume` as it lives in HBM, in the format of ne_b200_host_build_bricks (two-call
 * pattern: NULL arrays to get dims = {bx, by, bz, n_slots}). Dense grids are bricked by GPU kernels during
 * ne_b200_scene_upload; this hook lets a test check them against the host builder bit for bit. */
int ne_b200_test_read_bricks(ne_b200_ctx* ctx, int volume, int32_t dims[4], int32_t* table, float* inv_majorant, float* pool,
                             float* max_density);

/* on != 0: ne_b200_test_bsdf, ne_b200_test_sample_one_light and ne_b200_test_li_tape run the medium shading production
 * renders use ("FAST medium shading", csrc/ne_device.cuh: hardware reciprocal / rsqrt / sine / cosine in the phase-function
 * code instead of IEEE divisions; geometry untouched) so that the same reference vectors hold it to the same tolerances
 * (BSDF.h:100-142 through VolumeBSDF / HG / IsotropicPhaseFunction, Medium.h:73-130). The instances / hits given to those
 * hooks must then carry a medium. Renders: on by default, NE_B200_EXACT_SHADING=1 selects the reference-order arithmetic. */
int ne_b200_test_set_fast_shading(ne_b200_ctx* ctx, int on);

/* Philox4x32-10 uniforms: out[i] = u(seed, pixel, sample, dimension i), for KATs of the generator. */
int ne_b200_test_philox(ne_b200_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* NE_B200_H */

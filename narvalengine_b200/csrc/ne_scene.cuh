// ne_scene.cuh — the flattened scene as it lives in HBM (DESIGN.md "Data layout"). Built by ne_host.cpp from the
// ne_b200_scene_desc the caller hands over; read-only for every kernel.
#pragma once
#include "ne_math.cuh"

namespace ne {

enum { PRIM_RECTANGLE = 0, PRIM_SPHERE = 1, PRIM_POINT = 2, PRIM_VOLUME = 3, PRIM_MESH = 4 };
enum { MAT_MICROFACET = 0, MAT_EMITTER = 1, MAT_VOLUME = 2, MAT_DIRECTIONAL = 3, MAT_INFINITE = 4 };
enum { TEX_R32F = 0, TEX_RG32F = 1, TEX_RGB32F = 2, TEX_RGBA32F = 3, TEX_RGBA8 = 4 };
enum { BRICK = 8, CORE_VOX = 512, BRICK_VOX = 729 };  // bricks are stored with a +1 apron: 9x9x9 voxels

struct DTexture {
	int w, h, format, wrap_u, wrap_v;
	const void* texels;
};

// Brick-sparse density grid (GridMedia's Texture, src/materials/GridMedia.h). cells[bz][by][bx] = {slot, invMaj}:
// slot = brick record or -1 (nothing but zeros within reach); invMaj = float bits of 1 / (max voxel over
// [8b-1, 8b+8]^3), the reciprocal per-brick majorant, 0 when the brick has nothing to collide with. One 8-byte
// load per brick entered. pool[slot*729 + 81*z + 9*y + x] holds the 9^3 voxels [8b, 8b+8]^3 (apron layout: the
// trilinear stencil of any cell of the brick is inside its own record).
struct DVolume {
	int W, H, D;
	int bx, by, bz;
	int n_slots;
	const int2* cells;
	const float* pool;
	// per-brick majorants as IEEE halves in a table with a one-brick apron, (bx+2) x (by+2) x (bz+2) entries (layout and
	// encoding: ne_tracking.cuh, BrickTracker): majorant = float(half) * maj_scale, rounded UP (any bound >= the brick's
	// maximum keeps tracking unbiased); negative = empty brick / outside. 2 bytes per brick: 79 KB for a 256^3 grid, copied
	// to shared memory by the tracking kernels when it fits
	const unsigned short* maj16;
	float maj_scale;  // 2^-k
	float max_density, inv_max_density;  // GridMedia::invMaxDensity, GridMedia.cpp:12
};

struct DMaterial {
	int type;
	int albedo_tex, roughness_tex, metallic_tex, normal_tex, has_normal_flag;
	float li[3];
	float sigma_s[3], sigma_a[3];
	float density_mult;
	int phase;  // 0 isotropic, 1 HG
	float g;
	int volume;       // DVolume index or -1 (HomogeneousMedia)
	int has_bsdf;     // Material::bsdf != nullptr
	int transmissive; // bsdf->hasType(BxDF_TRANSMISSION)
	int has_light, has_medium;
	int infinite;      // InfiniteAreaLight (lights/InfiniteAreaLight.h): Le / sampleLi from the lat-long map `env`, Li() = 0
	int env;           // index into DScene::env
	int directional;   // DirectionalLight (lights/DirectionalLight.cpp): Le = Li() = li, sampleLi returns 0 (Q23); never hit
	float direction[3];
	int light_owner;  // fold index of the instance whose primitive Light::primitive points at (Q7: last one built)
};

// 2-wide BVH node, 64 B: both children's boxes + child links. child < 0: leaf, ~child = first triangle slot,
// count in cnt. Triangles are stored reordered as 3 x float4 (v0,v1,v2; .w of v0 = original triangle index bits).
struct __align__(16) BvhNode {
	float lo0[3], hi0[3];
	float lo1[3], hi1[3];
	int child0, child1;
	int cnt0, cnt1;
};

struct DMesh {
	int n_tris, n_verts, n_nodes;
	const BvhNode* nodes;
	const float4* tri;        // 3 float4 per triangle slot (BVH order)
	const float* pos;         // 3 per vertex (importer order), for Q30 uv lookup
	const float* uv;          // 2 per vertex or nullptr
	const uint32_t* idx;      // 3 per triangle (importer order)
	float bbmin[3], bbmax[3]; // Model::boundingBox (aabbMin/aabbMax over vertices, Model.cpp:160-163,368)
	int root_leaf_cnt;        // >0: the whole mesh is one leaf (n_tris <= leaf size)
};

struct DInstance {
	float M[16], Mi[16];  // InstancedModel::transformToWCS / invTransformToWCS
	int type, material, collision, mesh;
	float radius;
	float point[3];
	float scale[3];       // getScale(M), src/utils/Math.h:874-912 (Rectangle::pdf)
	int desc_index;       // index in the caller's primitive array
};

// InfiniteAreaLight's Distribution2D (utils/Sampling.h:69-112) over map.g * sin(theta): h conditional distributions of
// nc = h entries each (Q27: built with n = height, whatever the width) + the marginal over the h rows.
struct DEnvDist {
	int tex;              // the lat-long map (DScene::tex)
	int w, h, nc;
	const float* cFunc;   // [h][nc]
	const float* cCdf;    // [h][nc + 1]
	const float* cInt;    // [h] funcInt of each conditional
	const float* mFunc;   // [h]
	const float* mCdf;    // [h + 1]
	float mInt;
};

struct DScene {
	int n_inst;    // fold order: instancedModels..., lights...
	int n_models;  // first n_models entries are Scene::instancedModels
	int n_lights;  // the rest are Scene::lights
	int n_mat, n_vol;
	const DInstance* inst;
	const DMaterial* mat;
	const DTexture* tex;
	const DVolume* vol;
	const DMesh* mesh;
	const DEnvDist* env;
	int n_infinite;  // light instances with an InfiniteAreaLight material
	int n_directional;  // light instances with a DirectionalLight material (their Le is added to camera rays)
	int has_medium;  // any instance with a medium material (intersectTr can only return true then, Q12)
};

struct DCamera {
	V3 position, lower_left, horizontal, vertical, side, up;
	float lens_radius;
};

// RayIntersection, src/primitives/Ray.h:8-15. inst < 0: nothing.
struct Hit {
	V3 p, n;
	float u, v;
	float tNear, tFar;
	int inst;
	int prim;
};

// Work counters (ne_b200_counters). Accumulated per thread, reduced per warp, one atomic per warp.
struct DCounters {
	unsigned long long paths, extend_rays, shadow_rays, delta_steps, ratio_steps, brick_visits, bvh_nodes, tri_tests, prim_tests,
		scatter_events, surface_events;
	// folded in by the wavefront's last kernel (the render graph runs without the host): device time per stage kind
	// (%globaltimer stamps between stages), iterations, kernel launches, request-array overflow flag
	unsigned long long stage_ns[4], iterations, launches, overflow;
};

}  // namespace ne

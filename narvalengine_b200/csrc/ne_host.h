// ne_host.h — host-side scene preparation (no CUDA calls): glm-order transform/camera arithmetic, binned-SAH BVH
// build, brick-sparse grid build. Implemented in ne_host.cpp.
#pragma once
#include <cstdint>
#include <vector>

#include "ne_b200.h"
#include "ne_scene.cuh"

namespace ne {

// getTransform / glm::inverse / getScale / Camera::Camera restated in glm's operation order.
void host_make_transform(const float pos[3], const float rotDeg[3], const float scale[3], float M[16], float Mi[16]);
void host_get_scale(const float M[16], float s[3]);
void host_camera_make(const float from[3], const float at[3], const float up[3], float vfov, float aspect, float aperture, float focus,
                      ne_b200_camera* out);

struct HostBvh {
	std::vector<BvhNode> nodes;
	std::vector<float> tri;  // 12 floats per slot: v0.xyz, bits(orig index), v1.xyz, 0, v2.xyz, 0
	float bbmin[3], bbmax[3];
};
// Binned-SAH (16 bins) 2-wide BVH, <= 4 triangles per leaf, pre-order node array (root = node 0).
void host_build_bvh(const float* positions, int nVerts, const uint32_t* indices, int nTris, HostBvh& out);

struct HostBricks {
	int W, H, D, bx, by, bz;
	std::vector<int32_t> table;
	std::vector<float> pool;
	std::vector<float> bmaj;  // per-brick majorant density
	std::vector<float> binv;  // its reciprocal (0 = empty)
	float maxDensity;
};
// From a dense W*H*D grid or from OpenVDB-style 8^3 leaves (see ne_b200_volume).
void host_build_bricks(const ne_b200_volume& v, HostBricks& out);

}  // namespace ne

// flatten.h — interface of the reference-side adapter (oracle/ref/flatten.cpp): NarvalEngine's Scene* / Camera -> the C
// ABI's PODs. Compiled against the reference headers; see flatten.cpp.
#pragma once
#include <cstdint>
#include <vector>

#include "ne_b200.h"

namespace narvalengine {
class Scene;
class Camera;
}

namespace narval_b200_adapter {

// The descriptor plus the arrays it points at that do not already live in the engine (mesh position / uv / index arrays
// rebuilt from the Triangle objects). Texture and grid pointers alias the engine's own buffers.
struct FlatScene {
	std::vector<ne_b200_texture> textures;
	std::vector<ne_b200_volume> volumes;
	std::vector<ne_b200_material> materials;
	std::vector<ne_b200_primitive> primitives;
	std::vector<std::vector<float>> meshPositions, meshUvs;
	std::vector<std::vector<uint32_t>> meshIndices;
	ne_b200_scene_desc desc{};
};

void flatten(narvalengine::Scene* scene, FlatScene& out);
ne_b200_camera toPod(const narvalengine::Camera& c);

}  // namespace narval_b200_adapter

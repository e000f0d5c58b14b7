// flatten.cpp — THE REFERENCE-SIDE ADAPTER, compiled against the reference's own headers (test infrastructure here; in
// NarvalEngine this file is what a maintainer adds next to include/ne_b200_offline_engine.hpp).
//
//   narval_b200_adapter::flatten(narvalengine::Scene*)  ->  ne_b200_scene_desc   (the object graph SceneReader built,
//                                                            walked in fold order: Scene::instancedModels, Scene::lights)
//   narval_b200_adapter::toPod(const narvalengine::Camera&) -> ne_b200_camera     (the seven vectors getRayPassingThrough reads)
//
// Seam 2 of SURVEY 8b: OfflineEngine(Camera, SceneSettings, Scene*) (src/core/OfflineEngine.h:23-47) hands over the
// engine's own Camera and Scene*; these two functions turn them into the C ABI's PODs. What is read, per type:
//   InstancedModel (primitives/InstancedModel.h:11-33)  transformToWCS / invTransformToWCS (glm::mat4, column-major like
//                                                       the ABI), isCollisionEnabled, model
//   Model (primitives/Model.h:65-83)                    primitives / lights (one analytic primitive for JSON primitives;
//                                                       Triangle[] over the stride-11 vertex buffer for assimp meshes,
//                                                       Model.h:56-62), materials
//   Rectangle / Sphere / Point / AABB                   recognised by dynamic_cast; Sphere::radius; Point vertex
//   Triangle (primitives/Triangle.h:14)                 vertexData[3] point INTO the model's vertex buffer: vertex index =
//                                                       (pointer - lowest pointer) / stride, so the index buffer and the
//                                                       vertex order come back exactly (Q30 depends on them)
//   Material (materials/Material.h:9-69)                bsdf->bxdf[0] (GlossyBSDF / VolumeBSDF + its PhaseFunction), light
//                                                       (DiffuseLight::li, DirectionalLight::le/direction, InfiniteAreaLight::tex),
//                                                       medium (GridMedia: scattering, absorption, densityMultiplier, grid;
//                                                       HomogeneousMedia likewise without a grid), textures[ctz(name)]
//   Texture (materials/Texture.h:77-89)                 width, height, depth, texFormat, samplerFlags, mem
// Pointers in the descriptor alias the engine's own buffers (textures, grids) or vectors owned by FlatScene (mesh arrays):
// the library copies everything to HBM during ne_b200_scene_upload, so FlatScene may be dropped right after.
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "core/BSDF.h"
#include "core/Camera.h"
#include "core/GlossyBSDF.h"
#include "core/Scene.h"
#include "core/VolumeBSDF.h"
#include "primitives/AABB.h"
#include "primitives/Point.h"
#include "primitives/Rectangle.h"
#include "primitives/Sphere.h"
#include "primitives/Triangle.h"
#include "lights/DiffuseLight.h"
#include "lights/DirectionalLight.h"
#include "lights/InfiniteAreaLight.h"  // uses Sphere without including it
#include "materials/GridMedia.h"
#include "materials/HomogeneousMedia.h"

#include "flatten.h"

namespace narval_b200_adapter {
using namespace narvalengine;

namespace {

int formatOf(TextureLayout f) {
	switch (f) {
	case R32F: return NE_B200_TEX_R32F;
	case RG32F: return NE_B200_TEX_RG32F;
	case RGB32F: return NE_B200_TEX_RGB32F;
	case RGBA32F: return NE_B200_TEX_RGBA32F;
	default: return NE_B200_TEX_RGBA8;
	}
}

struct Flattener {
	FlatScene& out;
	std::map<const Texture*, int> texIndex, volIndex;
	std::map<const Material*, int> matIndex;
	explicit Flattener(FlatScene& o) : out(o) {}

	int texture(const Texture* t) {
		if (!t) return -1;
		auto it = texIndex.find(t);
		if (it != texIndex.end()) return it->second;
		ne_b200_texture d{};
		d.width = int(t->width);
		d.height = int(t->height);
		d.format = formatOf(t->texFormat);
		// Texture::wrapTextureCoordinates (materials/Texture.cpp:115-129): MIRROR per axis, anything else clamps
		d.wrap_u = (t->samplerFlags & NE_TEX_SAMPLER_U_MASK) == NE_TEX_SAMPLER_U_MIRROR ? NE_B200_WRAP_MIRROR : NE_B200_WRAP_CLAMP;
		d.wrap_v = (t->samplerFlags & NE_TEX_SAMPLER_V_MASK) == NE_TEX_SAMPLER_V_MIRROR ? NE_B200_WRAP_MIRROR : NE_B200_WRAP_CLAMP;
		d.texels = t->mem.data;
		out.textures.push_back(d);
		return texIndex[t] = int(out.textures.size()) - 1;
	}
	int volume(const Texture* grid) {
		auto it = volIndex.find(grid);
		if (it != volIndex.end()) return it->second;
		ne_b200_volume v{};
		v.width = int(grid->width);
		v.height = int(grid->height);
		v.depth = int(grid->depth);
		v.dense = static_cast<const float*>(grid->mem.data);  // Texture(W,H,D,R32F): x fastest, then y, then z (GridMedia.h:33-36)
		out.volumes.push_back(v);
		return volIndex[grid] = int(out.volumes.size()) - 1;
	}
	int material(Material* m) {
		if (!m) return -1;
		auto it = matIndex.find(m);
		if (it != matIndex.end()) return it->second;
		ne_b200_material d{};
		d.albedo_tex = d.roughness_tex = d.metallic_tex = d.normal_tex = d.env_tex = -1;
		d.volume = -1;
		if (m->light) {
			if (auto* inf = dynamic_cast<InfiniteAreaLight*>(m->light)) {
				d.type = NE_B200_MAT_INFINITE;
				d.env_tex = texture(inf->tex);
			} else if (auto* dir = dynamic_cast<DirectionalLight*>(m->light)) {
				d.type = NE_B200_MAT_DIRECTIONAL;
				for (int k = 0; k < 3; k++) { d.li[k] = dir->le[k]; d.direction[k] = dir->direction[k]; }
			} else {
				d.type = NE_B200_MAT_EMITTER;
				for (int k = 0; k < 3; k++) d.li[k] = m->light->li[k];
			}
		} else if (m->medium) {
			d.type = NE_B200_MAT_VOLUME;
			if (auto* g = dynamic_cast<GridMedia*>(m->medium)) {
				for (int k = 0; k < 3; k++) { d.scattering[k] = g->scattering[k]; d.absorption[k] = g->absorption[k]; }
				d.density_multiplier = g->densityMultiplier;
				d.volume = volume(g->grid);
			} else if (auto* h = dynamic_cast<HomogeneousMedia*>(m->medium)) {
				for (int k = 0; k < 3; k++) { d.scattering[k] = h->scattering[k]; d.absorption[k] = h->absorption[k]; }
				d.density_multiplier = h->density;
			}
			d.phase = NE_B200_PHASE_ISOTROPIC;
			if (m->bsdf && m->bsdf->bxdf[0])
				if (auto* vb = dynamic_cast<VolumeBSDF*>(m->bsdf->bxdf[0]))
					if (auto* hg = dynamic_cast<HG*>(vb->phaseFunction)) { d.phase = NE_B200_PHASE_HG; d.g = hg->g; }
		} else {
			d.type = NE_B200_MAT_MICROFACET;  // GlossyBSDF over GGX + Schlick (SceneReader.cpp:131-138)
			d.albedo_tex = m->hasTexture(ALBEDO) || m->textures[ctz(ALBEDO)] ? texture(m->textures[ctz(ALBEDO)]) : -1;
			d.metallic_tex = m->textures[ctz(METALLIC)] ? texture(m->textures[ctz(METALLIC)]) : -1;
			d.roughness_tex = m->textures[ctz(ROUGHNESS)] ? texture(m->textures[ctz(ROUGHNESS)]) : -1;
			d.normal_tex = m->textures[ctz(NORMAL)] ? texture(m->textures[ctz(NORMAL)]) : -1;
			// Material::hasTexture tests `textureTypes`, which addTexture OVERWRITES with the last name added (Q24): a normal map
			// bends normals only when NORMAL was the last addTexture call
			d.has_normal_flag = (d.normal_tex >= 0 && m->hasTexture(NORMAL)) ? 1 : 0;
		}
		out.materials.push_back(d);
		return matIndex[m] = int(out.materials.size()) - 1;
	}

	void instance(InstancedModel* im, bool inLights) {
		ne_b200_primitive p{};
		std::memcpy(p.to_world, &im->transformToWCS[0][0], 64);   // glm::mat4 is column-major, as the ABI expects
		std::memcpy(p.to_object, &im->invTransformToWCS[0][0], 64);
		p.collision = im->isCollisionEnabled ? 1 : 0;
		Model* m = im->model;
		const std::vector<Primitive*>& prims = inLights ? m->lights : m->primitives;
		Primitive* first = prims.empty() ? nullptr : prims[0];
		Material* mat = !m->materials.empty() ? m->materials[0] : (first ? first->material : nullptr);
		p.material = material(mat);
		if (first && dynamic_cast<Triangle*>(first)) {
			// an assimp mesh: Triangle::vertexData[k] point into the model's interleaved vertex buffer (stride = model->stride floats:
			// position 3, normal 3, tangent 3, uv 2; Model.h:56-62)
			p.type = NE_B200_PRIM_MESH;
			const int stride = m->stride;
			const float* base = nullptr;
			const float* top = nullptr;
			for (Primitive* q : prims) {
				Triangle* t = static_cast<Triangle*>(q);
				for (int k = 0; k < 3; k++) {
					if (!base || t->vertexData[k] < base) base = t->vertexData[k];
					if (!top || t->vertexData[k] > top) top = t->vertexData[k];
				}
			}
			const int nVerts = int((top - base) / stride) + 1;
			out.meshPositions.emplace_back(size_t(nVerts) * 3);
			out.meshUvs.emplace_back(size_t(nVerts) * 2);
			out.meshIndices.emplace_back(prims.size() * 3);
			std::vector<float>& P = out.meshPositions.back();
			std::vector<float>& U = out.meshUvs.back();
			std::vector<uint32_t>& I = out.meshIndices.back();
			for (int v = 0; v < nVerts; v++) {
				const float* src = base + size_t(v) * stride;
				P[3 * v] = src[0]; P[3 * v + 1] = src[1]; P[3 * v + 2] = src[2];
				U[2 * v] = src[9]; U[2 * v + 1] = src[10];
			}
			for (size_t t = 0; t < prims.size(); t++)
				for (int k = 0; k < 3; k++) I[3 * t + k] = uint32_t((static_cast<Triangle*>(prims[t])->vertexData[k] - base) / stride);
			p.n_vertices = nVerts;
			p.n_triangles = int(prims.size());
			p.positions = P.data();
			p.uvs = U.data();
			p.indices = I.data();
		} else if (first && dynamic_cast<Rectangle*>(first)) {
			p.type = NE_B200_PRIM_RECTANGLE;
		} else if (auto* s = dynamic_cast<Sphere*>(first)) {
			p.type = NE_B200_PRIM_SPHERE;
			p.radius = s->radius;
		} else if (auto* q = dynamic_cast<Point*>(first)) {
			p.type = NE_B200_PRIM_POINT;
			std::memcpy(p.point, q->vertexData[0], 12);
		} else {
			p.type = NE_B200_PRIM_VOLUME;  // the AABB proxy of a medium (SceneReader.cpp:541-610)
		}
		out.primitives.push_back(p);
	}
};

}  // namespace

void flatten(Scene* scene, FlatScene& out) {
	out = FlatScene();
	Flattener f(out);
	// fold order matters (Scene::intersectScene, core/Scene.cpp:30-56): models first, then lights, each in list order. The
	// library re-derives the same split from the material types, so the order WITHIN each list is what must be kept.
	for (InstancedModel* im : scene->instancedModels) f.instance(im, false);
	for (InstancedModel* im : scene->lights) f.instance(im, true);
	out.desc.n_textures = int(out.textures.size());
	out.desc.textures = out.textures.data();
	out.desc.n_volumes = int(out.volumes.size());
	out.desc.volumes = out.volumes.data();
	out.desc.n_materials = int(out.materials.size());
	out.desc.materials = out.materials.data();
	out.desc.n_primitives = int(out.primitives.size());
	out.desc.primitives = out.primitives.data();
	out.desc.sort_and_group = 0;  // the lists are already in the order the editor left them (SceneEditor::sortAndGroup ran or not)
}

ne_b200_camera toPod(const Camera& c) {
	ne_b200_camera o{};
	auto put = [](float* d, const glm::vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
	put(o.position, c.position);
	put(o.lower_left, c.lowerLeft);
	put(o.horizontal, c.horizontal);
	put(o.vertical, c.vertical);
	put(o.side, c.side);
	put(o.up, c.up);
	o.lens_radius = c.lensRadius;
	return o;
}

}  // namespace narval_b200_adapter

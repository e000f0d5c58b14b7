// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by or shipped with the product library.
//
// C-ABI harness around the UNMODIFIED reference translation units (compiled in place from /root/reference/src by
// oracle/Makefile into oracle/_ref/). It builds the reference's own Scene/Model/Material/Camera objects from the
// same POD scene description the product takes (include/ne_b200.h), following the recipe of
// src/io/SceneReader.cpp:67-675 (that TU itself needs rapidjson and cannot be compiled here), and exposes the
// hot-path functions so tests/ and bench.py's cpu_baseline can call them with a seeded narvalengine::mt.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <sstream>
#include <iostream>
#include <fstream>
#include <stack>
#include <queue>

// GridMedia::interpolatedDensity is private; access specifiers do not change object layout with g++.
#define private public
#include "materials/GridMedia.h"
#undef private
#include "core/Scene.h"
#include "core/Camera.h"
#include "core/OfflineEngine.h"
#include "core/GlossyBSDF.h"
#include "core/VolumeBSDF.h"
#include "core/Microfacet.h"
#include "core/ResourceManager.h"
#include "integrators/VolumetricPathIntegrator.h"
#include "lights/DiffuseLight.h"
#include "lights/DirectionalLight.h"
#include "primitives/Rectangle.h"
#include "primitives/Sphere.h"
#include "primitives/Point.h"
#include "flatten.h"
#include "primitives/AABB.h"
#include "primitives/Model.h"
#include "primitives/InstancedModel.h"
#include "materials/HomogeneousMedia.h"
#include "lights/InfiniteAreaLight.h"  // uses Sphere without including it

#include "ne_b200.h"

using namespace narvalengine;

// ---------------------------------------------------------------------------------------------------------------
// ResourceManager shim (core/ResourceManager.cpp needs OpenVDB + assimp and is not compiled). Name -> object maps.
// ---------------------------------------------------------------------------------------------------------------
namespace narvalengine {
ResourceManager* ResourceManager::self = nullptr;
ResourceManager::ResourceManager() {}
ResourceManager::~ResourceManager() {}
ResourceManager* ResourceManager::getSelf() {
	if (!self) self = new ResourceManager();
	return self;
}
StringID ResourceManager::replaceMaterial(std::string name, Material* m) {
	StringID id = genStringID(name.c_str());
	materials[id] = m;
	return id;
}
StringID ResourceManager::setMaterial(std::string name, Material* m) { return replaceMaterial(name, m); }
Material* ResourceManager::getMaterial(StringID id) { return materials.count(id) ? materials[id] : nullptr; }
Material* ResourceManager::getMaterial(std::string name) { return getMaterial(genStringID(name.c_str())); }
StringID ResourceManager::setTexture(std::string name, Texture* t) {
	StringID id = genStringID(name.c_str());
	textures[id] = t;
	return id;
}
StringID ResourceManager::replaceTexture(std::string name, Texture* t) { return setTexture(name, t); }
Texture* ResourceManager::getTexture(StringID id) { return textures.count(id) ? textures[id] : nullptr; }
Texture* ResourceManager::getTexture(std::string name) { return getTexture(genStringID(name.c_str())); }
StringID ResourceManager::loadTexture(std::string name, std::string, bool) { return genStringID(name.c_str()); }
StringID ResourceManager::replaceModel(std::string name, Model* m) {
	StringID id = genStringID(name.c_str());
	models[id] = m;
	return id;
}
StringID ResourceManager::setModel(std::string name, Model* m) { return replaceModel(name, m); }
Model* ResourceManager::getModel(StringID id) { return models.count(id) ? models[id] : nullptr; }
Model* ResourceManager::getModel(std::string name) { return getModel(genStringID(name.c_str())); }
}

namespace {

struct RefScene {
	Scene* scene = nullptr;
	std::vector<Material*> materials;
	std::vector<Texture*> textures;     // 2-D material textures
	std::vector<Texture*> volumes;      // 3-D density textures
	std::vector<InstancedModel*> fold;  // instancedModels..., lights...
	int uid = 0;
};

std::atomic<int> g_scene_uid{0};

glm::mat4 toMat4(const float* m) {
	glm::mat4 r;
	for (int c = 0; c < 4; c++)
		for (int k = 0; k < 4; k++) r[c][k] = m[c * 4 + k];
	return r;
}
void fromMat4(const glm::mat4& m, float* out) {
	for (int c = 0; c < 4; c++)
		for (int k = 0; k < 4; k++) out[c * 4 + k] = m[c][k];
}
glm::vec3 v3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
void put3(float* o, glm::vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

TextureLayout layoutOf(int f) {
	switch (f) {
	case NE_B200_TEX_R32F: return R32F;
	case NE_B200_TEX_RG32F: return RG32F;
	case NE_B200_TEX_RGB32F: return RGB32F;
	case NE_B200_TEX_RGBA32F: return RGBA32F;
	default: return RGBA8;
	}
}
uint32_t bytesPerTexel(int f) {
	switch (f) {
	case NE_B200_TEX_R32F: return 4;
	case NE_B200_TEX_RG32F: return 8;
	case NE_B200_TEX_RGB32F: return 12;
	case NE_B200_TEX_RGBA32F: return 16;
	default: return 4;
	}
}

// Densify a leaf-brick volume the way tools::copyToDense leaves it (ResourceManager.cpp:186-214): background 0.
std::vector<float> densify(const ne_b200_volume& v) {
	size_t n = size_t(v.width) * v.height * v.depth;
	std::vector<float> g(n, 0.0f);
	for (int l = 0; l < v.n_leaves; l++) {
		const int* o = v.leaf_origin + 3 * l;
		const float* val = v.leaf_values + 512 * size_t(l);
		for (int z = 0; z < 8; z++)
			for (int y = 0; y < 8; y++)
				for (int x = 0; x < 8; x++) {
					int X = o[0] + x, Y = o[1] + y, Z = o[2] + z;
					if (X < 0 || Y < 0 || Z < 0 || X >= v.width || Y >= v.height || Z >= v.depth) continue;
					g[size_t(v.width) * v.height * Z + size_t(v.width) * Y + X] = val[64 * z + 8 * y + x];
				}
	}
	return g;
}

// SceneReader::processMaterial, src/io/SceneReader.cpp:67-222
Material* buildMaterial(RefScene& rs, const ne_b200_material& m, const std::string& name) {
	Material* mat = new Material();
	if (m.type == NE_B200_MAT_MICROFACET) {
		// addTexture order of SceneReader.cpp:126-130: ALBEDO, METALLIC, ROUGHNESS, [NORMAL]
		if (m.albedo_tex >= 0) { rs.textures[m.albedo_tex]->textureName = ALBEDO; mat->addTexture(ALBEDO, rs.textures[m.albedo_tex]); }
		if (m.metallic_tex >= 0) { rs.textures[m.metallic_tex]->textureName = METALLIC; mat->addTexture(METALLIC, rs.textures[m.metallic_tex]); }
		if (m.roughness_tex >= 0) { rs.textures[m.roughness_tex]->textureName = ROUGHNESS; mat->addTexture(ROUGHNESS, rs.textures[m.roughness_tex]); }
		if (m.normal_tex >= 0) { rs.textures[m.normal_tex]->textureName = NORMAL; mat->addTexture(NORMAL, rs.textures[m.normal_tex]); }
		GGXDistribution* ggxD = new GGXDistribution();
		ggxD->alpha = 0.25f;  // overwritten on every BSDF call (GlossyBSDF.cpp:12,19,26)
		GlossyBSDF* glossy = new GlossyBSDF(ggxD, new FresnelSchilck());
		mat->bsdf = new BSDF();
		mat->bsdf->addBxdf(glossy);
	} else if (m.type == NE_B200_MAT_EMITTER) {
		DiffuseLight* l = new DiffuseLight();
		l->li = v3(m.li);
		mat->light = l;
	} else if (m.type == NE_B200_MAT_INFINITE) {
		// SceneReader.cpp:169-186
		Texture* tex = rs.textures[m.env_tex];
		tex->textureName = TEX_1;
		InfiniteAreaLight* l = new InfiniteAreaLight(tex);
		mat->light = l;
		mat->addTexture(TEX_1, tex);
	} else if (m.type == NE_B200_MAT_DIRECTIONAL) {
		// SceneReader.cpp:156-168: only `le` and `direction` are set (li stays 0, Q23)
		DirectionalLight* l = new DirectionalLight();
		l->le = v3(m.li);
		l->direction = v3(m.direction);
		mat->light = l;
	} else if (m.type == NE_B200_MAT_VOLUME) {
		PhaseFunction* pf = m.phase == NE_B200_PHASE_HG ? (PhaseFunction*)new HG(m.g) : (PhaseFunction*)new IsotropicPhaseFunction();
		Medium* medium;
		if (m.volume >= 0)
			medium = new GridMedia(v3(m.scattering), v3(m.absorption), rs.volumes[m.volume], m.density_multiplier);
		else
			medium = new HomogeneousMedia(v3(m.scattering), v3(m.absorption), m.density_multiplier);
		mat->medium = medium;
		mat->bsdf = new BSDF();
		mat->bsdf->addBxdf(new VolumeBSDF(pf));
		if (m.volume >= 0) mat->addTexture(TEX_1, rs.volumes[m.volume]);
	} else {
		delete mat;
		return nullptr;
	}
	ResourceManager::getSelf()->replaceMaterial(name, mat);
	return mat;
}

// SceneReader::processPrimitives, src/io/SceneReader.cpp:224-648 (geometry literals per :407-468 and :541-562).
InstancedModel* buildPrimitive(RefScene& rs, const ne_b200_primitive& p, const std::string& name, const std::string& matName, bool& isLight) {
	Material* material = p.material >= 0 ? rs.materials[p.material] : nullptr;
	isLight = material && material->light != nullptr;
	std::vector<Primitive*> primitives, lights;
	std::vector<Material*> materials;
	std::vector<Mesh> meshes;
	VertexLayout vertexLayout;
	Model* model = nullptr;

	if (p.type == NE_B200_PRIM_MESH) {
		// Hand-filled aiScene so that Model(const aiScene*, path, materialName) and BVH::init run unchanged.
		aiScene* sc = new aiScene();
		aiMesh* mesh = new aiMesh();
		mesh->mNumVertices = p.n_vertices;
		mesh->mNumFaces = p.n_triangles;
		mesh->mVertices = new aiVector3D[p.n_vertices];
		mesh->mNormals = new aiVector3D[p.n_vertices];
		mesh->mTangents = new aiVector3D[p.n_vertices];
		mesh->mTextureCoords[0] = new aiVector3D[p.n_vertices];
		for (int i = 0; i < p.n_vertices; i++) {
			mesh->mVertices[i].x = p.positions[3 * i];
			mesh->mVertices[i].y = p.positions[3 * i + 1];
			mesh->mVertices[i].z = p.positions[3 * i + 2];
			if (p.uvs) { mesh->mTextureCoords[0][i].x = p.uvs[2 * i]; mesh->mTextureCoords[0][i].y = p.uvs[2 * i + 1]; }
		}
		mesh->mFaces = new aiFace[p.n_triangles];
		for (int f = 0; f < p.n_triangles; f++) {
			mesh->mFaces[f].mNumIndices = 3;
			mesh->mFaces[f].mIndices = new unsigned int[3]{p.indices[3 * f], p.indices[3 * f + 1], p.indices[3 * f + 2]};
		}
		sc->mNumMeshes = 1;
		sc->mMeshes = new aiMesh*[1]{mesh};
		sc->mNumMaterials = 1;
		sc->mMaterials = new aiMaterial*[1]{new aiMaterial()};
		aiNode* root = new aiNode();
		root->mNumMeshes = 1;
		root->mMeshes = new unsigned int[1]{0};
		sc->mRootNode = root;
		model = new Model(sc, "", material ? matName : std::string(""));
		isLight = false;  // Model::lights is never filled by the assimp path (SceneReader.cpp:245-267)
	} else if (p.type == NE_B200_PRIM_POINT || p.type == NE_B200_PRIM_SPHERE) {
		vertexLayout.init();
		vertexLayout.add(VertexAttrib::Position, VertexAttribType::Float, 3);
		vertexLayout.end();
		float* vertexData = new float[3];
		uint32_t* indexData = new uint32_t[1]{0};
		Primitive* prim;
		uint32_t primSize;
		if (p.type == NE_B200_PRIM_POINT) {
			vertexData[0] = p.point[0]; vertexData[1] = p.point[1]; vertexData[2] = p.point[2];
			Point* pt = new Point[1];
			pt[0].vertexData[0] = vertexData;
			prim = pt; primSize = sizeof(Point);
		} else {
			vertexData[0] = vertexData[1] = vertexData[2] = 0;
			Sphere* s = new Sphere[1];
			s[0].vertexData[0] = vertexData;
			s[0].radius = p.radius;
			prim = s; primSize = sizeof(Sphere);
		}
		if (isLight) { lights.push_back(prim); material->light->primitive = prim; }
		else primitives.push_back(prim);
		materials.push_back(material);
		prim->material = material;
		model = new Model(MemoryBuffer{vertexData, 12}, MemoryBuffer{indexData, 4},
			primitives.size() ? MemoryBuffer{prim, primSize} : MemoryBuffer{}, primitives,
			lights.size() ? MemoryBuffer{prim, primSize} : MemoryBuffer{}, lights, materials, meshes, vertexLayout);
	} else if (p.type == NE_B200_PRIM_RECTANGLE) {
		vertexLayout.init();
		vertexLayout.add(VertexAttrib::Position, VertexAttribType::Float, 3);
		vertexLayout.add(VertexAttrib::Normal, VertexAttribType::Float, 3);
		vertexLayout.add(VertexAttrib::Tangent, VertexAttribType::Float, 3);
		vertexLayout.add(VertexAttrib::TexCoord0, VertexAttribType::Float, 2);
		vertexLayout.end();
		const float pos[4][3] = {{-0.5f, -0.5f, 0}, {0.5f, -0.5f, 0}, {0.5f, 0.5f, 0}, {-0.5f, 0.5f, 0}};
		const float uv[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
		float* vertexData = new float[44];
		for (int i = 0; i < 4; i++) {
			float* v = vertexData + 11 * i;
			v[0] = pos[i][0]; v[1] = pos[i][1]; v[2] = pos[i][2];
			v[3] = 0; v[4] = 0; v[5] = -1.0f;
			v[6] = 1.0f; v[7] = 0; v[8] = 0;
			v[9] = uv[i][0]; v[10] = uv[i][1];
		}
		uint32_t* indexData = new uint32_t[6]{0, 1, 2, 0, 2, 3};
		Rectangle* r = new Rectangle[1];
		r[0].vertexData[0] = &vertexData[0];
		r[0].vertexData[1] = &vertexData[22];
		r[0].normal = glm::vec3(0, 0, -1.0f);
		r->material = material;
		if (isLight) { lights.push_back(r); material->light->primitive = r; }
		else primitives.push_back(r);
		materials.push_back(material);
		model = new Model(MemoryBuffer{vertexData, 44 * 4}, MemoryBuffer{indexData, 24},
			primitives.size() ? MemoryBuffer{r, sizeof(Rectangle)} : MemoryBuffer{}, primitives,
			lights.size() ? MemoryBuffer{r, sizeof(Rectangle)} : MemoryBuffer{}, lights, materials, meshes, vertexLayout);
		r[0].vertexLayout = &model->vertexLayout;
	} else if (p.type == NE_B200_PRIM_VOLUME) {
		vertexLayout.init();
		vertexLayout.add(VertexAttrib::Position, VertexAttribType::Float, 3);
		vertexLayout.end();
		float* vertexData = new float[24];
		const float c[8][3] = {{-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {.5f, .5f, -.5f}, {-.5f, .5f, -.5f},
		                       {.5f, .5f, .5f}, {-.5f, .5f, .5f}, {-.5f, -.5f, .5f}, {.5f, -.5f, .5f}};
		for (int i = 0; i < 8; i++) for (int k = 0; k < 3; k++) vertexData[3 * i + k] = c[i][k];
		uint32_t* indexData = new uint32_t[36]();
		AABB* aabb = new AABB[1];
		aabb[0].vertexData[0] = &vertexData[0];
		aabb[0].vertexData[1] = &vertexData[12];
		aabb[0].material = material;
		materials.push_back(material);
		primitives.push_back(aabb);
		isLight = false;
		model = new Model(MemoryBuffer{vertexData, 96}, MemoryBuffer{indexData, 144}, MemoryBuffer{aabb, sizeof(AABB)},
			primitives, MemoryBuffer{}, lights, materials, meshes, vertexLayout);
	} else
		return nullptr;

	StringID id = ResourceManager::getSelf()->replaceModel(name, model);
	InstancedModel* im = new InstancedModel(model, id, toMat4(p.to_world));
	im->invTransformToWCS = toMat4(p.to_object);  // the descriptor's matrices are used verbatim on both sides
	im->isCollisionEnabled = p.collision != 0;
	return im;
}

RayIntersection makeIsect(RefScene* rs, const ne_b200_hit& h) {
	RayIntersection ri;
	ri.hitPoint = v3(h.hit_point);
	ri.normal = v3(h.normal);
	ri.uv = glm::vec2(h.uv[0], h.uv[1]);
	ri.tNear = h.t_near;
	ri.tFar = h.t_far;
	ri.instancedModel = nullptr;
	ri.primitive = nullptr;
	if (h.instance >= 0 && h.instance < (int)rs->fold.size()) {
		InstancedModel* im = rs->fold[h.instance];
		ri.instancedModel = im;
		Model* m = im->model;
		if (m->bvh.nodeCount > 0)
			ri.primitive = &((Triangle*)m->memoryBufferPrimitives.data)[h.primitive];
		else if (m->primitives.size())
			ri.primitive = m->primitives[0];
		else if (m->lights.size())
			ri.primitive = m->lights[0];
	}
	return ri;
}

void fillHit(RefScene* rs, bool did, const RayIntersection& ri, ne_b200_hit* o) {
	memset(o, 0, sizeof(*o));
	o->hit = did ? 1 : 0;
	o->instance = -1;
	if (!did) return;
	put3(o->hit_point, ri.hitPoint);
	put3(o->normal, ri.normal);
	o->uv[0] = ri.uv.x; o->uv[1] = ri.uv.y;
	o->t_near = ri.tNear;
	o->t_far = ri.tFar;
	for (size_t i = 0; i < rs->fold.size(); i++)
		if (rs->fold[i] == ri.instancedModel) o->instance = (int)i;
	if (ri.primitive && ri.primitive->material && ri.primitive->material->light) o->is_light = 1;
	if (ri.instancedModel && ri.instancedModel->model->bvh.nodeCount > 0)
		o->primitive = int((Triangle*)ri.primitive - (Triangle*)ri.instancedModel->model->memoryBufferPrimitives.data);
}

// Number of 32-bit draws separating two engine states (std::uniform_real_distribution<float> over mt19937 takes
// exactly one draw per float).
int drawsBetween(std::mt19937 from, const std::mt19937& to, int limit = 1 << 22) {
	for (int i = 0; i <= limit; i++) {
		if (from == to) return i;
		from.discard(1);
	}
	return -1;
}

Camera makeCamera(const float* lookFrom, const float* lookAt, const float* up, float vfov, float aspect, float aperture, float focus) {
	return Camera(v3(lookFrom), v3(lookAt), v3(up), vfov, aspect, aperture, focus);
}

}  // namespace

extern "C" {

int neref_version() { return 1; }
// 1 when built with the thread_local patch of the global RNG (oracle/Makefile), 0 for the as-shipped shared RNG.
int neref_thread_local_rng() {
#ifdef NE_ORACLE_TLS_RNG
	return 1;
#else
	return 0;
#endif
}

void neref_seed(uint32_t k) { narvalengine::mt.seed(k); narvalengine::dist.reset(); }

// The uniform tape narvalengine::random() yields after mt.seed(k) (Math.h:51-66).
void neref_tape(uint32_t k, int n, float* out) {
	neref_seed(k);
	for (int i = 0; i < n; i++) out[i] = narvalengine::random();
}

void* neref_scene_create(const ne_b200_scene_desc* d) {
	RefScene* rs = new RefScene();
	rs->uid = g_scene_uid++;
	rs->scene = new Scene();
	std::string pfx = "s" + std::to_string(rs->uid) + ".";
	for (int i = 0; i < d->n_textures; i++) {
		const ne_b200_texture& t = d->textures[i];
		int flags = 0;
		flags |= t.wrap_u == NE_B200_WRAP_MIRROR ? NE_TEX_SAMPLER_U_MIRROR : NE_TEX_SAMPLER_U_CLAMP;
		flags |= t.wrap_v == NE_B200_WRAP_MIRROR ? NE_TEX_SAMPLER_V_MIRROR : NE_TEX_SAMPLER_V_CLAMP;
		flags |= NE_TEX_SAMPLER_W_CLAMP;
		uint32_t bytes = bytesPerTexel(t.format) * t.width * t.height;
		rs->textures.push_back(new Texture(t.width, t.height, layoutOf(t.format), flags, MemoryBuffer{(void*)t.texels, bytes}));
	}
	for (int i = 0; i < d->n_volumes; i++) {
		const ne_b200_volume& v = d->volumes[i];
		int flags = NE_TEX_SAMPLER_UVW_CLAMP | NE_TEX_SAMPLER_MIN_MAG_LINEAR;
		uint32_t bytes = uint32_t(size_t(v.width) * v.height * v.depth * 4);
		if (v.dense)
			rs->volumes.push_back(new Texture(v.width, v.height, v.depth, R32F, flags, MemoryBuffer{(void*)v.dense, bytes}));
		else {
			std::vector<float> g = densify(v);
			rs->volumes.push_back(new Texture(v.width, v.height, v.depth, R32F, flags, MemoryBuffer{g.data(), bytes}));
		}
	}
	for (int i = 0; i < d->n_materials; i++) {
		Material* m = buildMaterial(*rs, d->materials[i], pfx + "mat" + std::to_string(i));
		if (!m) { fprintf(stderr, "neref: unsupported material type %d\n", d->materials[i].type); return nullptr; }
		rs->materials.push_back(m);
	}
	for (int i = 0; i < d->n_primitives; i++) {
		bool isLight = false;
		const ne_b200_primitive& p = d->primitives[i];
		InstancedModel* im = buildPrimitive(*rs, p, pfx + "prim" + std::to_string(i), pfx + "mat" + std::to_string(p.material), isLight);
		if (!im) { fprintf(stderr, "neref: unsupported primitive type %d\n", p.type); return nullptr; }
		if (isLight) rs->scene->lights.push_back(im);
		else rs->scene->instancedModels.push_back(im);
	}
	if (d->sort_and_group) {
		// SceneEditor::sortAndGroup, src/SceneEditor.cpp:2039-2054: media go to the end, order otherwise kept.
		std::vector<InstancedModel*> a, b;
		for (InstancedModel* im : rs->scene->instancedModels) {
			bool medium = false;
			for (Material* m : im->model->materials) if (m && m->medium) medium = true;
			(medium ? b : a).push_back(im);
		}
		a.insert(a.end(), b.begin(), b.end());
		rs->scene->instancedModels = a;
	}
	rs->fold = rs->scene->instancedModels;
	rs->fold.insert(rs->fold.end(), rs->scene->lights.begin(), rs->scene->lights.end());
	return rs;
}

void neref_scene_destroy(void* h) {
	RefScene* rs = (RefScene*)h;
	if (!rs) return;
	for (Texture* t : rs->volumes) { delete[] (uint8_t*)t->mem.data; t->mem.data = nullptr; }
	// The rest of the object graph is small and intentionally leaked (Scene's destructor would double-delete
	// shared buffers); this is a test harness.
}

void neref_scene_counts(void* h, int* nModels, int* nLights) {
	RefScene* rs = (RefScene*)h;
	*nModels = (int)rs->scene->instancedModels.size();
	*nLights = (int)rs->scene->lights.size();
}

// -- pure functions ---------------------------------------------------------------------------------------------
void neref_get_transform(const float* pos, const float* rotDeg, const float* scale, float* toWorld, float* toObject) {
	glm::mat4 m = getTransform(v3(pos), v3(rotDeg), v3(scale));
	fromMat4(m, toWorld);
	fromMat4(glm::inverse(m), toObject);
}
void neref_onb(const float* n, float* v, float* u) {
	glm::vec3 vv, uu;
	generateOrthonormalCS(v3(n), vv, uu);
	put3(v, vv); put3(u, uu);
}
void neref_get_scale(const float* m, float* s) { put3(s, getScale(toMat4(m))); }
float neref_area_to_solid_angle(float pdfArea, const float* n, const float* p1, const float* p2) { return convertAreaToSolidAngle(pdfArea, v3(n), v3(p1), v3(p2)); }
float neref_power_heuristic(float a, float b) { return powerHeuristic(a, b); }
float neref_roughness_to_alpha(float r) { return roughnessToAlpha(r); }
void neref_sample_unit_sphere(float e1, float e2, float* out) { put3(out, sampleUnitSphere(e1, e2)); }
float neref_ggx_D(float alpha, const float* h) { GGXDistribution g; g.alpha = alpha; return g.D(v3(h)); }
float neref_ggx_G(float alpha, const float* wo, const float* wi) { GGXDistribution g; g.alpha = alpha; return g.G(v3(wo), v3(wi)); }
float neref_ggx_pdf(float alpha, const float* wi, const float* h) { GGXDistribution g; g.alpha = alpha; return g.pdf(v3(wi), v3(h)); }
float neref_fresnel(float c) { FresnelSchilck f; return f.eval(c).x; }
float neref_hg_eval(float g, const float* in, const float* out) { HG hg(g); return hg.eval(v3(in), v3(out)); }
void neref_hg_sample(float g, uint32_t seed, float* out) { neref_seed(seed); HG hg(g); put3(out, hg.sample(glm::vec3(0, 0, 1))); }
// -- primitive-level entry points: the objects the reference's own unit tests exercise (unitTests/tests.cpp) ------
void neref_to_lcs(const float* v, const float* ns, const float* ss, const float* ts, float* out) { put3(out, toLCS(v3(v), v3(ns), v3(ss), v3(ts))); }
void neref_to_world(const float* v, const float* ns, const float* ss, const float* ts, float* out) { put3(out, toWorld(v3(v), v3(ns), v3(ss), v3(ts))); }
// Triangle::intersect (primitives/Triangle.cpp:49-81) on a stand-alone triangle of 3 position-only vertices.
int neref_triangle_intersect(const float* v0, const float* v1, const float* v2, const float* o, const float* d, float* tNearFar, float* hitPoint, float* normal) {
	float data[9] = {v0[0], v0[1], v0[2], v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]};
	Triangle t(&data[0], &data[3], &data[6]);
	RayIntersection ri;
	bool did = t.intersect(Ray(v3(o), v3(d)), ri);
	tNearFar[0] = ri.tNear; tNearFar[1] = ri.tFar;
	put3(hitPoint, ri.hitPoint); put3(normal, ri.normal);
	return did ? 1 : 0;
}
void neref_triangle_barycentric(const float* v0, const float* v1, const float* v2, const float* p, float* out) {
	float data[9] = {v0[0], v0[1], v0[2], v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]};
	Triangle t(&data[0], &data[3], &data[6]);
	put3(out, t.barycentricCoordinates(v3(p), v3(v0), v3(v1), v3(v2)));
}
int neref_point_in_triangle_range(const float* p, const float* a, const float* b, const float* c) { return isPointInsideTriangleRange(v3(p), v3(a), v3(b), v3(c)) ? 1 : 0; }
// AABB::intersect (primitives/AABB.cpp:48-79).
int neref_aabb_intersect(const float* bmin, const float* bmax, const float* o, const float* d, float* tNearFar, float* hitPoint, float* normal) {
	AABB box(v3(bmin), v3(bmax));
	RayIntersection ri;
	bool did = box.intersect(Ray(v3(o), v3(d)), ri);
	tNearFar[0] = ri.tNear; tNearFar[1] = ri.tFar;
	put3(hitPoint, ri.hitPoint); put3(normal, ri.normal);
	return did ? 1 : 0;
}
// IsotropicPhaseFunction::sample / pdf after mt.seed(seed) (materials/Medium.h:43-63).
float neref_isotropic_sample(uint32_t seed, float* out) {
	neref_seed(seed);
	IsotropicPhaseFunction p;
	glm::vec3 s = p.sample(glm::vec3(0, 0, 1));
	put3(out, s);
	return p.pdf(glm::vec3(0, 0, 1), s);
}
void neref_tonemap(const float* in, int n, float* out) {
	// OfflineEngine::postProcessing (OfflineEngine.cpp:39-52) through a real OfflineEngine object.
	static OfflineEngine* eng = nullptr;
	if (!eng) {
		SceneSettings st; st.resolution = glm::ivec2(1, 1); st.spp = 1; st.bounces = 1;
		eng = new OfflineEngine(Camera(), st, nullptr);
	}
	for (int i = 0; i < n; i++) put3(out + 3 * i, eng->postProcessing(v3(in + 3 * i)));
}

void neref_camera_make(const float* lookFrom, const float* lookAt, const float* up, float vfov, float aspect, float aperture, float focus, ne_b200_camera* out) {
	Camera c = makeCamera(lookFrom, lookAt, up, vfov, aspect, aperture, focus);
	*out = narval_b200_adapter::toPod(c);  // the reference-side adapter (oracle/ref/flatten.cpp)
}

// The reference-side adapter on a Scene* built by this harness: Scene* -> ne_b200_scene_desc (oracle/ref/flatten.cpp).
// The returned handle owns the descriptor; free it with neref_flat_free.
void* neref_flatten(void* h) {
	RefScene* rs = (RefScene*)h;
	auto* f = new narval_b200_adapter::FlatScene();
	narval_b200_adapter::flatten(rs->scene, *f);
	return f;
}
const ne_b200_scene_desc* neref_flat_desc(void* f) { return &((narval_b200_adapter::FlatScene*)f)->desc; }
void neref_flat_free(void* f) { delete (narval_b200_adapter::FlatScene*)f; }
// getRayPassingThrough for n (x,y) pairs in sequence after mt.seed(seed).
void neref_camera_rays(const float* lookFrom, const float* lookAt, const float* up, float vfov, float aspect, float aperture, float focus,
                       uint32_t seed, int n, const float* xy, float* o, float* d) {
	Camera c = makeCamera(lookFrom, lookAt, up, vfov, aspect, aperture, focus);
	neref_seed(seed);
	for (int i = 0; i < n; i++) {
		Ray r = c.getRayPassingThrough(xy[2 * i], xy[2 * i + 1]);
		put3(o + 3 * i, r.o); put3(d + 3 * i, r.d);
	}
}

// -- scene functions --------------------------------------------------------------------------------------------
void neref_intersect(void* h, int n, const float* o, const float* d, float tMin, float tMax, ne_b200_hit* out) {
	RefScene* rs = (RefScene*)h;
	for (int i = 0; i < n; i++) {
		RayIntersection ri;
		ri.instancedModel = nullptr;
		bool did = rs->scene->intersectScene(Ray(v3(o + 3 * i), v3(d + 3 * i)), ri, tMin, tMax);
		fillHit(rs, did, ri, out + i);
	}
}

// BSDF::eval / pdf / sample at a synthetic hit on fold-order instance `instance`. seeds==nullptr: no sampling.
void neref_bsdf(void* h, int n, int instance, const float* incoming, const float* scattered, const float* normals, const float* uvs,
                const uint32_t* seeds, float* eval, float* pdf, float* sampled) {
	RefScene* rs = (RefScene*)h;
	for (int i = 0; i < n; i++) {
		ne_b200_hit hh{};
		hh.instance = instance;
		memcpy(hh.normal, normals + 3 * i, 12);
		if (uvs) { hh.uv[0] = uvs[2 * i]; hh.uv[1] = uvs[2 * i + 1]; }
		RayIntersection ri = makeIsect(rs, hh);
		BSDF* b = ri.primitive->material->bsdf;
		glm::vec3 in = v3(incoming + 3 * i), sc = v3(scattered + 3 * i);
		if (eval) put3(eval + 3 * i, b->eval(in, sc, ri));
		if (pdf) pdf[i] = b->pdf(in, sc, ri.normal, ri);
		if (seeds && sampled) { neref_seed(seeds[i]); put3(sampled + 3 * i, b->sample(in, ri.normal, ri)); }
	}
}

static GridMedia* gridOf(RefScene* rs, int instance, RayIntersection& ri) {
	ne_b200_hit hh{};
	hh.instance = instance;
	ri = makeIsect(rs, hh);
	return dynamic_cast<GridMedia*>(ri.primitive->material->medium);
}

void neref_grid_tr(void* h, int n, int instance, const float* o, const float* d, const float* tNear, const float* tFar,
                   const uint32_t* seeds, float* tr, int32_t* used) {
	RefScene* rs = (RefScene*)h;
	RayIntersection ri;
	GridMedia* g = gridOf(rs, instance, ri);
	for (int i = 0; i < n; i++) {
		neref_seed(seeds[i]);
		std::mt19937 start = narvalengine::mt;
		ri.tNear = tNear[i]; ri.tFar = tFar[i];
		glm::vec3 t = g->Tr(Ray(v3(o + 3 * i), v3(d + 3 * i)), ri);
		tr[i] = t.x;
		if (used) used[i] = drawsBetween(start, narvalengine::mt);
	}
}

void neref_grid_sample(void* h, int n, int instance, const float* o, const float* d, const float* tNear, const float* tFar,
                       const uint32_t* seeds, float* transmittance, float* so, float* sd, int32_t* used) {
	RefScene* rs = (RefScene*)h;
	RayIntersection ri;
	GridMedia* g = gridOf(rs, instance, ri);
	for (int i = 0; i < n; i++) {
		neref_seed(seeds[i]);
		std::mt19937 start = narvalengine::mt;
		ri.tNear = tNear[i]; ri.tFar = tFar[i];
		Ray sc;
		glm::vec3 t = g->sample(Ray(v3(o + 3 * i), v3(d + 3 * i)), sc, ri);
		put3(transmittance + 3 * i, t);
		put3(so + 3 * i, sc.o); put3(sd + 3 * i, sc.d);
		if (used) used[i] = drawsBetween(start, narvalengine::mt);
	}
}

void neref_density(void* h, int n, int instance, const float* pts, float* out, float* invMax) {
	RefScene* rs = (RefScene*)h;
	RayIntersection ri;
	GridMedia* g = gridOf(rs, instance, ri);
	for (int i = 0; i < n; i++)
		out[i] = g->interpolatedDensity(g->fromOCStoGCS(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
	if (invMax) *invMax = g->invMaxDensity;
}

void neref_li(void* h, int n, const float* o, const float* d, int bounces, const uint32_t* seeds, float* radiance, int32_t* used) {
	RefScene* rs = (RefScene*)h;
	rs->scene->settings.bounces = bounces;
	VolumetricPathIntegrator integ;
	for (int i = 0; i < n; i++) {
		neref_seed(seeds[i]);
		std::mt19937 start = narvalengine::mt;
		glm::vec3 L = integ.Li(Ray(v3(o + 3 * i), v3(d + 3 * i)), rs->scene);
		put3(radiance + 3 * i, L);
		if (used) used[i] = drawsBetween(start, narvalengine::mt);
	}
}

void neref_sample_one_light(void* h, int n, const float* incomingDirs, const ne_b200_hit* hits, const uint32_t* seeds, float* radiance, int32_t* used) {
	RefScene* rs = (RefScene*)h;
	VolumetricPathIntegrator integ;
	for (int i = 0; i < n; i++) {
		RayIntersection ri = makeIsect(rs, hits[i]);
		Ray in(ri.hitPoint - v3(incomingDirs + 3 * i), v3(incomingDirs + 3 * i));
		neref_seed(seeds[i]);
		std::mt19937 start = narvalengine::mt;
		glm::vec3 L = integ.uniformSampleOneLight(in, ri, rs->scene);
		put3(radiance + 3 * i, L);
		if (used) used[i] = drawsBetween(start, narvalengine::mt);
	}
}

// Full-frame render with the reference's own Camera / Li / postProcessing and OfflineEngine::renderTile's pixel
// and sample loops (OfflineEngine.cpp:61-71) over the WHOLE image (the shipped 40x10 tiling drops edge pixels,
// Q28). Rows are interleaved over `nthreads` std::threads. With the thread_local RNG build every thread seeds
// its own engine; with the as-shipped build all threads share (and race on) the single global engine.
// Only rows rowBegin, rowBegin+rowStep, ... < rowEnd are rendered (a bounded, frame-covering sample for timing).
// linear / tonemapped: W*H*3 floats, may be NULL. Returns wall seconds of the render loop.
double neref_render(void* h, const float* lookFrom, const float* lookAt, const float* up, float vfov, float aperture, float focus,
                    int W, int H, int spp, int bounces, uint32_t seed, int nthreads, int rowBegin, int rowEnd, int rowStep, float* linear, float* tonemapped) {
	RefScene* rs = (RefScene*)h;
	rs->scene->settings.bounces = bounces;
	rs->scene->settings.spp = spp;
	rs->scene->settings.resolution = glm::ivec2(W, H);
	Camera cam = makeCamera(lookFrom, lookAt, up, vfov, float(W) / float(H), aperture, focus);
	SceneSettings st = rs->scene->settings;
	OfflineEngine eng(cam, st, rs->scene);
	if (nthreads < 1) nthreads = 1;
	if (rowEnd <= 0 || rowEnd > H) rowEnd = H;
	if (rowBegin < 0) rowBegin = 0;
	if (rowStep < 1) rowStep = 1;
	auto t0 = std::chrono::steady_clock::now();
	auto worker = [&](int tid) {
#ifdef NE_ORACLE_TLS_RNG
		narvalengine::mt.seed(seed * 7919u + 104729u * (uint32_t)tid);
#else
		if (tid == 0) narvalengine::mt.seed(seed);
#endif
		Integrator* integ = eng.pathIntegrator->clone();
		for (int y = rowBegin + tid * rowStep; y < rowEnd; y += nthreads * rowStep)
			for (int x = 0; x < W; x++) {
				glm::vec3 color(0, 0, 0);
				for (int s = 0; s < spp; s++) {
					float u = float(x + narvalengine::random()) / W;
					float v = float(y + narvalengine::random()) / H;
					Ray r = cam.getRayPassingThrough(u, v);
					color += integ->Li(r, rs->scene);
				}
				color = color / float(spp);
				if (linear) put3(linear + 3 * (size_t(W) * y + x), color);
				if (tonemapped) put3(tonemapped + 3 * (size_t(W) * y + x), eng.postProcessing(color));
			}
		delete integ;
	};
	if (nthreads == 1) worker(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
		for (auto& t : th) t.join();
	}
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// The reference's own tile entry point, OfflineEngine::renderTile (OfflineEngine.cpp:54-76), single tile.
void neref_render_tile(void* h, const float* lookFrom, const float* lookAt, const float* up, float vfov, float aperture, float focus,
                       int W, int H, int spp, int bounces, uint32_t seed, int tile, float* pixelsOut) {
	RefScene* rs = (RefScene*)h;
	rs->scene->settings.bounces = bounces;
	rs->scene->settings.spp = spp;
	rs->scene->settings.resolution = glm::ivec2(W, H);
	Camera cam = makeCamera(lookFrom, lookAt, up, vfov, float(W) / float(H), aperture, focus);
	OfflineEngine eng(cam, rs->scene->settings, rs->scene);
	memset((void*)eng.pixels, 0, sizeof(glm::vec3) * W * H);
	neref_seed(seed);
	std::atomic<bool> done{false};
	eng.renderTile(cam, tile, done);
	memcpy(pixelsOut, eng.pixels, sizeof(glm::vec3) * W * H);
}

}  // extern "C"

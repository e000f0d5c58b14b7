// Minimal PODs with assimp's field names, enough for primitives/Model.cpp to compile unmodified
// and for the harness to hand-fill an aiScene (so Model(const aiScene*,...) + BVH::init run as shipped).
#pragma once
#include <cstring>
#include <string>
struct aiVector3D { float x = 0, y = 0, z = 0; };
struct aiString {
	std::string s;
	const char* C_Str() const { return s.c_str(); }
};
enum aiTextureType {
	aiTextureType_NONE = 0, aiTextureType_DIFFUSE, aiTextureType_SPECULAR, aiTextureType_AMBIENT, aiTextureType_EMISSIVE,
	aiTextureType_HEIGHT, aiTextureType_NORMALS, aiTextureType_SHININESS, aiTextureType_OPACITY, aiTextureType_DISPLACEMENT,
	aiTextureType_LIGHTMAP, aiTextureType_REFLECTION, aiTextureType_BASE_COLOR, aiTextureType_NORMAL_CAMERA,
	aiTextureType_EMISSION_COLOR, aiTextureType_METALNESS, aiTextureType_DIFFUSE_ROUGHNESS, aiTextureType_AMBIENT_OCCLUSION,
	aiTextureType_UNKNOWN
};
#define AI_MATKEY_TEXTURE(type, N) "$tex.file", type, N
struct aiTexture {};
struct aiMaterial {
	unsigned int GetTextureCount(aiTextureType) const { return 0; }
	int Get(const char*, unsigned int, unsigned int, aiString&) const { return -1; }
	int GetTexture(aiTextureType, unsigned int, aiString*) const { return -1; }
	aiString GetName() const { return aiString(); }
};
struct aiFace { unsigned int mNumIndices = 0; unsigned int* mIndices = nullptr; };
struct aiMesh {
	unsigned int mNumVertices = 0, mNumFaces = 0, mMaterialIndex = 0;
	aiVector3D* mVertices = nullptr; aiVector3D* mNormals = nullptr; aiVector3D* mTangents = nullptr;
	aiVector3D* mTextureCoords[8] = {};
	aiFace* mFaces = nullptr;
};
struct aiNode {
	unsigned int mNumMeshes = 0; unsigned int* mMeshes = nullptr;
	unsigned int mNumChildren = 0; aiNode** mChildren = nullptr;
};
struct aiScene {
	unsigned int mNumMeshes = 0; aiMesh** mMeshes = nullptr;
	unsigned int mNumMaterials = 0; aiMaterial** mMaterials = nullptr;
	aiNode* mRootNode = nullptr;
	const aiTexture* GetEmbeddedTexture(const char*) const { return nullptr; }
};

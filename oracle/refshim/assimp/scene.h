// Stand-in for the Assimp subset the reference's Scene::Load uses (rt.hpp:1640-1730). Written for this repository; see
// ../README.md. Backed by this repository's own OBJ reader (nanogi_b200/host/obj_loader.hpp): first sub-mesh only, fan
// triangulation, de-indexed vertices (three per face) — the reference de-indexes again for Embree and reads the per-vertex
// arrays through Faces, so joined or not joined vertices give the same triangles, normals and uvs.
#pragma once
#include <cstddef>
struct aiVector3D { float x, y, z; };
struct aiFace { unsigned int mNumIndices; unsigned int* mIndices; };
struct aiMesh {
    unsigned int mNumVertices = 0, mNumFaces = 0;
    aiVector3D* mVertices = nullptr; aiVector3D* mNormals = nullptr;
    aiVector3D* mTextureCoords[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    aiFace* mFaces = nullptr;
    bool HasNormals() const { return mNormals != nullptr && mNumVertices > 0; }
    bool HasTextureCoords(unsigned int i) const { return i < 8 && mTextureCoords[i] != nullptr && mNumVertices > 0; }
};
struct aiScene { unsigned int mNumMeshes = 0; aiMesh** mMeshes = nullptr; };

// Stand-in: Assimp::Importer::ReadFile / ApplyPostProcessing / GetErrorString (rt.hpp:1645-1672). See scene.h.
#pragma once
#include <string>
#include <vector>
#include "postprocess.h"
#include "scene.h"
#include "../../../nanogi_b200/host/obj_loader.hpp"
namespace Assimp {
class Importer {
public:
    const aiScene* ReadFile(const char* path, unsigned int) {
        path_ = path;
        try {
            try { fill(ngi::LoadObj(path_, false, false, false), true); }
            catch (const std::runtime_error& e) {
                if (std::string(e.what()).find("has no normals") == std::string::npos) throw;
                fill(ngi::LoadObj(path_, true, false, true), false);          // positions only: the file carries no normals
            }
        } catch (const std::exception& e) { error_ = e.what(); return nullptr; }
        return &scene_;
    }
    const aiScene* ApplyPostProcessing(unsigned int flags) {
        if (!has_file_normals_ && (flags & (aiProcess_GenNormals | aiProcess_GenSmoothNormals)))
            fill(ngi::LoadObj(path_, (flags & aiProcess_GenNormals) != 0, (flags & aiProcess_GenSmoothNormals) != 0, true), true);
        return &scene_;
    }
    const char* GetErrorString() const { return error_.c_str(); }
private:
    void fill(const ngi::TriMesh& m, bool normals) {
        const size_t nv = m.positions.size() / 3;
        v_.resize(nv); n_.resize(nv); t_.clear();
        for (size_t i = 0; i < nv; i++) { v_[i] = {m.positions[3 * i], m.positions[3 * i + 1], m.positions[3 * i + 2]}; n_[i] = {m.normals[3 * i], m.normals[3 * i + 1], m.normals[3 * i + 2]}; }
        if (!m.texcoords.empty()) { t_.resize(nv); for (size_t i = 0; i < nv; i++) t_[i] = {m.texcoords[2 * i], m.texcoords[2 * i + 1], 0.0f}; }
        idx_.resize(nv); faces_.resize(nv / 3);
        for (size_t i = 0; i < nv; i++) idx_[i] = (unsigned)i;
        for (size_t f = 0; f < nv / 3; f++) { faces_[f].mNumIndices = 3; faces_[f].mIndices = &idx_[3 * f]; }
        has_file_normals_ = normals;
        mesh_ = aiMesh();
        mesh_.mNumVertices = (unsigned)nv; mesh_.mNumFaces = (unsigned)(nv / 3);
        mesh_.mVertices = v_.data(); mesh_.mNormals = normals ? n_.data() : nullptr;
        mesh_.mTextureCoords[0] = t_.empty() ? nullptr : t_.data();
        mesh_.mFaces = faces_.data();
        meshes_[0] = &mesh_;
        scene_.mNumMeshes = nv ? 1 : 0; scene_.mMeshes = meshes_;
    }
    std::string path_, error_;
    bool has_file_normals_ = false;
    std::vector<aiVector3D> v_, n_, t_;
    std::vector<unsigned> idx_;
    std::vector<aiFace> faces_;
    aiMesh mesh_; aiMesh* meshes_[1] = {nullptr}; aiScene scene_;
};
}

// obj_loader.hpp — Wavefront OBJ reader producing the DE-INDEXED float triangle soup the render module
// consumes. Stands in for the reference's Assimp import (include/nanogi/rt.hpp:1640-1730):
//   * only the first sub-mesh is used (rt.hpp:1679 `scene->mMeshes[0]`): faces of the first
//     (object/group, material) combination that owns faces; later ones are ignored with a warning;
//   * polygons are triangulated as a fan from their first vertex (Assimp's aiProcess_Triangulate
//     does the same for convex quads, which is all the shipped fixtures contain);
//   * Assimp stores float32 vertices/normals; the reference widens them to double (rt.hpp:1685-1696)
//     and narrows back to float for Embree (rt.hpp:2119-2121) — so float here loses nothing;
//   * meshes WITHOUT normals need a `postprocess` block: generate_normals -> flat face normals,
//     generate_smooth_normals -> area-weighted average over coincident positions
//     (aiProcess_GenNormals / aiProcess_GenSmoothNormals, rt.hpp:1655-1663). With neither the
//     reference dereferences a null mNormals (rt.hpp:1688); we fail with a message instead.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ngi {

struct TriMesh {
    std::vector<float> positions;  // [nTri][3][3]
    std::vector<float> normals;    // [nTri][3][3]
    std::vector<float> texcoords;  // [nTri][3][2] or empty
    size_t num_tris() const { return positions.size() / 9; }
    int ignored_submeshes = 0;
};

namespace detail {
inline const char* skip_ws(const char* p) { while (*p == ' ' || *p == '\t') p++; return p; }
inline bool parse_floats(const char* p, float* out, int n, int* got = nullptr) {
    int k = 0;
    for (; k < n; k++) {
        p = skip_ws(p);
        char* end = nullptr;
        const float v = std::strtof(p, &end);
        if (end == p) break;
        out[k] = v; p = end;
    }
    if (got) *got = k;
    return k == n;
}
}  // namespace detail

inline TriMesh LoadObj(const std::string& path, bool gen_normals, bool gen_smooth_normals, bool has_postprocess) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Unable to open file \"" + path + "\"");
    std::vector<float> V, VN, VT;
    struct Corner { int v, vt, vn; };
    std::vector<Corner> corners;       // 3 per triangle, mesh 0 only
    std::string line;
    int group_id = 0;                  // bumps on o / g / usemtl
    int mesh0_group = -1;
    int ignored = 0; int last_ignored_group = -1;
    bool any_vt = true, any_vn = true;
    int lineno = 0;
    while (std::getline(in, line)) {
        lineno++;
        const char* p = detail::skip_ws(line.c_str());
        if (*p == 0 || *p == '#' || *p == '\r') continue;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float f[3];
            if (!detail::parse_floats(p + 1, f, 3)) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": bad vertex");
            V.insert(V.end(), f, f + 3);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            float f[3];
            if (!detail::parse_floats(p + 2, f, 3)) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": bad normal");
            VN.insert(VN.end(), f, f + 3);
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            float f[3] = {0, 0, 0}; int got = 0;
            detail::parse_floats(p + 2, f, 3, &got);
            if (got < 2) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": bad texcoord");
            VT.insert(VT.end(), f, f + 2);
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            if (mesh0_group < 0) mesh0_group = group_id;
            if (group_id != mesh0_group) {
                if (last_ignored_group != group_id) { ignored++; last_ignored_group = group_id; }
                continue;
            }
            std::vector<Corner> poly;
            p += 1;
            while (true) {
                p = detail::skip_ws(p);
                if (*p == 0 || *p == '\r' || *p == '\n') break;
                Corner c{0, 0, 0};
                char* end = nullptr;
                long vi = std::strtol(p, &end, 10);
                if (end == p) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": bad face");
                c.v = (int)vi; p = end;
                if (*p == '/') {
                    p++;
                    if (*p != '/') { long t = std::strtol(p, &end, 10); if (end != p) { c.vt = (int)t; p = end; } }
                    if (*p == '/') { p++; long nn = std::strtol(p, &end, 10); if (end != p) { c.vn = (int)nn; p = end; } }
                }
                // OBJ indices are 1-based; negative = relative to the current end
                const int nv = (int)(V.size() / 3), nt = (int)(VT.size() / 2), nn = (int)(VN.size() / 3);
                c.v = c.v > 0 ? c.v - 1 : (c.v < 0 ? nv + c.v : -1);
                c.vt = c.vt > 0 ? c.vt - 1 : (c.vt < 0 ? nt + c.vt : -1);
                c.vn = c.vn > 0 ? c.vn - 1 : (c.vn < 0 ? nn + c.vn : -1);
                if (c.v < 0 || c.v >= nv) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": vertex index out of range");
                if (c.vt >= nt || c.vn >= nn) throw std::runtime_error(path + ":" + std::to_string(lineno) + ": attribute index out of range");
                poly.push_back(c);
            }
            if (poly.size() < 3) continue;  // points / lines are dropped by aiProcess_Triangulate consumers
            for (size_t k = 1; k + 1 < poly.size(); k++) {
                const Corner tri[3] = {poly[0], poly[k], poly[k + 1]};
                for (auto& c : tri) { corners.push_back(c); if (c.vt < 0) any_vt = false; if (c.vn < 0) any_vn = false; }
            }
        } else if ((p[0] == 'o' || p[0] == 'g') && (p[1] == ' ' || p[1] == '\t' || p[1] == 0 || p[1] == '\r')) {
            group_id++;
        } else if (std::strncmp(p, "usemtl", 6) == 0) {
            group_id++;
        }
        // mtllib, s (smoothing groups) and everything else carry no geometry
    }
    if (corners.empty()) throw std::runtime_error("No mesh is found in " + path);  // rt.hpp:1649-1653

    TriMesh m;
    m.ignored_submeshes = ignored;
    const size_t nt = corners.size() / 3;
    m.positions.resize(nt * 9);
    for (size_t i = 0; i < corners.size(); i++) std::memcpy(&m.positions[i * 3], &V[(size_t)corners[i].v * 3], 3 * sizeof(float));
    if (any_vt && !VT.empty()) {
        m.texcoords.resize(nt * 6);
        for (size_t i = 0; i < corners.size(); i++) std::memcpy(&m.texcoords[i * 2], &VT[(size_t)corners[i].vt * 2], 2 * sizeof(float));
    }
    m.normals.resize(nt * 9);
    if (any_vn && !VN.empty()) {
        for (size_t i = 0; i < corners.size(); i++) std::memcpy(&m.normals[i * 3], &VN[(size_t)corners[i].vn * 3], 3 * sizeof(float));
        return m;
    }
    if (!has_postprocess || (!gen_normals && !gen_smooth_normals))
        throw std::runtime_error("mesh " + path + " has no normals and no postprocess.generate_normals / generate_smooth_normals");
    // face normals (unnormalised = 2*area weighted)
    std::vector<float> fn(nt * 3);
    for (size_t t = 0; t < nt; t++) {
        const float* a = &m.positions[t * 9]; const float* b = a + 3; const float* c = a + 6;
        const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        fn[t * 3 + 0] = e1[1] * e2[2] - e1[2] * e2[1];
        fn[t * 3 + 1] = e1[2] * e2[0] - e1[0] * e2[2];
        fn[t * 3 + 2] = e1[0] * e2[1] - e1[1] * e2[0];
    }
    auto normalize3 = [](float* v) {
        const float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; }
    };
    if (gen_smooth_normals) {
        // accumulate per OBJ position index (coincident positions that share an index; Assimp joins
        // positions within an epsilon — files that duplicate positions get per-copy normals here)
        std::vector<float> acc(V.size(), 0.f);
        for (size_t i = 0; i < corners.size(); i++) for (int k = 0; k < 3; k++) acc[(size_t)corners[i].v * 3 + k] += fn[(i / 3) * 3 + k];
        for (size_t i = 0; i < corners.size(); i++) {
            float n[3] = {acc[(size_t)corners[i].v * 3], acc[(size_t)corners[i].v * 3 + 1], acc[(size_t)corners[i].v * 3 + 2]};
            normalize3(n);
            std::memcpy(&m.normals[i * 3], n, sizeof(n));
        }
    } else {
        for (size_t t = 0; t < nt; t++) {
            float n[3] = {fn[t * 3], fn[t * 3 + 1], fn[t * 3 + 2]};
            normalize3(n);
            for (int k = 0; k < 3; k++) std::memcpy(&m.normals[t * 9 + k * 3], n, sizeof(n));
        }
    }
    return m;
}

}  // namespace ngi

// scene_loader.hpp — host-side mirror of `Scene::Load` (reference include/nanogi/rt.hpp:1519-2154):
// schema.yml YAML -> primitives + OBJ meshes -> the flattened POD `NgiSceneDesc` that crosses the
// C ABI (include/nanogi_gpu.h). Same keys, same defaults, same error behaviour (bool + logged message).
// The Embree commit (rt.hpp:2085-2143) is what ngi_gpu_scene_create replaces.
#pragma once
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/nanogi_gpu.h"
#include "logger.hpp"
#include "obj_loader.hpp"
#include "yaml_lite.hpp"

namespace ngi {

namespace {
const int AppConfigVersionMin = 3;  // rt.hpp:1484-1485
const int AppConfigVersionMax = 5;
}

struct HostScene {
    std::vector<float> positions, normals, texcoords;
    std::vector<NgiPrimitive> prims;
    // TexR textures (rt.hpp:1741-1765 LoadTexture: one Texture per distinct path)
    struct HostTexture { std::string path; int width = 0, height = 0; std::vector<float> rgb; };
    std::vector<HostTexture> textures;
    mutable std::vector<NgiTexture> texture_descs;
    bool any_uv = false, all_uv = true;
    int sensor_prim = -1;
    std::vector<int> light_prims;
    std::string error;

    NgiSceneDesc desc() const {
        NgiSceneDesc d{};
        d.struct_size = sizeof(NgiSceneDesc);
        d.num_prims = (uint32_t)prims.size();
        d.num_tris = positions.size() / 9;
        d.positions = positions.data();
        d.normals = normals.data();
        d.texcoords = (any_uv && !texcoords.empty()) ? texcoords.data() : nullptr;   // meshes without uv carry zeros
        d.prims = prims.data();
        texture_descs.clear();
        for (const HostTexture& t : textures) { NgiTexture nt; nt.width = t.width; nt.height = t.height; nt.rgb = t.rgb.data(); texture_descs.push_back(nt); }
        d.num_textures = (uint32_t)texture_descs.size();
        d.textures = texture_descs.empty() ? nullptr : texture_descs.data();
        return d;
    }

    // LoadTexture, rt.hpp:1741-1765 (textures are shared by path)
    int LoadTexture(const std::string& path) {
        for (size_t i = 0; i < textures.size(); i++) if (textures[i].path == path) return (int)i;
        HostTexture t; t.path = path;
        std::string err;
        if (!LoadImageRGB(path, t.width, t.height, t.rgb, err)) { error = err; NGI_LOG_ERROR(error); return -1; }
        textures.push_back(std::move(t));
        return (int)textures.size() - 1;
    }

    static void ParseVec3(const yaml::Node& node, double out[3]) {  // rt.hpp:61-64
        if (node.size() < 3) node.fail("expected a sequence of 3 numbers");
        for (int i = 0; i < 3; i++) out[i] = node[i].as_double();
    }
    static void cross3(const double a[3], const double b[3], double o[3]) {
        o[0] = a[1] * b[2] - b[1] * a[2]; o[1] = a[2] * b[0] - b[2] * a[0]; o[2] = a[0] * b[1] - b[0] * a[1];
    }
    static void normalize3(double v[3]) {
        const double inv = 1.0 / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
    }
    static std::string dirname(const std::string& p) {
        const size_t s = p.find_last_of('/');
        return s == std::string::npos ? std::string() : p.substr(0, s + 1);
    }

    // Scene::Load(path, aspect), rt.hpp:1519
    bool Load(const std::string& path, double aspect) {
        try {
            std::ifstream in(path);
            if (!in) { error = "bad file: " + path; NGI_LOG_ERROR("YAML exception: " + error); return false; }
            std::stringstream ss; ss << in.rdbuf();
            const yaml::Node scene = yaml::Load(ss.str());
            const yaml::Node& sceneNode = scene["scene"];
            const std::string basePath = dirname(path);

            const int version = (int)scene["version"].as_int();                       // rt.hpp:1535-1540
            if (version < AppConfigVersionMin || AppConfigVersionMax < version) {
                error = "Invalid config version [Min " + std::to_string(AppConfigVersionMin) + ", Max " + std::to_string(AppConfigVersionMax) + ", Actual " + std::to_string(version) + "]";
                NGI_LOG_INFO(error);
                return false;
            }

            NGI_LOG_INFO("Load primitives");
            NGI_LOG_INDENTER();
            const yaml::Node& primitivesNode = sceneNode["primitives"];
            for (size_t i = 0; i < primitivesNode.size(); i++) {
                NgiPrimitive prim{};
                prim.first_tri = -1; prim.d_tex = -1; prim.g_tex = -1;
                bool mesh_has_uv = false;                                               // THIS primitive's mesh carries texture coordinates
                const yaml::Node& primitiveNode = primitivesNode[i];

                // ---- type, rt.hpp:1583-1616 ----
                const yaml::Node& typeNode = primitiveNode["type"];
                for (size_t j = 0; j < typeNode.size(); j++) {
                    const std::string s = typeNode[j].as_string();
                    if      (s == "D") prim.type |= NGI_TYPE_D;
                    else if (s == "G") prim.type |= NGI_TYPE_G;
                    else if (s == "S") prim.type |= NGI_TYPE_S;
                    else if (s == "L") prim.type |= NGI_TYPE_L;
                    else if (s == "E") prim.type |= NGI_TYPE_E;
                }
                if (prim.type == 0 || ((prim.type & NGI_TYPE_L) > 0 && (prim.type & NGI_TYPE_E) > 0)) {
                    error = "Invalid primitive type"; NGI_LOG_ERROR(error); return false;
                }
                if ((prim.type & NGI_TYPE_E) > 0) sensor_prim = (int)prims.size();
                if ((prim.type & NGI_TYPE_L) > 0) light_prims.push_back((int)prims.size());

                // ---- mesh, rt.hpp:1624-1738 ----
                if (primitiveNode["mesh"]) {
                    const yaml::Node& meshNode = primitiveNode["mesh"];
                    const yaml::Node& pp = meshNode["postprocess"];
                    const std::string localPath = meshNode["path"].as_string();
                    const bool gn = pp && pp["generate_normals"] ? pp["generate_normals"].as_bool() : false;
                    const bool gsn = pp && pp["generate_smooth_normals"] ? pp["generate_smooth_normals"].as_bool() : false;
                    TriMesh mesh;
                    try { mesh = LoadObj(basePath + localPath, gn, gsn, pp.defined()); }
                    catch (const std::exception& e) { error = e.what(); NGI_LOG_ERROR(error); return false; }
                    if (mesh.ignored_submeshes > 0)
                        NGI_LOG_WARN(localPath + ": " + std::to_string(mesh.ignored_submeshes) + " sub-mesh(es) after the first are ignored (reference rt.hpp:1679)");
                    prim.first_tri = (int32_t)(positions.size() / 9);
                    prim.num_tris = (int32_t)mesh.num_tris();
                    positions.insert(positions.end(), mesh.positions.begin(), mesh.positions.end());
                    normals.insert(normals.end(), mesh.normals.begin(), mesh.normals.end());
                    mesh_has_uv = !mesh.texcoords.empty();
                    if (!mesh.texcoords.empty()) { any_uv = true; texcoords.resize((size_t)prim.first_tri * 6, 0.f); texcoords.insert(texcoords.end(), mesh.texcoords.begin(), mesh.texcoords.end()); }
                    else all_uv = false;
                }

                // ---- params, rt.hpp:1767-2043 ----
                const yaml::Node& paramsNode = primitiveNode["params"];
                if ((prim.type & NGI_TYPE_L) > 0) {
                    const yaml::Node& LNode = paramsNode["L"];
                    const std::string type = LNode["type"].as_string();
                    if (type == "area") {
                        prim.l_type = NGI_L_AREA;
                        ParseVec3(LNode["area"]["Le"], prim.l_le);
                        if (prim.first_tri < 0) { error = "Area light must be associated with mesh"; NGI_LOG_ERROR(error); return false; }
                    } else if (type == "point") {
                        prim.l_type = NGI_L_POINT;
                        ParseVec3(LNode["point"]["Le"], prim.l_le);
                        ParseVec3(LNode["point"]["position"], prim.l_vec);
                    } else if (type == "directional") {                                // rt.hpp:1856
                        prim.l_type = NGI_L_DIRECTIONAL;
                        ParseVec3(LNode["directional"]["Le"], prim.l_le);
                        ParseVec3(LNode["directional"]["direction"], prim.l_vec);
                    }
                }
                if ((prim.type & NGI_TYPE_E) > 0) {
                    const yaml::Node& ENode = paramsNode["E"];
                    const std::string type = ENode["type"].as_string();
                    if (type == "pinhole") {                                           // rt.hpp:1882-1902
                        const yaml::Node& pinholeNode = ENode["pinhole"];
                        double Eye[3], Center[3], Up[3];
                        ParseVec3(pinholeNode["view"]["eye"], Eye);
                        ParseVec3(pinholeNode["view"]["center"], Center);
                        ParseVec3(pinholeNode["view"]["up"], Up);
                        prim.e_type = NGI_E_PINHOLE;
                        ParseVec3(pinholeNode["We"], prim.e_we);
                        for (int k = 0; k < 3; k++) { prim.e_position[k] = Eye[k]; prim.e_vz[k] = Eye[k] - Center[k]; }
                        prim.e_fov = pinholeNode["perspective"]["fov"].as_double() * (3.14159265358979323846264338327950288 / 180.0);
                        normalize3(prim.e_vz);
                        cross3(Up, prim.e_vz, prim.e_vx); normalize3(prim.e_vx);
                        cross3(prim.e_vz, prim.e_vx, prim.e_vy);
                        prim.e_aspect = aspect;
                    } else if (type == "area") {                                       // rt.hpp:1910-1929
                        prim.e_type = NGI_E_AREA;
                        ParseVec3(ENode["area"]["We"], prim.e_we);
                        // the reference looks at the sensor's OWN mesh (rt.hpp:1919-1924): a sensor mesh without uv would map every splat to pixel 0
                        if (prim.first_tri < 0 || !mesh_has_uv) { error = "Raw sensor must be associated with mesh with UV coordinates"; NGI_LOG_ERROR(error); return false; }
                    }
                }
                if ((prim.type & NGI_TYPE_D) > 0) {                                    // rt.hpp:1940-1957
                    const yaml::Node& DNode = paramsNode["D"];
                    if (DNode["R"]) ParseVec3(DNode["R"], prim.d_r);
                    else if (DNode["TexR"]) { if ((prim.d_tex = LoadTexture(basePath + DNode["TexR"].as_string())) < 0) return false; }   // rt.hpp:1947-1951
                    else { error = "D requires R or TexR"; NGI_LOG_ERROR(error); return false; }
                }
                if ((prim.type & NGI_TYPE_G) > 0) {                                    // rt.hpp:1965-1985
                    const yaml::Node& GNode = paramsNode["G"];
                    ParseVec3(GNode["Eta"], prim.g_eta);
                    ParseVec3(GNode["K"], prim.g_k);
                    prim.g_roughness = GNode["Roughness"].as_double();
                    if (GNode["R"]) ParseVec3(GNode["R"], prim.g_r);
                    else if (GNode["TexR"]) { if ((prim.g_tex = LoadTexture(basePath + GNode["TexR"].as_string())) < 0) return false; }   // rt.hpp:1975-1979
                    else { error = "G requires R or TexR"; NGI_LOG_ERROR(error); return false; }
                }
                if ((prim.type & NGI_TYPE_S) > 0) {                                    // rt.hpp:1993-2040
                    const yaml::Node& SNode = paramsNode["S"];
                    const std::string type = SNode["type"].as_string();
                    if (type == "reflection") {
                        prim.s_type = NGI_S_REFLECTION;
                        ParseVec3(SNode["reflection"]["R"], prim.s_r);
                    } else if (type == "refraction") {
                        prim.s_type = NGI_S_REFRACTION;
                        ParseVec3(SNode["refraction"]["R"], prim.s_r);
                        prim.s_eta1 = SNode["refraction"]["eta1"].as_double();
                        prim.s_eta2 = SNode["refraction"]["eta2"].as_double();
                    } else if (type == "fresnel") {
                        prim.s_type = NGI_S_FRESNEL;
                        ParseVec3(SNode["fresnel"]["R"], prim.s_r);
                        prim.s_eta1 = SNode["fresnel"]["eta1"].as_double();
                        prim.s_eta2 = SNode["fresnel"]["eta2"].as_double();
                    }
                }
                prims.push_back(prim);
            }
            if (any_uv) texcoords.resize(positions.size() / 9 * 6, 0.f);
        } catch (const std::exception& e) {
            error = e.what();
            NGI_LOG_ERROR("YAML exception: " + error);                                 // rt.hpp:2147-2151
            return false;
        }
        return true;
    }
};

}  // namespace ngi

// ref_gpu_bridge.hpp — the reference-side binding of INTEGRATION.md §B, compiled.
//
// TEST INFRASTRUCTURE (like the rest of oracle/): oracle/build_ref.sh compiles the reference's own src/nanogi.cpp a second time into
// oracle/_ref/nanogi_ref_gpu with this header included and ONE statement inserted in front of the `switch (Type)` of
// Renderer::Render (reference src/nanogi.cpp:203):
//
//     if (nanogi_gpu_bridge::Selected()) { nanogi_gpu_bridge::RenderOnGpu(scene, (int)Type, Params.NumSamples, Params.MaxNumVertices,
//                                                                         Params.Width, Params.Height, film); return; }
//
// Everything else of that binary — Run, the boost::program_options CLI, Scene::Load (YAML + Assimp + textures), SaveImage — is the
// reference's code. A maintainer's patch adds a `--device cpu|gpu` / `--gpus N` option instead of the two environment variables
// used here (NANOGI_DEVICE=gpu, NANOGI_GPUS=N, NANOGI_SEED=s), which keep the insertion to one line.
//
// The code below is ours, written against the reference's public structs (Scene, Primitive::Params, Mesh, Texture:
// include/nanogi/rt.hpp:148-165, :379-480, :1494-1499) and the C ABI of include/nanogi_gpu.h.
#pragma once
#include <nanogi_gpu.h>

#include <chrono>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

namespace nanogi_gpu_bridge {

inline bool Selected() {
    const char* d = std::getenv("NANOGI_DEVICE");
    return d && std::string(d) == "gpu";
}

inline void Copy3(double (&dst)[3], const glm::dvec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

// rendererType: RendererType's own values (src/nanogi.cpp:51-59) = NGI_RENDERER_* for pt / ptdirect / lt / ltdirect / bdpt
inline bool RenderOnGpu(const nanogi::Scene& scene, int rendererType, long long numSamples, int maxNumVertices, int width, int height,
                        std::vector<glm::dvec3>& film)
{
    using namespace nanogi;
    film.assign((size_t)width * height, glm::dvec3(0.0));
    if (rendererType > NGI_RENDERER_BDPT) { NGI_LOG_ERROR("renderer not supported by the GPU module (ptmnee)"); return false; }

    // (1) flatten Scene -> NgiSceneDesc exactly like Scene::Load hands geometry to Embree (rt.hpp:2095-2136): primitives in YAML
    //     order, faces in loader order, three fresh float vertices per face (the float cast of rt.hpp:2119-2121)
    std::vector<float> pos, nrm, uv;
    bool anyUv = false;
    for (const auto& prim : scene.Primitives) if (prim->MeshRef && !prim->MeshRef->Texcoords.empty()) anyUv = true;
    std::vector<NgiPrimitive> prims(scene.Primitives.size());
    auto texIndex = [&](const Texture* t) -> int32_t {
        if (!t) return -1;
        for (size_t k = 0; k < scene.Textures.size(); k++) if (scene.Textures[k].get() == t) return (int32_t)k;
        return -1;
    };
    for (size_t i = 0; i < scene.Primitives.size(); i++) {
        const Primitive& prim = *scene.Primitives[i];
        NgiPrimitive& o = prims[i];
        o = NgiPrimitive{};
        o.type = prim.Type;                                    // PrimitiveType bits = NGI_TYPE_* (rt.hpp:338-351)
        o.first_tri = -1; o.num_tris = 0; o.d_tex = -1; o.g_tex = -1;
        if (prim.MeshRef) {
            const Mesh& m = *prim.MeshRef;
            o.first_tri = (int32_t)(pos.size() / 9);
            o.num_tris = (int32_t)(m.Faces.size() / 3);
            for (unsigned int idx : m.Faces) {
                for (int k = 0; k < 3; k++) {
                    pos.push_back((float)m.Positions[3 * idx + k]);
                    nrm.push_back(m.Normals.empty() ? 0.0f : (float)m.Normals[3 * idx + k]);
                }
                if (anyUv) for (int k = 0; k < 2; k++) uv.push_back(m.Texcoords.empty() ? 0.0f : (float)m.Texcoords[2 * idx + k]);
            }
        }
        const auto& P = prim.Params;
        if (prim.Type & PrimitiveType::D) { Copy3(o.d_r, P.D.R); o.d_tex = texIndex(P.D.TexR); }
        if (prim.Type & PrimitiveType::G) {
            Copy3(o.g_r, P.G.R); o.g_tex = texIndex(P.G.TexR); Copy3(o.g_eta, P.G.Eta); Copy3(o.g_k, P.G.K); o.g_roughness = P.G.Roughness;
        }
        if (prim.Type & PrimitiveType::S) {
            o.s_type = (int32_t)P.S.Type;                      // SType values = NGI_S_* (rt.hpp:366-371)
            if (P.S.Type == SType::Reflection) Copy3(o.s_r, P.S.Reflection.R);
            else if (P.S.Type == SType::Refraction) { Copy3(o.s_r, P.S.Refraction.R); o.s_eta1 = P.S.Refraction.Eta1; o.s_eta2 = P.S.Refraction.Eta2; }
            else { Copy3(o.s_r, P.S.Fresnel.R); o.s_eta1 = P.S.Fresnel.Eta1; o.s_eta2 = P.S.Fresnel.Eta2; }
        }
        if (prim.Type & PrimitiveType::L) {
            o.l_type = (int32_t)P.L.Type;                      // LType values = NGI_L_* (rt.hpp:353-358)
            if (P.L.Type == LType::Area) Copy3(o.l_le, P.L.Area.Le);
            else if (P.L.Type == LType::Point) { Copy3(o.l_le, P.L.Point.Le); Copy3(o.l_vec, P.L.Point.Position); }
            else { Copy3(o.l_le, P.L.Directional.Le); Copy3(o.l_vec, P.L.Directional.Direction); }
        }
        if (prim.Type & PrimitiveType::E) {
            o.e_type = (int32_t)P.E.Type;                      // EType values = NGI_E_* (rt.hpp:360-364)
            if (P.E.Type == EType::Pinhole) {
                Copy3(o.e_position, P.E.Pinhole.Position);
                Copy3(o.e_vx, P.E.Pinhole.Vx); Copy3(o.e_vy, P.E.Pinhole.Vy); Copy3(o.e_vz, P.E.Pinhole.Vz);
                o.e_fov = P.E.Pinhole.Fov; o.e_aspect = P.E.Pinhole.Aspect; Copy3(o.e_we, P.E.Pinhole.We);
            } else {
                Copy3(o.e_we, P.E.Area.We);
            }
        }
    }
    std::vector<NgiTexture> textures(scene.Textures.size());
    for (size_t k = 0; k < scene.Textures.size(); k++) {
        textures[k].width = scene.Textures[k]->Width; textures[k].height = scene.Textures[k]->Height;
        textures[k].rgb = scene.Textures[k]->Data.data();
    }
    NgiSceneDesc d{};
    d.struct_size = sizeof d;
    d.num_prims = (uint32_t)prims.size();
    d.num_tris = pos.size() / 9;
    d.positions = pos.data(); d.normals = nrm.data(); d.texcoords = anyUv ? uv.data() : nullptr;
    d.prims = prims.data();
    d.num_textures = (uint32_t)textures.size(); d.textures = textures.empty() ? nullptr : textures.data();

    // (2) the scene on every GPU (built once, broadcast), then ONE render call: samples sharded by index, one NCCL film reduce
    int gpus = 1;
    if (const char* g = std::getenv("NANOGI_GPUS")) gpus = std::max(1, std::atoi(g));
    unsigned long long seed = (unsigned long long)std::time(nullptr);          // release builds seed from the clock (src/nanogi.cpp:190)
    if (const char* s = std::getenv("NANOGI_SEED")) seed = std::strtoull(s, nullptr, 10);
    void* group = nullptr;
    if (ngi_gpu_group_create(&d, nullptr, gpus, &group) != NGI_OK) { NGI_LOG_ERROR(std::string("GPU module: ") + ngi_gpu_last_error()); return false; }
    NgiRenderParams rp{};
    rp.struct_size = sizeof rp;
    rp.renderer = rendererType; rp.num_samples = numSamples; rp.sample_offset = 0;
    rp.film_norm_samples = numSamples;                         // film *= W*H/N (src/nanogi.cpp:436)
    rp.max_num_vertices = maxNumVertices; rp.width = width; rp.height = height; rp.seed = seed;
    std::vector<float> rgb((size_t)width * height * 3);
    NgiRenderStats st{};
    const int rc = ngi_gpu_group_render(group, &rp, rgb.data(), &st);
    ngi_gpu_group_destroy(group);
    if (rc != NGI_OK) { NGI_LOG_ERROR(std::string("GPU module: ") + ngi_gpu_last_error()); return false; }

    // (3) film: float RGB, row 0 = bottom — the layout of the reference's vector<dvec3> (rt.hpp:135-140)
    for (size_t i = 0; i < film.size(); i++) film[i] = glm::dvec3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    NGI_LOG_INFO("GPU module: " + std::to_string(st.paths) + " samples, " + std::to_string(st.extend_rays + st.shadow_rays) + " rays, " +
                 std::to_string(st.gpu_seconds) + " s on " + std::to_string(gpus) + " GPU(s)");
    return true;
}

}  // namespace nanogi_gpu_bridge

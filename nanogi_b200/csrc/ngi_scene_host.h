// ngi_scene_host.h — host-side preparation of the device scene from the POD NgiSceneDesc.
//
// Loader-side derivations the reference performs in Scene::Load before the Embree commit:
//   * sensor = last E primitive, light list in YAML order      (reference include/nanogi/rt.hpp:1606-1615)
//   * per-light triangle-area CDF + InvArea, computed in fp64   (rt.hpp:1747-1765, basic.hpp:448-461)
//   * directional-light bounding disk                           (rt.hpp:2067-2073)
// plus the narrowing of the fp64 parameters to the fp32 device structs. Pure host C++ (no CUDA calls):
// ngi_capi.cu uploads the arrays; tests/hostsim reuses them for the CPU simulator.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nanogi_gpu.h"
#include "ngi_shade.h"

struct NgiHostArrays {
    std::vector<float4> shade_tris;     // [n_real][5]: only filled on request (the CPU simulator); the CUDA module builds it on the device
    std::vector<int> tri_prim;          // [n_real] primitive index of every triangle
    std::vector<NgiDevPrim> prims;
    std::vector<unsigned> light_prims;
    std::vector<float> cdf;
    std::vector<float> shade_uv;        // [n_real][6] or empty
    std::vector<NgiDevTex> textures;
    std::vector<float> tex_data;        // concatenated RGB texels
    NgiDevSensor sensor;
    unsigned n_real = 0;
    std::string error;
};

inline f3 ngi_f3_from(const double* v) { return mk3((float)v[0], (float)v[1], (float)v[2]); }

// CreateTriangleAreaDist, rt.hpp:1747-1765 (fp64), Distribution1D::Normalize basic.hpp:453-461: appends the normalised
// area CDF (leading 0) of the triangle range to `cdf_out`, returns InvArea
inline float ngi_area_cdf(const NgiSceneDesc* d, const int first_tri, const int num_tris, std::vector<float>& cdf_out, int& cdf_offset) {
    std::vector<double> cdf(1, 0.0);
    double sumArea = 0;
    for (int t = 0; t < num_tris; t++) {
        const float* q = d->positions + ((size_t)first_tri + t) * 9;
        const double e1[3] = {(double)q[3] - q[0], (double)q[4] - q[1], (double)q[5] - q[2]};
        const double e2[3] = {(double)q[6] - q[0], (double)q[7] - q[1], (double)q[8] - q[2]};
        const double cx = e1[1] * e2[2] - e2[1] * e1[2], cy = e1[2] * e2[0] - e2[2] * e1[0], cz = e1[0] * e2[1] - e2[0] * e1[1];
        const double area = std::sqrt(cx * cx + cy * cy + cz * cz) * 0.5;
        cdf.push_back(cdf.back() + area);
        sumArea += area;
    }
    const double invSum = 1.0 / cdf.back();
    cdf_offset = (int)cdf_out.size();
    for (double v : cdf) cdf_out.push_back((float)(v * invSum));
    return (float)(1.0 / sumArea);
}

// shading record of triangle t: 3 positions + 3 vertex normals + primitive index, 5 x float4 (one item of k_shade_setup)
NGI_HD void ngi_shade_setup(const float* __restrict__ positions, const float* __restrict__ normals, const int* __restrict__ tri_prim,
                            const size_t t, float4* __restrict__ shade) {
    const float* p = positions + t * 9;
    const float* q = normals + t * 9;
    float4* r = shade + t * 5;
    r[0] = make_float4(p[0], p[1], p[2], p[3]);
    r[1] = make_float4(p[4], p[5], p[6], p[7]);
    r[2] = make_float4(p[8], q[0], q[1], q[2]);
    r[3] = make_float4(q[3], q[4], q[5], q[6]);
    r[4] = make_float4(q[7], q[8], u2f((unsigned)tri_prim[t]), 0.0f);
}

inline bool ngi_prepare_scene(const NgiSceneDesc* d, NgiHostArrays& out, const bool host_shade_tris = false) {
    if (!d || d->struct_size != sizeof(NgiSceneDesc)) { out.error = "NgiSceneDesc.struct_size mismatch (ABI)"; return false; }
    if (d->num_tris > 0 && (!d->positions || !d->normals)) { out.error = "positions / normals are NULL"; return false; }
    if (d->num_prims == 0 || !d->prims) { out.error = "scene has no primitives"; return false; }
    if (d->num_tris >= 0x7FFFFFF0ull) { out.error = "too many triangles (max 2^31 - 16)"; return false; }
    const size_t n = (size_t)d->num_tris;
    out.n_real = (unsigned)n;
    std::vector<int>& triPrim = out.tri_prim;
    triPrim.assign(n, -1);
    out.prims.resize(d->num_prims);
    int sensor = -1;
    // scene bounds: only the bounding disk of a directional light needs them (rt.hpp:2067-2073)
    double bmin[3] = {1e300, 1e300, 1e300}, bmax[3] = {-1e300, -1e300, -1e300};
    bool any_directional = false;
    for (uint32_t i = 0; i < d->num_prims; i++) if ((d->prims[i].type & NGI_TYPE_L) && d->prims[i].l_type == NGI_L_DIRECTIONAL) any_directional = true;
    if (any_directional)
        for (size_t i = 0; i < n * 3; i++)
            for (int k = 0; k < 3; k++) {
                const double v = d->positions[i * 3 + k];
                if (v < bmin[k]) bmin[k] = v;
                if (v > bmax[k]) bmax[k] = v;
            }
    for (uint32_t i = 0; i < d->num_prims; i++) {
        const NgiPrimitive& s = d->prims[i];
        NgiDevPrim p;
        std::memset(&p, 0, sizeof(p));
        p.type = s.type;
        p.first_tri = s.first_tri; p.num_tris = s.first_tri >= 0 ? s.num_tris : 0;
        p.l_type = s.l_type; p.s_type = s.s_type; p.cdf_offset = -1;
        if (p.first_tri >= 0 && ((size_t)p.first_tri + (size_t)p.num_tris > n || p.num_tris < 0)) { out.error = "primitive triangle range out of bounds"; return false; }
        p.d_tex = p.g_tex = -1;
        if (s.d_tex >= 0 || s.g_tex >= 0) {                                            // D.TexR / G.TexR, rt.hpp:1947-1951, :1975-1979
            if ((s.d_tex >= 0 && (uint32_t)s.d_tex >= d->num_textures) || (s.g_tex >= 0 && (uint32_t)s.g_tex >= d->num_textures)) { out.error = "texture index out of range"; return false; }
            if (!d->texcoords) { out.error = "a textured primitive needs texture coordinates (NgiSceneDesc.texcoords)"; return false; }
            p.d_tex = s.d_tex; p.g_tex = s.g_tex;
        }
        p.d_r = ngi_f3_from(s.d_r);
        p.g_r = ngi_f3_from(s.g_r); p.g_eta = ngi_f3_from(s.g_eta); p.g_k = ngi_f3_from(s.g_k); p.g_rough = (float)s.g_roughness;
        p.s_r = ngi_f3_from(s.s_r); p.s_eta1 = (float)s.s_eta1; p.s_eta2 = (float)s.s_eta2;
        p.l_le = ngi_f3_from(s.l_le); p.l_vec = ngi_f3_from(s.l_vec);
        for (int t = 0; t < p.num_tris; t++) triPrim[(size_t)p.first_tri + t] = (int)i;
        if (s.type & NGI_TYPE_E) {
            if (s.e_type != NGI_E_PINHOLE && s.e_type != NGI_E_AREA) { out.error = "unknown sensor type"; return false; }
            if (s.e_type == NGI_E_AREA && (p.num_tris <= 0 || !d->texcoords)) {        // rt.hpp:1919-1924
                out.error = "Raw sensor must be associated with mesh with UV coordinates"; return false;
            }
            sensor = (int)i;                                                           // rt.hpp:1606-1610 (last one wins)
        }
        if (s.type & NGI_TYPE_L) {
            out.light_prims.push_back(i);                                              // rt.hpp:1612-1615
            if (s.l_type == NGI_L_AREA) {
                if (p.num_tris <= 0) { out.error = "Area light must be associated with mesh"; return false; }   // rt.hpp:1826-1830
                p.l_inv_area = ngi_area_cdf(d, p.first_tri, p.num_tris, out.cdf, p.cdf_offset);
            } else if (s.l_type == NGI_L_DIRECTIONAL) {                                // rt.hpp:2067-2073
                double c[3], r2 = 0;
                for (int k = 0; k < 3; k++) { c[k] = (bmax[k] + bmin[k]) * 0.5; r2 += (c[k] - bmax[k]) * (c[k] - bmax[k]); }
                const double radius = std::sqrt(r2) * 1.01;
                p.l_center = mk3((float)c[0], (float)c[1], (float)c[2]);
                p.l_radius = (float)radius;
                p.l_inv_area = (float)(1.0 / (2.0 * 3.14159265358979323846 * radius * radius));
            }
        }
        out.prims[i] = p;
    }
    for (uint32_t i = 0; i < d->num_textures; i++) {
        const NgiTexture& t = d->textures[i];
        if (t.width <= 0 || t.height <= 0 || !t.rgb) { out.error = "invalid texture"; return false; }
        NgiDevTex T; T.offset = (int)(out.tex_data.size() / 3); T.width = t.width; T.height = t.height; T.pad = 0;
        out.textures.push_back(T);
        out.tex_data.insert(out.tex_data.end(), t.rgb, t.rgb + (size_t)t.width * t.height * 3);
    }
    if (sensor < 0) { out.error = "scene has no sensor (E) primitive"; return false; }
    const bool area_sensor = d->prims[sensor].e_type == NGI_E_AREA;
    if (d->texcoords && (d->num_textures > 0 || area_sensor)) out.shade_uv.assign(d->texcoords, d->texcoords + n * 6);
    {
        const NgiPrimitive& s = d->prims[sensor];
        NgiDevSensor& E = out.sensor;
        std::memset(&E, 0, sizeof(E));
        E.kind = area_sensor ? NGI_ET_AREA : NGI_ET_PINHOLE;
        E.cdf_offset = -1;
        if (area_sensor) {
            E.first_tri = out.prims[sensor].first_tri; E.num_tris = out.prims[sensor].num_tris;
            E.inv_area = ngi_area_cdf(d, E.first_tri, E.num_tris, out.cdf, E.cdf_offset);
            E.we = ngi_f3_from(s.e_we);
        }
        E.px = s.e_position[0]; E.py = s.e_position[1]; E.pz = s.e_position[2];
        E.vx = ngi_f3_from(s.e_vx); E.vy = ngi_f3_from(s.e_vy); E.vz = ngi_f3_from(s.e_vz);
        const double tanFov = std::tan(s.e_fov * 0.5);
        E.tan_fov = (float)tanFov;
        E.aspect = (float)s.e_aspect;
        E.inv_a = (float)(1.0 / (tanFov * tanFov * s.e_aspect * 4.0));                  // rt.hpp:974
        E.prim = sensor;
    }
    for (size_t t = 0; t < n; t++)
        if (triPrim[t] < 0) { out.error = "triangle " + std::to_string(t) + " belongs to no primitive"; return false; }
    if (host_shade_tris) {
        out.shade_tris.resize(n * 5);
        for (size_t t = 0; t < n; t++) ngi_shade_setup(d->positions, d->normals, triPrim.data(), t, out.shade_tris.data());
    }
    return true;
}

// conservative box padding: the triangle test's rounding is ~ulp(|o - v0|), so every
// triangle box grows by 2^-16 of the largest coordinate magnitude in play (scene bounds, sensor position)
inline float ngi_box_pad(const float smin[3], const float smax[3], const NgiDevSensor& E) {
    float mag = 1e-30f;
    for (int k = 0; k < 3; k++) { mag = fmaxf(mag, fabsf(smin[k])); mag = fmaxf(mag, fabsf(smax[k])); }
    mag = fmaxf(mag, (float)fmax(fabs(E.px), fmax(fabs(E.py), fabs(E.pz))));
    return mag * (1.0f / 65536.0f);
}

// ngi_shade.h — fp32 device restatement of nanogi's Primitive / emitter / BSDF functions.
//
// Replaces, for the pt / ptdirect path (and lt / ltdirect), the sampling and evaluation code of `struct Primitive`
// (reference include/nanogi/rt.hpp:379-1470) and the helpers at rt.hpp:55-140, :282-302, :2321-2374.
// Each function cites the lines it restates. Arithmetic is fp32 (the reference is fp64) except the path
// vertex position, which stays fp64 exactly like the reference's `geom.p` (see ngi_wave.h). All the
// reference's quirks are kept on purpose: one-sided lights about the SHADING normal,
// the Beckmann shadow-masking typo (rt.hpp:1428-1429), the possibly-negative G pdf (rt.hpp:1250), the
// (eta_i/eta_t)^2 radiance scaling, "wo not written" == path termination.
#pragma once
#include "ngi_math.h"

// ---- device-resident scene --------------------------------------------------------------------
enum { NGI_D = 1, NGI_G = 2, NGI_S = 4, NGI_L = 8, NGI_E = 16, NGI_BSDF = 7, NGI_EMITTER = 24 };
enum { NGI_LT_AREA = 0, NGI_LT_POINT = 1, NGI_LT_DIRECTIONAL = 2 };
enum { NGI_ST_REFLECTION = 0, NGI_ST_REFRACTION = 1, NGI_ST_FRESNEL = 2 };

struct NgiDevPrim {           // 144 bytes, 16-byte aligned rows
    int type, first_tri, num_tris, l_type;
    int s_type, cdf_offset, d_tex, g_tex;   // *_tex: index into NgiDevScene::textures, -1 = constant colour
    f3 d_r;   float g_rough;
    f3 g_r;   float s_eta1;
    f3 g_eta; float s_eta2;
    f3 g_k;   float l_inv_area;
    f3 s_r;   float l_radius;
    f3 l_le;  float pad2;
    f3 l_vec; float pad3;     // L.Point.Position or L.Directional.Direction
    f3 l_center; float pad4;
};

enum { NGI_ET_AREA = 0, NGI_ET_PINHOLE = 1 };
struct NgiDevSensor {         // E.Pinhole, rt.hpp:422-429 / E.Area, rt.hpp:412-420
    double px, py, pz;
    f3 vx, vy, vz;
    float tan_fov;            // tan(Fov/2)
    float aspect;
    float inv_a;              // 1 / (tan^2 * aspect * 4), rt.hpp:974
    int prim;
    int kind;                 // NGI_ET_*
    int first_tri, num_tris, cdf_offset;   // E.Area: the sensor mesh and its area CDF (rt.hpp:1926-1927)
    float inv_area;           // E.Area.InvArea
    f3 we;                    // E.Area.We (rt.hpp:947-953); the pinhole's We is parsed but unused (rt.hpp:955-978)
};

// per-triangle shading record, indexed by GLOBAL triangle id: 5 x float4 = 80 B
//   r0 = (v0.xyz, v1.x) r1 = (v1.yz, v2.xy) r2 = (v2.z, n0.xyz) r3 = (n1.xyz, n2.x) r4 = (n2.yz, prim bits, -)
// nearest-neighbour RGB texture (Texture, rt.hpp:157-270): texel (x, y) at tex_data[3 * (offset + y * width + x)], row 0 = top
struct NgiDevTex { int offset, width, height, pad; };

struct NgiDevScene {
    const uint4* nodes8; const float4* tris8;
    const float4* nodes2; const float4* tris2;
    const float4* shade_tris;
    const NgiDevPrim* prims;
    const unsigned* light_prims;
    const float* cdf;          // concatenated per-light normalised CDFs (leading 0 each)
    const float* shade_uv;     // [n_tris][3 verts][2] texture coordinates by global triangle id, or NULL
    const NgiDevTex* textures; const float* tex_data;
    unsigned n_tris, n_lights;
    NgiDevSensor sensor;
};

struct NgiGeom {               // SurfaceGeometry, rt.hpp:282-302 (p lives in fp64 in the path state)
    f3 sn, gn, dpdu, dpdv;
    f3 albedo;                 // R of the BSDF evaluated at this point: D.R / G.R or their TexR at geom.uv (rt.hpp:1022, :1045)
};
// R of the lobe `type` resolves to (precedence D > G like the evaluation itself) without a texture
NGI_HD f3 ngi_constant_albedo(const NgiDevPrim& P, const int type) { return (type & NGI_D) ? P.d_r : P.g_r; }

// ---- helpers ----------------------------------------------------------------------------------
// rt.hpp:55-59
NGI_HD void ngi_orthonormal_basis(const f3 a, f3& b, f3& c) {
    c = fabsf(a.x) > fabsf(a.y) ? normalize(mk3(a.z, 0.0f, -a.x)) : normalize(mk3(0.0f, a.z, -a.y));
    b = normalize(cross(c, a));
}
NGI_HD void ngi_tangent_space(NgiGeom& g) { ngi_orthonormal_basis(g.sn, g.dpdu, g.dpdv); }   // rt.hpp:295-300
NGI_HD f3 ngi_to_local(const NgiGeom& g, const f3 w) { return mk3(dot(g.dpdu, w), dot(g.dpdv, w), dot(g.sn, w)); }
NGI_HD f3 ngi_to_world(const NgiGeom& g, const f3 l) { return g.dpdu * l.x + g.dpdv * l.y + g.sn * l.z; }
// rt.hpp:71-75
NGI_HD float ngi_local_tan(const f3 v) { const float t = 1.0f - v.z * v.z; return t <= 0.0f ? 0.0f : ngi_divf(ngi_sqrtf(t), v.z); }

// rt.hpp:87-103
NGI_HD void ngi_concentric_disk(const float u0, const float u1, float& sx, float& sy) {
    const float vx = 2.0f * u0 - 1.0f, vy = 2.0f * u1 - 1.0f;
    if (vx == 0.0f && vy == 0.0f) { sx = 0.0f; sy = 0.0f; return; }
    float r, theta;
    if (vx > -vy) {
        if (vx > vy) { r = vx; theta = ngi_divf((NGI_PI_F * 0.25f) * vy, vx); }
        else         { r = vy; theta = (NGI_PI_F * 0.25f) * (2.0f - ngi_divf(vx, vy)); }
    } else {
        if (vx < vy) { r = -vx; theta = (NGI_PI_F * 0.25f) * (4.0f + ngi_divf(vy, vx)); }
        else         { r = -vy; theta = (NGI_PI_F * 0.25f) * (6.0f - ngi_divf(vx, vy)); }
    }
    float s, c;
    ngi_sincosf(theta, &s, &c);                                                               // theta in [-pi/4, 7 pi/4]
    sx = r * c; sy = r * s;
}
// rt.hpp:105-109
NGI_HD f3 ngi_cosine_hemisphere(const float u0, const float u1) {
    float sx, sy;
    ngi_concentric_disk(u0, u1, sx, sy);
    return mk3(sx, sy, ngi_sqrtf(fmaxf(0.0f, 1.0f - sx * sx - sy * sy)));
}
// rt.hpp:116-122
NGI_HD f3 ngi_uniform_sphere(const float u0, const float u1) {
    const float z = 1.0f - 2.0f * u0;
    const float r = ngi_sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    ngi_sincosf(2.0f * NGI_PI_F * u1, &s, &c);
    return mk3(r * c, r * s, z);
}
// rt.hpp:135-140
NGI_HD int ngi_pixel_index(const float rx, const float ry, const int w, const int h) {
    const int pX = clampi((int)(rx * (float)w), 0, w - 1);
    const int pY = clampi((int)(ry * (float)h), 0, h - 1);
    return pY * w + pX;
}

// RasterPosition (pinhole), rt.hpp:1344-1378
NGI_HD bool ngi_raster_position(const NgiDevSensor& E, const f3 wo, float& rx, float& ry, float& cosTheta) {
    const f3 woEye = mk3(dot(E.vx, wo), dot(E.vy, wo), dot(E.vz, wo));
    if (woEye.z >= 0.0f) return false;
    rx = (ngi_divf(ngi_divf(ngi_divf(-woEye.x, woEye.z), E.tan_fov), E.aspect) + 1.0f) * 0.5f;   // (SFU quotients move a sample across a
    ry = (ngi_divf(ngi_divf(-woEye.y, woEye.z), E.tan_fov) + 1.0f) * 0.5f;                        //  pixel edge it is within ~3e-7 of, like any fp32 rounding)
    cosTheta = -woEye.z;
    if (rx < 0.0f || rx > 1.0f || ry < 0.0f || ry > 1.0f) return false;
    return true;
}
// EvaluateDirection / EvaluateDirectionPDF for E.pinhole: 1/(cos^3 A), rt.hpp:955-978, :1189-1212
NGI_HD float ngi_pinhole_importance(const NgiDevSensor& E, const f3 wo, float& rx, float& ry) {
    float cosTheta;
    if (!ngi_raster_position(E, wo, rx, ry, cosTheta)) return 0.0f;
    const float inv = ngi_rcpf(cosTheta);
    return inv * inv * inv * E.inv_a;
}
// SampleDirection for E.pinhole, rt.hpp:733-740
NGI_HD f3 ngi_pinhole_sample(const NgiDevSensor& E, const float u0, const float u1) {
    const float rx = 2.0f * u0 - 1.0f, ry = 2.0f * u1 - 1.0f;
    const f3 woEye = normalize(mk3(E.aspect * E.tan_fov * rx, E.tan_fov * ry, -1.0f));
    return E.vx * woEye.x + E.vy * woEye.y + E.vz * woEye.z;
}

// ---- type G helpers, rt.hpp:1407-1442 ---------------------------------------------------------
NGI_HD float ngi_beckmann(const float rough, const f3 H) {                                   // :1407-1414
    if (H.z <= 0.0f) return 0.0f;
    const float ex = ngi_divf(ngi_local_tan(H), rough);
    const float t1 = expf(-(ex * ex));
    const float c2 = H.z * H.z;
    const float t2 = NGI_PI_F * rough * rough * (c2 * c2);
    return ngi_divf(t1, t2);
}
NGI_HD float ngi_shadow_masking(const f3 wi, const f3 wo, const f3 H) {                       // :1423-1431 (typo kept)
    const float n_dot_H = H.z, n_dot_wo = wo.z, n_dot_wi = wi.z;
    const float wo_dot_H = fabsf(dot(wo, H));
    const float wi_dot_H = fabsf(dot(wo, H));  // sic: the reference uses `wo` here too
    return fminf(1.0f, fminf(ngi_divf(2.0f * n_dot_H * n_dot_wo, wo_dot_H), ngi_divf(2.0f * n_dot_H * n_dot_wi, wi_dot_H)));
}
NGI_HD f3 ngi_fr_conductor(const f3 eta, const f3 k, const float cosThetaI) {                 // :1433-1442
    const f3 e2k2 = eta * eta + k * k;
    const f3 tmp = e2k2 * (cosThetaI * cosThetaI);
    const f3 twoEtaCos = eta * (2.0f * cosThetaI);
    const f3 rParl2 = (tmp - twoEtaCos + 1.0f) / (tmp + twoEtaCos + 1.0f);
    const f3 rPerp2 = (e2k2 - twoEtaCos + cosThetaI * cosThetaI) / (e2k2 + twoEtaCos + cosThetaI * cosThetaI);
    return (rParl2 + rPerp2) * 0.5f;
}
// rt.hpp:1450-1466
NGI_HD float ngi_fresnel(const float wiDotN, const float etaI, const float etaT) {
    const float eta = ngi_divf(etaI, etaT);
    const float cosThetaTSq = 1.0f - eta * eta * (1.0f - wiDotN * wiDotN);
    if (cosThetaTSq <= 0.0f) return 1.0f;
    const float absCosThetaI = fabsf(wiDotN);
    const float absCosThetaT = ngi_sqrtf(cosThetaTSq);
    const float rhoS = ngi_divf(etaI * absCosThetaI - etaT * absCosThetaT, etaI * absCosThetaI + etaT * absCosThetaT);
    const float rhoT = ngi_divf(etaI * absCosThetaT - etaT * absCosThetaI, etaI * absCosThetaT + etaT * absCosThetaI);
    return (rhoS * rhoS + rhoT * rhoT) * 0.5f;
}

// ---- Primitive::SampleDirection for BSDF types, rt.hpp:747-903 --------------------------------
// returns false when the reference returns without writing `wo` (=> fs = 0 => the path ends)
NGI_HD bool ngi_sample_bsdf(const NgiDevPrim& P, const int type, const NgiGeom& g, const f3 wi, const float u0, const float u1,
                            const float uComp, f3& wo) {
    const f3 localWi = ngi_to_local(g, wi);
    if (type & NGI_D) {                                                                       // :749-761
        if (localWi.z <= 0.0f) return false;
        wo = ngi_to_world(g, ngi_cosine_hemisphere(u0, u1));
        return true;
    }
    if (type & NGI_G) {                                                                       // :769-796
        if (localWi.z <= 0.0f) return false;
        const float tanThetaHSqr = -P.g_rough * P.g_rough * logf(1.0f - u0);
        const float cosThetaH = ngi_rsqrtf(1.0f + tanThetaHSqr);
        const float sinThetaH = ngi_sqrtf(fmaxf(0.0f, 1.0f - cosThetaH * cosThetaH));
        float s, c;
        ngi_sincosf(2.0f * NGI_PI_F * u1, &s, &c);
        const f3 H = mk3(sinThetaH * c, sinThetaH * s, cosThetaH);
        const f3 localWo = -localWi - H * (2.0f * dot(-localWi, H));
        if (localWo.z <= 0.0f) return false;
        wo = ngi_to_world(g, localWo);
        return true;
    }
    if (type & NGI_S) {
        if (P.s_type == NGI_ST_REFLECTION) {                                                  // :808-820
            if (localWi.z <= 0.0f) return false;
            wo = ngi_to_world(g, mk3(-localWi.x, -localWi.y, localWi.z));
            return true;
        }
        float etaI = P.s_eta1, etaT = P.s_eta2;
        if (localWi.z < 0.0f) { const float t = etaI; etaI = etaT; etaT = t; }
        const float wiDotN = localWi.z;
        const float eta = ngi_divf(etaI, etaT);
        const float cosThetaTSq = 1.0f - eta * eta * (1.0f - wiDotN * wiDotN);
        if (P.s_type == NGI_ST_REFRACTION) {                                                  // :828-859
            if (cosThetaTSq <= 0.0f) { wo = ngi_to_world(g, mk3(-localWi.x, -localWi.y, localWi.z)); return true; }
            const float cosThetaT = ngi_sqrtf(cosThetaTSq) * (wiDotN > 0.0f ? -1.0f : 1.0f);
            wo = ngi_to_world(g, mk3(-eta * localWi.x, -eta * localWi.y, cosThetaT));
            return true;
        }
        if (P.s_type == NGI_ST_FRESNEL) {                                                     // :867-900
            const float Fr = ngi_fresnel(wiDotN, etaI, etaT);
            if (uComp <= Fr) {
                wo = ngi_to_world(g, mk3(-localWi.x, -localWi.y, localWi.z));
            } else {
                const float cosThetaT = ngi_sqrtf(cosThetaTSq) * (wiDotN > 0.0f ? -1.0f : 1.0f);
                wo = ngi_to_world(g, mk3(-eta * localWi.x, -eta * localWi.y, cosThetaT));
            }
            return true;
        }
    }
    return false;  // type 0 (a pure [L] primitive after `& ~Emitter`): assert(0) in the reference, rt.hpp:909
}

// ---- Primitive::EvaluateDirection (BSDF types) + EvaluateDirectionPDF ----
// rt.hpp:990-1140 and :1219-1328. `transLE` = TransportDirection::LE (light paths: lt / ltdirect), a literal at every
// call site; the eye-path renderers pass false (EL).
NGI_HD f3 ngi_eval_bsdf(const NgiDevPrim& P, const int type, const NgiGeom& g, const f3 wi, const f3 wo, const bool forceDegenerated,
                        float& pdf, const bool transLE = false) {
    pdf = 0.0f;
    if (!(type & NGI_BSDF)) return mk3(0.0f);
    const f3 localWi = ngi_to_local(g, wi), localWo = ngi_to_local(g, wo);
    // shadingNormalCorrection, :994-1005 (EL => 1 unless the sides disagree; LE => wiDotNs * woDotNg / (woDotNs * wiDotNg))
    const float wiDotNg = dot(wi, g.gn), woDotNg = dot(wo, g.gn);
    float snc = (wiDotNg * localWi.z <= 0.0f || woDotNg * localWo.z <= 0.0f) ? 0.0f : 1.0f;
    if (transLE && snc != 0.0f) snc = ngi_divf(localWi.z * woDotNg, localWo.z * wiDotNg);
    if (type & NGI_D) {                                                                       // :1013-1024, :1221-1231
        if (localWi.z <= 0.0f || localWo.z <= 0.0f) return mk3(0.0f);
        pdf = NGI_INV_PI_F;
        return g.albedo * (NGI_INV_PI_F * snc);
    }
    if (type & NGI_G) {                                                                       // :1032-1047, :1239-1251
        if (localWi.z <= 0.0f || localWo.z <= 0.0f) return mk3(0.0f);
        const f3 H = normalize(localWi + localWo);
        const float D = ngi_beckmann(P.g_rough, H);
        const float G = ngi_shadow_masking(localWi, localWo, H);
        const f3 F = ngi_fr_conductor(P.g_eta, P.g_k, dot(localWi, H));
        pdf = ngi_divf(ngi_divf(D * H.z, 4.0f * dot(localWo, H)), localWo.z);
        return g.albedo * F * (ngi_divf(ngi_divf(D * G, 4.0f * localWi.z), localWo.z) * snc);
    }
    if (type & NGI_S) {
        if (!forceDegenerated) return mk3(0.0f);                                              // :1057-1060, :1261-1264
        if (P.s_type == NGI_ST_REFLECTION) {                                                  // :1066-1076, :1270-1280
            if (localWi.z <= 0.0f || localWo.z <= 0.0f) return mk3(0.0f);
            pdf = 1.0f;
            return P.s_r * snc;
        }
        float etaI = P.s_eta1, etaT = P.s_eta2;
        if (localWi.z < 0.0f) { const float t = etaI; etaI = etaT; etaT = t; }
        const float eta = ngi_divf(etaI, etaT);
        const float refr = transLE ? 1.0f : eta;                                              // refrCorrection, :1095 / :1130
        if (P.s_type == NGI_ST_REFRACTION) {                                                  // :1084-1098, :1288-1291
            pdf = 1.0f;
            return P.s_r * (snc * refr * refr);
        }
        if (P.s_type == NGI_ST_FRESNEL) {                                                     // :1106-1134, :1299-1325
            const float Fr = ngi_fresnel(localWi.z, etaI, etaT);
            if (localWi.z * localWo.z >= 0.0f) { pdf = Fr; return P.s_r * (Fr * snc); }
            pdf = 1.0f - Fr;
            return P.s_r * ((1.0f - Fr) * snc * refr * refr);
        }
    }
    return mk3(0.0f);
}

// ---- Distribution1D::SampleReuse over a light's area CDF, basic.hpp:469-475 --------------------
NGI_HD int ngi_cdf_sample_reuse(const float* __restrict__ cdf, const int count /* entries incl. leading 0 */, const float u, float& u2) {
    // upper_bound(cdf, u) - 1
    int lo = 0, hi = count;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ngi_ldg(cdf + mid) <= u) lo = mid + 1; else hi = mid;
    }
    const int i = clampi(lo - 1, 0, count - 2);
    const float c0 = ngi_ldg(cdf + i), c1 = ngi_ldg(cdf + i + 1);
    u2 = fminf(ngi_divf(u - c0, c1 - c0), 1.0f);                                              // (< 1 by construction; the SFU quotient can exceed it by an ulp)
    return i;
}

struct NgiLightSample {
    f3 p;        // sampled position (fp32)
    f3 n;        // sn = gn of the sampled point (face normal, rt.hpp:519-523)
    f3 le;
    float pdf;   // pdfL * pdfPL
    int degenerate;   // point light
    int valid;        // 0 for lights that contribute nothing on this path (directional, rt.hpp:934-937)
    int prim;         // index of the light primitive
    int l_type;
};

// SampleTriangleMesh (rt.hpp:488-525): triangle by area CDF with sample reuse, uniform point on it, face normal.
// `tri` returns the global triangle id (for the uv of an E.area sensor point). `pd` (optional) receives the position
// interpolated in fp64 like the reference (rt.hpp:508): a point that becomes a ray ORIGIN (first vertex of a light path,
// E.area sensor point) must lie on its triangle to well below the 1e-4 ray epsilon — the fp32 interpolation is off the
// plane by 1-2 ulp (1.2e-4 at Cornell coordinates), enough for the first ray to hit its own light.
NGI_HD void ngi_sample_triangle_mesh(const NgiDevScene& sc, const int first_tri, const int num_tris, const int cdf_offset,
                                     const float u0, const float u1, f3& p, f3& n, int& tri, float& bx, float& by, double* pd = nullptr) {
    float u2;
    const int i = ngi_cdf_sample_reuse(sc.cdf + cdf_offset, num_tris + 1, u0, u2);
    const float s = ngi_sqrtf(fmaxf(0.0f, u2));                                                   // UniformSampleTriangle, rt.hpp:129-133
    bx = 1.0f - s; by = u1 * s;
    tri = first_tri + i;
    const float4* r = sc.shade_tris + 5 * (size_t)tri;
    const float4 r0 = ngi_ldg(r), r1 = ngi_ldg(r + 1), r2 = ngi_ldg(r + 2);
    const f3 p1 = mk3(r0.x, r0.y, r0.z), p2 = mk3(r0.w, r1.x, r1.y), p3 = mk3(r1.z, r1.w, r2.x);
    p = p1 * (1.0f - bx - by) + p2 * bx + p3 * by;                                            // rt.hpp:508
    n = normalize(cross(p2 - p1, p3 - p1));                                                   // rt.hpp:521-522
    if (pd) {
        const double b1 = (double)bx, b2 = (double)by, b0 = 1.0 - b1 - b2;
        pd[0] = (double)p1.x * b0 + (double)p2.x * b1 + (double)p3.x * b2;
        pd[1] = (double)p1.y * b0 + (double)p2.y * b1 + (double)p3.y * b2;
        pd[2] = (double)p1.z * b0 + (double)p2.z * b1 + (double)p3.z * b2;
    }
}

// Scene::SampleEmitter (rt.hpp:2321-2336) + Primitive::SamplePosition for L (rt.hpp:483-563) + the pdfs.
// `emission` = the sample starts a light path (lt / ltdirect): directional lights then get their disk position
// (rt.hpp:549-562); as an NEE sample they contribute nothing (rt.hpp:934-937) and stay invalid.
NGI_HD NgiLightSample ngi_sample_light(const NgiDevScene& sc, const float uPick, const float u0, const float u1, const bool emission = false,
                                       double* pd = nullptr) {
    NgiLightSample ls;
    ls.valid = 0; ls.degenerate = 0; ls.pdf = 1.0f; ls.p = mk3(0.0f); ls.n = mk3(0.0f); ls.le = mk3(0.0f);
    const int n = (int)sc.n_lights;
    const int li = clampi((int)(uPick * (float)n), 0, n - 1);
    ls.prim = (int)ngi_ldg(sc.light_prims + li);
    const NgiDevPrim& L = sc.prims[ls.prim];
    const float pdfL = ngi_rcpf((float)n);                                                      // rt.hpp:2338-2344
    ls.le = L.l_le;
    ls.l_type = L.l_type;
    if (L.l_type == NGI_LT_AREA) {
        int tri; float bx, by;
        ngi_sample_triangle_mesh(sc, L.first_tri, L.num_tris, L.cdf_offset, u0, u1, ls.p, ls.n, tri, bx, by, pd);
        ls.pdf = pdfL * L.l_inv_area;
        ls.valid = 1;
        return ls;
    } else if (L.l_type == NGI_LT_POINT) {
        ls.p = L.l_vec; ls.degenerate = 1; ls.pdf = pdfL; ls.valid = 1;                       // rt.hpp:542-547, :654-657
    } else if (emission) {                                                                    // rt.hpp:549-562
        float dx, dy;
        ngi_concentric_disk(u0, u1, dx, dy);
        f3 b, c;
        ngi_orthonormal_basis(L.l_vec, b, c);
        ls.n = L.l_vec;
        ls.p = L.l_center - L.l_vec * L.l_radius + b * (dx * L.l_radius) + c * (dy * L.l_radius);
        ls.pdf = pdfL * L.l_inv_area;
        ls.valid = 1;
    }
    // directional as an NEE sample: EvaluateDirection(..., forceDegenerated=false) = 0 (rt.hpp:934-937) => contributes nothing
    if (pd) { pd[0] = (double)ls.p.x; pd[1] = (double)ls.p.y; pd[2] = (double)ls.p.z; }
    return ls;
}

// ---- E.area sensor (rt.hpp:573-577, :947-953, :1179-1187, :1386-1391) ---------------------------
// geom.uv of a point on triangle `tri` with barycentrics (bx -> 2nd vertex, by -> 3rd), rt.hpp:512-517 / :2221-2227
NGI_HD void ngi_uv_at(const NgiDevScene& sc, const unsigned tri, const float bx, const float by, double& tu, double& tv) {
    const float* t = sc.shade_uv + 6 * (size_t)tri;
    const double w = (double)(1.0f - bx - by);
    tu = (double)ngi_ldg(t + 0) * w + (double)ngi_ldg(t + 2) * (double)bx + (double)ngi_ldg(t + 4) * (double)by;
    tv = (double)ngi_ldg(t + 1) * w + (double)ngi_ldg(t + 3) * (double)bx + (double)ngi_ldg(t + 5) * (double)by;
}
// RasterPosition(E.area) = geom.uv (rt.hpp:1386-1391) -> PixelIndex (rt.hpp:135-140)
NGI_HD int ngi_area_sensor_pixel(const NgiDevScene& sc, const unsigned tri, const float bx, const float by, const int w, const int h) {
    double tu, tv;
    ngi_uv_at(sc, tri, bx, by, tu, tv);
    const int pX = clampi((int)(tu * (double)w), 0, w - 1);
    const int pY = clampi((int)(tv * (double)h), 0, h - 1);
    return pY * w + pX;
}

// ngi_wave.h — per-slot body of the wavefront "logic" stage.
//
// Replaces the per-sample loops ProcessSample_PT (reference src/nanogi.cpp:446-607) and
// ProcessSample_PTDirect (src/nanogi.cpp:609-802) plus the sample scheduler of RenderProcess
// (src/nanogi.cpp:225-440). One reference loop iteration is split at the ray query:
//
//   logic(slot):   [tail of iteration k]   miss -> end | pt: emission on L hit (:566-577) | RR (:581-591) |
//                                          vertex update (:597-603) | vertex cap (:485 / :647)
//                  [head of iteration k+1] ptdirect: NEE sample -> shadow queue (:654-712) |
//                                          SampleDirection + fs/pdf (:492-544 / :716-754) -> extend ray
//                  a slot whose path ended pulls the next sample index and starts at the eye vertex
//   extend(slot):  Scene::Intersect's ray query (include/nanogi/rt.hpp:2162-2182)   [trace kernel]
//   shadow(entry): Scene::Visible (rt.hpp:2251-2261) + film accumulation (:706)     [trace kernel]
//
// Random numbers: Philox4x32-10, key = seed, counter = (sample index, vertex index, block):
//   block 0 = {direction u0, direction u1, uComp, RR}, block 1 = {light pick, light u0, light u1, -},
//   block 2 = {sensor pick (one sensor: unused), sensor u0, sensor u1, -} (E.area sensors; ltdirect at every vertex).
// lt / ltdirect (src/nanogi.cpp:804-1131) run on the same stages: the "eye" stage then starts a
// path on a sampled light point, the surface stage connects to the sensor instead of a light (ltdirect) and the
// classify stage splats hits of an E.area sensor (lt).
// The vertex position is carried in fp64 and advanced as p += d * (double)t exactly like the reference
// (rt.hpp:2197), then narrowed to fp32 for the ray query (rt.hpp:2166-2168); everything else is fp32.
#pragma once
#include "ngi_bvh.h"
#include "ngi_shade.h"

#define NGI_INFO_ALIVE 1u
#define NGI_INFO_RR_SURVIVE 2u

// Path state of one slot: two 32-byte records + the hit record. 32 bytes = one DRAM / L2 sector: the stage that (re)starts a
// path writes whole sectors, so L2 never has to fetch a sector from HBM just to merge a partial write into it. (The first
// layout was one array per field — sample, px, py, pz at 8 B, thr_pix and dir_info at 16 B. The eye stage writes only the
// slots whose path ended, ~half of them in a random pattern, so nearly every sector it touched was written partially:
// profiles/r01_ncu_c2_final.txt shows it READING 98 B from DRAM per entry — fills — and writing 155 B for 112 B of payload.)
struct alignas(32) NgiSlotA { unsigned long long sample; double px, py, pz; };   // sample index (Philox counter), vertex position (fp64)
struct alignas(32) NgiSlotB { float4 thr_pix; float4 dir_info; };                // throughput.xyz | pixel (or light prim) bits; extend-ray
                                                                                 // direction.xyz | info bits (alive | rr_survive<<1 | numVertices<<8)
struct NgiWaveParams {
    NgiSlotA* sa;
    NgiSlotB* sb;
    float4* hit;                  // t, u, v, global triangle id (written by the extend kernel)
    // shadow queue: 3 x float4 per entry = (o.xyz, tmax) (d.xyz, pixel) (C.xyz, -)
    float4* shadow_q;
    unsigned* extend_q;           // slot ids of the extend rays of this iteration (compacted by the logic stage)
    unsigned* iter_counters;      // [0] shadow entries this iteration, [1] extend rays this iteration
    unsigned* fetch_cursors;      // [0] shadow, [1] extend: dynamic-fetch cursors of the persistent trace kernels
    unsigned* surface_q;          // slots that continue at a surface vertex this iteration (written by the classify stage)
    unsigned* regen_q;            // slots whose path ended: regenerated from the sample counter by the eye stage
    unsigned* stage_counters;     // [0] surface_q entries, [1] regen_q entries
    unsigned long long* next_sample;
    float* film;                  // [H][W][3], row 0 = bottom
    unsigned capacity;            // slots
    int renderer;                 // 0 pt, 1 ptdirect
    int max_verts;
    int width, height;
    unsigned long long sample_end;   // exclusive
    unsigned seed_lo, seed_hi;
    float film_scale;             // W*H / film_norm_samples (src/nanogi.cpp:436), folded into every splat
};

// ---- atomics / queue allocation (warp-aggregated on the device) ---------------------------------
NGI_HD unsigned long long ngi_fetch_sample(unsigned long long* ctr) {
#if defined(__CUDA_ARCH__)
    const unsigned mask = __activemask();
    const int leader = __ffs((int)mask) - 1;
    const int lane = (int)(threadIdx.x & 31u);
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
#else
    return (*ctr)++;
#endif
}
NGI_HD unsigned ngi_queue_alloc(unsigned* ctr) {
#if defined(__CUDA_ARCH__)
    const unsigned mask = __activemask();
    const int leader = __ffs((int)mask) - 1;
    const int lane = (int)(threadIdx.x & 31u);
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned)__popc(mask & ((1u << lane) - 1u));
#else
    return (*ctr)++;
#endif
}
NGI_HD void ngi_film_add(float* film, const int pixel, const f3 c) {
#if defined(__CUDA_ARCH__)
    atomicAdd(film + 3 * (size_t)pixel + 0, c.x);
    atomicAdd(film + 3 * (size_t)pixel + 1, c.y);
    atomicAdd(film + 3 * (size_t)pixel + 2, c.z);
#else
    film[3 * (size_t)pixel + 0] += c.x; film[3 * (size_t)pixel + 1] += c.y; film[3 * (size_t)pixel + 2] += c.z;
#endif
}

// What one path vertex hands to the queues: an optional shadow-queue entry (NEE) and whether the slot traces an
// extend ray. The per-slot functions only FILL this; the caller allocates the queue entries — the CUDA kernels with
// one atomic per block and queue (ngi_gpu.cu), the CPU simulator one by one (ngi_emit).
struct NgiVertexOut {
    bool shadow, extend;
    f3 so, sd, sC; float stmax; int spixel;
};
NGI_HD void ngi_write_shadow(const NgiWaveParams& wp, const unsigned e, const NgiVertexOut& v) {
    float4* q = wp.shadow_q + 3 * (size_t)e;
    q[0] = make_float4(v.so.x, v.so.y, v.so.z, v.stmax);
    q[1] = make_float4(v.sd.x, v.sd.y, v.sd.z, u2f((unsigned)v.spixel));
    q[2] = make_float4(v.sC.x, v.sC.y, v.sC.z, 0.0f);
}
NGI_HD void ngi_emit(const NgiWaveParams& wp, const unsigned slot, const NgiVertexOut& v) {
    if (v.shadow) ngi_write_shadow(wp, ngi_queue_alloc(wp.iter_counters + 0), v);
    if (v.extend) wp.extend_q[ngi_queue_alloc(wp.iter_counters + 1)] = slot;           // compacted extend queue (+ exact ray count)
}

// ---- fp64 view of the traced fp32 direction: d + U(-1/2, 1/2) * ulp(d) per component ---------------
NGI_HD double ngi_dither1(const float d, const unsigned r10) {
    const unsigned e = f2u(d) & 0x7F800000u;
    const float ulp = e > (24u << 23) ? u2f(e - (23u << 23)) : 0.0f;
    return (double)d + ((double)r10 * (1.0 / 1024.0) - 0.5 + (0.5 / 1024.0)) * (double)ulp;
}
NGI_HD void ngi_dither_direction(const f3 d, const float4 hit, double& dx, double& dy, double& dz) {
    unsigned h = (f2u(hit.x) * 0x9E3779B1u) ^ (f2u(hit.y) * 0x85EBCA77u) ^ (f2u(hit.z) * 0xC2B2AE3Du) ^ f2u(hit.w);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    dx = ngi_dither1(d.x, h & 1023u);
    dy = ngi_dither1(d.y, (h >> 10) & 1023u);
    dz = ngi_dither1(d.z, (h >> 20) & 1023u);
}

// ---- TexR lookup at a hit: geom.uv (rt.hpp:2221-2227, fp64 like the reference so that texel boundaries fall in the
// same place) and Texture::Evaluate (rt.hpp:262-268: nearest texel, fract wrap) ---------------------------------
NGI_HD f3 ngi_texture_at_hit(const NgiDevScene& sc, const int tex, const unsigned tri, const float u, const float v) {
    double tu, tv;
    ngi_uv_at(sc, tri, u, v, tu, tv);
    const NgiDevTex T = sc.textures[tex];
    const int x = clampi((int)((tu - floor(tu)) * T.width), 0, T.width - 1);
    const int y = clampi((int)((tv - floor(tv)) * T.height), 0, T.height - 1);
    const float* px = sc.tex_data + 3 * ((size_t)T.offset + (size_t)y * T.width + x);
    return mk3(ngi_ldg(px), ngi_ldg(px + 1), ngi_ldg(px + 2));
}

// ---- surface reconstruction of a hit, rt.hpp:2190-2233 ------------------------------------------
NGI_HD int ngi_reconstruct(const NgiDevScene& sc, const unsigned tri, const float u, const float v, NgiGeom& g) {
    const float4* r = sc.shade_tris + 5 * (size_t)tri;
    const float4 r0 = ngi_ldg(r), r1 = ngi_ldg(r + 1), r2 = ngi_ldg(r + 2), r3 = ngi_ldg(r + 3), r4 = ngi_ldg(r + 4);
    const f3 p1 = mk3(r0.x, r0.y, r0.z), p2 = mk3(r0.w, r1.x, r1.y), p3 = mk3(r1.z, r1.w, r2.x);
    const f3 n1 = mk3(r2.y, r2.z, r2.w), n2 = mk3(r3.x, r3.y, r3.z), n3 = mk3(r3.w, r4.x, r4.y);
    g.gn = normalize(cross(p2 - p1, p3 - p1));                                                // :2206
    const float w = 1.0f - u - v;                                                             // (1.0f - u - v) in float, :2212
    g.sn = normalize(n1 * w + n2 * u + n3 * v);
    if (g.sn.x != g.sn.x || g.sn.y != g.sn.y || g.sn.z != g.sn.z) g.sn = g.gn;               // NaN fallback, :2213-2218
    ngi_tangent_space(g);                                                                     // :2233
    return (int)f2u(r4.z);
}

// ---- the logic stage ---------------------------------------------------------------------------
// Split in three per-slot functions so that the CUDA kernel can regroup the slots of a block between them
// (block-local compaction: dense warps per stage, see k_logic in ngi_gpu.cu) while the CPU simulator simply
// calls them back to back (ngi_logic_step):
//   classify  tail of reference iteration k that needs no surface geometry: miss / RR / vertex cap; `pt` adds
//             the emission of a hit light here. Returns whether the path continues at a surface vertex.
//   surface   head of iteration k+1 at a surface vertex: reconstruct, NEE, BSDF sample. Returns false when the
//             path ends at this vertex (the slot is then regenerated by `eye` in the same logic stage).
//   eye       starts the next sample at the eye vertex: NEE to the light-sample's pixel, camera ray.
#define NGI_CLASS_REGENERATE 0
#define NGI_CLASS_SURFACE 1

// A sampled sensor position: E->SamplePosition (rt.hpp:565-586). Pinhole: the eye point, degenerate. E.area: a point of
// the sensor mesh (face normal) plus the pixel its uv maps to (RasterPosition = geom.uv, rt.hpp:1386-1391).
struct NgiSensorPoint { double px, py, pz; f3 n; int pixel; int degenerate; };
NGI_HD NgiSensorPoint ngi_sample_sensor(const NgiDevScene& sc, const NgiWaveParams& wp, const float u0, const float u1) {
    const NgiDevSensor& E = sc.sensor;
    NgiSensorPoint sp;
    if (E.kind == NGI_ET_PINHOLE) {
        sp.px = E.px; sp.py = E.py; sp.pz = E.pz; sp.n = mk3(0.0f); sp.pixel = -1; sp.degenerate = 1;
    } else {
        f3 p; int tri; float bx, by; double pd[3];
        ngi_sample_triangle_mesh(sc, E.first_tri, E.num_tris, E.cdf_offset, u0, u1, p, sp.n, tri, bx, by, pd);
        sp.px = pd[0]; sp.py = pd[1]; sp.pz = pd[2];
        sp.pixel = ngi_area_sensor_pixel(sc, (unsigned)tri, bx, by, wp.width, wp.height);
        sp.degenerate = 0;
    }
    return sp;
}

// kinds of path vertex (a literal at every call site, so each flavour is specialised by the compiler)
#define NGI_VTX_SURFACE 0     /* a surface hit, any renderer                                              */
#define NGI_VTX_EYE 1         /* first vertex of pt / ptdirect: a point of the sensor                     */
#define NGI_VTX_LIGHT 2       /* first vertex of lt / ltdirect: a point of a light (src/nanogi.cpp:808-836) */

// one path vertex: optional connection (ptdirect: NEE to a light, src/nanogi.cpp:654-712; ltdirect: to the sensor,
// src/nanogi.cpp:1000-1052), direction sampling, extend-ray emission. `GEN` = false compiles the hot-path flavour — eye
// path (pt / ptdirect) with a pinhole sensor, everything else folded away at compile time — and `GEN` = true the generic
// one that looks at wp.renderer (lt / ltdirect = light path) and at the sensor kind (E.area) at run time.
// `aux` is what the slot carries in thr_pix.w: the pixel index of an eye path, the light primitive of a light path
// (ltdirect needs it at every vertex: the reference evaluates `L->EvaluatePositionPDF(geomE)`, src/nanogi.cpp:1017).
// For KIND != SURFACE `g` holds the emitter point's frame (sn = gn = face normal; unused for degenerate emitters) and
// `em_le` / `em_type` / `em_degenerate` describe the light the path starts on.
template <int KIND, bool GEN>
NGI_HD void ngi_vertex(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot,
                       const unsigned long long sample, f3 thr, int aux, const int nverts, const int type,
                       const NgiGeom& g, const f3 wi, const double px, const double py, const double pz, const int primIdx,
                       const f3 em_le, const int em_type, const int em_degenerate, NgiVertexOut& out) {
    out.shadow = false; out.extend = false;
    const NgiDevSensor& E = sc.sensor;
    const unsigned vtx = (unsigned)(nverts - 1);
    const NgiDevPrim& P = sc.prims[primIdx];
    const bool LT = GEN && wp.renderer >= 2;                    // light path
    const bool pinhole = !GEN || E.kind == NGI_ET_PINHOLE;

    if (!LT) {
        // ---- direct light sampling (ptdirect), nanogi.cpp:654-712 ----
        if (wp.renderer == 1 && sc.n_lights > 0) {
            unsigned rb[4];
            philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), vtx, 1u, wp.seed_lo, wp.seed_hi, rb);
            const NgiLightSample ls = ngi_sample_light(sc, u01(rb[0]), u01(rb[1]), u01(rb[2]));
            if (ls.valid) {
                const f3 diff = mk3((float)((double)ls.p.x - px), (float)((double)ls.p.y - py), (float)((double)ls.p.z - pz));
                const float dist2 = dot(diff, diff);
                const float dist = ngi_sqrtf(dist2);
                const f3 ppL = diff / dist;                                                   // :680
                f3 fsE; int index = aux;
                float g1 = 1.0f;
                if (KIND == NGI_VTX_EYE) {
                    if (pinhole) {
                        float rx = 0.0f, ry = 0.0f;
                        const float we = ngi_pinhole_importance(E, ppL, rx, ry);              // :681 (type E)
                        fsE = mk3(we);
                        index = ngi_pixel_index(rx, ry, wp.width, wp.height);                 // :698-703
                    } else {
                        const float ce = dot(g.sn, ppL);                                      // E.area: We iff cos > 0 (rt.hpp:947-953)
                        fsE = ce > 0.0f ? E.we : mk3(0.0f);
                        g1 = fabsf(ce);                                                       // the sensor point is not degenerate, rt.hpp:2371
                        // index: RasterPosition(ppL, geom) = geom.uv of the sensor point, already in `aux`
                    }
                } else {
                    float pdfUnused;
                    fsE = ngi_eval_bsdf(P, type, g, wi, ppL, false, pdfUnused);               // :681
                    g1 = fabsf(dot(g.sn, ppL));                                               // GeometryTerm, rt.hpp:2371
                }
                f3 fsL = ls.le;                                                               // :682
                float g2 = 1.0f;
                if (!ls.degenerate) {
                    const float cl = dot(ls.n, -ppL);
                    if (cl <= 0.0f) fsL = mk3(0.0f);                                          // rt.hpp:922-927
                    g2 = fabsf(cl);                                                           // rt.hpp:2372
                }
                const float G = ngi_divf(g1 * g2, dist2);                                     // :683
                const f3 C = thr * fsE * fsL * ngi_divf(G, ls.pdf);                                  // :686 (V applied by the shadow kernel)
                if (!is_zero(C)) {
                    out.shadow = true;
                    out.so = mk3((float)px, (float)py, (float)pz);                            // rt.hpp:2166-2168
                    out.sd = ppL; out.stmax = dist * (1.0f - NGI_EPS_F);                      // rt.hpp:2260
                    out.sC = C * wp.film_scale; out.spixel = index;
                }
            }
        }
    } else if (wp.renderer == 3) {
        // ---- direct sensor sampling (ltdirect), nanogi.cpp:1000-1052 ----
        unsigned rc[4];
        philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), vtx, 2u, wp.seed_lo, wp.seed_hi, rc);
        const NgiSensorPoint sp = ngi_sample_sensor(sc, wp, u01(rc[1]), u01(rc[2]));          // :1005-1016 (one sensor: pdfE = 1)
        const NgiDevPrim& L0 = sc.prims[aux];                                                 // the light this path started on
        const float pdfPE = L0.l_type == NGI_LT_POINT ? 1.0f : L0.l_inv_area;                 // :1017 — sic, the LIGHT's position pdf
        const f3 diff = mk3((float)(sp.px - px), (float)(sp.py - py), (float)(sp.pz - pz));
        const float dist2 = dot(diff, diff);
        const float dist = ngi_sqrtf(dist2);
        const f3 ppE = diff / dist;                                                           // :1026
        f3 fsL; float g1 = 1.0f;
        if (KIND == NGI_VTX_LIGHT) {                                                          // :1027, EvaluateDirection(L, forceDegenerated = false)
            if (em_type == NGI_LT_AREA) { const float cl = dot(g.sn, ppE); fsL = cl > 0.0f ? em_le : mk3(0.0f); g1 = fabsf(cl); }
            else if (em_type == NGI_LT_POINT) fsL = em_le;                                    // degenerate: no cosine in G
            else { fsL = mk3(0.0f); }                                                         // directional: 0 (rt.hpp:934-937)
        } else {
            float pdfUnused;
            fsL = ngi_eval_bsdf(P, type, g, wi, ppE, false, pdfUnused, true);                 // :1027 (TransportDirection::LE)
            g1 = fabsf(dot(g.sn, ppE));
        }
        f3 fsE; float g2 = 1.0f; int index;
        if (pinhole) {                                                                        // :1028, :1044
            float rx = 0.0f, ry = 0.0f;
            fsE = mk3(ngi_pinhole_importance(E, -ppE, rx, ry));
            index = ngi_pixel_index(rx, ry, wp.width, wp.height);
        } else {
            const float ce = dot(sp.n, -ppE);
            fsE = ce > 0.0f ? E.we : mk3(0.0f);
            g2 = fabsf(ce);
            index = sp.pixel;
        }
        const float G = ngi_divf(g1 * g2, dist2);                                             // :1029
        const f3 C = thr * fsL * fsE * ngi_divf(G, pdfPE);                                           // :1032 (LeP = 1, pdfE = 1; V by the shadow kernel)
        if (!is_zero(C)) {                                                                    // :1040
            out.shadow = true;
            out.so = mk3((float)px, (float)py, (float)pz);
            out.sd = ppE; out.stmax = dist * (1.0f - NGI_EPS_F);                              // Scene::Visible, rt.hpp:2251-2261
            out.sC = C * wp.film_scale; out.spixel = index;
        }
    }

    // ---- sample the next direction, nanogi.cpp:492-544 / :716-754 / :852-874 / :1061-1083 ----
    unsigned ra[4];
    philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), vtx, 0u, wp.seed_lo, wp.seed_hi, ra);
    f3 wo;
    bool ok;
    if (KIND == NGI_VTX_EYE) {
        if (pinhole) {
            wo = ngi_pinhole_sample(E, u01(ra[0]), u01(ra[1]));
            float rx, ry, ct;
            ok = ngi_raster_position(E, wo, rx, ry, ct);                                      // :504-523 / :728-733
            if (ok) aux = ngi_pixel_index(rx, ry, wp.width, wp.height);
            // fs / pdfD = We / pdf = 1 exactly (same expression on both sides)
        } else {
            const f3 lw = ngi_cosine_hemisphere(u01(ra[0]), u01(ra[1]));                      // rt.hpp:726-731
            wo = ngi_to_world(g, lw);
            ok = dot(g.sn, wo) > 0.0f && !is_zero(E.we);                                      // We iff cos > 0, pdf 1/pi (rt.hpp:947-953, :1179-1187)
            thr = thr * (E.we * NGI_PI_F);
            // RasterPosition = geom.uv: `aux` already holds the pixel of the sensor point
        }
    } else if (KIND == NGI_VTX_LIGHT) {                                                       // SampleDirection(L), rt.hpp:698-715; fs = Le
        if (em_type == NGI_LT_AREA) {
            wo = ngi_to_world(g, ngi_cosine_hemisphere(u01(ra[0]), u01(ra[1])));
            ok = dot(g.sn, wo) > 0.0f;                                                        // rt.hpp:922-927; pdf 1/pi (rt.hpp:1156-1161)
            thr = thr * (em_le * NGI_PI_F);
        } else if (em_type == NGI_LT_POINT) {
            wo = ngi_uniform_sphere(u01(ra[0]), u01(ra[1]));
            ok = true;
            thr = thr * (em_le * (4.0f * NGI_PI_F));                                          // pdf 1/(4 pi), rt.hpp:124-127
        } else {
            wo = g.sn;                                                                        // L.Directional.Direction, rt.hpp:711-715
            ok = true;
            thr = thr * em_le;                                                                // pdf 1 with forceDegenerated (rt.hpp:1172-1175)
        }
        if (is_zero(em_le)) ok = false;                                                       // fs == 0 -> break (:862 / :1071)
    } else {
        ok = ngi_sample_bsdf(P, type, g, wi, u01(ra[0]), u01(ra[1]), u01(ra[2]), wo);
        if (ok) {
            float pdfD;
            const f3 fs = ngi_eval_bsdf(P, type, g, wi, wo, true, pdfD, LT);                  // :531 / :741 / :861 / :1070
            if (is_zero(fs)) ok = false;                                                      // :532 / :742
            else thr = thr * (fs / pdfD);                                                     // :544 / :754
        }
    }
    if (!ok) return;
    const unsigned survive = (u01(ra[3]) > 0.5f) ? 0u : NGI_INFO_RR_SURVIVE;              // :583-587, decided up front
    NgiSlotA a; a.sample = sample; a.px = px; a.py = py; a.pz = pz;
    NgiSlotB b;
    b.thr_pix = make_float4(thr.x, thr.y, thr.z, u2f((unsigned)aux));
    b.dir_info = make_float4(wo.x, wo.y, wo.z, u2f(NGI_INFO_ALIVE | survive | ((unsigned)nverts << 8)));
    wp.sa[slot] = a;                                                                       // two whole sectors
    wp.sb[slot] = b;
    out.extend = true;
}

template <bool GEN>
NGI_HD int ngi_logic_classify(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot) {
    const float4 di = wp.sb[slot].dir_info;
    const unsigned info = f2u(di.w);
    if (!(info & NGI_INFO_ALIVE)) return NGI_CLASS_REGENERATE;
    const float4 h = wp.hit[slot];
    const unsigned tri = f2u(h.w);
    if (tri == NGI_MISS) return NGI_CLASS_REGENERATE;                                         // miss -> break, nanogi.cpp:557 / :767 / :887 / :1096
    if (wp.renderer == 0) {                                                                   // nanogi.cpp:566-577
        const int primIdx = (int)f2u(ngi_ldg(sc.shade_tris + 5 * (size_t)tri + 4).z);
        const NgiDevPrim& P = sc.prims[primIdx];
        if ((P.type & NGI_L) && P.l_type == NGI_LT_AREA) {
            // EvaluateDirection(L.area): Le iff cos_sn(-d) > 0 (rt.hpp:922-927); EvaluatePosition(area) = 1
            NgiGeom g;
            ngi_reconstruct(sc, tri, h.y, h.z, g);
            const f3 d = mk3(di.x, di.y, di.z);
            if (dot(g.sn, -d) > 0.0f) {
                const float4 tp = wp.sb[slot].thr_pix;
                ngi_film_add(wp.film, (int)f2u(tp.w), mk3(tp.x, tp.y, tp.z) * P.l_le * wp.film_scale);
            }
        }
    }
    if (GEN && wp.renderer == 2) {                                                            // lt: hit with the sensor, nanogi.cpp:899-920
        const int primIdx = (int)f2u(ngi_ldg(sc.shade_tris + 5 * (size_t)tri + 4).z);
        const NgiDevPrim& P = sc.prims[primIdx];
        // only an E.area sensor has a mesh to hit; RasterPosition = geom.uv (always succeeds), We iff cos_sn(-d) > 0
        if ((P.type & NGI_E) && sc.sensor.kind == NGI_ET_AREA && primIdx == sc.sensor.prim && sc.shade_uv) {
            NgiGeom g;
            ngi_reconstruct(sc, tri, h.y, h.z, g);
            const f3 d = mk3(di.x, di.y, di.z);
            if (dot(g.sn, -d) > 0.0f) {
                const float4 tp = wp.sb[slot].thr_pix;
                ngi_film_add(wp.film, ngi_area_sensor_pixel(sc, tri, h.y, h.z, wp.width, wp.height),
                             mk3(tp.x, tp.y, tp.z) * sc.sensor.we * wp.film_scale);
            }
        }
    }
    if (!(info & NGI_INFO_RR_SURVIVE)) return NGI_CLASS_REGENERATE;                           // nanogi.cpp:581-591
    const int nverts = (int)(info >> 8) + 1;                                                  // :603
    if (wp.max_verts != -1 && nverts >= wp.max_verts) return NGI_CLASS_REGENERATE;            // :485 / :647
    return NGI_CLASS_SURFACE;
}

template <bool GEN>
NGI_HD void ngi_logic_surface(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot, NgiVertexOut& out) {
    const float4 di = wp.sb[slot].dir_info;
    const unsigned info = f2u(di.w);
    const float4 h = wp.hit[slot];
    const f3 d = mk3(di.x, di.y, di.z);
    const float4 tp = wp.sb[slot].thr_pix;
    // isect.geom.p = ray.o + ray.d * (double)tfar, rt.hpp:2197. In the reference ray.d is the fp64
    // direction whose fp32 ROUNDING was traced (rt.hpp:2169-2171): the reconstructed point is off the
    // traced ray by (d64 - d32) * t, which together with the absolute 1e-4 epsilon decides how often
    // the next ray re-hits its own surface (measured on the Cornell box: 5.3 % of bounce rays with
    // that term, 5.0 % without). The device direction only exists in fp32, so the rounding residual
    // is re-created as a uniform +-ulp/2 dither hashed from the hit record (tests/test_sim_parity.py).
    double ddx, ddy, ddz;
    ngi_dither_direction(d, h, ddx, ddy, ddz);
    const NgiSlotA sa = wp.sa[slot];
    const double px = sa.px + ddx * (double)h.x;
    const double py = sa.py + ddy * (double)h.x;
    const double pz = sa.pz + ddz * (double)h.x;
    NgiGeom g;
    const int primIdx = ngi_reconstruct(sc, f2u(h.w), h.y, h.z, g);
    const int nverts = (int)(info >> 8) + 1;                                                  // :603
    const f3 thr = mk3(tp.x, tp.y, tp.z) * 2.0f;                                              // throughput /= rrProb, :591
    const NgiDevPrim& P = sc.prims[primIdx];
    const int type = P.type & ~NGI_EMITTER;                                                   // :601
    const int tex = (type & NGI_D) ? P.d_tex : P.g_tex;
    g.albedo = (tex >= 0 && sc.shade_uv) ? ngi_texture_at_hit(sc, tex, f2u(h.w), h.y, h.z) : ngi_constant_albedo(P, type);
    ngi_vertex<NGI_VTX_SURFACE, GEN>(sc, wp, slot, sa.sample, thr, (int)f2u(tp.w), nverts, type, g, -d /* :602 */, px, py, pz, primIdx,
                                    mk3(0.0f), 0, 0, out);
}

// the slot's path ended: start sample index `sample` at its first vertex — the eye vertex for pt / ptdirect
// (nanogi.cpp:450-479 / :613-641), a sampled light point for lt / ltdirect (nanogi.cpp:808-836 / :959-987); an index
// past the end of the shard leaves the slot idle. The CUDA kernel numbers its queue entries consecutively from
// the render's sample cursor (no atomics); the simulator draws them from the cursor one at a time.
template <bool GEN>
NGI_HD void ngi_logic_eye(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot, const unsigned long long sample,
                          NgiVertexOut& out) {
    const NgiDevSensor& E = sc.sensor;
    out.shadow = false; out.extend = false;
    if (sample < wp.sample_end) {
        NgiGeom g; g.sn = g.gn = g.dpdu = g.dpdv = g.albedo = mk3(0.0f);
        const bool LT = GEN && wp.renderer >= 2;
        if (!LT) {
            if (!GEN || E.kind == NGI_ET_PINHOLE) {
                // EvaluatePosition / pdfPE / pdfE = 1 for the pinhole
                ngi_vertex<NGI_VTX_EYE, GEN>(sc, wp, slot, sample, mk3(1.0f), -1, 1, NGI_E, g, mk3(0.0f), E.px, E.py, E.pz, E.prim,
                                               mk3(0.0f), 0, 0, out);
            } else {
                // E.area: a point of the sensor mesh (block 2 of vertex 0), throughput 1 / pdfPE = area (nanogi.cpp:461-471)
                unsigned rc[4];
                philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), 0u, 2u, wp.seed_lo, wp.seed_hi, rc);
                const NgiSensorPoint sp = ngi_sample_sensor(sc, wp, u01(rc[1]), u01(rc[2]));
                g.sn = g.gn = sp.n;
                ngi_tangent_space(g);
                ngi_vertex<NGI_VTX_EYE, true>(sc, wp, slot, sample, mk3(ngi_rcpf(E.inv_area)), sp.pixel, 1, NGI_E, g, mk3(0.0f), sp.px, sp.py, sp.pz,
                                               E.prim, mk3(0.0f), 0, 0, out);
            }
        } else if (sc.n_lights > 0) {
            unsigned rb[4];
            philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), 0u, 1u, wp.seed_lo, wp.seed_hi, rb);
            double pd[3];
            const NgiLightSample ls = ngi_sample_light(sc, u01(rb[0]), u01(rb[1]), u01(rb[2]), true, pd);   // :808-820 / :959-971
            g.sn = g.gn = ls.n;
            if (!ls.degenerate) ngi_tangent_space(g);
            // throughput = EvaluatePosition / pdfPL / pdfL (:830 / :980)
            ngi_vertex<NGI_VTX_LIGHT, true>(sc, wp, slot, sample, mk3(ngi_rcpf(ls.pdf)), ls.prim, 1, NGI_L, g, mk3(0.0f),
                                            pd[0], pd[1], pd[2], ls.prim, ls.le, ls.l_type, ls.degenerate, out);
        }
    }
    if (!out.extend) wp.sb[slot].dir_info = make_float4(0.0f, 0.0f, 0.0f, u2f(0u));        // idle slot
}

// all three for one slot (the CPU simulator's order; the CUDA kernels regroup slots between the stages)
template <bool GEN>
NGI_HD_NOINLINE void ngi_logic_step(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot) {
    NgiVertexOut out;
    if (ngi_logic_classify<GEN>(sc, wp, slot) == NGI_CLASS_SURFACE) {
        ngi_logic_surface<GEN>(sc, wp, slot, out);
        ngi_emit(wp, slot, out);
        if (out.extend) return;
    }
    ngi_logic_eye<GEN>(sc, wp, slot, ngi_fetch_sample(wp.next_sample), out);
    ngi_emit(wp, slot, out);
}

// ---- extend / shadow bodies (BVH8 = product path) -----------------------------------------------
NGI_HD void ngi_extend_step(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned slot) {
    const float4 di = wp.sb[slot].dir_info;
    if (!(f2u(di.w) & NGI_INFO_ALIVE)) return;
    const NgiSlotA sa = wp.sa[slot];
    const f3 o = mk3((float)sa.px, (float)sa.py, (float)sa.pz);                               // rt.hpp:2166-2168
    NgiHitRec h;
    ngi_trace_bvh8<false>(sc.nodes8, sc.tris8, o, mk3(di.x, di.y, di.z), NGI_EPS_F, NGI_INF_F, h);   // rt.hpp:2246-2249
    wp.hit[slot] = make_float4(h.t, h.u, h.v, u2f(h.tri));
}
NGI_HD void ngi_shadow_step(const NgiDevScene& sc, const NgiWaveParams& wp, const unsigned e) {
    const float4* q = wp.shadow_q + 3 * (size_t)e;
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    NgiHitRec h;
    const bool occluded = ngi_trace_bvh8<true>(sc.nodes8, sc.tris8, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), NGI_EPS_F, q0.w, h);
    if (!occluded) ngi_film_add(wp.film, (int)f2u(q1.w), mk3(q2.x, q2.y, q2.z));             // nanogi.cpp:706
}

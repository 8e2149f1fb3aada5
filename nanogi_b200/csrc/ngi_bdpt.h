// ngi_bdpt.h — bidirectional path tracing, one sample per thread.
//
// Replaces ProcessSample_BDPT (reference src/nanogi.cpp:1133-1186) and struct Path / PathVertex (reference
// include/nanogi/bdpt.hpp:38-539): SampleSubpath, Connect, EvaluateContribution = EvaluateUnweightContribution x
// EvaluatePowerHeuristicsMISWeightOpt (the weight the reference actually uses, bdpt.hpp:181-185), SelectionProb, RasterPosition,
// EvaluateCst, EvaluatePDF. SURVEY §8f row 4 ("then bdpt").
//
// Why not the wavefront machinery: a bdpt sample is two subpaths kept in full plus O(n^2) connections, each weighted by O(n)
// products of direction pdfs over the whole path — per-sample state of a few KB and data-dependent loop nests, not a stream of
// independent path vertices. It runs as one thread per sample with the subpaths in local memory and the per-ray form of the BVH8
// traversal (ngi_trace_bvh8); the same per-sample function is stepped by the CPU simulator (tests/hostsim). Subpaths are capped at
// NGI_BD_MAX_VERTS vertices each (Russian roulette with p = 0.5 makes a longer one a 6e-8 event; the reference has no cap).
//
// Arithmetic: fp32 like the other renderers; vertex positions fp64; the path pdfs that enter the MIS weight are products of up to
// 2 x NGI_BD_MAX_VERTS factors of magnitude 1e-6 (G at Cornell scale) — formed in fp64, or they would leave the fp32 range.
// Random numbers: Philox key = seed, counter = (sample, vertex, block): light subpath block 1 (emitter pick / position, vertex 0)
// and block 0 (direction u0 u1, uComp, RR — like lt); eye subpath block 2 (sensor position, vertex 0) and block 3.
#pragma once
#include "ngi_wave.h"

#define NGI_BD_MAX_VERTS 24

struct NgiBdVertex {          // PathVertex, bdpt.hpp:38-43 (the frame is rebuilt from sn when needed)
    double px, py, pz;
    f3 sn, gn;
    f3 albedo;                // R of the lobe at this point (constant or TexR at the hit's uv)
    int prim;
    int type;                 // PrimitiveType bits the vertex acts as
    int degenerate;           // point light / pinhole
    int pixel;                // E.area sensor point or hit: PixelIndex(geom.uv); -1 otherwise
};

struct NgiBdParams {
    float* film;
    int width, height, max_verts;
    unsigned seed_lo, seed_hi;
    float film_scale;
};

struct NgiBdCounters { unsigned long long extend, shadow; };

NGI_HD f3 ngi_bd_dir(const NgiBdVertex& from, const NgiBdVertex& to) {        // glm::normalize(to.p - from.p)
    const double dx = to.px - from.px, dy = to.py - from.py, dz = to.pz - from.pz;
    const double inv = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
    return mk3((float)(dx * inv), (float)(dy * inv), (float)(dz * inv));
}
NGI_HD NgiGeom ngi_bd_geom(const NgiBdVertex& v) {
    NgiGeom g; g.sn = v.sn; g.gn = v.gn; g.albedo = v.albedo; g.dpdu = g.dpdv = mk3(0.0f);
    if (!v.degenerate) ngi_tangent_space(g);
    return g;
}

// Primitive::EvaluateDirection for any vertex kind (rt.hpp:912-1148) and, in `pdf`, Primitive::EvaluateDirectionPDF (:1150-1336)
NGI_HD f3 ngi_bd_eval_direction(const NgiDevScene& sc, const NgiBdVertex& v, const int type, const f3 wi, const f3 wo, const bool transLE,
                                const bool forceDegenerated, float& pdf) {
    const NgiDevPrim& P = sc.prims[v.prim];
    pdf = 0.0f;
    if (type & NGI_L) {
        if (P.l_type == NGI_LT_AREA) { if (dot(v.sn, wo) <= 0.0f) return mk3(0.0f); pdf = NGI_INV_PI_F; return P.l_le; }       // :922-927, :1156-1161
        if (P.l_type == NGI_LT_POINT) { pdf = NGI_INV_PI_F * 0.25f; return P.l_le; }                                              // :929-932, :1163-1166
        pdf = forceDegenerated ? 1.0f : 0.0f;                                                                                      // :934-937, :1172-1175
        return forceDegenerated ? P.l_le : mk3(0.0f);
    }
    if (type & NGI_E) {
        const NgiDevSensor& E = sc.sensor;
        if (E.kind == NGI_ET_AREA) { if (dot(v.sn, wo) <= 0.0f) return mk3(0.0f); pdf = NGI_INV_PI_F; return E.we; }             // :947-953, :1179-1187
        float rx, ry;
        const float we = ngi_pinhole_importance(E, wo, rx, ry);                                                                   // :955-978, :1189-1212
        pdf = we;
        return mk3(we);
    }
    if (!(type & NGI_BSDF)) return mk3(0.0f);                                                                                      // assert(0), :1146
    const NgiGeom g = ngi_bd_geom(v);
    return ngi_eval_bsdf(P, type, g, wi, wo, forceDegenerated, pdf, transLE);
}
// Primitive::EvaluatePosition (rt.hpp:594-641)
NGI_HD float ngi_bd_eval_position(const NgiDevScene& sc, const NgiBdVertex& v, const int type, const bool forceDegenerated) {
    if (type & NGI_L) return sc.prims[v.prim].l_type == NGI_LT_POINT ? (forceDegenerated ? 1.0f : 0.0f) : 1.0f;
    if (type & NGI_E) return sc.sensor.kind == NGI_ET_PINHOLE ? (forceDegenerated ? 1.0f : 0.0f) : 1.0f;
    return 0.0f;
}
// Primitive::EvaluatePositionPDF(geom, true) x Scene::EvaluateEmitterPDF (rt.hpp:643-690, :2338-2352)
NGI_HD float ngi_bd_position_pdf(const NgiDevScene& sc, const NgiBdVertex& v, const int type) {
    if (type & NGI_L) {
        const NgiDevPrim& P = sc.prims[v.prim];
        return (P.l_type == NGI_LT_POINT ? 1.0f : P.l_inv_area) * (1.0f / (float)sc.n_lights);
    }
    if (type & NGI_E) return sc.sensor.kind == NGI_ET_PINHOLE ? 1.0f : sc.sensor.inv_area;
    return 0.0f;
}
// GeometryTerm (rt.hpp:2364-2374)
NGI_HD float ngi_bd_geometry_term(const NgiBdVertex& a, const NgiBdVertex& b) {
    const double dx = b.px - a.px, dy = b.py - a.py, dz = b.pz - a.pz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const double inv = 1.0 / sqrt(d2);
    const f3 w = mk3((float)(dx * inv), (float)(dy * inv), (float)(dz * inv));
    float t = 1.0f;
    if (!a.degenerate) t *= fabsf(dot(a.sn, w));
    if (!b.degenerate) t *= fabsf(dot(b.sn, w));
    return (float)((double)t / d2);
}

// ---- Path::SampleSubpath, bdpt.hpp:54-123. kind 0: light subpath (LE), 1: eye subpath (EL). Returns the vertex count. ----------
NGI_HD int ngi_bd_sample_subpath(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind,
                                 NgiBdVertex* V, NgiBdCounters& cnt) {
    const NgiDevSensor& E = sc.sensor;
    int n = 0;
    const int cap = bp.max_verts == -1 ? NGI_BD_MAX_VERTS : (bp.max_verts < NGI_BD_MAX_VERTS ? bp.max_verts : NGI_BD_MAX_VERTS);
    for (int step = 0; step < cap; step++) {
        if (step == 0) {
            NgiBdVertex v;
            v.albedo = mk3(0.0f); v.pixel = -1;
            if (kind == 0) {
                if (sc.n_lights == 0) return 0;
                unsigned rb[4];
                philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), 0u, 1u, bp.seed_lo, bp.seed_hi, rb);
                double pd[3];
                const NgiLightSample ls = ngi_sample_light(sc, u01(rb[0]), u01(rb[1]), u01(rb[2]), true, pd);     // :63-66
                v.px = pd[0]; v.py = pd[1]; v.pz = pd[2]; v.sn = v.gn = ls.n; v.prim = ls.prim; v.type = NGI_L; v.degenerate = ls.degenerate;
            } else if (E.kind == NGI_ET_PINHOLE) {
                v.px = E.px; v.py = E.py; v.pz = E.pz; v.sn = v.gn = mk3(0.0f); v.prim = E.prim; v.type = NGI_E; v.degenerate = 1;
            } else {
                unsigned rc[4];
                philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), 0u, 2u, bp.seed_lo, bp.seed_hi, rc);
                f3 p; int tri; float bx, by; double pd[3];
                ngi_sample_triangle_mesh(sc, E.first_tri, E.num_tris, E.cdf_offset, u01(rc[1]), u01(rc[2]), p, v.sn, tri, bx, by, pd);
                v.gn = v.sn; v.px = pd[0]; v.py = pd[1]; v.pz = pd[2]; v.prim = E.prim; v.type = NGI_E; v.degenerate = 0;
                v.pixel = ngi_area_sensor_pixel(sc, (unsigned)tri, bx, by, bp.width, bp.height);
            }
            V[n++] = v;
            continue;
        }
        const NgiBdVertex& pv = V[n - 1];
        const f3 wi = n > 1 ? ngi_bd_dir(pv, V[n - 2]) : mk3(0.0f);                                              // :77
        unsigned ra[4];
        philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), (unsigned)(step - 1), kind == 0 ? 0u : 3u, bp.seed_lo, bp.seed_hi, ra);
        const float u0 = u01(ra[0]), u1 = u01(ra[1]), uc = u01(ra[2]);
        const NgiDevPrim& P = sc.prims[pv.prim];
        f3 wo = mk3(0.0f);
        bool wrote = true;
        if (pv.type & NGI_L) {                                                                                   // SampleDirection, rt.hpp:698-715
            if (P.l_type == NGI_LT_AREA) { const NgiGeom g = ngi_bd_geom(pv); wo = ngi_to_world(g, ngi_cosine_hemisphere(u0, u1)); }
            else if (P.l_type == NGI_LT_POINT) wo = ngi_uniform_sphere(u0, u1);
            else wo = P.l_vec;
        } else if (pv.type & NGI_E) {                                                                            // rt.hpp:726-740
            if (E.kind == NGI_ET_PINHOLE) wo = ngi_pinhole_sample(E, u0, u1);
            else { const NgiGeom g = ngi_bd_geom(pv); wo = ngi_to_world(g, ngi_cosine_hemisphere(u0, u1)); }
        } else {
            const NgiGeom g = ngi_bd_geom(pv);
            wrote = ngi_sample_bsdf(P, pv.type, g, wi, u0, u1, uc, wo);
        }
        if (!wrote) break;                                                                                        // wo stays zero -> f == 0
        float pdfUnused;
        const f3 f = ngi_bd_eval_direction(sc, pv, pv.type, wi, wo, kind == 0, true, pdfUnused);                  // :81
        if (is_zero(f)) break;
        NgiHitRec h;
        cnt.extend++;
        const f3 o = mk3((float)pv.px, (float)pv.py, (float)pv.pz);
        if (!ngi_trace_bvh8<false>(sc.nodes8, sc.tris8, o, wo, NGI_EPS_F, NGI_INF_F, h)) break;                   // :92
        NgiBdVertex v;
        double ddx, ddy, ddz;
        ngi_dither_direction(wo, make_float4(h.t, h.u, h.v, u2f(h.tri)), ddx, ddy, ddz);                          // see ngi_logic_surface
        v.px = pv.px + ddx * (double)h.t; v.py = pv.py + ddy * (double)h.t; v.pz = pv.pz + ddz * (double)h.t;
        NgiGeom g;
        v.prim = ngi_reconstruct(sc, h.tri, h.u, h.v, g);
        v.sn = g.sn; v.gn = g.gn; v.degenerate = 0;
        const NgiDevPrim& HP = sc.prims[v.prim];
        v.type = HP.type & ~NGI_EMITTER;                                                                          // :102
        const int tex = (v.type & NGI_D) ? HP.d_tex : HP.g_tex;
        v.albedo = (tex >= 0 && sc.shade_uv) ? ngi_texture_at_hit(sc, tex, h.tri, h.u, h.v) : ngi_constant_albedo(HP, v.type);
        v.pixel = ((HP.type & NGI_E) && E.kind == NGI_ET_AREA && sc.shade_uv) ? ngi_area_sensor_pixel(sc, h.tri, h.u, h.v, bp.width, bp.height) : -1;
        V[n++] = v;
        if (u01(ra[3]) > 0.5f) break;                                                                             // :108-113
    }
    return n;
}

// A connected full path: views into the two subpaths (light vertices 0..s-1, then eye vertices t-1..0) with the type each vertex
// acts as (Connect overrides the end vertex's type when one subpath is empty, bdpt.hpp:139,153).
struct NgiBdPath {
    const NgiBdVertex* L; const NgiBdVertex* E;
    int s, t, n;
    int type_first, type_last;      // overrides (or the vertex's own type)
    NGI_HD const NgiBdVertex& v(int i) const { return i < s ? L[i] : E[n - 1 - i]; }
    NGI_HD int type(int i) const { return i == 0 ? type_first : (i == n - 1 ? type_last : v(i).type); }
};

// Path::EvaluateCst, bdpt.hpp:217-250, for a split of the SAME path at `s` light vertices
NGI_HD f3 ngi_bd_cst(const NgiDevScene& sc, const NgiBdPath& p, const int s) {
    const int n = p.n, t = n - s;
    float pdfUnused;
    if (s == 0 && t > 0) {
        const NgiBdVertex& v = p.v(0);
        return ngi_bd_eval_direction(sc, v, p.type(0), mk3(0.0f), ngi_bd_dir(v, p.v(1)), false, false, pdfUnused) * ngi_bd_eval_position(sc, v, p.type(0), false);
    }
    if (s > 0 && t == 0) {
        const NgiBdVertex& v = p.v(n - 1);
        return ngi_bd_eval_direction(sc, v, p.type(n - 1), mk3(0.0f), ngi_bd_dir(v, p.v(n - 2)), true, false, pdfUnused) * ngi_bd_eval_position(sc, v, p.type(n - 1), false);
    }
    const NgiBdVertex& vL = p.v(s - 1);
    const NgiBdVertex& vE = p.v(s);
    const f3 fsL = ngi_bd_eval_direction(sc, vL, p.type(s - 1), s - 2 >= 0 ? ngi_bd_dir(vL, p.v(s - 2)) : mk3(0.0f), ngi_bd_dir(vL, vE), true, false, pdfUnused);
    if (is_zero(fsL)) return mk3(0.0f);
    const f3 fsE = ngi_bd_eval_direction(sc, vE, p.type(s), s + 1 < n ? ngi_bd_dir(vE, p.v(s + 1)) : mk3(0.0f), ngi_bd_dir(vE, vL), false, false, pdfUnused);
    return fsL * fsE * ngi_bd_geometry_term(vL, vE);
}

// Path::EvaluatePDF, bdpt.hpp:491-535 (fp64 product, see the header comment)
NGI_HD double ngi_bd_pdf(const NgiDevScene& sc, const NgiBdPath& p, const int s) {
    if (is_zero(ngi_bd_cst(sc, p, s))) return 0.0;
    const int n = p.n, t = n - s;
    double pdf = 1.0;
    float pd;
    if (s > 0) {
        pdf *= (double)ngi_bd_position_pdf(sc, p.v(0), p.type(0));
        for (int i = 0; i < s - 1; i++) {
            const NgiBdVertex& vi = p.v(i);
            ngi_bd_eval_direction(sc, vi, p.type(i), i >= 1 ? ngi_bd_dir(vi, p.v(i - 1)) : mk3(0.0f), ngi_bd_dir(vi, p.v(i + 1)), true, true, pd);
            pdf *= (double)pd * (double)ngi_bd_geometry_term(vi, p.v(i + 1));
        }
    }
    if (t > 0) {
        pdf *= (double)ngi_bd_position_pdf(sc, p.v(n - 1), p.type(n - 1));
        for (int i = n - 1; i >= s + 1; i--) {
            const NgiBdVertex& vi = p.v(i);
            ngi_bd_eval_direction(sc, vi, p.type(i), i + 1 < n ? ngi_bd_dir(vi, p.v(i + 1)) : mk3(0.0f), ngi_bd_dir(vi, p.v(i - 1)), false, true, pd);
            pdf *= (double)pd * (double)ngi_bd_geometry_term(vi, p.v(i - 1));
        }
    }
    return pdf;
}

// Path::EvaluateUnweightContribution, bdpt.hpp:252-343
NGI_HD f3 ngi_bd_unweighted(const NgiDevScene& sc, const NgiBdPath& p) {
    const int n = p.n, s = p.s, t = p.t;
    float pd;
    f3 alphaL = mk3(1.0f);
    if (s > 0) {
        alphaL = mk3(ngi_bd_eval_position(sc, p.v(0), p.type(0), true) / ngi_bd_position_pdf(sc, p.v(0), p.type(0)));
        for (int i = 0; i < s - 1; i++) {
            const NgiBdVertex& v = p.v(i);
            const f3 f = ngi_bd_eval_direction(sc, v, p.type(i), i >= 1 ? ngi_bd_dir(v, p.v(i - 1)) : mk3(0.0f), ngi_bd_dir(v, p.v(i + 1)), true, true, pd);
            if (is_zero(f)) return mk3(0.0f);
            alphaL = alphaL * (f / pd);
        }
    }
    f3 alphaE = mk3(1.0f);
    if (t > 0) {
        alphaE = mk3(ngi_bd_eval_position(sc, p.v(n - 1), p.type(n - 1), true) / ngi_bd_position_pdf(sc, p.v(n - 1), p.type(n - 1)));
        for (int i = n - 1; i > s; i--) {
            const NgiBdVertex& v = p.v(i);
            const f3 f = ngi_bd_eval_direction(sc, v, p.type(i), i < n - 1 ? ngi_bd_dir(v, p.v(i + 1)) : mk3(0.0f), ngi_bd_dir(v, p.v(i - 1)), false, true, pd);
            if (is_zero(f)) return mk3(0.0f);
            alphaE = alphaE * (f / pd);
        }
    }
    const f3 cst = ngi_bd_cst(sc, p, s);
    if (is_zero(cst)) return mk3(0.0f);
    return alphaL * cst * alphaE;
}

// one bdpt sample: ProcessSample_BDPT, src/nanogi.cpp:1133-1186
NGI_HD_NOINLINE void ngi_bdpt_sample(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, NgiBdVertex* VL, NgiBdVertex* VE,
                                     NgiBdCounters& cnt) {
    const int nL = ngi_bd_sample_subpath(sc, bp, sample, 0, VL, cnt);                        // :1137
    const int nE = ngi_bd_sample_subpath(sc, bp, sample, 1, VE, cnt);                        // :1138
    if (nL == 0 || nE == 0) return;
    for (int n = 2; n <= nE + nL; n++) {                                                     // :1148
        if (bp.max_verts != -1 && n > bp.max_verts) continue;
        const int minS = n - nE > 0 ? n - nE : 0, maxS = nL < n ? nL : n;
        for (int s = minS; s <= maxS; s++) {
            const int t = n - s;
            NgiBdPath p; p.L = VL; p.E = VE; p.s = s; p.t = t; p.n = n;
            // Path::Connect, bdpt.hpp:125-177
            if (s == 0) {
                if (!(sc.prims[VE[t - 1].prim].type & NGI_L)) continue;
                p.type_first = NGI_L; p.type_last = VE[0].type;
            } else if (t == 0) {
                if (!(sc.prims[VL[s - 1].prim].type & NGI_E) || VL[s - 1].prim != sc.sensor.prim || VL[s - 1].pixel < 0) continue;   // only THE sensor (last E primitive)
                p.type_first = VL[0].type; p.type_last = NGI_E;
            } else {
                const NgiBdVertex& a = VL[s - 1];
                const NgiBdVertex& b = VE[t - 1];
                const double dx = b.px - a.px, dy = b.py - a.py, dz = b.pz - a.pz;           // Scene::Visible, rt.hpp:2251-2261
                const double len = sqrt(dx * dx + dy * dy + dz * dz);
                const f3 d = mk3((float)(dx / len), (float)(dy / len), (float)(dz / len));
                NgiHitRec h;
                cnt.shadow++;
                if (ngi_trace_bvh8<true>(sc.nodes8, sc.tris8, mk3((float)a.px, (float)a.py, (float)a.pz), d, NGI_EPS_F, (float)len * (1.0f - NGI_EPS_F), h)) continue;
                p.type_first = VL[0].type; p.type_last = VE[0].type;
            }
            const f3 Cstar = ngi_bd_unweighted(sc, p);                                       // EvaluateContribution, bdpt.hpp:181-185
            if (is_zero(Cstar)) continue;
            const double ps = ngi_bd_pdf(sc, p, s);                                          // EvaluatePowerHeuristicsMISWeightOpt, :362-380
            double invWeight = 0.0;
            for (int i = 0; i <= n; i++) {
                const double pi = ngi_bd_pdf(sc, p, i);
                if (pi > 0.0) { const double r = pi / ps; invWeight += r * r; }
            }
            double sel = 1.0;                                                                // SelectionProb, :187-205
            for (int i = 1; i < s - 1; i++) sel *= 0.5;
            for (int i = t - 2; i >= 1; i--) sel *= 0.5;
            const f3 C = Cstar * (float)(1.0 / (invWeight * sel));
            if (is_zero(C) || !(C.x == C.x && C.y == C.y && C.z == C.z)) continue;
            // Path::RasterPosition of the last vertex, bdpt.hpp:207-215
            const NgiBdVertex& last = p.v(n - 1);
            int pixel;
            if (sc.sensor.kind == NGI_ET_PINHOLE) {
                float rx, ry, ct;
                if (!ngi_raster_position(sc.sensor, ngi_bd_dir(last, p.v(n - 2)), rx, ry, ct)) continue;   // cannot happen when C != 0 (We = 0 off the raster)
                pixel = ngi_pixel_index(rx, ry, bp.width, bp.height);
            } else {
                pixel = last.pixel;
                if (pixel < 0) continue;
            }
            ngi_film_add(bp.film, pixel, C * bp.film_scale);
        }
    }
}

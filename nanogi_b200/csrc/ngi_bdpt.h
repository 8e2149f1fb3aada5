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

// bdpt's building blocks are real functions (one copy each): inlined everywhere, the first k_bdpt was 14 700 SASS instructions
// (235 KB) and spent 38 of every 39 stalled issue slots waiting for instructions (profiles/r01_ncu_bdpt_v1.txt)
#if defined(__CUDACC__)
#define NGI_BD_FN __host__ __device__ __noinline__
#else
#define NGI_BD_FN inline
#endif

// one copy of each per-ray traversal for all of bdpt
NGI_BD_FN bool ngi_bd_trace_closest(const NgiDevScene& sc, const f3 o, const f3 d, NgiHitRec& h) {
    return ngi_trace_bvh8<false>(sc.nodes8, sc.tris8, o, d, NGI_EPS_F, NGI_INF_F, h);
}
NGI_BD_FN bool ngi_bd_trace_any(const NgiDevScene& sc, const f3 o, const f3 d, const float tmax) {
    NgiHitRec h;
    return ngi_trace_bvh8<true>(sc.nodes8, sc.tris8, o, d, NGI_EPS_F, tmax, h);
}

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
// (the *_inl forms are for the wavefront kernels, which call each building block from one or two places; the NGI_BD_FN wrappers are
// the single out-of-line copies the per-thread megakernel needs, see above)
NGI_HD f3 ngi_bd_eval_direction_inl(const NgiDevScene& sc, const NgiBdVertex& v, const int type, const f3 wi, const f3 wo, const bool transLE,
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
NGI_BD_FN f3 ngi_bd_eval_direction(const NgiDevScene& sc, const NgiBdVertex& v, const int type, const f3 wi, const f3 wo, const bool transLE,
                                   const bool forceDegenerated, float& pdf) {
    return ngi_bd_eval_direction_inl(sc, v, type, wi, wo, transLE, forceDegenerated, pdf);
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

// direction and GeometryTerm (rt.hpp:2364-2374) of the edge a -> b
NGI_HD void ngi_bd_edge(const NgiBdVertex& a, const NgiBdVertex& b, f3& w, float& G) {
    const double dx = b.px - a.px, dy = b.py - a.py, dz = b.pz - a.pz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const double inv = 1.0 / sqrt(d2);
    w = mk3((float)(dx * inv), (float)(dy * inv), (float)(dz * inv));
    float g = 1.0f;
    if (!a.degenerate) g *= fabsf(dot(a.sn, w));
    if (!b.degenerate) g *= fabsf(dot(b.sn, w));
    G = (float)((double)g / d2);
}
// does EvaluateDirection(forceDegenerated = false) of a vertex acting as `type` equal the forceDegenerated = true value? It differs
// only for specular lobes (-> 0, rt.hpp:1057-1060) and directional lights (-> 0, :934-937); precedence D > G > S
NGI_HD bool ngi_bd_nondegenerate_vertex(const NgiDevScene& sc, const NgiBdVertex& v, const int type) {
    if (type & NGI_L) return sc.prims[v.prim].l_type != NGI_LT_DIRECTIONAL;
    if (type & NGI_E) return true;
    return (type & (NGI_D | NGI_G)) != 0;
}

// ---- Path::SampleSubpath, bdpt.hpp:54-123, in three pieces shared by the per-thread loop below and the wavefront kernels
// (ngi_bdpt_wave.h). kind 0: light subpath (LE), 1: eye subpath (EL). --------------------------------------------------------
// vertex 0 (bdpt.hpp:60-75): a sampled light point / the sensor point. false: the subpath is empty (no lights).
NGI_HD bool ngi_bd_vertex0_inl(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind, NgiBdVertex& v) {
    const NgiDevSensor& E = sc.sensor;
    v.albedo = mk3(0.0f); v.pixel = -1;
    if (kind == 0) {
        if (sc.n_lights == 0) return false;
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
    return true;
}
NGI_BD_FN bool ngi_bd_vertex0(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind, NgiBdVertex& v) {
    return ngi_bd_vertex0_inl(sc, bp, sample, kind, v);
}
// the direction sampled at the subpath's last vertex `pv` in iteration `step` >= 1 (bdpt.hpp:77-90); `prev` = the vertex before it
// (nullptr at vertex 0). false: the subpath ends here. `rr` = the uniform of the Russian roulette that follows the hit (:108-113).
template <bool INL>
NGI_HD bool ngi_bd_sample_direction_t(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind, const int step,
                                      const NgiBdVertex& pv, const NgiBdVertex* prev, f3& wo, float& rr) {
    const NgiDevSensor& E = sc.sensor;
    const f3 wi = prev ? ngi_bd_dir(pv, *prev) : mk3(0.0f);                                                  // :77
    unsigned ra[4];
    philox4x32_10((unsigned)sample, (unsigned)(sample >> 32), (unsigned)(step - 1), kind == 0 ? 0u : 3u, bp.seed_lo, bp.seed_hi, ra);
    const float u0 = u01(ra[0]), u1 = u01(ra[1]), uc = u01(ra[2]);
    rr = u01(ra[3]);
    const NgiDevPrim& P = sc.prims[pv.prim];
    wo = mk3(0.0f);
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
    if (!wrote) return false;                                                                                 // wo stays zero -> f == 0
    float pdfUnused;
    const f3 f = INL ? ngi_bd_eval_direction_inl(sc, pv, pv.type, wi, wo, kind == 0, true, pdfUnused)         // :81
                     : ngi_bd_eval_direction(sc, pv, pv.type, wi, wo, kind == 0, true, pdfUnused);
    return !is_zero(f);
}
NGI_BD_FN bool ngi_bd_sample_direction(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind, const int step,
                                       const NgiBdVertex& pv, const NgiBdVertex* prev, f3& wo, float& rr) {
    return ngi_bd_sample_direction_t<false>(sc, bp, sample, kind, step, pv, prev, wo, rr);
}
// the vertex at the hit `h` of the ray (pv, wo) (bdpt.hpp:92-106)
NGI_HD void ngi_bd_hit_vertex_inl(const NgiDevScene& sc, const NgiBdParams& bp, const double ppx, const double ppy, const double ppz, const f3 wo,
                                  const NgiHitRec& h, NgiBdVertex& v) {
    double ddx, ddy, ddz;
    ngi_dither_direction(wo, make_float4(h.t, h.u, h.v, u2f(h.tri)), ddx, ddy, ddz);                          // see ngi_logic_surface
    v.px = ppx + ddx * (double)h.t; v.py = ppy + ddy * (double)h.t; v.pz = ppz + ddz * (double)h.t;
    NgiGeom g;
    v.prim = ngi_reconstruct(sc, h.tri, h.u, h.v, g);
    v.sn = g.sn; v.gn = g.gn; v.degenerate = 0;
    const NgiDevPrim& HP = sc.prims[v.prim];
    v.type = HP.type & ~NGI_EMITTER;                                                                          // :102
    const int tex = (v.type & NGI_D) ? HP.d_tex : HP.g_tex;
    v.albedo = (tex >= 0 && sc.shade_uv) ? ngi_texture_at_hit(sc, tex, h.tri, h.u, h.v) : ngi_constant_albedo(HP, v.type);
    v.pixel = ((HP.type & NGI_E) && sc.sensor.kind == NGI_ET_AREA && sc.shade_uv) ? ngi_area_sensor_pixel(sc, h.tri, h.u, h.v, bp.width, bp.height) : -1;
}
NGI_BD_FN void ngi_bd_hit_vertex(const NgiDevScene& sc, const NgiBdParams& bp, const double ppx, const double ppy, const double ppz, const f3 wo,
                                 const NgiHitRec& h, NgiBdVertex& v) {
    ngi_bd_hit_vertex_inl(sc, bp, ppx, ppy, ppz, wo, h, v);
}
NGI_HD int ngi_bd_vertex_cap(const NgiBdParams& bp) {
    return bp.max_verts == -1 ? NGI_BD_MAX_VERTS : (bp.max_verts < NGI_BD_MAX_VERTS ? bp.max_verts : NGI_BD_MAX_VERTS);
}
// the whole subpath, one vertex after the other (per-thread form). Returns the vertex count.
NGI_BD_FN int ngi_bd_sample_subpath(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, const int kind,
                                 NgiBdVertex* V, NgiBdCounters& cnt) {
    const int cap = ngi_bd_vertex_cap(bp);
    if (cap < 1 || !ngi_bd_vertex0(sc, bp, sample, kind, V[0])) return 0;
    int n = 1;
    for (int step = 1; step < cap; step++) {
        const NgiBdVertex& pv = V[n - 1];
        f3 wo; float rr;
        if (!ngi_bd_sample_direction(sc, bp, sample, kind, step, pv, n > 1 ? &V[n - 2] : nullptr, wo, rr)) break;
        NgiHitRec h;
        cnt.extend++;
        const f3 o = mk3((float)pv.px, (float)pv.py, (float)pv.pz);
        if (!ngi_bd_trace_closest(sc, o, wo, h)) break;                                                            // :92
        ngi_bd_hit_vertex(sc, bp, pv.px, pv.py, pv.pz, wo, h, V[n]);
        n++;
        if (rr > 0.5f) break;                                                                                     // :108-113
    }
    return n;
}

// A connected full path: views into the two subpaths (light vertices 0..s-1, then eye vertices t-1..0) with the type each vertex
// acts as (Connect overrides the end vertex's type when one subpath is empty, bdpt.hpp:139,153).
struct NgiBdPath {
    const NgiBdVertex* L; const NgiBdVertex* E;
    size_t stride;                  // distance between consecutive vertices of a subpath (1: per-thread arrays; the wavefront
                                    // kernels keep vertex k of every walker together, ngi_bdpt_wave.h)
    int s, t, n;
    int type_first, type_last;      // overrides (or the vertex's own type)
    NGI_HD const NgiBdVertex& v(int i) const { return i < s ? L[(size_t)i * stride] : E[(size_t)(n - 1 - i) * stride]; }
    NGI_HD int type(int i) const { return i == 0 ? type_first : (i == n - 1 ? type_last : v(i).type); }
};

// ---- contribution of one connected path --------------------------------------------------------------------------------
// The reference evaluates C = EvaluateUnweightContribution(s) * EvaluatePowerHeuristicsMISWeightOpt(s) with
//     w_s = 1 / sum_i (p_i / p_s)^2,   p_i = EvaluatePDF(i) = [EvaluateCst(i) != 0] * PL(i) * PE(i)
// where PL(i) / PE(i) are the products of (direction pdf x geometry term) along the first i / last n - i vertices
// (bdpt.hpp:491-535), and it recomputes every product from scratch: O(n) evaluations per strategy, O(n^2) per connection,
// O(n^4) per sample. The SAME numbers come out of one forward and one backward sweep over the path — prefix products — in
// O(n) evaluations per connection: the first version of this kernel restated the reference's loops literally and ran at
// 3.7 Mpaths/s on C2 (a warp lives as long as its longest path and the cost grew with the fourth power of its length).
struct NgiBdScratch {
    f3 w[2 * NGI_BD_MAX_VERTS];            // w[k] = direction vertex k -> k + 1
    float G[2 * NGI_BD_MAX_VERTS];         // GeometryTerm(v_k, v_k+1)
    f3 ff[2 * NGI_BD_MAX_VERTS];           // forward:  EvaluateDirection(v_k; from k-1, to k+1; LE, forceDegenerated) and its pdf
    float fpdf[2 * NGI_BD_MAX_VERTS];
    f3 bf[2 * NGI_BD_MAX_VERTS];           // backward: EvaluateDirection(v_k; from k+1, to k-1; EL, forceDegenerated) and its pdf
    float bpdf[2 * NGI_BD_MAX_VERTS];
    double PL[2 * NGI_BD_MAX_VERTS + 1], PE[2 * NGI_BD_MAX_VERTS + 1];
};

// Path::EvaluateCst(i), bdpt.hpp:217-250, from the cached edges
NGI_HD f3 ngi_bd_cst(const NgiDevScene& sc, const NgiBdPath& p, const NgiBdScratch& q, const int i) {
    const int n = p.n;
    float pdfUnused;
    if (i == 0) {
        const NgiBdVertex& v = p.v(0);
        return ngi_bd_eval_direction(sc, v, p.type(0), mk3(0.0f), q.w[0], false, false, pdfUnused) * ngi_bd_eval_position(sc, v, p.type(0), false);
    }
    if (i == n) {
        const NgiBdVertex& v = p.v(n - 1);
        return ngi_bd_eval_direction(sc, v, p.type(n - 1), mk3(0.0f), -q.w[n - 2], true, false, pdfUnused) * ngi_bd_eval_position(sc, v, p.type(n - 1), false);
    }
    const f3 fsL = ngi_bd_eval_direction(sc, p.v(i - 1), p.type(i - 1), i - 2 >= 0 ? -q.w[i - 2] : mk3(0.0f), q.w[i - 1], true, false, pdfUnused);
    if (is_zero(fsL)) return mk3(0.0f);
    const f3 fsE = ngi_bd_eval_direction(sc, p.v(i), p.type(i), i + 1 < n ? q.w[i] : mk3(0.0f), -q.w[i - 1], false, false, pdfUnused);
    return fsL * fsE * q.G[i - 1];
}

// [EvaluateCst(i) != 0] for a strategy other than the sampled one, WITHOUT new evaluations: cst(i) = fsL * G * fsE with
// forceDegenerated = false, which differs from the forceDegenerated = true values already in ff / bf only for specular lobes
// (-> 0, rt.hpp:1057-1060), directional lights (-> 0, :934-937) and the position terms of point lights / the pinhole (-> 0,
// :594-641); emitters do not depend on the transport direction.
NGI_HD bool ngi_bd_nondegenerate(const NgiDevScene& sc, const NgiBdPath& p, const int k) {
    return ngi_bd_nondegenerate_vertex(sc, p.v(k), p.type(k));
}
NGI_HD bool ngi_bd_cst_nonzero(const NgiDevScene& sc, const NgiBdPath& p, const NgiBdScratch& q, const int i) {
    const int n = p.n;
    if (i == 0) return ngi_bd_eval_position(sc, p.v(0), p.type(0), false) != 0.0f && ngi_bd_nondegenerate(sc, p, 0) && !is_zero(q.ff[0]);
    if (i == n) return ngi_bd_eval_position(sc, p.v(n - 1), p.type(n - 1), false) != 0.0f && ngi_bd_nondegenerate(sc, p, n - 1) && !is_zero(q.bf[n - 1]);
    return ngi_bd_nondegenerate(sc, p, i - 1) && ngi_bd_nondegenerate(sc, p, i) && !is_zero(q.ff[i - 1]) && !is_zero(q.bf[i]) && q.G[i - 1] != 0.0f;
}

// EvaluateContribution(s) / SelectionProb(s) of the connected path, bdpt.hpp:181-205, :252-343, :362-380, :491-535
NGI_BD_FN f3 ngi_bd_contribution(const NgiDevScene& sc, const NgiBdPath& p, NgiBdScratch& q) {
    const int n = p.n, s = p.s, t = p.t;
    for (int k = 0; k + 1 < n; k++) {                      // edges: direction and geometry term, once
        ngi_bd_edge(p.v(k), p.v(k + 1), q.w[k], q.G[k]);
    }
    // unweighted contribution first (most connections end here with zero): alphaL * cst(s) * alphaE, bdpt.hpp:252-343
    const f3 cstS = ngi_bd_cst(sc, p, q, s);
    if (is_zero(cstS)) return mk3(0.0f);
    for (int k = 0; k + 1 < n; k++)
        q.ff[k] = ngi_bd_eval_direction(sc, p.v(k), p.type(k), k >= 1 ? -q.w[k - 1] : mk3(0.0f), q.w[k], true, true, q.fpdf[k]);
    for (int k = n - 1; k >= 1; k--)
        q.bf[k] = ngi_bd_eval_direction(sc, p.v(k), p.type(k), k + 1 < n ? q.w[k] : mk3(0.0f), -q.w[k - 1], false, true, q.bpdf[k]);
    const float pA0 = ngi_bd_position_pdf(sc, p.v(0), p.type(0)), pAn = ngi_bd_position_pdf(sc, p.v(n - 1), p.type(n - 1));
    f3 alphaL = mk3(1.0f), alphaE = mk3(1.0f);
    if (s > 0) {
        alphaL = mk3(ngi_bd_eval_position(sc, p.v(0), p.type(0), true) / pA0);
        for (int k = 0; k < s - 1; k++) { if (is_zero(q.ff[k])) return mk3(0.0f); alphaL = alphaL * (q.ff[k] / q.fpdf[k]); }
    }
    if (t > 0) {
        alphaE = mk3(ngi_bd_eval_position(sc, p.v(n - 1), p.type(n - 1), true) / pAn);
        for (int k = n - 1; k > s; k--) { if (is_zero(q.bf[k])) return mk3(0.0f); alphaE = alphaE * (q.bf[k] / q.bpdf[k]); }
    }
    const f3 Cstar = alphaL * cstS * alphaE;
    if (is_zero(Cstar)) return mk3(0.0f);
    // path pdfs of every strategy from prefix / suffix products (fp64), bdpt.hpp:491-535
    q.PL[0] = 1.0; q.PL[1] = (double)pA0;
    for (int i = 2; i <= n; i++) q.PL[i] = q.PL[i - 1] * (double)q.fpdf[i - 2] * (double)q.G[i - 2];
    q.PE[n] = 1.0; q.PE[n - 1] = (double)pAn;
    for (int i = n - 2; i >= 0; i--) q.PE[i] = q.PE[i + 1] * (double)q.bpdf[i + 1] * (double)q.G[i];
    const double ps = q.PL[s] * q.PE[s];                                     // cst(s) != 0 was established above
    double invWeight = 0.0;                                                  // EvaluatePowerHeuristicsMISWeightOpt, bdpt.hpp:362-380
    for (int i = 0; i <= n; i++) {
        const double pi = q.PL[i] * q.PE[i];
        if (!(pi > 0.0)) continue;
        if (i != s && !ngi_bd_cst_nonzero(sc, p, q, i)) continue;            // EvaluatePDF returns 0 when EvaluateCst(i) == 0
        const double r = pi / ps;
        invWeight += r * r;
    }
    double sel = 1.0;                                                        // SelectionProb, bdpt.hpp:187-205
    for (int i = 1; i < s - 1; i++) sel *= 0.5;
    for (int i = t - 2; i >= 1; i--) sel *= 0.5;
    return Cstar * (float)(1.0 / (invWeight * sel));
}

// One (n, s) strategy of a sample: Path::Connect (bdpt.hpp:125-177) + contribution + film splat (src/nanogi.cpp:1164-1181), in
// two pieces around the visibility query: ngi_bd_connect_ray says whether the strategy needs one (and which), ngi_bd_connect_finish
// evaluates and splats a strategy that passed. `stride`: see NgiBdPath.
// Connect's tests that need no ray (bdpt.hpp:133-137, :147-151): can strategy (n, s) exist at all?
NGI_HD bool ngi_bd_strategy_possible(const NgiDevScene& sc, const NgiBdVertex* VL, const NgiBdVertex* VE, const int n, const int s, const size_t stride = 1) {
    const int t = n - s;
    if (s == 0) return (sc.prims[VE[(size_t)(t - 1) * stride].prim].type & NGI_L) != 0;
    if (t == 0) { const NgiBdVertex& a = VL[(size_t)(s - 1) * stride]; return (sc.prims[a.prim].type & NGI_E) != 0 && a.prim == sc.sensor.prim && a.pixel >= 0; }   // only THE sensor (last E primitive)
    return true;
}
// Scene::Visible's ray between the two end vertices a -> b (rt.hpp:2251-2261); only for s > 0 and t > 0
NGI_HD void ngi_bd_connect_ray(const double ax, const double ay, const double az, const double bx, const double by, const double bz,
                               f3& o, f3& d, float& tmax) {
    const double dx = bx - ax, dy = by - ay, dz = bz - az;
    const double len = sqrt(dx * dx + dy * dy + dz * dz);
    o = mk3((float)ax, (float)ay, (float)az);
    d = mk3((float)(dx / len), (float)(dy / len), (float)(dz / len));
    tmax = (float)len * (1.0f - NGI_EPS_F);
}
NGI_BD_FN void ngi_bd_connect_finish(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdVertex* VL, const NgiBdVertex* VE, const size_t stride,
                                     const int n, const int s, NgiBdScratch& q) {
    const int t = n - s;
    NgiBdPath p; p.L = VL; p.E = VE; p.stride = stride; p.s = s; p.t = t; p.n = n;
    if (s == 0) { p.type_first = NGI_L; p.type_last = VE[0].type; }
    else if (t == 0) { p.type_first = VL[0].type; p.type_last = NGI_E; }
    else { p.type_first = VL[0].type; p.type_last = VE[0].type; }
    const f3 C = ngi_bd_contribution(sc, p, q);
    if (is_zero(C) || !(C.x == C.x && C.y == C.y && C.z == C.z)) return;
    // Path::RasterPosition of the last vertex, bdpt.hpp:207-215
    int pixel;
    if (sc.sensor.kind == NGI_ET_PINHOLE) {
        float rx, ry, ct;
        if (!ngi_raster_position(sc.sensor, -q.w[n - 2], rx, ry, ct)) return;        // cannot happen when C != 0 (We = 0 off the raster)
        pixel = ngi_pixel_index(rx, ry, bp.width, bp.height);
    } else {
        pixel = p.v(n - 1).pixel;
        if (pixel < 0) return;
    }
    ngi_film_add(bp.film, pixel, C * bp.film_scale);
}
// per-thread form: tests, visibility query and evaluation in one go. Not inlined: the megakernel calls it from a per-lane state
// machine (k_bdpt), the simulator from the plain loop nest below.
NGI_BD_FN void ngi_bd_connect(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdVertex* VL, const NgiBdVertex* VE, const int n, const int s,
                                    NgiBdScratch& q, NgiBdCounters& cnt) {
    const int t = n - s;
    if (!ngi_bd_strategy_possible(sc, VL, VE, n, s)) return;
    if (s > 0 && t > 0) {
        f3 o, d; float tmax;
        ngi_bd_connect_ray(VL[s - 1].px, VL[s - 1].py, VL[s - 1].pz, VE[t - 1].px, VE[t - 1].py, VE[t - 1].pz, o, d, tmax);
        cnt.shadow++;
        if (ngi_bd_trace_any(sc, o, d, tmax)) return;
    }
    ngi_bd_connect_finish(sc, bp, VL, VE, 1, n, s, q);
}

// the (n, s) strategies of a sample in the reference's order (src/nanogi.cpp:1148-1160): advances the cursor, returns false when done
NGI_HD bool ngi_bd_next_strategy(const NgiBdParams& bp, const int nL, const int nE, int& n, int& s) {
    while (true) {
        const int maxS = nL < n ? nL : n;
        if (n >= 2 && s < maxS) { s++; return true; }
        n++;
        if (n > nE + nL || (bp.max_verts != -1 && n > bp.max_verts)) return false;
        s = (n - nE > 0 ? n - nE : 0) - 1;
    }
}

// one bdpt sample: ProcessSample_BDPT, src/nanogi.cpp:1133-1186 (the simulator's form; k_bdpt interleaves the strategies of
// different samples across the lanes of a warp)
NGI_BD_FN void ngi_bdpt_sample(const NgiDevScene& sc, const NgiBdParams& bp, const unsigned long long sample, NgiBdVertex* VL, NgiBdVertex* VE,
                                     NgiBdScratch& q, NgiBdCounters& cnt) {
    const int nL = ngi_bd_sample_subpath(sc, bp, sample, 0, VL, cnt);                        // :1137
    const int nE = ngi_bd_sample_subpath(sc, bp, sample, 1, VE, cnt);                        // :1138
    if (nL == 0 || nE == 0) return;
    int n = 1, s = 0;
    while (ngi_bd_next_strategy(bp, nL, nE, n, s)) ngi_bd_connect(sc, bp, VL, VE, n, s, q, cnt);
}

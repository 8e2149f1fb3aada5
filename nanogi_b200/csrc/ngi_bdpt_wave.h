// ngi_bdpt_wave.h — bidirectional path tracing as a wavefront: the per-item bodies of the k_bdw_* kernels.
//
// Replaces ProcessSample_BDPT (reference src/nanogi.cpp:1133-1186) over struct Path (reference include/nanogi/bdpt.hpp:38-539)
// for whole BATCHES of samples; the arithmetic of every piece is ngi_bdpt.h's (the per-thread form, kept as the cross-check
// renderer). Why: the per-thread megakernel k_bdpt ran 4.0 of 32 lanes with 15 % of the issue slots busy
// (profiles/r01_ncu_bdpt_v4.txt) — lanes of a warp sit in different phases of different samples, every ray is traced by the lane
// that needs it, and the kernel is 7 900 instructions long. Here a batch of B samples goes through dense stages instead:
//
//   subpaths    2 B "walkers" (walker = 2 * sample + kind; kind 0 light, 1 eye), one vertex per iteration:
//               k_bdw_start   vertex 0 + the first direction                               -> ray queue 1
//               k_bdw_extend  closest hits of ray queue k (persistent warp-cooperative trace, ngi_trace_warp.cuh)
//               k_bdw_step    hit -> vertex k, Russian roulette, next direction            -> ray queue k + 1
//                             + the subpath cache of vertex k - 1 (NgiBdCache below)
//               (the queue shrinks by >= 2x per iteration; vertex k of every walker is stored together: V[k][walker])
//   strategies  k_bdw_count   per sample: how many (n, s) strategies need a visibility ray / need none, and (block scan + one
//                             atomic per block) where its items go; the host reads the two totals (the only host sync of a batch)
//               k_bdw_expand  writes the strategy items (sample, n, s): connecting ones first, ray-less ones behind them
//               k_bdw_shadow  Scene::Visible of every connecting strategy (persistent any-hit trace); an occluded item's
//                             key becomes NGI_BDW_DEAD
//               radix sort (cub) of the items by (n, s): the lanes of a warp then run the same loop trip counts in the
//                             contribution stage (unsorted it ran 6.2 of 32 lanes, profiles/r01_ncu_bdw_v1.txt); dead items go last
//               k_bdw_contrib one item per lane: contribution x MIS weight from the subpath caches (4 BSDF evaluations) -> film
//
// Same Philox counters as the per-thread form, so both produce the same samples; film sums differ only in the order of the
// float atomics.
#pragma once
#include "ngi_bdpt.h"

// What the contribution stage needs of a subpath vertex k besides the vertex itself, so that one connection costs four BSDF
// evaluations instead of O(n): everything ngi_bd_contribution (ngi_bdpt.h) computes per connection that depends on ONE subpath only.
// "forward" = in the direction the subpath was sampled (k - 1 -> k -> k + 1) with the subpath's own transport mode (light: LE, eye:
// EL), "reverse" = k + 1 -> k -> k - 1 with the opposite mode; both with forceDegenerated = true (bdpt.hpp:491-535).
// 64 bytes = two 32-byte sectors: the second one holds everything the MIS sweep reads of a vertex that is not an end vertex.
struct alignas(64) NgiBdCache {
    f3 w;               // edge k -> k + 1: direction                                                       (k <= len - 2)
    f3 A;               // alpha: EvaluatePosition / pdfA x prod_{j<k} f_j / pdf_j (bdpt.hpp:252-343)
    float pad0, pad1;
    double P;           // pdfA x prod_{j<k} fp_j G_j: the density of sampling vertices 0..k (EvaluatePDF's subpath factor)
    float G;            // GeometryTerm of the edge k -> k + 1                                              (k <= len - 2)
    float rp;           // reverse pdf at k                                                                 (1 <= k <= len - 2)
    unsigned flags;     // NGI_BDC_*
    f3 back;            // direction k -> k - 1 (zero at vertex 0): the `wi` of every evaluation at k along its own subpath
};
#define NGI_BDC_ND 1u        /* ngi_bd_nondegenerate_vertex as its own type */
#define NGI_BDC_F_NZ 2u      /* forward value != 0 */
#define NGI_BDC_R_NZ 4u      /* reverse value != 0 */
#define NGI_BDC_POS_NZ 8u    /* EvaluatePosition(forceDegenerated = false) != 0 (vertex 0) */

struct NgiBdWave {
    NgiBdVertex* V;                  // [cap][walkers]
    NgiBdCache* C;                   // [cap][walkers]
    unsigned* nverts;                // [walkers] vertices of each subpath
    float4* rays[2];                 // ray queue k lives in rays[k & 1]: (o.xyz, rr uniform) (d.xyz, walker bits)
    float4* hits;                    // (t, u, v, triangle) per entry of the queue being traced
    unsigned* counts;                // [NGI_BD_MAX_VERTS + 1] entries of ray queue k
    unsigned* cursors;               // dynamic-fetch cursors of the trace launches: [k] ray queue k, [31] the visibility rays
    unsigned long long* offsets;     // [batch] where the items of each sample start: (ray items | ray-less items << 32)
    uint2* items;                    // x = sample within the batch, y = n | s << 8 (NGI_BDW_DEAD: occluded); [0, n_ray_items) connect with
                                     // a visibility ray, [n_ray_items, n_ray_items + n_rayless) need none
    uint2* items_sorted;             // the same items ordered by y
    unsigned long long first;        // first sample index of the batch
    unsigned batch;                  // samples in this batch
    unsigned walkers;                // 2 * batch capacity = stride between V[k] and V[k + 1]
    unsigned n_ray_items, n_rayless; // totals (known to the host after k_bdw_count)
};

#define NGI_BDW_DEAD 0xFFFFu

NGI_HD const NgiBdVertex* ngi_bdw_subpath(const NgiBdWave& wv, const unsigned sample_in_batch, const int kind) { return wv.V + 2u * sample_in_batch + (unsigned)kind; }

// k_bdw_start: vertex 0 of walker w and the direction sampled there. true: (o, wo, rr) is an entry of ray queue 1.
NGI_HD bool ngi_bdw_start(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const unsigned w, const int cap, f3& o, f3& wo, float& rr) {
    const unsigned long long sample = wv.first + (w >> 1);
    const int kind = (int)(w & 1u);
    NgiBdVertex v;
    if (cap < 1 || !ngi_bd_vertex0_inl(sc, bp, sample, kind, v)) { wv.nverts[w] = 0u; return false; }
    wv.V[w] = v;
    wv.nverts[w] = 1u;
    {
        NgiBdCache c;
        const float pA = ngi_bd_position_pdf(sc, v, v.type);
        c.w = mk3(0.0f); c.G = 0.0f; c.rp = 0.0f; c.back = mk3(0.0f); c.pad0 = c.pad1 = 0.0f;
        c.A = mk3(ngi_bd_eval_position(sc, v, v.type, true) / pA);
        c.P = (double)pA;
        c.flags = (ngi_bd_nondegenerate_vertex(sc, v, v.type) ? NGI_BDC_ND : 0u) | (ngi_bd_eval_position(sc, v, v.type, false) != 0.0f ? NGI_BDC_POS_NZ : 0u);
        wv.C[w] = c;
    }
    if (cap < 2) return false;
    if (!ngi_bd_sample_direction_t<true>(sc, bp, sample, kind, 1, v, nullptr, wo, rr)) return false;
    o = mk3((float)v.px, (float)v.py, (float)v.pz);
    return true;
}

// k_bdw_step: entry (r0, r1) of ray queue `step` and its hit -> vertex `step` of the walker, Russian roulette, next direction.
// true: (o, wo, rr) is an entry of ray queue step + 1.
NGI_HD bool ngi_bdw_step(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const int step, const int cap, const float4 r0, const float4 r1,
                         const float4 hit, unsigned& w, f3& o, f3& wo, float& rr) {
    w = f2u(r1.w);
    if (f2u(hit.w) == NGI_MISS) return false;                                                                     // bdpt.hpp:92
    const NgiBdVertex pv = wv.V[(size_t)(step - 1) * wv.walkers + w];
    NgiHitRec h; h.t = hit.x; h.u = hit.y; h.v = hit.z; h.tri = f2u(hit.w);
    NgiBdVertex v;
    ngi_bd_hit_vertex_inl(sc, bp, pv.px, pv.py, pv.pz, mk3(r1.x, r1.y, r1.z), h, v);
    wv.V[(size_t)step * wv.walkers + w] = v;
    wv.nverts[w] = (unsigned)step + 1u;
    {   // the previous vertex now has a successor: complete its cache record and start the new vertex's
        const bool light = (w & 1u) == 0u;
        NgiBdCache* cpp = wv.C + (size_t)(step - 1) * wv.walkers + w;
        NgiBdCache cp = *cpp;
        ngi_bd_edge(pv, v, cp.w, cp.G);
        const f3 back = cp.back;                                                                   // towards vertex step - 2
        float fp;
        const f3 f = ngi_bd_eval_direction_inl(sc, pv, pv.type, back, cp.w, light, true, fp);
        if (!is_zero(f)) cp.flags |= NGI_BDC_F_NZ;
        if (step >= 2) {
            const f3 r = ngi_bd_eval_direction_inl(sc, pv, pv.type, cp.w, back, !light, true, cp.rp);
            if (!is_zero(r)) cp.flags |= NGI_BDC_R_NZ;
        }
        *cpp = cp;
        NgiBdCache cn;
        cn.w = mk3(0.0f); cn.G = 0.0f; cn.rp = 0.0f; cn.pad0 = cn.pad1 = 0.0f;
        cn.back = -cp.w;
        cn.A = (is_zero(cp.A) || is_zero(f)) ? mk3(0.0f) : cp.A * (f / fp);
        cn.P = cp.P * (double)fp * (double)cp.G;
        cn.flags = ngi_bd_nondegenerate_vertex(sc, v, v.type) ? NGI_BDC_ND : 0u;
        wv.C[(size_t)step * wv.walkers + w] = cn;
    }
    if (r0.w > 0.5f) return false;                                                                                // :108-113
    if (step + 1 >= cap) return false;
    if (!ngi_bd_sample_direction_t<true>(sc, bp, wv.first + (w >> 1), (int)(w & 1u), step + 1, v, &pv, wo, rr)) return false;
    o = mk3((float)v.px, (float)v.py, (float)v.pz);
    return true;
}

// k_bdw_count / k_bdw_expand: the strategies of one sample in the reference's order (src/nanogi.cpp:1148-1160). With items == nullptr
// only counts; otherwise writes the connecting strategies from items + ray_at and the ray-less ones that pass Connect's tests
// (bdpt.hpp:133-137, :147-151) from items + n_ray_items + rayless_at.
NGI_HD void ngi_bdw_strategies(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const unsigned i, unsigned& n_ray, unsigned& n_rayless,
                               const bool write, unsigned ray_at, unsigned rayless_at) {
    n_ray = 0u; n_rayless = 0u;
    // the increments below are written as "+ one" with a value the compiler cannot prove constant: ptxas 12.9 (sm_100a) otherwise
    // keeps the two counters in UNIFORM registers although the branches that advance them diverge (k_bdw_expand's first build:
    // UIADD3 UR4, UR4, 1 inside the divergent region), so a lane's index advanced whenever ANY lane of its warp wrote an item
    const unsigned one = 1u + (i >> 31);
    const int nL = (int)wv.nverts[2u * i], nE = (int)wv.nverts[2u * i + 1u];
    if (nL == 0 || nE == 0) return;
    const NgiBdVertex* VL = ngi_bdw_subpath(wv, i, 0);
    const NgiBdVertex* VE = ngi_bdw_subpath(wv, i, 1);
    int n = 1, s = 0;
    while (ngi_bd_next_strategy(bp, nL, nE, n, s)) {
        const uint2 item = make_uint2(i, (unsigned)n | ((unsigned)s << 8));
        if (s > 0 && n - s > 0) {
            if (write) wv.items[ray_at + n_ray] = item;
            n_ray += one;
        } else if (ngi_bd_strategy_possible(sc, VL, VE, n, s, wv.walkers)) {
            if (write) wv.items[(size_t)wv.n_ray_items + rayless_at + n_rayless] = item;
            n_rayless += one;
        }
    }
}

// k_bdw_shadow: the visibility ray of a connecting strategy
NGI_HD void ngi_bdw_item_ray(const NgiBdWave& wv, const uint2 item, f3& o, f3& d, float& tmax) {
    const int n = (int)(item.y & 0xFFu), s = (int)((item.y >> 8) & 0xFFu), t = n - s;
    const NgiBdVertex* a = ngi_bdw_subpath(wv, item.x, 0) + (size_t)(s - 1) * wv.walkers;
    const NgiBdVertex* b = ngi_bdw_subpath(wv, item.x, 1) + (size_t)(t - 1) * wv.walkers;
    ngi_bd_connect_ray(a->px, a->py, a->pz, b->px, b->py, b->pz, o, d, tmax);
}

// k_bdw_contrib: contribution, MIS weight and splat of one strategy that passed Connect — ngi_bd_contribution's numbers
// (bdpt.hpp:181-205, :252-343, :362-380, :491-535) from the subpath caches: only the two END vertices a = y_{s-1} (light subpath)
// and b = z_{t-1} (eye subpath) have values that depend on the connection (their forward / reverse evaluations across the
// connecting edge); every other factor of every strategy's path pdf is a cached per-vertex number. Path indices: x_0..x_{n-1} =
// y_0..y_{s-1}, z_{t-1}..z_0. With s == 0 the vertex b acts as the light (type L), with t == 0 a acts as the sensor (type E).
NGI_HD void ngi_bdw_contrib(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const uint2 item) {
    if (item.y == NGI_BDW_DEAD) return;
    const int n = (int)(item.y & 0xFFu), s = (int)((item.y >> 8) & 0xFFu), t = n - s;
    const size_t W = wv.walkers;
    const NgiBdVertex* VL = ngi_bdw_subpath(wv, item.x, 0);
    const NgiBdVertex* VE = ngi_bdw_subpath(wv, item.x, 1);
    const NgiBdCache* CL = wv.C + 2u * item.x;
    const NgiBdCache* CE = CL + 1;
    const f3 zero = mk3(0.0f);

    // the end vertices, the directions to their subpath predecessors and the connection-specific evaluations
    NgiBdVertex a = NgiBdVertex(), b = NgiBdVertex();
    int typeA = 0, typeB = 0;
    bool ndA = false, ndB = false;
    f3 wiA = zero, wiB = zero, wc = zero;
    float Gc = 0.0f;
    f3 ffA = zero, bfA = zero, ffB = zero, bfB = zero;
    float fpA = 0.0f, bpA = 0.0f, fpB = 0.0f, bpB = 0.0f;
    f3 alphaL = mk3(1.0f), alphaE = mk3(1.0f);
    double PLs = 1.0, PEs = 1.0;
    if (s >= 1) {
        a = VL[(size_t)(s - 1) * W];
        typeA = t == 0 ? NGI_E : a.type;
        ndA = ngi_bd_nondegenerate_vertex(sc, a, typeA);
        const NgiBdCache& ca = CL[(size_t)(s - 1) * W];
        wiA = ca.back; alphaL = ca.A; PLs = ca.P;
    }
    if (t >= 1) {
        b = VE[(size_t)(t - 1) * W];
        typeB = s == 0 ? NGI_L : b.type;
        ndB = ngi_bd_nondegenerate_vertex(sc, b, typeB);
        const NgiBdCache& cb = CE[(size_t)(t - 1) * W];
        wiB = cb.back; alphaE = cb.A; PEs = cb.P;
    }
    f3 cstS;                                                                                   // EvaluateCst(s), bdpt.hpp:217-250
    if (s >= 1 && t >= 1) {
        ngi_bd_edge(a, b, wc, Gc);
        ffA = ngi_bd_eval_direction_inl(sc, a, typeA, wiA, wc, true, true, fpA);                   // ff[s-1]
        if (!ndA || is_zero(ffA)) return;
        bfB = ngi_bd_eval_direction_inl(sc, b, typeB, wiB, -wc, false, true, bpB);                 // bf[s]
        cstS = ndB ? ffA * bfB * Gc : zero;
    } else if (s == 0) {
        ffB = ngi_bd_eval_direction_inl(sc, b, typeB, zero, wiB, true, true, fpB);                 // ff[0]: b emits towards z_{t-2}
        cstS = ndB ? ffB * ngi_bd_eval_position(sc, b, typeB, false) : zero;
    } else {
        bfA = ngi_bd_eval_direction_inl(sc, a, typeA, zero, wiA, false, true, bpA);                // bf[n-1]: a senses from y_{s-2}
        cstS = ndA ? bfA * ngi_bd_eval_position(sc, a, typeA, false) : zero;
    }
    if (is_zero(cstS)) return;
    const f3 Cstar = alphaL * cstS * alphaE;
    if (is_zero(Cstar)) return;
    if (s >= 1 && t >= 1) {                                                                    // the other two, only needed for the weight
        if (s >= 2) bfA = ngi_bd_eval_direction_inl(sc, a, typeA, wc, wiA, false, true, bpA);      // bf[s-1]
        if (t >= 2) ffB = ngi_bd_eval_direction_inl(sc, b, typeB, -wc, wiB, true, true, fpB);      // ff[s]
    }

    // EvaluatePowerHeuristicsMISWeightOpt: sum over the strategies i of (p_i / p_s)^2, p_i = PL[i] PE[i] if EvaluateCst(i) != 0
    const double ps = PLs * PEs;
    double invWeight = ps > 0.0 ? 1.0 : 0.0;
    if (t >= 1) {       // strategies i = s + 1 .. n: the light side grows along the eye subpath; j = n - i eye vertices remain
        double PL = s == 0 ? (double)ngi_bd_position_pdf(sc, b, typeB) : PLs * (double)fpA * (double)Gc;
        for (int j = t - 1; j >= 0; j--) {
            const double PE = j >= 1 ? CE[(size_t)(j - 1) * W].P : 1.0;
            const double pi = PL * PE;
            const NgiBdCache cj = CE[(size_t)j * W];                                           // vertex z_j = x_{i-1}
            if (pi > 0.0) {
                bool nz;
                if (j == 0) {                                                                  // i == n: z_0 reached by the light path
                    if (t == 1) nz = ngi_bd_eval_position(sc, b, typeB, false) != 0.0f && ndB && !is_zero(bfB);
                    else nz = (cj.flags & (NGI_BDC_POS_NZ | NGI_BDC_ND | NGI_BDC_F_NZ)) == (NGI_BDC_POS_NZ | NGI_BDC_ND | NGI_BDC_F_NZ);
                } else {                                                                       // connects z_j and z_{j-1}
                    const bool ndj = j == t - 1 ? ndB : (cj.flags & NGI_BDC_ND) != 0u;
                    const bool ffj = j == t - 1 ? !is_zero(ffB) : (cj.flags & NGI_BDC_R_NZ) != 0u;
                    const NgiBdCache& c1 = CE[(size_t)(j - 1) * W];
                    nz = ndj && ffj && (c1.flags & (NGI_BDC_ND | NGI_BDC_F_NZ)) == (NGI_BDC_ND | NGI_BDC_F_NZ) && c1.G != 0.0f;
                }
                if (nz) { const double r = pi / ps; invWeight += r * r; }
            }
            if (j >= 1) PL = PL * (double)(j == t - 1 ? fpB : cj.rp) * (double)CE[(size_t)(j - 1) * W].G;
        }
    }
    if (s >= 1) {       // strategies i = s - 1 .. 0: the eye side grows along the light subpath
        double PE = t == 0 ? (double)ngi_bd_position_pdf(sc, a, typeA) : PEs * (double)bpB * (double)Gc;
        for (int k = s - 1; k >= 0; k--) {
            const double PL = k >= 1 ? CL[(size_t)(k - 1) * W].P : 1.0;
            const double pi = PL * PE;
            const NgiBdCache ck = CL[(size_t)k * W];                                           // vertex y_k = x_i
            if (pi > 0.0) {
                bool nz;
                if (k == 0) {                                                                  // i == 0: y_0 reached by the eye path
                    if (s == 1) nz = ngi_bd_eval_position(sc, a, typeA, false) != 0.0f && ndA && !is_zero(ffA);
                    else nz = (ck.flags & (NGI_BDC_POS_NZ | NGI_BDC_ND | NGI_BDC_F_NZ)) == (NGI_BDC_POS_NZ | NGI_BDC_ND | NGI_BDC_F_NZ);
                } else {                                                                       // connects y_{k-1} and y_k
                    const bool ndk = k == s - 1 ? ndA : (ck.flags & NGI_BDC_ND) != 0u;
                    const bool bfk = k == s - 1 ? !is_zero(bfA) : (ck.flags & NGI_BDC_R_NZ) != 0u;
                    const NgiBdCache& c1 = CL[(size_t)(k - 1) * W];
                    nz = ndk && bfk && (c1.flags & (NGI_BDC_ND | NGI_BDC_F_NZ)) == (NGI_BDC_ND | NGI_BDC_F_NZ) && c1.G != 0.0f;
                }
                if (nz) { const double r = pi / ps; invWeight += r * r; }
            }
            if (k >= 1) PE = PE * (double)(k == s - 1 ? bpA : ck.rp) * (double)CL[(size_t)(k - 1) * W].G;
        }
    }
    double sel = 1.0;                                                                          // SelectionProb, bdpt.hpp:187-205
    for (int i = 1; i < s - 1; i++) sel *= 0.5;
    for (int i = t - 2; i >= 1; i--) sel *= 0.5;
    const f3 C = Cstar * (float)(1.0 / (invWeight * sel));
    if (is_zero(C) || !(C.x == C.x && C.y == C.y && C.z == C.z)) return;
    int pixel;                                                                                 // Path::RasterPosition, bdpt.hpp:207-215
    if (sc.sensor.kind == NGI_ET_PINHOLE) {
        const f3 toPrev = t >= 2 ? CE[0].w : (t == 1 ? -wc : wiA);                             // from x_{n-1} towards x_{n-2}
        float rx, ry, ct;
        if (!ngi_raster_position(sc.sensor, toPrev, rx, ry, ct)) return;
        pixel = ngi_pixel_index(rx, ry, bp.width, bp.height);
    } else {
        pixel = t >= 1 ? VE[0].pixel : a.pixel;
        if (pixel < 0) return;
    }
    ngi_film_add(bp.film, pixel, C * bp.film_scale);
}

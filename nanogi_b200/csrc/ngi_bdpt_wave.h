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
//               (the queue shrinks by >= 2x per iteration; vertex k of every walker is stored together: V[k][walker])
//   strategies  k_bdw_count   per sample: how many (n, s) strategies need a visibility ray / need none
//               exclusive scan (cub) -> offsets; the host reads the two totals (the only host sync of a batch)
//               k_bdw_expand  writes the strategy items (sample, n, s): connecting ones first, ray-less ones behind them
//               k_bdw_shadow  Scene::Visible of every connecting strategy (persistent any-hit trace); an occluded item's
//                             key becomes NGI_BDW_DEAD
//               radix sort (cub) of the items by (n, s): the lanes of a warp then run the same loop trip counts in the
//                             contribution stage (unsorted it ran 6.2 of 32 lanes, profiles/r01_ncu_bdw_v1.txt); dead items go last
//               k_bdw_contrib one item per lane: contribution x MIS weight (ngi_bd_connect_finish) -> film
//
// Same Philox counters as the per-thread form, so both produce the same samples; film sums differ only in the order of the
// float atomics.
#pragma once
#include "ngi_bdpt.h"

struct NgiBdWave {
    NgiBdVertex* V;                  // [cap][walkers]
    unsigned* nverts;                // [walkers] vertices of each subpath
    float4* rays[2];                 // ray queue k lives in rays[k & 1]: (o.xyz, rr uniform) (d.xyz, walker bits)
    float4* hits;                    // (t, u, v, triangle) per entry of the queue being traced
    unsigned* counts;                // [NGI_BD_MAX_VERTS + 1] entries of ray queue k
    unsigned* cursors;               // [NGI_BD_MAX_VERTS + 1] dynamic-fetch cursors of the trace launches; [NGI_BD_MAX_VERTS] = shadow
    unsigned long long* offsets;     // [batch] where the items of each sample start: (ray items | ray-less items << 32)
    uint2* items;                    // x = sample within the batch, y = n | s << 8 (NGI_BDW_DEAD: occluded); [0, n_ray_items) connect with
                                     // a visibility ray, [n_ray_items, n_ray_items + n_rayless) need none
    uint2* items_sorted;             // the same items ordered by y
    unsigned long long first;        // first sample index of the batch
    unsigned batch;                  // samples in this batch
    unsigned walkers;                // 2 * batch capacity = stride between V[k] and V[k + 1]
    unsigned n_ray_items, n_rayless; // totals (known to the host after the scan)
};

#define NGI_BDW_DEAD 0xFFFFu

NGI_HD const NgiBdVertex* ngi_bdw_subpath(const NgiBdWave& wv, const unsigned sample_in_batch, const int kind) { return wv.V + 2u * sample_in_batch + (unsigned)kind; }

// k_bdw_start: vertex 0 of walker w and the direction sampled there. true: (o, wo, rr) is an entry of ray queue 1.
NGI_HD bool ngi_bdw_start(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const unsigned w, const int cap, f3& o, f3& wo, float& rr) {
    const unsigned long long sample = wv.first + (w >> 1);
    const int kind = (int)(w & 1u);
    NgiBdVertex v;
    if (cap < 1 || !ngi_bd_vertex0(sc, bp, sample, kind, v)) { wv.nverts[w] = 0u; return false; }
    wv.V[w] = v;
    wv.nverts[w] = 1u;
    if (cap < 2) return false;
    if (!ngi_bd_sample_direction(sc, bp, sample, kind, 1, v, nullptr, wo, rr)) return false;
    o = mk3((float)v.px, (float)v.py, (float)v.pz);
    return true;
}

// k_bdw_step: entry (r0, r1) of ray queue `step` and its hit -> vertex `step` of the walker, Russian roulette, next direction.
// true: (o, wo, rr) is an entry of ray queue step + 1.
NGI_HD bool ngi_bdw_step(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const int step, const int cap, const float4 r0, const float4 r1,
                         const float4 hit, unsigned& w, f3& o, f3& wo, float& rr) {
    w = f2u(r1.w);
    if (f2u(hit.w) == NGI_MISS) return false;                                                                     // bdpt.hpp:92
    const NgiBdVertex* pvp = wv.V + (size_t)(step - 1) * wv.walkers + w;
    NgiBdVertex pv;                                                                                               // only its position is needed
    pv.px = pvp->px; pv.py = pvp->py; pv.pz = pvp->pz;
    NgiHitRec h; h.t = hit.x; h.u = hit.y; h.v = hit.z; h.tri = f2u(hit.w);
    NgiBdVertex v;
    ngi_bd_hit_vertex(sc, bp, pv.px, pv.py, pv.pz, mk3(r1.x, r1.y, r1.z), h, v);
    wv.V[(size_t)step * wv.walkers + w] = v;
    wv.nverts[w] = (unsigned)step + 1u;
    if (r0.w > 0.5f) return false;                                                                                // :108-113
    if (step + 1 >= cap) return false;
    if (!ngi_bd_sample_direction(sc, bp, wv.first + (w >> 1), (int)(w & 1u), step + 1, v, &pv, wo, rr)) return false;
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

// k_bdw_contrib: contribution, MIS weight and splat of one strategy that passed Connect
NGI_HD void ngi_bdw_contrib(const NgiDevScene& sc, const NgiBdParams& bp, const NgiBdWave& wv, const uint2 item, NgiBdScratch& q) {
    if (item.y == NGI_BDW_DEAD) return;
    const int n = (int)(item.y & 0xFFu), s = (int)((item.y >> 8) & 0xFFu);
    ngi_bd_connect_finish(sc, bp, ngi_bdw_subpath(wv, item.x, 0), ngi_bdw_subpath(wv, item.x, 1), wv.walkers, n, s, q);
}

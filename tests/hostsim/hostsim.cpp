// hostsim.cpp — TEST-ONLY single-threaded simulator of the device code.
//
// This container has no GPU, so the per-item bodies of every CUDA kernel (nanogi_b200/csrc/ngi_*.h, all
// `__host__ __device__`) are also compiled here with g++ and driven by plain loops in the order the
// kernels run. It lets `-m "not gpu"` tests check the device algorithms (LBVH build, BVH8 collapse and
// traversal, fp32 shading, wavefront logic, Philox streams) against the oracle before any GPU time is
// spent. It is NOT part of the product and is NOT a CPU fallback: libnanogi_gpu.so never links it and
// nothing outside tests/ loads it. Compile with -ffp-contract=off (see tests/hostsim/Makefile).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

static unsigned long long g_trace_count[2] = {0, 0};     // [0] BVH8 node steps, [1] triangle tests (sim_trace_stats)
#define NGI_TRACE_COUNT(what) (g_trace_count[what]++)
#include "../../nanogi_b200/csrc/ngi_build.h"
#include "../../nanogi_b200/csrc/ngi_bvh.h"
#include "../../nanogi_b200/csrc/ngi_scene_host.h"
#include "../../nanogi_b200/csrc/ngi_wave.h"
#include "../../nanogi_b200/csrc/ngi_bdpt.h"
#include "../../nanogi_b200/csrc/ngi_bdpt_wave.h"

namespace {

struct SimScene {
    NgiHostArrays ha;
    unsigned n = 0;  // padded triangle count
    std::vector<float4> rec_in, lo, hi;        // input order records, node boxes [2n-1]
    std::vector<float4> tris2, nodes2, tris8;
    std::vector<uint4> nodes8;
    std::vector<int> left, right;
    std::vector<unsigned> cnt;
    std::vector<NgiDpRow> dp;                  // SAH-optimal collapse decisions (ngi_dp_node)
    unsigned n_nodes8 = 0, depth8 = 0;
    float smin[3], smax[3], pad = 0;
    NgiDevScene dev;
};

thread_local std::string g_err;

}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char* sim_last_error() { return g_err.c_str(); }

__attribute__((visibility("default"))) void* sim_scene_create(const NgiSceneDesc* desc) {
    SimScene* s = new SimScene;
    if (!ngi_prepare_scene(desc, s->ha, true)) { g_err = s->ha.error; delete s; return nullptr; }
    const unsigned nr = s->ha.n_real;
    const unsigned n = nr < 2 ? 2 : nr;
    s->n = n;
    for (int k = 0; k < 3; k++) { s->smin[k] = 3.0e38f; s->smax[k] = -3.0e38f; }
    for (size_t i = 0; i < (size_t)nr * 3; i++)
        for (int k = 0; k < 3; k++) { s->smin[k] = fminf(s->smin[k], desc->positions[i * 3 + k]); s->smax[k] = fmaxf(s->smax[k], desc->positions[i * 3 + k]); }
    if (nr == 0) for (int k = 0; k < 3; k++) { s->smin[k] = 0; s->smax[k] = 0; }
    s->pad = ngi_box_pad(s->smin, s->smax, s->ha.sensor);
    const f3 anchor = mk3(s->smin[0], s->smin[1], s->smin[2]);
    s->rec_in.resize((size_t)n * 3);
    std::vector<float4> tlo(n), thi(n);
    for (unsigned i = 0; i < n; i++) ngi_tri_setup(desc->positions, i, nr, s->pad, anchor, s->rec_in.data(), tlo.data(), thi.data());
    // Morton + sort
    const f3 mmin = mk3(s->smin[0] - s->pad, s->smin[1] - s->pad, s->smin[2] - s->pad);
    f3 sinv;
    sinv.x = 1.0f / fmaxf(s->smax[0] - s->smin[0] + 2 * s->pad, 1e-30f);
    sinv.y = 1.0f / fmaxf(s->smax[1] - s->smin[1] + 2 * s->pad, 1e-30f);
    sinv.z = 1.0f / fmaxf(s->smax[2] - s->smin[2] + 2 * s->pad, 1e-30f);
    std::vector<unsigned long long> keys(n);
    for (unsigned i = 0; i < n; i++) {
        const f3 c = mk3(0.5f * (tlo[i].x + thi[i].x), 0.5f * (tlo[i].y + thi[i].y), 0.5f * (tlo[i].z + thi[i].z));
        keys[i] = ngi_morton63(c, mmin, sinv);
    }
    std::vector<unsigned> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return keys[a] < keys[b]; });
    std::vector<unsigned long long> skeys(n);
    s->tris2.resize((size_t)n * 3);
    s->lo.resize(2 * (size_t)n - 1); s->hi.resize(2 * (size_t)n - 1);
    for (unsigned k = 0; k < n; k++) {
        const unsigned i = order[k];
        skeys[k] = keys[i];
        for (int j = 0; j < 3; j++) s->tris2[(size_t)k * 3 + j] = s->rec_in[(size_t)i * 3 + j];
        s->lo[n - 1 + k] = tlo[i]; s->hi[n - 1 + k] = thi[i];
    }
    const float c_node = 1.0f, c_prim = getenv("NGI_SAH_CPRIM") ? (float)atof(getenv("NGI_SAH_CPRIM")) : NGI_SAH_C_PRIM;
    const bool greedy = getenv("NGI_COLLAPSE_GREEDY") != nullptr;
    s->dp.assign(n - 1, NgiDpRow());
    if (getenv("NGI_SIM_SAH")) {
        // EXPERIMENT (build-quality yardstick, not a product path): top-down binned-SAH binary tree over the same leaves,
        // collapsed by the same ngi_collapse_node — how much traversal cost is left in the PLOC topology?
        s->left.assign(n - 1, 0); s->right.assign(n - 1, 0); s->cnt.assign(n - 1, 0u);
        std::vector<unsigned> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::vector<float4>& lo = s->lo; std::vector<float4>& hi = s->hi;
        int next_inner = 1;
        struct Task { int node; unsigned b, e; };
        std::vector<Task> st{{0, 0u, n}};
        auto area = [](const float* mn, const float* mx) { const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2]; return dx * dy + dy * dz + dz * dx; };
        while (!st.empty()) {
            const Task t = st.back(); st.pop_back();
            float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f}, cmn[3] = {3e38f, 3e38f, 3e38f}, cmx[3] = {-3e38f, -3e38f, -3e38f};
            for (unsigned i = t.b; i < t.e; i++) {
                const float4 a = lo[n - 1 + idx[i]], b = hi[n - 1 + idx[i]];
                const float l3[3] = {a.x, a.y, a.z}, h3[3] = {b.x, b.y, b.z};
                for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], l3[k]); mx[k] = fmaxf(mx[k], h3[k]); const float c = 0.5f * (l3[k] + h3[k]); cmn[k] = fminf(cmn[k], c); cmx[k] = fmaxf(cmx[k], c); }
            }
            lo[t.node] = make_float4(mn[0], mn[1], mn[2], 0); hi[t.node] = make_float4(mx[0], mx[1], mx[2], 0);
            s->cnt[t.node] = t.e - t.b;
            const int NB = 32;
            int bestAxis = -1, bestBin = -1; float bestCost = 3e38f;
            for (int ax = 0; ax < 3; ax++) {
                const float ext = cmx[ax] - cmn[ax];
                if (!(ext > 0)) continue;
                float bmn[NB][3], bmx[NB][3]; int bc[NB];
                for (int b = 0; b < NB; b++) { bc[b] = 0; for (int k = 0; k < 3; k++) { bmn[b][k] = 3e38f; bmx[b][k] = -3e38f; } }
                const float sc = NB / ext;
                for (unsigned i = t.b; i < t.e; i++) {
                    const float4 a = lo[n - 1 + idx[i]], b4 = hi[n - 1 + idx[i]];
                    const float l3[3] = {a.x, a.y, a.z}, h3[3] = {b4.x, b4.y, b4.z};
                    const int b = std::min(NB - 1, (int)((0.5f * (l3[ax] + h3[ax]) - cmn[ax]) * sc));
                    bc[b]++;
                    for (int k = 0; k < 3; k++) { bmn[b][k] = fminf(bmn[b][k], l3[k]); bmx[b][k] = fmaxf(bmx[b][k], h3[k]); }
                }
                float rA[NB]; int rC[NB];
                float amn[3] = {3e38f, 3e38f, 3e38f}, amx[3] = {-3e38f, -3e38f, -3e38f}; int c = 0;
                for (int b = NB - 1; b > 0; b--) { for (int k = 0; k < 3; k++) { amn[k] = fminf(amn[k], bmn[b][k]); amx[k] = fmaxf(amx[k], bmx[b][k]); } c += bc[b]; rA[b] = c ? area(amn, amx) : 0; rC[b] = c; }
                for (int k = 0; k < 3; k++) { amn[k] = 3e38f; amx[k] = -3e38f; }
                c = 0;
                for (int b = 0; b < NB - 1; b++) {
                    for (int k = 0; k < 3; k++) { amn[k] = fminf(amn[k], bmn[b][k]); amx[k] = fmaxf(amx[k], bmx[b][k]); }
                    c += bc[b];
                    if (c == 0 || rC[b + 1] == 0) continue;
                    const float cost = area(amn, amx) * c + rA[b + 1] * rC[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestBin = b; }
                }
            }
            unsigned mid;
            if (bestAxis < 0) mid = (t.b + t.e) / 2;
            else {
                const float sc = NB / (cmx[bestAxis] - cmn[bestAxis]);
                auto it = std::partition(idx.begin() + t.b, idx.begin() + t.e, [&](unsigned id) {
                    const float4 a = lo[n - 1 + id], b4 = hi[n - 1 + id];
                    const float c = bestAxis == 0 ? 0.5f * (a.x + b4.x) : bestAxis == 1 ? 0.5f * (a.y + b4.y) : 0.5f * (a.z + b4.z);
                    return std::min(NB - 1, (int)((c - cmn[bestAxis]) * sc)) <= bestBin; });
                mid = (unsigned)(it - idx.begin());
                if (mid == t.b || mid == t.e) mid = (t.b + t.e) / 2;
            }
            auto child = [&](unsigned b, unsigned e) -> int {
                if (e - b == 1) return (int)(n - 1 + idx[b]);
                const int id = next_inner++;
                st.push_back({id, b, e});
                return id;
            };
            s->left[t.node] = child(t.b, mid);
            s->right[t.node] = child(mid, t.e);
        }
        // children carry larger ids than their parent here: rows bottom-up = ids downwards
        for (int id = (int)n - 2; id >= 0; id--) {
            const float4 a = lo[id], b = hi[id];
            const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
            ngi_dp_node(s->dp.data(), lo.data(), hi.data(), (int)n, id, s->left[id], s->right[id], s->cnt[id], dx * dy + dy * dz + dz * dx, c_node, c_prim);
        }
    } else {
    // PLOC rounds, same per-item functions as the CUDA kernels (k_ploc_*)
    s->left.assign(n - 1, 0); s->right.assign(n - 1, 0); s->cnt.assign(n - 1, 0u);
    {
        std::vector<int> cid[2] = {std::vector<int>(n), std::vector<int>(n)}, nn(n);
        std::vector<float4> clo[2] = {std::vector<float4>(n), std::vector<float4>(n)}, chi[2] = {std::vector<float4>(n), std::vector<float4>(n)};
        std::vector<unsigned> keep(n), pos(n);
        for (unsigned k = 0; k < n; k++) { cid[0][k] = (int)(n - 1 + k); clo[0][k] = s->lo[n - 1 + k]; chi[0][k] = s->hi[n - 1 + k]; }
        unsigned C = n, merges = 0;
        int cur = 0;
        while (C > 1) {
            for (unsigned i = 0; i < C; i++) nn[i] = ngi_ploc_nearest(clo[cur].data(), chi[cur].data(), (int)C, (int)i);
            unsigned acc = 0;
            for (unsigned i = 0; i < C; i++) { keep[i] = ngi_ploc_keep(nn.data(), (int)i); pos[i] = acc; acc += keep[i]; }
            NgiPlocCtx pc;
            pc.nn = nn.data(); pc.pos = pos.data(); pc.cid_in = cid[cur].data(); pc.clo_in = clo[cur].data(); pc.chi_in = chi[cur].data();
            pc.cid_out = cid[cur ^ 1].data(); pc.clo_out = clo[cur ^ 1].data(); pc.chi_out = chi[cur ^ 1].data();
            pc.lo = s->lo.data(); pc.hi = s->hi.data(); pc.left = s->left.data(); pc.right = s->right.data(); pc.cnt = s->cnt.data();
            pc.n = (int)n; pc.next_id = (int)(n - 2) - (int)merges;
            pc.dp = s->dp.data(); pc.c_node = c_node; pc.c_prim = c_prim; pc.depth = nullptr;
            for (unsigned i = 0; i < C; i++) ngi_ploc_merge(pc, (int)i);
            merges += C - acc; C = acc; cur ^= 1;
        }
    }
    }
    s->nodes2.resize((size_t)(n - 1) * 4);
    for (int i = 0; i < (int)n - 1; i++) ngi_pack2(s->lo.data(), s->hi.data(), s->left.data(), s->right.data(), (int)n, i, s->nodes2.data());
    // collapse
    s->nodes8.assign((size_t)n * 5, make_uint4(0, 0, 0, 0));
    s->tris8.resize((size_t)n * 3);
    unsigned counters[3] = {1, 0, 0};
    std::vector<NgiBuildTask> cur{{0, 0u}}, next(n);
    NgiCollapseCtx c;
    c.lo = s->lo.data(); c.hi = s->hi.data(); c.left = s->left.data(); c.right = s->right.data(); c.cnt = s->cnt.data();
    c.tris2 = s->tris2.data(); c.n = (int)n; c.nodes8 = s->nodes8.data(); c.tris8 = s->tris8.data(); c.counters = counters;
    c.dp = greedy ? nullptr : s->dp.data();
    unsigned depth = 0;
    while (!cur.empty()) {
        depth++;
        counters[2] = 0;
        c.out_tasks = next.data();
        for (auto& t : cur) ngi_collapse_node(c, t);
        cur.assign(next.begin(), next.begin() + counters[2]);
    }
    s->n_nodes8 = counters[0];
    s->depth8 = depth;
    if (counters[1] != n) { g_err = "collapse lost triangles: " + std::to_string(counters[1]) + " of " + std::to_string(n); delete s; return nullptr; }
    {   // traversal form: valid24 masks + triangles at their fixed places (k_expand8)
        std::vector<float4> fixed((size_t)s->n_nodes8 * 24 * 3, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        for (unsigned i = 0; i < s->n_nodes8; i++) ngi_expand_node(s->nodes8.data(), s->tris8.data(), i, s->nodes8.data(), fixed.data());
        s->tris8.swap(fixed);
    }
    NgiDevScene& d = s->dev;
    d.nodes8 = s->nodes8.data(); d.tris8 = s->tris8.data(); d.nodes2 = s->nodes2.data(); d.tris2 = s->tris2.data();
    d.shade_tris = s->ha.shade_tris.data(); d.prims = s->ha.prims.data(); d.light_prims = s->ha.light_prims.data();
    d.shade_uv = s->ha.shade_uv.empty() ? nullptr : s->ha.shade_uv.data();
    d.textures = s->ha.textures.data(); d.tex_data = s->ha.tex_data.data();
    d.cdf = s->ha.cdf.data(); d.n_tris = n; d.n_lights = (unsigned)s->ha.light_prims.size(); d.sensor = s->ha.sensor;
    return s;
}

__attribute__((visibility("default"))) void sim_scene_destroy(void* h) { delete (SimScene*)h; }

// out = {n padded, nodes8, depth8, nodes2, pad}
__attribute__((visibility("default"))) void sim_scene_info(void* h, double* out) {
    SimScene* s = (SimScene*)h;
    out[0] = s->n; out[1] = s->n_nodes8; out[2] = s->depth8; out[3] = s->n - 1; out[4] = s->pad;
}

__attribute__((visibility("default"))) int sim_trace(void* h, const NgiRay* rays, uint64_t n, NgiHit* hits, int any_hit, int accel) {
    SimScene* s = (SimScene*)h;
    for (uint64_t i = 0; i < n; i++) {
        const NgiRay& r = rays[i];
        const f3 o = mk3(r.o[0], r.o[1], r.o[2]), d = mk3(r.d[0], r.d[1], r.d[2]);
        NgiHitRec hr; bool hit;
        if (accel == 0) hit = any_hit ? ngi_trace_bvh8<true>(s->dev.nodes8, s->dev.tris8, o, d, r.tmin, r.tmax, hr) : ngi_trace_bvh8<false>(s->dev.nodes8, s->dev.tris8, o, d, r.tmin, r.tmax, hr);
        else if (accel == 1) hit = any_hit ? ngi_trace_bvh2<true>(s->dev.nodes2, s->dev.tris2, o, d, r.tmin, r.tmax, hr) : ngi_trace_bvh2<false>(s->dev.nodes2, s->dev.tris2, o, d, r.tmin, r.tmax, hr);
        else hit = any_hit ? ngi_trace_brute<true>(s->dev.tris2, s->n, o, d, r.tmin, r.tmax, hr) : ngi_trace_brute<false>(s->dev.tris2, s->n, o, d, r.tmin, r.tmax, hr);
        NgiHit out; out.t = hit ? hr.t : 0.0f; out.u = hit ? hr.u : 0.0f; out.v = hit ? hr.v : 0.0f;
        out.tri = hit ? (any_hit ? 0u : hr.tri) : NGI_NO_HIT;
        hits[i] = out;
    }
    return 0;
}

// traversal cost of the BVH8 on a ray batch (build-quality metric): out = {node steps, triangle tests}
__attribute__((visibility("default"))) int sim_trace_stats(void* h, const NgiRay* rays, uint64_t n, int any_hit, double* out) {
    SimScene* s = (SimScene*)h;
    g_trace_count[0] = g_trace_count[1] = 0;
    for (uint64_t i = 0; i < n; i++) {
        const NgiRay& r = rays[i];
        const f3 o = mk3(r.o[0], r.o[1], r.o[2]), d = mk3(r.d[0], r.d[1], r.d[2]);
        NgiHitRec hr;
        if (any_hit) ngi_trace_bvh8<true>(s->dev.nodes8, s->dev.tris8, o, d, r.tmin, r.tmax, hr);
        else ngi_trace_bvh8<false>(s->dev.nodes8, s->dev.tris8, o, d, r.tmin, r.tmax, hr);
    }
    out[0] = (double)g_trace_count[0]; out[1] = (double)g_trace_count[1];
    return 0;
}

// the wavefront loop, kernels run in the order k_iter_begin, k_logic, k_extend, k_shadow
// stats = {paths, extend rays, shadow rays, iterations}
__attribute__((visibility("default"))) int sim_render(void* h, const NgiRenderParams* rp, float* film, double* stats) {
    SimScene* s = (SimScene*)h;
    const size_t npx = (size_t)rp->width * rp->height;
    std::fill(film, film + npx * 3, 0.0f);
    if (stats) std::fill(stats, stats + 8, 0.0);
    if (rp->max_num_vertices != -1 && rp->max_num_vertices < 2) return 0;
    if (rp->renderer == NGI_RENDERER_BDPT) {
        NgiBdParams bp;
        bp.film = film; bp.width = rp->width; bp.height = rp->height; bp.max_verts = rp->max_num_vertices;
        bp.seed_lo = (unsigned)rp->seed; bp.seed_hi = (unsigned)(rp->seed >> 32);
        bp.film_scale = rp->film_norm_samples > 0 ? (float)((double)npx / (double)rp->film_norm_samples) : 1.0f;
        NgiBdCounters cnt; cnt.extend = 0; cnt.shadow = 0;
        std::vector<NgiBdScratch> q(1);
        if (rp->flags & NGI_RENDER_BDPT_PER_THREAD) {       // k_bdpt: one sample after the other
            std::vector<NgiBdVertex> VL(NGI_BD_MAX_VERTS), VE(NGI_BD_MAX_VERTS);
            for (long long i = 0; i < rp->num_samples; i++) ngi_bdpt_sample(s->dev, bp, (unsigned long long)(rp->sample_offset + i), VL.data(), VE.data(), q[0], cnt);
            if (stats) { stats[0] = (double)rp->num_samples; stats[1] = (double)cnt.extend; stats[2] = (double)cnt.shadow; stats[3] = 1; }
            return 0;
        }
        // the wavefront stages (ngi_bdpt_wave.h) in the order the k_bdw_* kernels run, batch by batch
        const int cap = ngi_bd_vertex_cap(bp);
        const unsigned B = rp->wave_capacity ? rp->wave_capacity : 4096;
        NgiBdWave wv;
        std::memset(&wv, 0, sizeof(wv));
        wv.walkers = 2 * B;
        std::vector<NgiBdVertex> V((size_t)std::max(cap, 1) * wv.walkers);
        std::vector<NgiBdCache> C((size_t)std::max(cap, 1) * wv.walkers);
        wv.C = C.data();
        std::vector<unsigned> nverts(wv.walkers);
        std::vector<float4> rays0((size_t)2 * wv.walkers), rays1((size_t)2 * wv.walkers), hits(wv.walkers);
        std::vector<unsigned long long> offsets(B + 1);
        std::vector<uint2> items, items_sorted;
        wv.V = V.data(); wv.nverts = nverts.data(); wv.rays[0] = rays0.data(); wv.rays[1] = rays1.data(); wv.hits = hits.data();
        wv.offsets = offsets.data();
        unsigned long long batches = 0;
        for (long long b0 = 0; b0 < rp->num_samples; b0 += B, batches++) {
            wv.first = (unsigned long long)(rp->sample_offset + b0);
            wv.batch = (unsigned)std::min<long long>(B, rp->num_samples - b0);
            unsigned count = 0;
            for (unsigned w = 0; w < 2 * wv.batch; w++) {                                      // k_bdw_start
                f3 o, wo; float rr;
                if (ngi_bdw_start(s->dev, bp, wv, w, cap, o, wo, rr)) {
                    wv.rays[1][2 * (size_t)count] = make_float4(o.x, o.y, o.z, rr); wv.rays[1][2 * (size_t)count + 1] = make_float4(wo.x, wo.y, wo.z, u2f(w)); count++;
                }
            }
            for (int step = 1; step < cap && count > 0; step++) {
                const float4* rq = wv.rays[step & 1];
                float4* nq = wv.rays[(step + 1) & 1];
                for (unsigned i = 0; i < count; i++) {                                         // k_bdw_extend
                    NgiHitRec h;
                    const bool hit = ngi_trace_bvh8<false>(s->dev.nodes8, s->dev.tris8, mk3(rq[2 * i].x, rq[2 * i].y, rq[2 * i].z), mk3(rq[2 * i + 1].x, rq[2 * i + 1].y, rq[2 * i + 1].z), NGI_EPS_F, NGI_INF_F, h);
                    hits[i] = hit ? make_float4(h.t, h.u, h.v, u2f(h.tri)) : make_float4(0.0f, 0.0f, 0.0f, u2f(NGI_MISS));
                }
                cnt.extend += count;
                unsigned next = 0;
                for (unsigned i = 0; i < count; i++) {                                         // k_bdw_step
                    unsigned w; f3 o, wo; float rr;
                    if (ngi_bdw_step(s->dev, bp, wv, step, cap, rq[2 * i], rq[2 * i + 1], hits[i], w, o, wo, rr)) {
                        nq[2 * (size_t)next] = make_float4(o.x, o.y, o.z, rr); nq[2 * (size_t)next + 1] = make_float4(wo.x, wo.y, wo.z, u2f(w)); next++;
                    }
                }
                count = next;
            }
            unsigned long long acc = 0;
            for (unsigned i = 0; i < wv.batch; i++) {                                          // k_bdw_count + scan
                unsigned nr, nl;
                ngi_bdw_strategies(s->dev, bp, wv, i, nr, nl, false, 0u, 0u);
                offsets[i] = acc; acc += (unsigned long long)nr | ((unsigned long long)nl << 32);
            }
            offsets[wv.batch] = acc;
            wv.n_ray_items = (unsigned)(acc & 0xFFFFFFFFull); wv.n_rayless = (unsigned)(acc >> 32);
            const size_t n_items = (size_t)wv.n_ray_items + wv.n_rayless;
            items.assign(n_items + 1, make_uint2(0u, 0u));
            wv.items = items.data();
            for (unsigned i = 0; i < wv.batch; i++) {                                          // k_bdw_expand
                unsigned nr, nl;
                ngi_bdw_strategies(s->dev, bp, wv, i, nr, nl, true, (unsigned)(offsets[i] & 0xFFFFFFFFull), (unsigned)(offsets[i] >> 32));
            }
            for (unsigned i = 0; i < wv.n_ray_items; i++) {                                    // k_bdw_shadow
                f3 o, d; float tmax; NgiHitRec h;
                ngi_bdw_item_ray(wv, items[i], o, d, tmax);
                if (ngi_trace_bvh8<true>(s->dev.nodes8, s->dev.tris8, o, d, NGI_EPS_F, tmax, h)) items[i].y = NGI_BDW_DEAD;
            }
            cnt.shadow += wv.n_ray_items;
            items_sorted.assign(items.begin(), items.begin() + n_items);                       // radix sort by y (stable)
            std::stable_sort(items_sorted.begin(), items_sorted.end(), [](const uint2& a, const uint2& b) { return a.y < b.y; });
            wv.items_sorted = items_sorted.data();
            for (size_t i = 0; i < n_items; i++) ngi_bdw_contrib(s->dev, bp, wv, items_sorted[i]);   // k_bdw_contrib
        }
        if (stats) { stats[0] = (double)rp->num_samples; stats[1] = (double)cnt.extend; stats[2] = (double)cnt.shadow; stats[3] = (double)batches; }
        return 0;
    }
    const unsigned P = rp->wave_capacity ? rp->wave_capacity : 4096;
    std::vector<float4> hit(P), shadow_q((size_t)P * 2 * 3);
    std::vector<NgiSlotA> sa(P); std::vector<NgiSlotB> sb(P);
    std::memset(sb.data(), 0, sizeof(NgiSlotB) * P);
    unsigned iter_counters[2] = {0, 0}, fetch_cursors[2] = {0, 0};
    std::vector<unsigned> extend_q(P);
    unsigned long long next_sample = (unsigned long long)rp->sample_offset;
    NgiWaveParams wp;
    wp.sa = sa.data(); wp.sb = sb.data();
    wp.hit = hit.data(); wp.shadow_q = shadow_q.data(); wp.iter_counters = iter_counters;
    wp.extend_q = extend_q.data(); wp.fetch_cursors = fetch_cursors;
    wp.next_sample = &next_sample; wp.film = film; wp.capacity = P; wp.renderer = rp->renderer; wp.max_verts = rp->max_num_vertices;
    wp.width = rp->width; wp.height = rp->height; wp.sample_end = (unsigned long long)(rp->sample_offset + rp->num_samples);
    wp.seed_lo = (unsigned)rp->seed; wp.seed_hi = (unsigned)(rp->seed >> 32);
    wp.film_scale = rp->film_norm_samples > 0 ? (float)((double)npx / (double)rp->film_norm_samples) : 1.0f;
    unsigned long long extend = 0, shadow = 0, iters = 0;
    unsigned long long cost[4] = {0, 0, 0, 0};   // BVH8 node steps / triangle tests of the extend rays, then of the shadow rays (stats[4..7])
    while (true) {
        iter_counters[0] = iter_counters[1] = 0;
        if (rp->renderer >= NGI_RENDERER_LT || s->dev.sensor.kind == NGI_ET_AREA) for (unsigned i = 0; i < P; i++) ngi_logic_step<true>(s->dev, wp, i);
        else for (unsigned i = 0; i < P; i++) ngi_logic_step<false>(s->dev, wp, i);
        extend += iter_counters[1]; shadow += iter_counters[0];
        iters++;
        if (iter_counters[1] == 0 && iter_counters[0] == 0 && next_sample >= wp.sample_end) break;
        g_trace_count[0] = g_trace_count[1] = 0;
        for (unsigned q = 0; q < iter_counters[1]; q++) ngi_extend_step(s->dev, wp, extend_q[q]);   // compacted extend queue
        cost[0] += g_trace_count[0]; cost[1] += g_trace_count[1];
        g_trace_count[0] = g_trace_count[1] = 0;
        for (unsigned e = 0; e < iter_counters[0]; e++) ngi_shadow_step(s->dev, wp, e);
        cost[2] += g_trace_count[0]; cost[3] += g_trace_count[1];
    }
    if (stats) {
        stats[0] = (double)rp->num_samples; stats[1] = (double)extend; stats[2] = (double)shadow; stats[3] = (double)iters;
        for (int k = 0; k < 4; k++) stats[4 + k] = (double)cost[k];
    }
    return 0;
}

// same contract as ngi_gpu_eval_bsdf
__attribute__((visibility("default"))) int sim_eval_bsdf(void* h, const float* q, const float* wo_in, uint64_t n, int force_degenerated, float* out) {
    SimScene* s = (SimScene*)h;
    for (uint64_t i = 0; i < n; i++) {
        const float* a = q + 16 * i;
        const NgiDevPrim& P = s->dev.prims[(int)a[0]];
        const int type = (int)a[1];
        NgiGeom g; g.sn = mk3(a[2], a[3], a[4]); g.gn = mk3(a[5], a[6], a[7]);
        ngi_tangent_space(g);
        g.albedo = ngi_constant_albedo(P, type);
        const f3 wi = mk3(a[8], a[9], a[10]);
        f3 wo = mk3(0.0f); bool valid = true;
        if (a[14] != 0.0f) wo = mk3(wo_in[3 * i], wo_in[3 * i + 1], wo_in[3 * i + 2]);
        else valid = ngi_sample_bsdf(P, type, g, wi, a[11], a[12], a[13], wo);
        float pdf = 0; f3 fs = mk3(0.0f);
        if (valid) fs = ngi_eval_bsdf(P, type, g, wi, wo, force_degenerated != 0, pdf);
        float* o = out + 8 * i;
        o[0] = wo.x; o[1] = wo.y; o[2] = wo.z; o[3] = fs.x; o[4] = fs.y; o[5] = fs.z; o[6] = pdf; o[7] = valid ? 1.0f : 0.0f;
    }
    return 0;
}

__attribute__((visibility("default"))) void sim_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

}  // extern "C"

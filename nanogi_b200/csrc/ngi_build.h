// ngi_build.h — per-item bodies of the GPU BVH build kernels.
//
// Replaces Embree's scene commit (reference include/nanogi/rt.hpp:2085-2143: rtcNewScene / rtcNewTriangleMesh /
// rtcCommit). Pipeline, all on the device:
//   1. triangle records + padded AABBs + scene bounds            (k_tri_setup)
//   2. 63-bit Morton codes of the AABB centres                    (k_morton)        -> cub radix sort
//   3. binary topology over the Morton order by PLOC — parallel locally-ordered clustering (Meister & Bittner
//      2017): every round each cluster looks NGI_PLOC_RADIUS neighbours left and right for the partner that
//      minimises the surface area of the union, mutual nearest neighbours merge, the cluster array is compacted
//      (k_ploc_nearest, k_ploc_flag, cub scan, k_ploc_merge). Node boxes come out of the merges (no refit).
//      An LBVH (Karras 2012, ngi_karras_node) was the first version; its traversal cost on the 1M-triangle scene
//      was ~3.3x the Cornell box's per ray, mostly from the room's large triangles inflating Morton-adjacent nodes.
//   5. 64-byte two-child traversal nodes (BVH2, cross-check)      (k_pack2)
//   6. top-down collapse to compressed 8-wide nodes, level by level with atomically allocated child /
//      triangle ranges (k_collapse8); octant slot assignment + outward-rounded 8-bit quantisation.
#pragma once
#include "ngi_math.h"

struct NgiBuildTask { int node2; unsigned node8; };

// ---- 1. triangle record + padded box --------------------------------------------------------------
// record = (v0.xyz | id) (e1.xyz) (e2.xyz) with e1 = v1 - v0, e2 = v2 - v0 as single IEEE subtractions (the
// oracle precomputes the same). Slots >= n_real are degenerate filler triangles (never hit: det == 0).
NGI_HD void ngi_tri_setup(const float* __restrict__ positions, unsigned i, unsigned n_real, float pad, f3 anchor,
                          float4* __restrict__ rec, float4* __restrict__ lo, float4* __restrict__ hi) {
    if (i < n_real) {
        const float* p = positions + 9 * (size_t)i;
        const f3 v0 = mk3(p[0], p[1], p[2]), v1 = mk3(p[3], p[4], p[5]), v2 = mk3(p[6], p[7], p[8]);
        rec[3 * (size_t)i + 0] = make_float4(v0.x, v0.y, v0.z, u2f(i));
        rec[3 * (size_t)i + 1] = make_float4(NGI_SUB(v1.x, v0.x), NGI_SUB(v1.y, v0.y), NGI_SUB(v1.z, v0.z), 0.0f);
        rec[3 * (size_t)i + 2] = make_float4(NGI_SUB(v2.x, v0.x), NGI_SUB(v2.y, v0.y), NGI_SUB(v2.z, v0.z), 0.0f);
        lo[i] = make_float4(fminf(v0.x, fminf(v1.x, v2.x)) - pad, fminf(v0.y, fminf(v1.y, v2.y)) - pad, fminf(v0.z, fminf(v1.z, v2.z)) - pad, 0.0f);
        hi[i] = make_float4(fmaxf(v0.x, fmaxf(v1.x, v2.x)) + pad, fmaxf(v0.y, fmaxf(v1.y, v2.y)) + pad, fmaxf(v0.z, fmaxf(v1.z, v2.z)) + pad, 0.0f);
    } else {
        rec[3 * (size_t)i + 0] = make_float4(anchor.x, anchor.y, anchor.z, u2f(0xFFFFFFFFu));
        rec[3 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        rec[3 * (size_t)i + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        lo[i] = make_float4(anchor.x - pad, anchor.y - pad, anchor.z - pad, 0.0f);
        hi[i] = make_float4(anchor.x + pad, anchor.y + pad, anchor.z + pad, 0.0f);
    }
}

// ---- 2. Morton -----------------------------------------------------------------------------------
NGI_HD unsigned long long ngi_expand21(unsigned long long v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
NGI_HD unsigned long long ngi_morton63(f3 c, f3 smin, f3 sinv) {
    const float fx = fminf(fmaxf((c.x - smin.x) * sinv.x, 0.0f), 1.0f);
    const float fy = fminf(fmaxf((c.y - smin.y) * sinv.y, 0.0f), 1.0f);
    const float fz = fminf(fmaxf((c.z - smin.z) * sinv.z, 0.0f), 1.0f);
    const unsigned long long x = (unsigned long long)(fx * 2097151.0f);
    const unsigned long long y = (unsigned long long)(fy * 2097151.0f);
    const unsigned long long z = (unsigned long long)(fz * 2097151.0f);
    return (ngi_expand21(x) << 2) | (ngi_expand21(y) << 1) | ngi_expand21(z);
}

// ---- 3. Karras topology --------------------------------------------------------------------------
// Nodes are numbered 0..n-2 (inner) and n-1..2n-2 (leaf k = n-1+k, k = position in Morton order).
NGI_HD int ngi_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + ngi_clz32((unsigned)i ^ (unsigned)j);
    return ngi_clz64(a ^ b);
}
NGI_HD void ngi_karras_node(const unsigned long long* __restrict__ keys, int n, int i, int& left, int& right, int& first, int& last) {
    const int d = (ngi_delta(keys, n, i, i + 1) - ngi_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = ngi_delta(keys, n, i, i - d);
    int lmax = 2;
    while (ngi_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (ngi_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = ngi_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (ngi_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    first = lo; last = hi;
}

// ---- 3b. PLOC ------------------------------------------------------------------------------------
// Node numbering: inner nodes 0..n-2 with the ROOT = 0 (ids are handed out downwards from n-2, the last merge
// gets 0), leaf k = n-1+k (k = position in Morton order). cnt[] = triangles below an inner node.
#define NGI_PLOC_RADIUS 16

NGI_HD float ngi_union_half_area(const float4 a0, const float4 a1, const float4 b0, const float4 b1) {
    const float dx = fmaxf(a1.x, b1.x) - fminf(a0.x, b0.x), dy = fmaxf(a1.y, b1.y) - fminf(a0.y, b0.y), dz = fmaxf(a1.z, b1.z) - fminf(a0.z, b0.z);
    return dx * dy + dy * dz + dz * dx;
}
// nearest neighbour of cluster i inside the window; ties -> lowest index (guarantees a mutual pair every round)
NGI_HD int ngi_ploc_nearest(const float4* __restrict__ clo, const float4* __restrict__ chi, const int C, const int i) {
    const float4 a0 = clo[i], a1 = chi[i];
    int best = -1;
    float bestA = NGI_INF_F;
    const int j0 = i - NGI_PLOC_RADIUS < 0 ? 0 : i - NGI_PLOC_RADIUS;
    const int j1 = i + NGI_PLOC_RADIUS > C - 1 ? C - 1 : i + NGI_PLOC_RADIUS;
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const float A = ngi_union_half_area(a0, a1, clo[j], chi[j]);
        if (A < bestA || best < 0) { bestA = A; best = j; }
    }
    return best;
}
// 1 = cluster i survives the round (alone or as the merged pair), 0 = it is absorbed by its lower-index partner
NGI_HD unsigned ngi_ploc_keep(const int* __restrict__ nn, const int i) {
    const int j = nn[i];
    return (j >= 0 && nn[j] == i && j < i) ? 0u : 1u;
}
// ---- SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1) -------------------------------------------------------
// C(n, i) = cheapest way to turn the binary subtree of n into a forest of at most i wide-BVH subtrees:
//     C(n, 1) = min( leaf:      A_n * P_n * c_prim          (P_n <= NGI_LEAF_MAX_TRIS triangles in one leaf slot),
//                    internal:  A_n * c_node + D(n, 8) )     (n becomes an 8-wide node)
//     C(n, i) = min( D(n, i), C(n, i - 1) )                  i = 2 .. 7
//     D(n, j) = min over 0 < k < j of  C(left, k) + C(right, j - k)
// One row per binary inner node, filled bottom-up: PLOC creates a node after both of its children (ngi_ploc_merge), so the row is
// computed right where the node is made. The collapse then follows the recorded decisions (ngi_collapse_node) instead of opening
// the child with the largest area until the node is full. A node step costs the same whether 3 or 8 slots are occupied, so the
// model's constant per-node cost is exact for this traversal; the greedy collapse left the nodes of the 1 M-triangle scene 3.9 / 8
// full on average (tools/bvh_quality.py).
#define NGI_LEAF_MAX_TRIS 3
#define NGI_DP_INF 3.0e38f
#ifndef NGI_SAH_C_PRIM
#define NGI_SAH_C_PRIM 0.5f        /* cost of a triangle test relative to an 8-wide node step (C3 k_extend: 1.385 / 1.379 / 1.366 ms at 0.2 / 0.3 / 0.5, s41) */
#endif
struct alignas(16) NgiDpRow {
    float c[7];        // c[i - 1] = C(n, i)
    unsigned dec;      // bit 0: C(n, 1) is the leaf; bits 1 + 3 (j - 2) .. +2, j = 2 .. 8: the split k of D(n, j), or 0 = "C(n, j) is C(n, j - 1)"
};
NGI_HD unsigned ngi_dp_split(const unsigned dec, const int j) { return (dec >> (1 + 3 * (j - 2))) & 7u; }
// m: binary node id; leaves (m >= n - 1) cost A * c_prim whatever the budget
NGI_HD float ngi_dp_get(const NgiDpRow* __restrict__ dp, const float4* __restrict__ lo, const float4* __restrict__ hi, const int n, const int m,
                        const int i, const float c_prim) {
    if (m >= n - 1) {
        const float4 a = lo[m], b = hi[m];
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        return (dx * dy + dy * dz + dz * dx) * c_prim;
    }
    return dp[m].c[i - 1];
}
NGI_HD void ngi_dp_node(NgiDpRow* __restrict__ dp, const float4* __restrict__ lo, const float4* __restrict__ hi, const int n, const int id,
                        const int a, const int b, const unsigned count, const float area, const float c_node, const float c_prim) {
    float ca[7], cb[7];
#pragma unroll
    for (int i = 1; i <= 7; i++) { ca[i - 1] = ngi_dp_get(dp, lo, hi, n, a, i, c_prim); cb[i - 1] = ngi_dp_get(dp, lo, hi, n, b, i, c_prim); }
    float D[9]; unsigned K[9];
#pragma unroll
    for (int j = 2; j <= 8; j++) {
        float best = NGI_DP_INF; unsigned bk = 1;
#pragma unroll
        for (int k = 1; k < j; k++) {
            if (k > 7 || j - k > 7) continue;
            const float v = ca[k - 1] + cb[j - k - 1];
            if (v < best) { best = v; bk = (unsigned)k; }
        }
        D[j] = best; K[j] = bk;
    }
    NgiDpRow r;
    const float leaf = count <= NGI_LEAF_MAX_TRIS ? area * (float)count * c_prim : NGI_DP_INF;
    const float inner = area * c_node + D[8];
    r.dec = leaf <= inner ? 1u : 0u;
    r.c[0] = leaf <= inner ? leaf : inner;
#pragma unroll
    for (int i = 2; i <= 7; i++) {
        if (D[i] < r.c[i - 2]) { r.c[i - 1] = D[i]; r.dec |= K[i] << (1 + 3 * (i - 2)); }
        else r.c[i - 1] = r.c[i - 2];
    }
    r.dec |= K[8] << (1 + 3 * 6);
    dp[id] = r;
}

struct NgiPlocCtx {
    const int* nn; const unsigned* pos;          // nearest neighbour, exclusive scan of the keep flags
    const int* cid_in; const float4* clo_in; const float4* chi_in;
    int* cid_out; float4* clo_out; float4* chi_out;
    float4* lo; float4* hi;                      // [2n-1] node boxes
    int* left; int* right; unsigned* cnt;        // [n-1]
    unsigned* depth;                             // [n-1] height of the binary subtree (the BVH2 cross-check traversal has a 64-entry stack)
    int n;                                       // leaves
    int next_id;                                 // id of the first merge of this round (ids go downwards)
    NgiDpRow* dp;                                // [n-1] collapse decisions (ngi_dp_node), or NULL
    float c_node, c_prim;
};
NGI_HD void ngi_ploc_merge(const NgiPlocCtx& c, const int i) {
    const int j = c.nn[i];
    const bool mutual = j >= 0 && c.nn[j] == i;
    if (mutual && j < i) return;                 // absorbed
    const unsigned p = c.pos[i];
    if (mutual) {
        const int rank = j - (int)c.pos[j];      // absorbed clusters in front of j = merges in front of this one
        const int id = c.next_id - rank;
        const int a = c.cid_in[i], b = c.cid_in[j];
        const float4 a0 = c.clo_in[i], a1 = c.chi_in[i], b0 = c.clo_in[j], b1 = c.chi_in[j];
        const float4 u0 = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.0f);
        const float4 u1 = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.0f);
        c.left[id] = a; c.right[id] = b;
        c.cnt[id] = (a >= c.n - 1 ? 1u : c.cnt[a]) + (b >= c.n - 1 ? 1u : c.cnt[b]);
        if (c.depth) {
            const unsigned da = a >= c.n - 1 ? 0u : c.depth[a], db = b >= c.n - 1 ? 0u : c.depth[b];
            c.depth[id] = 1u + (da > db ? da : db);
        }
        c.lo[id] = u0; c.hi[id] = u1;
        c.cid_out[p] = id; c.clo_out[p] = u0; c.chi_out[p] = u1;
        if (c.dp) {
            const float dx = u1.x - u0.x, dy = u1.y - u0.y, dz = u1.z - u0.z;
            ngi_dp_node(c.dp, c.lo, c.hi, c.n, id, a, b, c.cnt[id], dx * dy + dy * dz + dz * dx, c.c_node, c.c_prim);
        }
    } else {
        c.cid_out[p] = c.cid_in[i]; c.clo_out[p] = c.clo_in[i]; c.chi_out[p] = c.chi_in[i];
    }
}

// ---- 5. BVH2 traversal node ----------------------------------------------------------------------
NGI_HD void ngi_pack2(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ left,
                      const int* __restrict__ right, int n, int i, float4* __restrict__ out) {
    const int l = left[i], r = right[i];
    const float4 l0 = lo[l], l1 = hi[l], r0 = lo[r], r1 = hi[r];
    const int lc = l >= n - 1 ? ~(l - (n - 1)) : l;
    const int rc = r >= n - 1 ? ~(r - (n - 1)) : r;
    out[4 * (size_t)i + 0] = make_float4(l0.x, l0.y, l0.z, l1.x);
    out[4 * (size_t)i + 1] = make_float4(l1.y, l1.z, r0.x, r0.y);
    out[4 * (size_t)i + 2] = make_float4(r0.z, r1.x, r1.y, r1.z);
    out[4 * (size_t)i + 3] = make_float4(u2f((unsigned)lc), u2f((unsigned)rc), 0.0f, 0.0f);
}

// ---- 6. collapse to BVH8 -------------------------------------------------------------------------

struct NgiCollapseCtx {
    // BVH2 (node numbering as above)
    const float4* lo; const float4* hi;      // [2n-1] padded boxes
    const int* left; const int* right;       // [n-1]
    const unsigned* cnt;                     // [n-1] triangles below an inner node
    const float4* tris2;                     // [n][3] Morton-ordered triangle records
    int n;
    // BVH8 output
    uint4* nodes8;                           // [<= n][5]
    float4* tris8;                           // [n][3]
    unsigned* counters;                      // [0] = nodes allocated, [1] = triangles allocated, [2] = next-level task count
    NgiBuildTask* out_tasks;
    const NgiDpRow* dp;                      // SAH-optimal collapse decisions, or NULL = greedy (open the largest child until 8)
};

NGI_HD unsigned ngi_atomic_add(unsigned* p, unsigned v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const unsigned o = *p; *p = o + v; return o;
#endif
}

NGI_HD unsigned ngi_pack4(const unsigned* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); }

NGI_HD float ngi_half_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

NGI_HD_NOINLINE void ngi_collapse_node(const NgiCollapseCtx& c, const NgiBuildTask task) {
    const int n = c.n;
    int child[8];
    bool dp_leaf[8];                          // DP: this child becomes a leaf slot (else: every child with <= NGI_LEAF_MAX_TRIS triangles does)
    int cnt = 0;
    if (c.dp) {
        // follow the recorded decisions: the 8 child slots of this node are distributed over the two binary subtrees
        int st_m[16], st_j[16]; int sp = 0;
        const unsigned k8 = ngi_dp_split(c.dp[task.node2].dec, 8);
        st_m[sp] = c.right[task.node2]; st_j[sp++] = 8 - (int)k8;
        st_m[sp] = c.left[task.node2]; st_j[sp++] = (int)k8;
        while (sp > 0) {
            const int m = st_m[--sp]; int j = st_j[sp];
            if (m >= n - 1) { dp_leaf[cnt] = true; child[cnt++] = m; continue; }
            const unsigned dec = c.dp[m].dec;
            unsigned k = 0;
            while (j > 1 && (k = ngi_dp_split(dec, j)) == 0u) j--;             // C(m, j) = C(m, j - 1)
            if (j == 1) { dp_leaf[cnt] = (dec & 1u) != 0u; child[cnt++] = m; continue; }
            st_m[sp] = c.right[m]; st_j[sp++] = j - (int)k;
            st_m[sp] = c.left[m]; st_j[sp++] = (int)k;
        }
    } else {
    cnt = 2;
    child[0] = c.left[task.node2];
    child[1] = c.right[task.node2];
    // greedily open the inner child with the largest surface area until 8 children
    while (cnt < 8) {
        int best = -1; float bestA = -1.0f;
        for (int k = 0; k < cnt; k++) {
            const int nd = child[k];
            if (nd >= n - 1) continue;  // single-triangle leaf
            if (c.cnt[nd] <= NGI_LEAF_MAX_TRIS) continue;  // small subtree stays one leaf slot
            const float a = ngi_half_area(c.lo[nd], c.hi[nd]);
            if (a > bestA) { bestA = a; best = k; }
        }
        if (best < 0) break;
        const int nd = child[best];
        child[best] = c.left[nd];
        child[cnt++] = c.right[nd];
    }
    for (int k = 0; k < cnt; k++) dp_leaf[k] = child[k] >= n - 1 || c.cnt[child[k]] <= NGI_LEAF_MAX_TRIS;
    }
    // node frame: origin one quantisation step below the node box and 252 steps of usable range, so that the
    // outward-rounded planes (widened by 2^-7 step, see below) never need clamping on either side
    const float4 nlo = c.lo[task.node2], nhi = c.hi[task.node2];
    const float blo[3] = {nlo.x, nlo.y, nlo.z}, phi[3] = {nhi.x, nhi.y, nhi.z};
    float plo[3];
    int eb[3]; double scale[3];
    for (int a = 0; a < 3; a++) {
        const double ext = (double)phi[a] - (double)blo[a];
        int e = -126;
        if (ext > 0) {
            int k; (void)frexp(ext / 252.0, &k);  // ext/252 = m * 2^k, m in [0.5, 1)  =>  2^k * 252 >= ext
            e = k;
        }
        if (e < -126) e = -126;
        if (e > 112) e = 112;   // the traversal forms 2^(e+15) as a float exponent
        while (e < 112 && ext > 252.0 * ldexp(1.0, e)) e++;
        eb[a] = e + 127;
        scale[a] = ldexp(1.0, e);
        float o = (float)((double)blo[a] - scale[a]);
        if ((double)o > (double)blo[a] - scale[a]) o = nextafterf(o, -NGI_INF_F);   // round the origin downwards
        plo[a] = o;
    }
    // classify children
    bool inner[8]; int ntri[8]; int leafTri[8][NGI_LEAF_MAX_TRIS];
    unsigned nInner = 0, nTris = 0;
    float cx[8], cy[8], cz[8];
    for (int k = 0; k < cnt; k++) {
        const int nd = child[k];
        if (nd >= n - 1) { inner[k] = false; ntri[k] = 1; leafTri[k][0] = nd - (n - 1); }
        else if (dp_leaf[k]) {
            // gather the (at most NGI_LEAF_MAX_TRIS) leaves of the small subtree, left first
            inner[k] = false; ntri[k] = 0;
            int st[NGI_LEAF_MAX_TRIS + 1]; int sp = 0;
            st[sp++] = nd;
            while (sp > 0) {
                const int x = st[--sp];
                if (x >= n - 1) leafTri[k][ntri[k]++] = x - (n - 1);
                else { st[sp++] = c.right[x]; st[sp++] = c.left[x]; }
            }
        } else { inner[k] = true; ntri[k] = 0; }
        if (inner[k]) nInner++; else nTris += (unsigned)ntri[k];
        const float4 l = c.lo[nd], h = c.hi[nd];
        cx[k] = 0.5f * (l.x + h.x) - 0.5f * (nlo.x + nhi.x);
        cy[k] = 0.5f * (l.y + h.y) - 0.5f * (nlo.y + nhi.y);
        cz[k] = 0.5f * (l.z + h.z) - 0.5f * (nlo.z + nhi.z);
    }
    const unsigned childBase = nInner ? ngi_atomic_add(&c.counters[0], nInner) : 0u;
    const unsigned triBase = nTris ? ngi_atomic_add(&c.counters[1], nTris) : 0u;
    const unsigned taskBase = nInner ? ngi_atomic_add(&c.counters[2], nInner) : 0u;
    // greedy octant slot assignment: slot bit a set <=> child lies towards +axis a
    int slotOf[8]; bool slotUsed[8]; bool done[8];
    for (int k = 0; k < 8; k++) { slotUsed[k] = false; done[k] = false; slotOf[k] = -1; }
    for (int it = 0; it < cnt; it++) {
        float bestC = -3.0e38f; int bk = -1, bs = -1;
        for (int k = 0; k < cnt; k++) {
            if (done[k]) continue;
            for (int s = 0; s < 8; s++) {
                if (slotUsed[s]) continue;
                const float cost = ((s & 1) ? cx[k] : -cx[k]) + ((s & 2) ? cy[k] : -cy[k]) + ((s & 4) ? cz[k] : -cz[k]);
                if (cost > bestC) { bestC = cost; bk = k; bs = s; }
            }
        }
        done[bk] = true; slotUsed[bs] = true; slotOf[bk] = bs;
    }
    int childAt[8];
    for (int s = 0; s < 8; s++) childAt[s] = -1;
    for (int k = 0; k < cnt; k++) childAt[slotOf[k]] = k;

    unsigned meta[8], q[6][8];
    unsigned imask = 0, innerRank = 0, triOff = 0;
    for (int s = 0; s < 8; s++) {
        meta[s] = 0;
        for (int a = 0; a < 6; a++) q[a][s] = 0;
        const int k = childAt[s];
        if (k < 0) continue;
        const int nd = child[k];
        const float4 l = c.lo[nd], h = c.hi[nd];
        const float clo[3] = {l.x, l.y, l.z}, chi[3] = {h.x, h.y, h.z};
        for (int a = 0; a < 3; a++) {
            // outward rounding, widened by 2^-7 step: covers the <= 2^-9 step decode error of ngi_q1 (ngi_bvh.h)
            double ql = floor(((double)clo[a] - (double)plo[a]) / scale[a] - 0.0078125);
            double qh = ceil(((double)chi[a] - (double)plo[a]) / scale[a] + 0.0078125);
            if (ql < 0) ql = 0;
            if (ql > 255) ql = 255;
            if (qh < 0) qh = 0;
            if (qh > 255) qh = 255;
            q[a][s] = (unsigned)ql; q[3 + a][s] = (unsigned)qh;
        }
        if (inner[k]) {
            imask |= 1u << s;
            meta[s] = (1u << 5) | (24u + (unsigned)s);
            NgiBuildTask t; t.node2 = nd; t.node8 = childBase + innerRank;
            c.out_tasks[taskBase + innerRank] = t;
            innerRank++;
        } else {
            meta[s] = (((1u << ntri[k]) - 1u) << 5) | triOff;
            for (int j = 0; j < ntri[k]; j++) {
                const size_t src = (size_t)leafTri[k][j] * 3, dst = (size_t)(triBase + triOff + j) * 3;
                c.tris8[dst] = c.tris2[src]; c.tris8[dst + 1] = c.tris2[src + 1]; c.tris8[dst + 2] = c.tris2[src + 2];
            }
            triOff += (unsigned)ntri[k];
        }
    }
    uint4* out = c.nodes8 + 5 * (size_t)task.node8;
    out[0] = make_uint4(f2u(plo[0]), f2u(plo[1]), f2u(plo[2]), (unsigned)eb[0] | ((unsigned)eb[1] << 8) | ((unsigned)eb[2] << 16) | (imask << 24));
    out[1] = make_uint4(childBase, triBase, ngi_pack4(meta), ngi_pack4(meta + 4));
    out[2] = make_uint4(ngi_pack4(q[0]), ngi_pack4(q[0] + 4), ngi_pack4(q[1]), ngi_pack4(q[1] + 4));
    out[3] = make_uint4(ngi_pack4(q[2]), ngi_pack4(q[2] + 4), ngi_pack4(q[3]), ngi_pack4(q[3] + 4));
    out[4] = make_uint4(ngi_pack4(q[4]), ngi_pack4(q[4] + 4), ngi_pack4(q[5]), ngi_pack4(q[5] + 4));
}

// ---- 7. traversal form of a node --------------------------------------------------------------------
// The collapse writes n1 = (child_base, tri_base, meta[0..3], meta[4..7]) with the leaf triangles packed from tri_base on
// (meta = unary count << 5 | offset). The traversal wants n1 = (child_base, -, valid24, -) and triangle j of slot s at the fixed
// place 24 * node + 3 * s + j (ngi_bvh.h). One item = one node; `out_nodes` may be `in_nodes`.
NGI_HD void ngi_expand_node(const uint4* __restrict__ in_nodes, const float4* __restrict__ tris_compact, const unsigned ni,
                            uint4* out_nodes, float4* __restrict__ tris_fixed) {
    const uint4 n0 = in_nodes[5 * (size_t)ni], n1 = in_nodes[5 * (size_t)ni + 1];
    const uint4 n2 = in_nodes[5 * (size_t)ni + 2], n3 = in_nodes[5 * (size_t)ni + 3], n4 = in_nodes[5 * (size_t)ni + 4];
    const unsigned imask = n0.w >> 24;
    unsigned valid = 0;
    for (int s = 0; s < 8; s++) {
        const unsigned meta = ((s < 4 ? n1.z : n1.w) >> (8 * (s & 3))) & 0xFFu;
        if (meta == 0u || ((imask >> s) & 1u)) continue;
        const int cnt = ngi_popc(meta >> 5);
        const unsigned off = meta & 31u;
        for (int j = 0; j < cnt; j++) {
            const size_t src = ((size_t)n1.y + off + (unsigned)j) * 3, dst = ((size_t)24 * ni + 3 * (unsigned)s + (unsigned)j) * 3;
            tris_fixed[dst] = tris_compact[src]; tris_fixed[dst + 1] = tris_compact[src + 1]; tris_fixed[dst + 2] = tris_compact[src + 2];
        }
        valid |= ((1u << cnt) - 1u) << (3 * s);
    }
    out_nodes[5 * (size_t)ni] = n0;
    out_nodes[5 * (size_t)ni + 1] = make_uint4(n1.x, 0u, valid, 0u);
    out_nodes[5 * (size_t)ni + 2] = n2; out_nodes[5 * (size_t)ni + 3] = n3; out_nodes[5 * (size_t)ni + 4] = n4;
}

// ngi_bvh.h — ray/triangle test and the two traversal kernels' per-ray bodies.
//
// Replaces Embree's rtcIntersect as called from Scene::Intersect / Scene::Visible
// (reference include/nanogi/rt.hpp:2162-2261, rtcIntersect at :2182). Geometry contract:
//   * float32 Moeller-Trumbore with ONE explicit expression tree (NGI_FMA/NGI_MUL/... are single IEEE
//     roundings; the CPU oracle evaluates the identical tree) -> bit-identical (t, u, v);
//   * no back-face culling (rt.hpp:2174-2178), accept iff tmin < t < tmax (strict);
//   * closest hit = lexicographic minimum of (t, global triangle id): order-free, so any conservative
//     acceleration structure returns the same answer; node culling is inclusive (t_box <= t_best);
//   * boxes are padded at build time, the BVH is a pure filter.
//
// Two structures are traversed:
//   BVH8  (product)  compressed 8-wide nodes, 80 B = 5 x 16-byte loads, quantised child boxes
//                    (layout after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs
//                    Through Compressed Wide BVHs", HPG 2017), 48-byte triangles = 3 x 16-byte loads;
//   BVH2  (cross-check) the GPU-built LBVH the BVH8 is collapsed from, 64-byte two-child nodes.
#pragma once
#include "ngi_math.h"

struct NgiHitRec { float t, u, v; unsigned tri; };
#define NGI_MISS 0xFFFFFFFFu

// triangle record: a = (v0.xyz, id bits), b = (e1.xyz, -), c = (e2.xyz, -) with e1 = v1 - v0, e2 = v2 - v0
NGI_HD bool ngi_tri_test(const float4 a, const float4 b, const float4 c, const f3 o, const f3 d, const float tmin, const float tmax,
                         float& t, float& u, float& v) {
    // P = d x e2
    const float px = NGI_FMA(d.y, c.z, -NGI_MUL(d.z, c.y));
    const float py = NGI_FMA(d.z, c.x, -NGI_MUL(d.x, c.z));
    const float pz = NGI_FMA(d.x, c.y, -NGI_MUL(d.y, c.x));
    const float det = NGI_FMA(b.x, px, NGI_FMA(b.y, py, NGI_MUL(b.z, pz)));
    if (!(det != 0.0f)) return false;
    const float inv = NGI_RCP(det);
    const float tx = NGI_SUB(o.x, a.x), ty = NGI_SUB(o.y, a.y), tz = NGI_SUB(o.z, a.z);
    const float uu = NGI_MUL(NGI_FMA(tx, px, NGI_FMA(ty, py, NGI_MUL(tz, pz))), inv);
    if (!(uu >= 0.0f && uu <= 1.0f)) return false;
    // Q = T x e1
    const float qx = NGI_FMA(ty, b.z, -NGI_MUL(tz, b.y));
    const float qy = NGI_FMA(tz, b.x, -NGI_MUL(tx, b.z));
    const float qz = NGI_FMA(tx, b.y, -NGI_MUL(ty, b.x));
    const float vv = NGI_MUL(NGI_FMA(d.x, qx, NGI_FMA(d.y, qy, NGI_MUL(d.z, qz))), inv);
    if (!(vv >= 0.0f && NGI_ADD(uu, vv) <= 1.0f)) return false;
    const float tt = NGI_MUL(NGI_FMA(c.x, qx, NGI_FMA(c.y, qy, NGI_MUL(c.z, qz))), inv);
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; u = uu; v = vv;
    return true;
}

// explicit closest-hit reduction: (t, id) lexicographic
NGI_HD void ngi_accept(NgiHitRec& best, float t, float u, float v, unsigned id) {
    if (t < best.t || (t == best.t && id < best.tri)) { best.t = t; best.u = u; best.v = v; best.tri = id; }
}

NGI_HD float ngi_safe_rcp_dir(float d) {
    // a zero / denormal component becomes +-2^-80 (keeps 0 * inf NaNs out of the slab test)
    const float eps = 8.27180613e-25f;
    const float dd = fabsf(d) > eps ? d : copysignf(eps, d);
#if defined(__CUDA_ARCH__)
    return __frcp_rn(dd);
#else
    return 1.0f / dd;
#endif
}

// ------------------------------------------------------------------------------------------------
// BVH2 (LBVH) — 64-byte nodes: both child boxes + child links. child >= 0: inner node, child < 0: leaf
// holding the single triangle ~child (index into the Morton-sorted triangle array).
//   n0 = (c0.min.xyz, c0.max.x)  n1 = (c0.max.yz, c1.min.xy)  n2 = (c1.min.z, c1.max.xyz)  n3 = (left, right, -, -)
// ------------------------------------------------------------------------------------------------
template <bool ANY_HIT>
NGI_HD bool ngi_trace_bvh2(const float4* __restrict__ nodes, const float4* __restrict__ tris, const f3 o, const f3 d,
                           const float tmin, const float tmax, NgiHitRec& out) {
    const float ix = ngi_safe_rcp_dir(d.x), iy = ngi_safe_rcp_dir(d.y), iz = ngi_safe_rcp_dir(d.z);
    NgiHitRec best; best.t = tmax; best.u = 0; best.v = 0; best.tri = NGI_MISS;
    bool found = false;
    int stack[64];
    int sp = 0;
    int cur = 0;
    while (true) {
        if (cur >= 0) {
            const float4 n0 = ngi_ldg(nodes + 4 * (size_t)cur), n1 = ngi_ldg(nodes + 4 * (size_t)cur + 1);
            const float4 n2 = ngi_ldg(nodes + 4 * (size_t)cur + 2), n3 = ngi_ldg(nodes + 4 * (size_t)cur + 3);
            const float limit = found ? best.t : tmax;
            float t0, t1, a, b;
            // child 0
            a = (n0.x - o.x) * ix; b = (n0.w - o.x) * ix; t0 = fminf(a, b); t1 = fmaxf(a, b);
            a = (n0.y - o.y) * iy; b = (n1.x - o.y) * iy; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
            a = (n0.z - o.z) * iz; b = (n1.y - o.z) * iz; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
            const float c0n = fmaxf(t0, tmin); const bool h0 = c0n <= fminf(t1, limit);
            // child 1
            a = (n1.z - o.x) * ix; b = (n2.y - o.x) * ix; t0 = fminf(a, b); t1 = fmaxf(a, b);
            a = (n1.w - o.y) * iy; b = (n2.z - o.y) * iy; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
            a = (n2.x - o.z) * iz; b = (n2.w - o.z) * iz; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
            const float c1n = fmaxf(t0, tmin); const bool h1 = c1n <= fminf(t1, limit);
            const int left = (int)f2u(n3.x), right = (int)f2u(n3.y);
            if (h0 && h1) {
                const bool swp = c1n < c0n;
                stack[sp++] = swp ? left : right;
                cur = swp ? right : left;
                continue;
            }
            if (h0) { cur = left; continue; }
            if (h1) { cur = right; continue; }
        } else {
            const size_t ti = (size_t)(~cur);
            const float4 a = ngi_ldg(tris + 3 * ti), b = ngi_ldg(tris + 3 * ti + 1), c = ngi_ldg(tris + 3 * ti + 2);
            float t, u, v;
            if (ngi_tri_test(a, b, c, o, d, tmin, tmax, t, u, v)) {
                if (ANY_HIT) { out.t = t; out.u = u; out.v = v; out.tri = 0; return true; }
                ngi_accept(best, t, u, v, f2u(a.w));
                found = true;
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    out = best;
    if (!found) { out.t = 0; out.tri = NGI_MISS; }
    return found;
}

// ------------------------------------------------------------------------------------------------
// BVH8 — compressed wide node, 80 bytes = 5 x uint4:
//   n0 = (p.x, p.y, p.z as float bits,  ex | ey<<8 | ez<<16 | imask<<24)       e* = biased float exponents
//   n1 = (child_base, -, valid24, -)
//   n2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])
//   n3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])
//   n4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// child box = p + q * 2^(e-127) per axis, q in [0,255] rounded outward.
// imask bit s = slot s is an inner child; inner children are stored contiguously from child_base in slot order. Children sit
// in slots so that (slot ^ octant) approximates front-to-back order (slot bit0/1/2 = child lies towards +x/+y/+z).
// Triangles have FIXED places: triangle j (j < 3) of the leaf in slot s of node ni is record 24 * ni + 3 * s + j of the triangle
// array, and valid24 bit 3 s + j says whether it exists. The places of inner and empty slots are never referenced (nor fetched:
// the array is only touched through the hit masks), so the padding costs address space, not bandwidth — and the node step needs no
// per-child bit arithmetic: an 8-bit box-hit mask h turns into the node group (h & imask, its bits permuted by the ray octant:
// one table lookup) and the triangle group (every bit of h repeated three times & valid24: one table lookup). Round 1 assembled
// both masks child by child from meta bytes (count << offset): 44 ALU-pipe instructions of a node step whose ALU pipe is the
// busiest unit of the trace kernels (profiles/r02_ncu_c3_tq.txt).
// (The collapse, ngi_collapse_node, still emits the compact round-1 form — n1 = (child_base, tri_base, meta[0..3], meta[4..7]),
// meta = unary count << 5 | offset — and ngi_expand_node rewrites it once the node count is known.)
// ------------------------------------------------------------------------------------------------
#define NGI_BVH8_STACK 40
// test-only instrumentation hook (tests/hostsim counts node steps / triangle tests per ray with it); empty in the product
#ifndef NGI_TRACE_COUNT
#define NGI_TRACE_COUNT(what)
#endif
#define NGI_BVH8_MAX_DEPTH 32     /* enforced by the build; traversal stacks are sized from it */

NGI_HD unsigned ngi_byte(unsigned w, int i) { return (w >> (8 * i)) & 0xFFu; }

// bit s of an 8-bit mask -> bit (s ^ o): the order in which a ray of octant-code o wants to visit the slots
NGI_HD unsigned ngi_perm8_calc(unsigned x, const unsigned o) {
    if (o & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
    if (o & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
    if (o & 4u) x = ((x & 0x0Fu) << 4) | ((x >> 4) & 0x0Fu);
    return x;
}
// bit s of an 8-bit mask -> bits 3 s, 3 s + 1, 3 s + 2
NGI_HD unsigned ngi_spread3_calc(unsigned x) {
    x = (x | (x << 8)) & 0x00F00Fu;
    x = (x | (x << 4)) & 0x0C30C3u;
    x = (x | (x << 2)) & 0x249249u;
    return x * 7u;
}
#if defined(__CUDACC__)
// the two lookups of the node step as tables in global memory (3 KB, L1-resident; read with ld.global.nc): filled once per module
// load by ngi_bvh_tables_init (ngi_gpu.cu)
__device__ unsigned char g_ngi_perm8[8 * 256];
__device__ unsigned g_ngi_spread3[256];
#endif
NGI_HD unsigned ngi_perm8(const unsigned x, const unsigned o) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__ldg(g_ngi_perm8 + (o << 8) + x);
#else
    return ngi_perm8_calc(x, o);
#endif
}
NGI_HD unsigned ngi_spread3(const unsigned x) {
#if defined(__CUDA_ARCH__)
    return __ldg(g_ngi_spread3 + x);
#else
    return ngi_spread3_calc(x);
#endif
}

// Quantised plane -> float without a conversion instruction. I2F.U8 runs on the XU pipe at 16 lanes/clk/SM and
// was the busiest pipe of the first traversal kernel (profiles/r01_ncu_c2_steady.txt: XU 71 %). Instead one PRMT
// drops byte i of `w` into mantissa bits 8..15 of 1.0f:  Q = 1 + q * 2^-15  (exact), and the plane distance
//     t = q * s + b  =  Q * (2^15 s) + (b - 2^15 s)
// costs that PRMT and one FMA per plane; (b - 2^15 s) is formed once per axis per node. Its rounding error is at
// most 2^-24 |b - 2^15 s| <= 2^-9 quantisation steps (plus the 2^-24 |b| every slab test has and the box padding
// covers); the build widens every quantised plane by 2^-7 step (ngi_collapse_node) so the decode stays conservative.
// `one` must be the bit pattern of 1.0f held in a REGISTER the compiler cannot constant-fold (the product kernels
// pass it as a kernel parameter, NgiTraceTuning::one_bits): PRMT takes one immediate, and with a literal here ptxas
// keeps the literal and materialises the four selectors in registers instead (+40 moves per node step).
NGI_HD float ngi_q1(unsigned w, int i, unsigned one) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, one, 0x7604u | ((unsigned)i << 4)));
#else
    return u2f(0x3F800000u | (((w >> (8 * i)) & 0xFFu) << 8));
#endif
}
// Two plane distances with ONE instruction: sm_100's packed FFMA2 (PTX fma.rn.f32x2) computes (q0, q1) * s + b with the scale
// and the offset as scalar operands; each half is an IEEE fma.rn, i.e. bit-identical to two fmaf(). The traversal kernels are
// issue-bound (profiles/r01_ncu_c3_final.txt: 68 % issue slots, the node step is half of all issued instructions and 48 of its
// ~260 instructions are these FMAs), so it removes 24 issue slots per node step — MEASURED AND REJECTED (default off): the
// register pairs it needs push the 56-register kernel into spills (6 LDL + 6 STL per step) and k_extend got 2-3 % slower on C2 and
// C3; at 64 registers (no spills, 32 instead of 36 warps per SM) 4 % slower (profiles/r01_sweep_ffma2.txt). The binding resource
// is the ALU pipe (PRMT / FMNMX / mask assembly), which FFMA2 does not touch.
#ifndef NGI_USE_FFMA2
#define NGI_USE_FFMA2 0
#endif
NGI_HD void ngi_fma2(const float q0, const float q1, const float s, const float b, float& r0, float& r1) {
#if defined(__CUDA_ARCH__) && NGI_USE_FFMA2
    unsigned long long q, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(q0), "f"(q1));
    asm("{\n\t.reg .b64 ss, bb;\n\tmov.b64 ss, {%2, %2};\n\tmov.b64 bb, {%3, %3};\n\tfma.rn.f32x2 %0, %1, ss, bb;\n\t}" : "=l"(r) : "l"(q), "f"(s), "f"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
#else
    r0 = fmaf(q0, s, b); r1 = fmaf(q1, s, b);
#endif
}

// per-ray constants of a BVH8 traversal
struct NgiRayCtx {
    f3 o, d;
    float idx, idy, idz;      // clamped reciprocal direction
    float tmin;
    unsigned octinv;          // 7 - octant of the direction
    unsigned one;             // 0x3F800000, see ngi_q1
    bool negx, negy, negz;    // taken from the clamped reciprocals so that -0.0f components stay self-consistent
};

NGI_HD void ngi_ray_ctx(NgiRayCtx& r, const f3 o, const f3 d, const float tmin) {
    r.o = o; r.d = d; r.tmin = tmin;
    r.idx = ngi_safe_rcp_dir(d.x); r.idy = ngi_safe_rcp_dir(d.y); r.idz = ngi_safe_rcp_dir(d.z);
    r.negx = r.idx < 0.0f; r.negy = r.idy < 0.0f; r.negz = r.idz < 0.0f;
    r.octinv = (r.negx ? 0u : 1u) | (r.negy ? 0u : 2u) | (r.negz ? 0u : 4u);
    r.one = 0x3F800000u;
}

// One node step: pops the front-most inner child of `ngroup`, (the caller pushes what is left of the group),
// intersects its 8 quantised child boxes with the ray and returns the new node group / triangle group.
NGI_HD void ngi_bvh8_pop_child(uint2& ngroup, const unsigned octinv, size_t& ni) {
    const unsigned imask = ngroup.y & 0xFFu;
    const int bit = ngi_bfind(ngroup.y);
    ngroup.y &= ~(1u << bit);
    const unsigned slot = ((unsigned)(bit - 24) ^ octinv) & 7u;
    const unsigned rel = (unsigned)ngi_popc(imask & ~(0xFFFFFFFFu << slot));
    ni = (size_t)ngroup.x + rel;
}

NGI_HD void ngi_bvh8_node_step(const uint4* __restrict__ nodes, const size_t ni, const NgiRayCtx& r, const float limit,
                               uint2& ngroup, uint2& tgroup) {
    const uint4 n0 = ngi_ldg(nodes + 5 * ni), n1 = ngi_ldg(nodes + 5 * ni + 1), n2 = ngi_ldg(nodes + 5 * ni + 2);
    const uint4 n3 = ngi_ldg(nodes + 5 * ni + 3), n4 = ngi_ldg(nodes + 5 * ni + 4);

    // per-axis scale 2^(e+15) / d and offset (p - o) / d - 2^(e+15) / d   (see ngi_q1)
    const float sx = u2f(((n0.w & 0xFFu) + 15u) << 23) * r.idx;
    const float sy = u2f((((n0.w >> 8) & 0xFFu) + 15u) << 23) * r.idy;
    const float sz = u2f((((n0.w >> 16) & 0xFFu) + 15u) << 23) * r.idz;
    const float bx = (u2f(n0.x) - r.o.x) * r.idx - sx, by = (u2f(n0.y) - r.o.y) * r.idy - sy, bz = (u2f(n0.z) - r.o.z) * r.idz - sz;
    const unsigned one = r.one;

    unsigned hit8 = 0;                    // bit s = the ray meets the box of slot s (empty slots: a degenerate box, masked out below)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        // near / far plane words per axis, chosen by the ray's sign
        const unsigned lox = h ? n2.y : n2.x, loy = h ? n2.w : n2.z, loz = h ? n3.y : n3.x;
        const unsigned hix = h ? n3.w : n3.z, hiy = h ? n4.y : n4.x, hiz = h ? n4.w : n4.z;
        const unsigned nx = r.negx ? hix : lox, fx = r.negx ? lox : hix;
        const unsigned ny = r.negy ? hiy : loy, fy = r.negy ? loy : hiy;
        const unsigned nz = r.negz ? hiz : loz, fz = r.negz ? loz : hiz;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            // children i and i + 1 together: six packed FMAs for the twelve plane distances
            float tnx[2], tfx[2], tny[2], tfy[2], tnz[2], tfz[2];
            ngi_fma2(ngi_q1(nx, i, one), ngi_q1(nx, i + 1, one), sx, bx, tnx[0], tnx[1]);
            ngi_fma2(ngi_q1(fx, i, one), ngi_q1(fx, i + 1, one), sx, bx, tfx[0], tfx[1]);
            ngi_fma2(ngi_q1(ny, i, one), ngi_q1(ny, i + 1, one), sy, by, tny[0], tny[1]);
            ngi_fma2(ngi_q1(fy, i, one), ngi_q1(fy, i + 1, one), sy, by, tfy[0], tfy[1]);
            ngi_fma2(ngi_q1(nz, i, one), ngi_q1(nz, i + 1, one), sz, bz, tnz[0], tnz[1]);
            ngi_fma2(ngi_q1(fz, i, one), ngi_q1(fz, i + 1, one), sz, bz, tfz[0], tfz[1]);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float tn = fmaxf(fmaxf(tnx[k], tny[k]), fmaxf(tnz[k], r.tmin));
                const float tf = fminf(fminf(tfx[k], tfy[k]), fminf(tfz[k], limit));
                if (tn <= tf) hit8 |= 1u << (4 * h + i + k);
            }
        }
    }
    const unsigned imask = n0.w >> 24;
    ngroup.x = n1.x;
    ngroup.y = (ngi_perm8(hit8 & imask, r.octinv) << 24) | imask;
    tgroup.x = 24u * (unsigned)ni;
    tgroup.y = ngi_spread3(hit8) & n1.z;
}

// per-ray traversal (one thread = one ray, no cooperation): the reference form of the algorithm. The product
// kernels in ngi_gpu.cu run the warp-cooperative version (ngi_trace_warp.cuh) over the same building blocks;
// because the closest hit is an order-free reduction both return bit-identical results.
template <bool ANY_HIT>
NGI_HD bool ngi_trace_bvh8(const uint4* __restrict__ nodes, const float4* __restrict__ tris, const f3 o, const f3 d,
                           const float tmin, const float tmax, NgiHitRec& out) {
    NgiRayCtx r;
    ngi_ray_ctx(r, o, d, tmin);
    NgiHitRec best; best.t = tmax; best.u = 0; best.v = 0; best.tri = NGI_MISS;
    bool found = false;

    uint2 stack[NGI_BVH8_STACK];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);  // root: base 0, pseudo-slot 7
    uint2 tgroup = make_uint2(0u, 0u);

    while (true) {
        if (ngroup.y > 0x00FFFFFFu) {
            size_t ni;
            ngi_bvh8_pop_child(ngroup, r.octinv, ni);
            if (ngroup.y > 0x00FFFFFFu) {
                if (sp < NGI_BVH8_STACK) stack[sp++] = ngroup;
            }
            NGI_TRACE_COUNT(0);
            ngi_bvh8_node_step(nodes, ni, r, best.t, ngroup, tgroup);   // best.t == tmax until a hit is found
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0u) {
            const int bit = ngi_bfind(tgroup.y);
            tgroup.y &= ~(1u << bit);
            const size_t ti = (size_t)tgroup.x + (unsigned)bit;
            const float4 a = ngi_ldg(tris + 3 * ti), b = ngi_ldg(tris + 3 * ti + 1), c = ngi_ldg(tris + 3 * ti + 2);
            float t, u, v;
            NGI_TRACE_COUNT(1);
            if (ngi_tri_test(a, b, c, o, d, tmin, tmax, t, u, v)) {
                if (ANY_HIT) { out.t = t; out.u = u; out.v = v; out.tri = 0; return true; }
                ngi_accept(best, t, u, v, f2u(a.w));
                found = true;
            }
        }

        if (ngroup.y <= 0x00FFFFFFu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    out = best;
    if (!found) { out.t = 0; out.tri = NGI_MISS; }
    return found;
}

// O(n) reference loop over the same triangle records (test aid: proves the BVHs are pure filters)
template <bool ANY_HIT>
NGI_HD bool ngi_trace_brute(const float4* __restrict__ tris, const unsigned n, const f3 o, const f3 d, const float tmin,
                            const float tmax, NgiHitRec& out) {
    NgiHitRec best; best.t = tmax; best.u = 0; best.v = 0; best.tri = NGI_MISS;
    bool found = false;
    for (unsigned i = 0; i < n; i++) {
        const float4 a = ngi_ldg(tris + 3 * (size_t)i), b = ngi_ldg(tris + 3 * (size_t)i + 1), c = ngi_ldg(tris + 3 * (size_t)i + 2);
        float t, u, v;
        if (ngi_tri_test(a, b, c, o, d, tmin, tmax, t, u, v)) {
            if (ANY_HIT) { out.t = t; out.u = u; out.v = v; out.tri = 0; return true; }
            ngi_accept(best, t, u, v, f2u(a.w));
            found = true;
        }
    }
    out = best;
    if (!found) { out.t = 0; out.tri = NGI_MISS; }
    return found;
}

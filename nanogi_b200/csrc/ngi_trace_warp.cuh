// ngi_trace_warp.cuh — warp-cooperative BVH8 traversal: the product form of the ray-query kernels.
//
// Replaces the per-ray calls Scene::Intersect / Scene::Visible -> rtcIntersect (reference
// include/nanogi/rt.hpp:2162-2261) for whole queues of rays.
//
// Why not "one thread = one ray to completion": the first version of k_extend ran with 6.0 of 32 lanes
// active on average (profiles/r01_ncu_c2_steady.txt) — rays of one warp need very different numbers of
// node steps and triangle tests, the compiler's reconvergence points let lanes drift between the node
// code and the triangle code, and a warp lives as long as its slowest ray. Here instead (after Aila & Laine
// 2009 "Understanding the efficiency of ray traversal on GPUs" and Ylitie, Karras, Laine 2017):
//   * persistent warps: grid = SM count x resident CTAs, every warp pulls rays from the queue in chunks
//     (one global atomic per chunk of rays) and hands them to lanes as they go idle (dynamic fetch);
//   * the loop body is warp-uniform and phase-structured — all lanes reconverge (__syncwarp) before the
//     node phase (one 8-wide node step per lane per round) and before the triangle phase (one triangle
//     test per lane per trip) — so each phase issues once for all lanes that need it;
//   * triangle postponing: when only a few lanes have triangles left while most could be doing node
//     steps, the leftover triangle group is pushed on the lane's stack and handled later.
// The closest hit is the order-free lexicographic minimum of (t, triangle id) and occlusion is an
// existence test, so the visiting order (and therefore all of the above) cannot change any result: the
// kernels stay bit-exact against the oracle and against the per-ray form ngi_trace_bvh8.
#pragma once
#include "ngi_bvh.h"

#ifndef NGI_TRACE_BLOCK
#define NGI_TRACE_BLOCK 64         /* threads per CTA of the persistent trace kernels (see ngi_gpu.cu) */
#endif
#define NGI_WARP_STACK 48          /* >= NGI_BVH8_MAX_DEPTH (node groups) + NGI_POSTPONE_MAX_SP (postponed triangle groups) */
#define NGI_POSTPONE_MAX_SP 16

struct NgiTraceTuning {
    int refill_min;     // refill idle lanes when at least this many are idle (or all are)
    int tri_min;        // postpone the triangle phase when fewer lanes than this have triangles pending
    unsigned one_bits;  // 0x3F800000 as a run-time value (keeps it in a register for PRMT, see ngi_q1)
    unsigned chunk;     // rays a warp takes from the queue per global atomic (upper bound)
    unsigned spread;    // != 0: short queues are spread over all warps — chunk = queue length / (2 x warps), at least 1. With a fixed chunk
                        // a queue of a few thousand rays keeps a few dozen warps busy for two ray latencies while the rest of the GPU idles
                        // (the nearly empty late steps of a bdpt batch: 50 us per launch for <= 30 k rays, profiles/r01_launches_bdw_final.txt)
};

// Source concept:
//   unsigned count() const;                       number of rays in the queue
//   unsigned* cursor() const;                     global fetch cursor (zeroed before the launch)
//   unsigned load(unsigned i, f3& o, f3& d, float& tmin, float& tmax);   returns a token handed back to store()
//   void store(unsigned token, bool found, const NgiHitRec& h);
template <bool ANY_HIT, class Source>
__device__ __forceinline__ void ngi_trace_warp_postpone(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Source src,
                                               const NgiTraceTuning tune) {
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n = src.count();

    // warp-uniform fetch state
    unsigned chunk_next = 0, chunk_end = 0;
    unsigned chunk_size = tune.chunk;
    if (tune.spread) {
        const unsigned per = n / (2u * gridDim.x * (blockDim.x >> 5));
        if (per < chunk_size) chunk_size = per ? per : 1u;
    }
    bool exhausted = (n == 0);

    // per-lane ray state
    bool active = false;
    unsigned item = 0;
    NgiRayCtx r;
    float tmax = 0.0f;
    NgiHitRec best; best.t = 0; best.u = 0; best.v = 0; best.tri = NGI_MISS;
    bool found = false;
    uint2 stack[NGI_WARP_STACK];
    int sp = 0;
    int defer = 0;            // consecutive postponements of this lane's triangle group (starvation guard)
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);

    while (true) {
        __syncwarp();
        // ---------------- dynamic fetch ----------------
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle != 0u) {
            const int nidle = __popc(idle);
            if (!exhausted && (nidle >= tune.refill_min || idle == FULL)) {
                if (chunk_next >= chunk_end) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(src.cursor(), chunk_size);
                    base = __shfl_sync(FULL, base, 0);
                    chunk_next = base;
                    chunk_end = base + chunk_size < n ? base + chunk_size : n;
                    if (base >= n) { exhausted = true; chunk_next = chunk_end = 0; }
                }
                if (!exhausted) {
                    const unsigned my = chunk_next + (unsigned)__popc(idle & lt_mask);
                    if (!active && my < chunk_end) {
                        f3 o, d; float tmin;
                        item = src.load(my, o, d, tmin, tmax);
                        ngi_ray_ctx(r, o, d, tmin);
                        r.one = tune.one_bits;
                        best.t = tmax; best.u = 0; best.v = 0; best.tri = NGI_MISS;
                        found = false; sp = 0; defer = 0;
                        ngroup = make_uint2(0u, 0x80000000u);   // root: base 0, pseudo-slot 7
                        tgroup = make_uint2(0u, 0u);
                        active = true;
                    }
                    chunk_next = chunk_next + (unsigned)nidle < chunk_end ? chunk_next + (unsigned)nidle : chunk_end;
                }
            }
            if (exhausted && __ballot_sync(FULL, active) == 0u) break;
        }

        // ---------------- node phase: one node step per lane ----------------
        if (active) {
            if (ngroup.y > 0x00FFFFFFu) {
                size_t ni;
                ngi_bvh8_pop_child(ngroup, r.octinv, ni);
                if (ngroup.y > 0x00FFFFFFu && sp < NGI_WARP_STACK) stack[sp++] = ngroup;   // never full: the build bounds the depth
                ngi_bvh8_node_step(nodes, ni, r, best.t, ngroup, tgroup);
            } else {
                tgroup = ngroup;                                      // a postponed triangle group came off the stack
                ngroup = make_uint2(0u, 0u);
            }
        }
        __syncwarp();

        // ---------------- triangle phase: one triangle per lane per trip ----------------
        while (true) {
            const bool has = active && tgroup.y != 0u;
            const unsigned m = __ballot_sync(FULL, has);
            if (m == 0u) break;
            if (__popc(m) < tune.tri_min) {
                // few lanes busy: postpone if more lanes could be doing node steps instead
                const unsigned node_ready = __ballot_sync(FULL, active && ngroup.y > 0x00FFFFFFu);
                const unsigned starved = __ballot_sync(FULL, has && defer >= 2);
                if (__popc(node_ready) > __popc(m) && starved == 0u) {
                    if (has && sp < NGI_POSTPONE_MAX_SP) { stack[sp++] = tgroup; tgroup.y = 0u; defer++; }
                    if (__ballot_sync(FULL, active && tgroup.y != 0u) == 0u) break;
                }
            }
            if (active && tgroup.y != 0u) {
                const int bit = ngi_bfind(tgroup.y);
                tgroup.y &= ~(1u << bit);
                defer = 0;
                const size_t ti = (size_t)tgroup.x + (unsigned)bit;
                const float4 a = ngi_ldg(tris + 3 * ti), b = ngi_ldg(tris + 3 * ti + 1), c = ngi_ldg(tris + 3 * ti + 2);
                float t, u, v;
                if (ngi_tri_test(a, b, c, r.o, r.d, r.tmin, tmax, t, u, v)) {
                    if (ANY_HIT) {
                        found = true; tgroup.y = 0u; ngroup.y = 0u; sp = 0;   // occluded: this ray is finished
                    } else {
                        ngi_accept(best, t, u, v, f2u(a.w));
                        found = true;
                    }
                }
            }
        }

        // ---------------- pop / retire ----------------
        if (active && ngroup.y <= 0x00FFFFFFu) {
            if (sp == 0) {
                src.store(item, found, best);
                active = false;
            } else {
                ngroup = stack[--sp];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Second form of the loop (NGI_TRACE_TQ, the product default): the triangle backlog lives in SHARED memory.
//
// What the first form (ngi_trace_warp_postpone above) loses, from profiles/r01_ncu_c3_extend_blocks.txt: 29.0 lanes hold a ray
// but only 23.9 take the node step — the other 5.1 have just popped a POSTPONED triangle group off their stack and sit the node
// phase out — and the triangle phase runs at 10.3 lanes. Here a lane keeps traversing while its triangle groups wait in a small
// per-lane LIFO in shared memory (NGI_TQ entries of 8 bytes, [entry][thread]: conflict-free), so
//   * the stack in local memory holds node groups only (depth <= NGI_BVH8_MAX_DEPTH) and every lane with a ray takes the node
//     steps of every round (unless its backlog is full);
//   * the triangle phase starts when at least tune.tri_min lanes have a group waiting (or fewer lanes could step nodes than test
//     triangles), and leaves again when it runs out of lanes: both phases run dense;
//   * a round takes NGI_NODE_REPS node steps per lane: every round pays ~150 warp instructions of fetch / vote / phase-change
//     bookkeeping (profiles/r02_ncu_c3_extend_blocks_s42.txt), two steps per round halve that share.
// A deferred triangle test only delays the shrinking of best.t (a few more node steps pass the culling test); the result is the
// same order-free minimum of (t, id).
// (Measured and rejected, profiles/r02_sweep_trace_tq.txt: software prefetch — the next node of every lane into L1 with
// prefetch.global.L1 = CCTL.E.PF1 made k_extend 2.3x SLOWER on C3, the first triangle of a new group or the ray records of a new
// chunk into L1 / L2 1 % slower; 40 instead of 36 warps per SM (48 registers, 13 spill instructions) 14 % slower; levels of the
// node stack in shared memory: no effect.)
//
// Everything a thread keeps in shared memory sits in ONE pool, [slot][thread] with 8-byte slots, addressed from a per-thread base
// register with immediate offsets (ld.shared / st.shared through inline PTX): with ordinary __shared__ arrays ptxas re-derived the
// addresses from S2R SR_TID.X / SR_CgaCtaId at every use inside the loop (9 S2R per round, ~10 % of the stall samples).
//   slots 0 .. NGI_TQ-1   triangle backlog entries (tri base, 24-bit mask)
//   slot  NGI_TQ          (queue token of the ray, u of its best hit)      needed only when the ray retires: with these three in
//   slot  NGI_TQ + 1      (v of its best hit, -)                           registers ptxas spilled four values around every node step
//   then one 8-byte slot per WARP: {next, end} of the chunk of the queue the warp is handing out (touched only when lanes are refilled)
#ifndef NGI_NODE_REPS
#define NGI_NODE_REPS 2   /* C3 k_extend per launch: 1.301 ms (1 step per round), 1.229 ms (2), 1.218 ms (3, but C2 +8 %): profiles/r02_sweep_trace_tq.txt */
#endif
#ifndef NGI_TQ
#define NGI_TQ 2          /* s40 sweep on C3: 2 entries 1.449 ms, 4 entries 1.466 ms, 6 entries 1.469 ms per k_extend launch */
#endif
#define NGI_POOL_SLOT_BYTES (NGI_TRACE_BLOCK * 8)
#define NGI_POOL_BYTES ((NGI_TQ + 2) * NGI_POOL_SLOT_BYTES + (NGI_TRACE_BLOCK / 32) * 8)

__device__ __forceinline__ void ngi_sts64(const unsigned addr, const uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y)); }
__device__ __forceinline__ uint2 ngi_lds64(const unsigned addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ void ngi_sts32(const unsigned addr, const unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ unsigned ngi_lds32(const unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

template <bool ANY_HIT, class Source>
__device__ __forceinline__ void ngi_trace_warp_tq(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Source src,
                                                  const NgiTraceTuning tune, unsigned char* pool) {
    const unsigned FULL = 0xFFFFFFFFu;
    unsigned sbase;           // shared-space address of this thread's slot 0; `mov` through asm so that it stays in a register
    {
        const unsigned a = (unsigned)__cvta_generic_to_shared(pool) + threadIdx.x * 8u;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(a));
    }
    const unsigned kAux0 = NGI_TQ * NGI_POOL_SLOT_BYTES, kAux1 = (NGI_TQ + 1) * NGI_POOL_SLOT_BYTES;
    const unsigned lane = threadIdx.x & 31u;
    // warp-uniform fetch state {next, end}, see above
    const unsigned cbase = (unsigned)__cvta_generic_to_shared(pool) + (NGI_TQ + 2) * NGI_POOL_SLOT_BYTES + (threadIdx.x >> 5) * 8u;
    if (lane == 0) ngi_sts64(cbase, make_uint2(0u, 0u));
    bool exhausted = (src.count() == 0);

    bool active = false;
    NgiRayCtx r;
    float tmax = 0.0f;
    float best_t = 0.0f; unsigned best_tri = NGI_MISS;
    bool found = false;
    uint2 stack[NGI_BVH8_MAX_DEPTH + 1];      // node groups only
    int sp = 0;
    int tqn = 0;              // triangle groups waiting in backlog slots [0, tqn) of this thread; tqn > 0 implies tgroup.y != 0
    unsigned skipped = 0;     // warp-uniform: consecutive rounds whose triangle phase was put off (starvation guard)
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);

    while (true) {
        __syncwarp();
        // ---------------- dynamic fetch ----------------
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle != 0u) {
            const int nidle = __popc(idle);
            if (!exhausted && (nidle >= tune.refill_min || idle == FULL)) {
                const uint2 ck = ngi_lds64((unsigned)__cvta_generic_to_shared(pool) + (NGI_TQ + 2) * NGI_POOL_SLOT_BYTES + (threadIdx.x >> 5) * 8u);
                unsigned chunk_next = ck.x, chunk_end = ck.y;
                if (chunk_next >= chunk_end) {
                    const unsigned n = src.count();
                    unsigned chunk_size = tune.chunk;
                    if (tune.spread) {
                        const unsigned per = n / (2u * gridDim.x * (blockDim.x >> 5));
                        if (per < chunk_size) chunk_size = per ? per : 1u;
                    }
                    unsigned base = 0;
                    if ((threadIdx.x & 31u) == 0) base = atomicAdd(src.cursor(), chunk_size);
                    base = __shfl_sync(FULL, base, 0);
                    chunk_next = base;
                    chunk_end = base + chunk_size < n ? base + chunk_size : n;
                    if (base >= n) { exhausted = true; chunk_next = chunk_end = 0; }
                }
                if (!exhausted) {
                    const unsigned my = chunk_next + (unsigned)__popc(idle & ((1u << (threadIdx.x & 31u)) - 1u));
                    if (!active && my < chunk_end) {
                        f3 o, d; float tmin;
                        const unsigned token = src.load(my, o, d, tmin, tmax);
                        ngi_sts32(sbase + kAux0, token);
                        ngi_ray_ctx(r, o, d, tmin);
                        r.one = tune.one_bits;
                        best_t = tmax; best_tri = NGI_MISS;
                        found = false; sp = 0; tqn = 0;
                        ngroup = make_uint2(0u, 0x80000000u);   // root: base 0, pseudo-slot 7
                        tgroup = make_uint2(0u, 0u);
                        active = true;
                    }
                    chunk_next = chunk_next + (unsigned)nidle < chunk_end ? chunk_next + (unsigned)nidle : chunk_end;
                }
                __syncwarp();
                if ((threadIdx.x & 31u) == 0)
                    ngi_sts64((unsigned)__cvta_generic_to_shared(pool) + (NGI_TQ + 2) * NGI_POOL_SLOT_BYTES + (threadIdx.x >> 5) * 8u, make_uint2(chunk_next, chunk_end));
            }
            if (exhausted && __ballot_sync(FULL, active) == 0u) break;
        }

        // ---------------- node phase: NGI_NODE_REPS node steps per lane ----------------
#pragma unroll 1
        for (int rep = 0; rep < NGI_NODE_REPS; rep++) {
            if (active && ngroup.y > 0x00FFFFFFu && tqn < NGI_TQ) {
                size_t ni;
                ngi_bvh8_pop_child(ngroup, r.octinv, ni);
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;            // never full: the build bounds the depth
                uint2 tnew;
                ngi_bvh8_node_step(nodes, ni, r, best_t, ngroup, tnew);
                if (tnew.y != 0u) {
                    if (tgroup.y == 0u) tgroup = tnew;
                    else { ngi_sts64(sbase + (unsigned)tqn * NGI_POOL_SLOT_BYTES, tnew); tqn++; }
                }
            }
            if (active && ngroup.y <= 0x00FFFFFFu && sp > 0) ngroup = stack[--sp];
        }
        __syncwarp();

        // ---------------- triangle phase: one triangle per lane per trip ----------------
        while (true) {
            const bool has = active && tgroup.y != 0u;
            const unsigned m = __ballot_sync(FULL, has);
            if (m == 0u) break;
            const int nm = __popc(m);
            if (nm < tune.tri_min) {
                // few lanes have triangles: go on with node steps if more lanes can take one (at most 4 rounds in a row, so that
                // a lane that has nothing but triangles left is not kept waiting)
                const unsigned ready = __ballot_sync(FULL, active && ngroup.y > 0x00FFFFFFu && tqn < NGI_TQ);
                if (__popc(ready) > nm && skipped < 4u) { skipped++; break; }
            }
            skipped = 0u;
            if (has) {
                const int bit = ngi_bfind(tgroup.y);
                tgroup.y &= ~(1u << bit);
                const size_t ti = (size_t)tgroup.x + (unsigned)bit;
                const float4 a = ngi_ldg(tris + 3 * ti), b = ngi_ldg(tris + 3 * ti + 1), c = ngi_ldg(tris + 3 * ti + 2);
                float t, u, v;
                if (ngi_tri_test(a, b, c, r.o, r.d, r.tmin, tmax, t, u, v)) {
                    if (ANY_HIT) {
                        found = true; tgroup.y = 0u; tqn = 0; ngroup.y = 0u; sp = 0;   // occluded: this ray is finished
                    } else {
                        const unsigned id = f2u(a.w);
                        if (t < best_t || (t == best_t && id < best_tri)) {      // ngi_accept: lexicographic minimum of (t, id)
                            best_t = t; best_tri = id;
                            ngi_sts32(sbase + kAux0 + 4u, f2u(u)); ngi_sts32(sbase + kAux1, f2u(v));
                        }
                        found = true;
                    }
                }
                if (tgroup.y == 0u && tqn > 0) { --tqn; tgroup = ngi_lds64(sbase + (unsigned)tqn * NGI_POOL_SLOT_BYTES); }
            }
        }

        // ---------------- retire ----------------
        if (active && ngroup.y <= 0x00FFFFFFu && sp == 0 && tgroup.y == 0u) {
            NgiHitRec best; best.t = best_t; best.tri = best_tri;
            best.u = 0.0f; best.v = 0.0f;
            if (!ANY_HIT && found) { best.u = u2f(ngi_lds32(sbase + kAux0 + 4u)); best.v = u2f(ngi_lds32(sbase + kAux1)); }
            src.store(ngi_lds32(sbase + kAux0), found, best);
            active = false;
        }
    }
}


// what the kernels call: NGI_TRACE_TQ = 1 (default) the shared-memory backlog form, 0 the postponing form (A/B builds)
#ifndef NGI_TRACE_TQ
#define NGI_TRACE_TQ 1
#endif
template <bool ANY_HIT, class Source>
__device__ __forceinline__ void ngi_trace_warp(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Source src,
                                               const NgiTraceTuning tune) {
#if NGI_TRACE_TQ
    __shared__ __align__(16) unsigned char s_pool[NGI_POOL_BYTES];
    ngi_trace_warp_tq<ANY_HIT>(nodes, tris, src, tune, s_pool);
#else
    ngi_trace_warp_postpone<ANY_HIT>(nodes, tris, src, tune);
#endif
}

// ngi_gpu.cu — CUDA kernels (sm_100a) and the C ABI of include/nanogi_gpu.h.
//
// Every kernel is a thin grid wrapper around a per-item body in ngi_build.h / ngi_bvh.h / ngi_wave.h.
// There is no CPU path in this library: without a CUDA device every entry point fails with
// NGI_ERR_NO_DEVICE.
//
// Kernel inventory
//   build   : k_bounds, k_tri_setup, k_morton, cub::DeviceRadixSort, k_gather_sorted, PLOC rounds (k_ploc_nearest,
//             k_ploc_flag, cub::DeviceScan, k_ploc_merge), k_pack2, k_collapse8 (one launch per BVH8 level), k_shade_upload is a plain memcpy
//   queries : k_trace<ACCEL, ANY_HIT>
//   render  : k_iter_begin, k_logic, k_extend, k_shadow   (one wavefront iteration = these four)
//   tests   : k_eval_bsdf
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/nanogi_gpu.h"
#include "ngi_build.h"
#include "ngi_bvh.h"
#include "ngi_scene_host.h"
#include "ngi_wave.h"
#include "ngi_bdpt.h"
#include "ngi_bdpt_wave.h"
#include "ngi_trace_warp.cuh"
#include "ngi_comm.h"

namespace {

thread_local std::string g_err;

int set_err(int code, const std::string& msg) { g_err = msg; return code; }

#define NGI_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            return set_err(e__ == cudaErrorMemoryAllocation ? NGI_ERR_OUT_OF_MEMORY : NGI_ERR_CUDA,     \
                           std::string(#call) + ": " + cudaGetErrorString(e__));                        \
        }                                                                                               \
    } while (0)

#ifndef NGI_STREAM_RAYS
#define NGI_STREAM_RAYS 1
#endif
constexpr int kBlock = 256;
#ifndef NGI_TRACE_BLOCK
#define NGI_TRACE_BLOCK 64   /* persistent trace kernels: small CTAs hand their SM slot back as soon as their 2 warps run dry */
#endif
constexpr int kTraceBlock = NGI_TRACE_BLOCK;
#ifndef NGI_TRACE_MIN_BLOCKS
#define NGI_TRACE_MIN_BLOCKS (1152 / NGI_TRACE_BLOCK)   /* 36 resident warps per SM = 56 registers per thread (the measured configuration) */
#endif
#ifndef NGI_SURFACE_MIN_BLOCKS
#define NGI_SURFACE_MIN_BLOCKS 4
#endif
#ifndef NGI_LOGIC_MIN_BLOCKS
#define NGI_LOGIC_MIN_BLOCKS 4   /* 64 registers + 44 B of spills: latency bound kernel, measured 13 % faster than 3 blocks x 80 registers */
#endif
inline unsigned grid_for(size_t n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }

// ================================================================================================
// build kernels
// ================================================================================================
__device__ __forceinline__ int float_to_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
    const int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
#if defined(__CUDA_ARCH__)
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}

// scene bounds over the raw vertex array: block reduction + 6 atomics per block
__global__ void __launch_bounds__(kBlock) k_bounds(const float* __restrict__ positions, size_t n_verts, int* __restrict__ bounds /*[6] ordered ints*/) {
    float mn[3] = {NGI_INF_F, NGI_INF_F, NGI_INF_F}, mx[3] = {-NGI_INF_F, -NGI_INF_F, -NGI_INF_F};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_verts; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) { const float v = positions[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (int off = 16; off > 0; off >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], off));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(bounds + k, float_to_ordered(mn[k])); atomicMax(bounds + 3 + k, float_to_ordered(mx[k])); }
    }
}

__global__ void __launch_bounds__(kBlock) k_tri_setup(const float* __restrict__ positions, unsigned n, unsigned n_real, float pad, f3 anchor,
                                                      float4* __restrict__ rec, float4* __restrict__ lo, float4* __restrict__ hi) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ngi_tri_setup(positions, i, n_real, pad, anchor, rec, lo, hi);
}

__global__ void __launch_bounds__(kBlock) k_shade_setup(const float* __restrict__ positions, const float* __restrict__ normals, const int* __restrict__ tri_prim,
                                                        unsigned n, float4* __restrict__ shade) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) ngi_shade_setup(positions, normals, tri_prim, t, shade);
}

__global__ void __launch_bounds__(kBlock) k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, unsigned n, f3 mmin, f3 sinv,
                                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = lo[i], b = hi[i];
    keys[i] = ngi_morton63(mk3(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)), mmin, sinv);
    vals[i] = i;
}

__global__ void __launch_bounds__(kBlock) k_gather_sorted(const unsigned* __restrict__ order, unsigned n, const float4* __restrict__ rec_in,
                                                          const float4* __restrict__ tlo, const float4* __restrict__ thi,
                                                          float4* __restrict__ tris2, float4* __restrict__ lo, float4* __restrict__ hi) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned i = order[k];
    tris2[3 * (size_t)k + 0] = rec_in[3 * (size_t)i + 0];
    tris2[3 * (size_t)k + 1] = rec_in[3 * (size_t)i + 1];
    tris2[3 * (size_t)k + 2] = rec_in[3 * (size_t)i + 2];
    lo[n - 1 + k] = tlo[i];
    hi[n - 1 + k] = thi[i];
}

// ---- PLOC rounds (ngi_build.h) ----
__global__ void __launch_bounds__(kBlock) k_ploc_init(int n, const float4* __restrict__ lo, const float4* __restrict__ hi, int* __restrict__ cid,
                                                      float4* __restrict__ clo, float4* __restrict__ chi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    cid[k] = n - 1 + k; clo[k] = lo[n - 1 + k]; chi[k] = hi[n - 1 + k];
}
__global__ void __launch_bounds__(kBlock) k_ploc_nearest(const float4* __restrict__ clo, const float4* __restrict__ chi, int C, int* __restrict__ nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) nn[i] = ngi_ploc_nearest(clo, chi, C, i);
}
__global__ void __launch_bounds__(kBlock) k_ploc_flag(const int* __restrict__ nn, int C, unsigned* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) keep[i] = ngi_ploc_keep(nn, i);
}
__global__ void __launch_bounds__(kBlock) k_ploc_merge(NgiPlocCtx c, int C, const unsigned* __restrict__ keep, unsigned* __restrict__ new_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    ngi_ploc_merge(c, i);
    if (i == C - 1) *new_count = c.pos[i] + keep[i];
}

__global__ void __launch_bounds__(kBlock) k_pack2(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ left,
                                                  const int* __restrict__ right, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) ngi_pack2(lo, hi, left, right, n, i, out);
}

__global__ void __launch_bounds__(128) k_collapse8(NgiCollapseCtx ctx, const NgiBuildTask* __restrict__ tasks, unsigned n_tasks) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tasks) ngi_collapse_node(ctx, tasks[i]);
}
__global__ void __launch_bounds__(kBlock) k_expand8(const uint4* __restrict__ in_nodes, const float4* __restrict__ tris_compact, unsigned n_nodes,
                                                    uint4* __restrict__ out_nodes, float4* __restrict__ tris_fixed) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) ngi_expand_node(in_nodes, tris_compact, i, out_nodes, tris_fixed);
}
// the node step's two lookup tables (ngi_bvh.h), once per device
int bvh_tables_init() {
    static std::mutex m;
    static bool done[64] = {};
    int dev = 0;
    NGI_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(m);
    if (dev >= 0 && dev < 64 && done[dev]) return NGI_OK;
    unsigned char perm[8 * 256]; unsigned spread[256];
    for (unsigned o = 0; o < 8; o++) for (unsigned x = 0; x < 256; x++) perm[(o << 8) + x] = (unsigned char)ngi_perm8_calc(x, o);
    for (unsigned x = 0; x < 256; x++) spread[x] = ngi_spread3_calc(x);
    NGI_CUDA(cudaMemcpyToSymbol(g_ngi_perm8, perm, sizeof(perm)));
    NGI_CUDA(cudaMemcpyToSymbol(g_ngi_spread3, spread, sizeof(spread)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return NGI_OK;
}

// ================================================================================================
// ray-query kernels (ngi_gpu_trace*)
// ================================================================================================
// BVH2 / brute force: cross-check structures, one thread per ray
template <int ACCEL, bool ANY_HIT>
__global__ void __launch_bounds__(kBlock) k_trace(NgiDevScene sc, const NgiRay* __restrict__ rays, size_t n, NgiHit* __restrict__ hits) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i), r1 = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 1);
    const f3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
    NgiHitRec h;
    bool hit;
    if (ACCEL == 1) hit = ngi_trace_bvh2<ANY_HIT>(sc.nodes2, sc.tris2, o, d, r0.w, r1.w, h);
    else hit = ngi_trace_brute<ANY_HIT>(sc.tris2, sc.n_tris, o, d, r0.w, r1.w, h);
    float4 out;
    out.x = hit ? h.t : 0.0f; out.y = hit ? h.u : 0.0f; out.z = hit ? h.v : 0.0f;
    out.w = __uint_as_float(hit ? (ANY_HIT ? 0u : h.tri) : NGI_NO_HIT);
    reinterpret_cast<float4*>(hits)[i] = out;
}

// BVH8 (product): persistent warps with dynamic fetch, see ngi_trace_warp.cuh
template <bool ANY_HIT>
struct QuerySource {
    const float4* rays; float4* hits; unsigned n; unsigned* cur;
    __device__ __forceinline__ unsigned count() const { return n; }
    __device__ __forceinline__ unsigned* cursor() const { return cur; }
    __device__ __forceinline__ unsigned load(unsigned i, f3& o, f3& d, float& tmin, float& tmax) const {
        const float4 r0 = __ldg(rays + 2 * (size_t)i), r1 = __ldg(rays + 2 * (size_t)i + 1);
        o = mk3(r0.x, r0.y, r0.z); d = mk3(r1.x, r1.y, r1.z); tmin = r0.w; tmax = r1.w;
        return i;
    }
    __device__ __forceinline__ void store(unsigned i, bool found, const NgiHitRec& h) const {
        float4 out;
        if (ANY_HIT) out = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(found ? 0u : NGI_NO_HIT));
        else out = found ? make_float4(h.t, h.u, h.v, __uint_as_float(h.tri)) : make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(NGI_NO_HIT));
        hits[i] = out;
    }
};
template <bool ANY_HIT>
__global__ void __launch_bounds__(kTraceBlock) k_trace8(NgiDevScene sc, QuerySource<ANY_HIT> src, NgiTraceTuning tune) {
    ngi_trace_warp<ANY_HIT>(sc.nodes8, sc.tris8, src, tune);
}

// ================================================================================================
// wavefront kernels
// ================================================================================================
struct NgiRenderCounters {
    unsigned long long next_sample;
    unsigned long long total_extend;
    unsigned long long total_shadow;
    unsigned long long iterations;
    unsigned iter[2];   // [0] shadow entries, [1] extend rays of the iteration in flight
    unsigned last[2];   // snapshot of the previous iteration (read by the host for termination)
    unsigned fetch[2];  // dynamic-fetch cursors of the persistent trace kernels ([0] shadow, [1] extend)
    unsigned stage[4];  // [0] surface queue entries, [1] regenerate queue entries
};

constexpr unsigned kIterLogCap = 8192;     // NGI_ITER_LOG: (shadow, extend) ray counts of the first iterations of a lane
__global__ void k_iter_begin(NgiRenderCounters* c, unsigned long long sample_end, unsigned* __restrict__ iter_log) {
    // the eye kernel of the previous iteration started samples next_sample .. next_sample + stage[1] - 1
    const unsigned long long ns = c->next_sample + c->stage[1];
    if (iter_log && c->iterations < kIterLogCap) { iter_log[2 * c->iterations] = c->iter[0]; iter_log[2 * c->iterations + 1] = c->iter[1]; }
    c->next_sample = ns < sample_end ? ns : sample_end;
    c->total_shadow += c->iter[0];
    c->total_extend += c->iter[1];
    c->last[0] = c->iter[0];
    c->last[1] = c->iter[1];
    c->iter[0] = 0u;
    c->iter[1] = 0u;
    c->fetch[0] = 0u;
    c->fetch[1] = 0u;
    c->stage[0] = 0u;
    c->stage[1] = 0u;
    c->iterations += 1ull;
}

// ---- fp64 film behind the fp32 atomics ------------------------------------------------------------------------------
// The reference accumulates in double (std::vector<glm::dvec3>, src/nanogi.cpp:203, :297). The kernels splat with fp32 atomics
// (RED.ADD.F32 is the fast path), and a fp32 sum stops growing once a splat falls below ulp(sum)/2, ~6e-8 of the sum: a pixel that
// collects > 1e7 splats (the eye-vertex connections of ptdirect / bdpt onto a small visible light at 1080p x 1024 spp) would come
// out too dark. So the fp32 film is only a staging buffer: once per wavefront iteration (bdpt: per batch) k_film_fold moves it into
// a fp64 accumulator — atomicExch fetches AND zeroes an element, so splats of kernels running concurrently on other streams are
// never lost — and k_film_finish writes the rounded sums back at the end of the render. 25 MB + 50 MB of traffic per fold at
// 1080p, < 1 % of an iteration.
__global__ void __launch_bounds__(kBlock) k_film_begin(float* __restrict__ film, double* __restrict__ acc, size_t n, int accumulate) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        acc[i] = accumulate ? (double)film[i] : 0.0;
        film[i] = 0.0f;
    }
}
__global__ void __launch_bounds__(kBlock) k_film_fold(float* __restrict__ film, double* __restrict__ acc, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = atomicExch(film + i, 0.0f);
        if (v != 0.0f) atomicAdd(acc + i, (double)v);
    }
}
__global__ void __launch_bounds__(kBlock) k_film_finish(float* __restrict__ film, const double* __restrict__ acc, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        film[i] = (float)(acc[i] + (double)film[i]);
}
constexpr unsigned kFilmGrid = 148u * 8u;

// Logic stage = three dense kernels over compacted queues (warp ballot/popc + one atomic per warp, ngi_queue_alloc):
//   k_classify  every slot: idle / miss / RR / vertex cap (+ pt emission) -> surface_q | regen_q
//   k_surface   surface_q: reconstruct, NEE, BSDF sample -> extend_q (+ shadow_q); paths that end here -> regen_q
//   k_eye       regen_q: next sample index, eye-vertex NEE, camera ray -> extend_q (+ shadow_q)
// History (profiles/): one thread per slot doing everything ran 15.6 of 32 lanes; a block-local regrouping through
// shared memory fixed the lanes (28-29 of 32) but spent 39 % of its stall samples at the two __syncthreads. Separate
// launches over global queues keep the dense warps and have no barrier.
constexpr unsigned kStageGrid = 148u * 8u;

// One thread per SLOT (coalesced state reads); the block compacts its 256 slots IN SLOT ORDER into the two queues and
// reserves queue space with one atomic per queue: the queues are sequences of slot-ordered chunks, so the surface / eye
// kernels read and write the path state almost as coalesced as a slot-indexed kernel would (pushing slot ids in atomic
// order instead made those kernels 2.3x slower: every state access became a scattered 16-byte touch).
// GEN = false: the hot path (pt / ptdirect with a pinhole sensor); GEN = true: the generic flavour (lt / ltdirect, E.area
// sensors) — separate instantiations keep the code and the register budget of the hot-path kernels unchanged
template <bool GEN>
__global__ void __launch_bounds__(kBlock) k_classify(NgiDevScene sc, NgiWaveParams wp) {
    __shared__ unsigned s_warp[2][kBlock / 32];
    __shared__ unsigned s_base[2];
    const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int cls = slot < wp.capacity ? ngi_logic_classify<GEN>(sc, wp, slot) : -1;
    const unsigned ms = __ballot_sync(0xFFFFFFFFu, cls == NGI_CLASS_SURFACE), mr = __ballot_sync(0xFFFFFFFFu, cls == NGI_CLASS_REGENERATE);
    if (lane == 0) { s_warp[0][warp] = (unsigned)__popc(ms); s_warp[1][warp] = (unsigned)__popc(mr); }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned acc = 0;
        for (int w = 0; w < kBlock / 32; w++) { const unsigned c = s_warp[threadIdx.x][w]; s_warp[threadIdx.x][w] = acc; acc += c; }
        s_base[threadIdx.x] = acc ? atomicAdd(wp.stage_counters + threadIdx.x, acc) : 0u;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    if (cls == NGI_CLASS_SURFACE) wp.surface_q[s_base[0] + s_warp[0][warp] + (unsigned)__popc(ms & lt)] = slot;
    else if (cls == NGI_CLASS_REGENERATE) wp.regen_q[s_base[1] + s_warp[1][warp] + (unsigned)__popc(mr & lt)] = slot;
}
// Queue space for a whole block with ONE atomic per queue: every thread passes whether it needs an entry in each of the
// NQ queues and gets its index back; entries of one block are contiguous and in thread order. (A warp-aggregated atomic
// per queue and warp — 3 x 35k same-address atomics per launch — was 70 % of the eye kernel's stall samples.)
template <int NQ>
__device__ __forceinline__ void block_reserve(unsigned* const (&counters)[NQ], const bool (&need)[NQ], unsigned (&index)[NQ],
                                              unsigned (*s_warp)[kBlock / 32], unsigned* s_base) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    unsigned m[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        m[q] = __ballot_sync(0xFFFFFFFFu, need[q]);
        if (lane == 0) s_warp[q][warp] = (unsigned)__popc(m[q]);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        unsigned acc = 0;
        for (int w = 0; w < kBlock / 32; w++) { const unsigned c = s_warp[threadIdx.x][w]; s_warp[threadIdx.x][w] = acc; acc += c; }
        unsigned* ctr = counters[0];
#pragma unroll
        for (int q = 1; q < NQ; q++) if (threadIdx.x == (unsigned)q) ctr = counters[q];          // (a select chain: indexing `counters` puts it in local memory)
        s_base[threadIdx.x] = acc ? atomicAdd(ctr, acc) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; q++) index[q] = s_base[q] + s_warp[q][warp] + (unsigned)__popc(m[q] & lt);
    __syncthreads();   // s_warp / s_base are reused by the next trip
}

template <bool GEN>
__global__ void __launch_bounds__(kBlock, NGI_SURFACE_MIN_BLOCKS) k_surface(NgiDevScene sc, NgiWaveParams wp) {
    __shared__ unsigned s_warp[3][kBlock / 32];
    __shared__ unsigned s_base[3];
    const unsigned n = wp.stage_counters[0];
    unsigned* const counters[3] = {wp.iter_counters + 0, wp.iter_counters + 1, wp.stage_counters + 1};
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        const unsigned e = base + threadIdx.x;
        NgiVertexOut out; out.shadow = false; out.extend = false;
        unsigned slot = 0;
        if (e < n) { slot = wp.surface_q[e]; ngi_logic_surface<GEN>(sc, wp, slot, out); }
        const bool need[3] = {out.shadow, out.extend, e < n && !out.extend};                      // a path that ended here is regenerated
        unsigned idx[3];
        block_reserve<3>(counters, need, idx, s_warp, s_base);
        if (need[0]) ngi_write_shadow(wp, idx[0], out);
        if (need[1]) wp.extend_q[idx[1]] = slot;
        if (need[2]) wp.regen_q[idx[2]] = slot;
    }
}
template <bool GEN>
__global__ void __launch_bounds__(kBlock, NGI_LOGIC_MIN_BLOCKS) k_eye(NgiDevScene sc, NgiWaveParams wp) {
    __shared__ unsigned s_warp[2][kBlock / 32];
    __shared__ unsigned s_base[2];
    const unsigned n = wp.stage_counters[1];
    const unsigned long long first = *wp.next_sample;      // advanced by n in the next k_iter_begin: entry e starts sample first + e
    unsigned* const counters[2] = {wp.iter_counters + 0, wp.iter_counters + 1};
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned e = base + threadIdx.x;
        NgiVertexOut out; out.shadow = false; out.extend = false;
        unsigned slot = 0;
        if (e < n) { slot = wp.regen_q[e]; ngi_logic_eye<GEN>(sc, wp, slot, first + e, out); }
        const bool need[2] = {out.shadow, out.extend};
        unsigned idx[2];
        block_reserve<2>(counters, need, idx, s_warp, s_base);
        if (need[0]) ngi_write_shadow(wp, idx[0], out);
        if (need[1]) wp.extend_q[idx[1]] = slot;
    }
}

// bdpt (ngi_bdpt.h): one sample per lane at a time, subpaths in local memory. The cost of a sample is heavy tailed (~ n^3 in
// its path length, geometric in n): run sample by sample, a warp waited for its longest path and profiles/r01_ncu_bdpt_v1.txt
// shows 3.0 of 32 lanes active. Here every lane is a small state machine — (re)start a sample when out of strategies, then
// evaluate ONE (n, s) strategy per trip — so that the warp's lockstep unit is a strategy, not a sample: a lane with a long path
// simply stays on it for more trips while its neighbours move on to their next samples.
__global__ void __launch_bounds__(128) k_bdpt(NgiDevScene sc, NgiBdParams bp, unsigned long long first, unsigned long long count,
                                              unsigned long long* __restrict__ ray_counters /* [0] extend, [1] shadow */) {
    NgiBdVertex VL[NGI_BD_MAX_VERTS], VE[NGI_BD_MAX_VERTS];
    NgiBdScratch q;
    NgiBdCounters cnt; cnt.extend = 0; cnt.shadow = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long next = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    int nL = 0, nE = 0, n = 1, s = 0;
    bool have = false;
    while (true) {
        // (re)start samples in BATCHES: sampling two subpaths costs far more than evaluating one strategy, and lanes run out of
        // strategies one at a time — refilling them as they went idle ran the sampling code with 1-3 lanes on almost every trip
        // (profiles/r01_ncu_bdpt_v1.txt: 3.0 of 32 lanes). Idle lanes wait until half the warp is idle (or nobody has work).
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have && next < count);
        const unsigned busy = __ballot_sync(0xFFFFFFFFu, have);
        if (idle == 0u && busy == 0u) break;
        if (idle != 0u && (__popc(idle) >= 16 || busy == 0u)) {
            if (!have && next < count) {
                nL = ngi_bd_sample_subpath(sc, bp, first + next, 0, VL, cnt);
                nE = ngi_bd_sample_subpath(sc, bp, first + next, 1, VE, cnt);
                next += stride;
                n = 1; s = 0;
                have = nL > 0 && nE > 0;
            }
        }
        __syncwarp();
        // advance to the next strategy that passes Connect's cheap tests, so that a trip does real work on most lanes
        if (have) {
            do { have = ngi_bd_next_strategy(bp, nL, nE, n, s); } while (have && !ngi_bd_strategy_possible(sc, VL, VE, n, s));
        }
        __syncwarp();
        if (have) ngi_bd_connect(sc, bp, VL, VE, n, s, q, cnt);
    }
    for (int off = 16; off > 0; off >>= 1) {
        cnt.extend += __shfl_xor_sync(0xFFFFFFFFu, cnt.extend, off);
        cnt.shadow += __shfl_xor_sync(0xFFFFFFFFu, cnt.shadow, off);
    }
    if ((threadIdx.x & 31u) == 0u) { atomicAdd(ray_counters + 0, cnt.extend); atomicAdd(ray_counters + 1, cnt.shadow); }
}

// Scene::Intersect's ray query (rt.hpp:2162-2182) for the compacted extend queue of this iteration
struct ExtendSource {
    NgiWaveParams wp;
    __device__ __forceinline__ unsigned count() const { return wp.iter_counters[1]; }
    __device__ __forceinline__ unsigned* cursor() const { return wp.fetch_cursors + 1; }
    __device__ __forceinline__ unsigned load(unsigned i, f3& o, f3& d, float& tmin, float& tmax) const {
        const unsigned slot = wp.extend_q[i];
#if NGI_STREAM_RAYS
        // ray records are read once: streaming loads (evict-first) leave L1 / L2 to the BVH
        const float4 di = __ldcs(&wp.sb[slot].dir_info);
        const double2* sa2 = reinterpret_cast<const double2*>(wp.sa + slot);               // {sample, px} {py, pz}
        const double2 a0 = __ldcs(sa2), a1 = __ldcs(sa2 + 1);
        o = mk3((float)a0.y, (float)a1.x, (float)a1.y);                                      // rt.hpp:2166-2168
#else
        const float4 di = wp.sb[slot].dir_info;
        const NgiSlotA* sa = wp.sa + slot;
        o = mk3((float)sa->px, (float)sa->py, (float)sa->pz);                                // rt.hpp:2166-2168
#endif
        d = mk3(di.x, di.y, di.z); tmin = NGI_EPS_F; tmax = NGI_INF_F;                        // rt.hpp:2246-2249
        return slot;
    }
    __device__ __forceinline__ void store(unsigned slot, bool found, const NgiHitRec& h) const {
        const float4 hv = found ? make_float4(h.t, h.u, h.v, u2f(h.tri)) : make_float4(0.0f, 0.0f, 0.0f, u2f(NGI_MISS));
#if NGI_STREAM_RAYS
        __stcs(wp.hit + slot, hv);
#else
        wp.hit[slot] = hv;
#endif
    }
};
// Scene::Visible (rt.hpp:2251-2261) for the shadow queue + film accumulation (src/nanogi.cpp:706)
struct ShadowSource {
    NgiWaveParams wp;
    __device__ __forceinline__ unsigned count() const { return wp.iter_counters[0]; }
    __device__ __forceinline__ unsigned* cursor() const { return wp.fetch_cursors + 0; }
    __device__ __forceinline__ unsigned load(unsigned e, f3& o, f3& d, float& tmin, float& tmax) const {
        const float4* q = wp.shadow_q + 3 * (size_t)e;
#if NGI_STREAM_RAYS
        const float4 q0 = __ldcs(q), q1 = __ldcs(q + 1);
#else
        const float4 q0 = q[0], q1 = q[1];
#endif
        o = mk3(q0.x, q0.y, q0.z); d = mk3(q1.x, q1.y, q1.z); tmin = NGI_EPS_F; tmax = q0.w;
        return e;
    }
    __device__ __forceinline__ void store(unsigned e, bool occluded, const NgiHitRec&) const {
        if (occluded) return;
        const float4* q = wp.shadow_q + 3 * (size_t)e;
        const float4 q1 = q[1], q2 = q[2];
        ngi_film_add(wp.film, (int)f2u(q1.w), mk3(q2.x, q2.y, q2.z));
    }
};
// per-ray forms of the two trace stages (NGI_RENDER_PER_RAY_TRACE): cross-check only, see tests/test_gpu_parity.py
__global__ void __launch_bounds__(kBlock) k_extend_per_ray(NgiDevScene sc, NgiWaveParams wp) {
    const unsigned n = wp.iter_counters[1];
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) ngi_extend_step(sc, wp, wp.extend_q[q]);
}
__global__ void __launch_bounds__(kBlock) k_shadow_per_ray(NgiDevScene sc, NgiWaveParams wp) {
    const unsigned n = wp.iter_counters[0];
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) ngi_shadow_step(sc, wp, e);
}
__global__ void __launch_bounds__(kTraceBlock, NGI_TRACE_MIN_BLOCKS) k_extend(NgiDevScene sc, NgiWaveParams wp, NgiTraceTuning tune) {
    ExtendSource src; src.wp = wp;
    ngi_trace_warp<false>(sc.nodes8, sc.tris8, src, tune);
}
__global__ void __launch_bounds__(kTraceBlock, NGI_TRACE_MIN_BLOCKS) k_shadow(NgiDevScene sc, NgiWaveParams wp, NgiTraceTuning tune) {
    ShadowSource src; src.wp = wp;
    ngi_trace_warp<true>(sc.nodes8, sc.tris8, src, tune);
}

// ================================================================================================
// wavefront bdpt (ngi_bdpt_wave.h): dense stages over a batch of samples
// ================================================================================================
constexpr int kBdwShadowCursor = 31;          // index into NgiBdWave::cursors
#ifndef NGI_BDW_STEP_MIN_BLOCKS
#define NGI_BDW_STEP_MIN_BLOCKS 1
#endif
__device__ __forceinline__ void bdw_push_ray(float4* q, const unsigned idx, const f3 o, const f3 wo, const float rr, const unsigned w) {
    q[2 * (size_t)idx] = make_float4(o.x, o.y, o.z, rr);
    q[2 * (size_t)idx + 1] = make_float4(wo.x, wo.y, wo.z, u2f(w));
}
__global__ void __launch_bounds__(kBlock) k_bdw_start(NgiDevScene sc, NgiBdParams bp, NgiBdWave wv, int cap) {
    __shared__ unsigned s_warp[1][kBlock / 32];
    __shared__ unsigned s_base[1];
    unsigned* const counters[1] = {wv.counts + 1};
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    f3 o = mk3(0.0f), wo = mk3(0.0f); float rr = 0.0f;
    bool push = false;
    if (w < 2u * wv.batch) push = ngi_bdw_start(sc, bp, wv, w, cap, o, wo, rr);
    const bool need[1] = {push};
    unsigned idx[1];
    block_reserve<1>(counters, need, idx, s_warp, s_base);
    if (push) bdw_push_ray(wv.rays[1], idx[0], o, wo, rr, w);
}
__global__ void __launch_bounds__(kBlock, NGI_BDW_STEP_MIN_BLOCKS) k_bdw_step(NgiDevScene sc, NgiBdParams bp, NgiBdWave wv, int step, int cap) {
    __shared__ unsigned s_warp[1][kBlock / 32];
    __shared__ unsigned s_base[1];
    const unsigned n = wv.counts[step];
    const float4* rq = wv.rays[step & 1];
    float4* nq = wv.rays[(step + 1) & 1];
    unsigned* const counters[1] = {wv.counts + step + 1};
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        const unsigned e = base + threadIdx.x;
        f3 o = mk3(0.0f), wo = mk3(0.0f); float rr = 0.0f; unsigned w = 0;
        bool push = false;
        if (e < n) push = ngi_bdw_step(sc, bp, wv, step, cap, rq[2 * (size_t)e], rq[2 * (size_t)e + 1], wv.hits[e], w, o, wo, rr);
        const bool need[1] = {push};
        unsigned idx[1];
        block_reserve<1>(counters, need, idx, s_warp, s_base);
        if (push) bdw_push_ray(nq, idx[0], o, wo, rr, w);
    }
}
struct BdwExtendSource {
    NgiBdWave wv; int step;
    __device__ __forceinline__ unsigned count() const { return wv.counts[step]; }
    __device__ __forceinline__ unsigned* cursor() const { return wv.cursors + step; }
    __device__ __forceinline__ unsigned load(unsigned i, f3& o, f3& d, float& tmin, float& tmax) const {
        const float4* q = wv.rays[step & 1] + 2 * (size_t)i;
        const float4 r0 = q[0], r1 = q[1];
        o = mk3(r0.x, r0.y, r0.z); d = mk3(r1.x, r1.y, r1.z); tmin = NGI_EPS_F; tmax = NGI_INF_F;
        return i;
    }
    __device__ __forceinline__ void store(unsigned i, bool found, const NgiHitRec& h) const {
        wv.hits[i] = found ? make_float4(h.t, h.u, h.v, u2f(h.tri)) : make_float4(0.0f, 0.0f, 0.0f, u2f(NGI_MISS));
    }
};
struct BdwShadowSource {
    NgiBdWave wv;
    __device__ __forceinline__ unsigned count() const { return wv.n_ray_items; }
    __device__ __forceinline__ unsigned* cursor() const { return wv.cursors + kBdwShadowCursor; }
    __device__ __forceinline__ unsigned load(unsigned i, f3& o, f3& d, float& tmin, float& tmax) const {
        const uint2 it = wv.items[i];
        ngi_bdw_item_ray(wv, it, o, d, tmax);
        tmin = NGI_EPS_F;
        return i;
    }
    __device__ __forceinline__ void store(unsigned i, bool occluded, const NgiHitRec&) const { if (occluded) wv.items[i].y = NGI_BDW_DEAD; }
};
__global__ void __launch_bounds__(kTraceBlock, NGI_TRACE_MIN_BLOCKS) k_bdw_extend(NgiDevScene sc, NgiBdWave wv, int step, NgiTraceTuning tune) {
    BdwExtendSource src; src.wv = wv; src.step = step;
    ngi_trace_warp<false>(sc.nodes8, sc.tris8, src, tune);
}
__global__ void __launch_bounds__(kTraceBlock, NGI_TRACE_MIN_BLOCKS) k_bdw_shadow(NgiDevScene sc, NgiBdWave wv, NgiTraceTuning tune) {
    BdwShadowSource src; src.wv = wv;
    ngi_trace_warp<true>(sc.nodes8, sc.tris8, src, tune);
}
// per sample: (connecting strategies | ray-less strategies << 32), and where its items go: block-level exclusive scan of the packed
// counts (neither half can carry: a batch has < 2^32 strategies) + one 64-bit atomic per block on the batch totals
constexpr int kBdwTotals = 60;                // ctl[60..61] = totals (u64: ray items | ray-less items << 32)
__global__ void __launch_bounds__(kBlock) k_bdw_count(NgiDevScene sc, NgiBdParams bp, NgiBdWave wv, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_warp[kBlock / 32];
    __shared__ unsigned long long s_base;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned nr = 0, nl = 0;
    if (i < wv.batch) ngi_bdw_strategies(sc, bp, wv, i, nr, nl, false, 0u, 0u);
    const unsigned long long v = (unsigned long long)nr | ((unsigned long long)nl << 32);
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int w = 0; w < kBlock / 32; w++) { const unsigned long long c = s_warp[w]; s_warp[w] = acc; acc += c; }
        s_base = acc ? atomicAdd(totals, acc) : 0ull;
    }
    __syncthreads();
    if (i < wv.batch) wv.offsets[i] = s_base + s_warp[warp] + (incl - v);
}
__global__ void __launch_bounds__(kBlock) k_bdw_expand(NgiDevScene sc, NgiBdParams bp, NgiBdWave wv) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wv.batch) return;
    const unsigned long long off = wv.offsets[i];
    unsigned nr, nl;
    ngi_bdw_strategies(sc, bp, wv, i, nr, nl, true, (unsigned)(off & 0xFFFFFFFFull), (unsigned)(off >> 32));
}
__global__ void __launch_bounds__(128) k_bdw_contrib(NgiDevScene sc, NgiBdParams bp, NgiBdWave wv) {
    const unsigned n = wv.n_ray_items + wv.n_rayless;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) ngi_bdw_contrib(sc, bp, wv, wv.items_sorted[e]);
}

__global__ void __launch_bounds__(kBlock) k_eval_bsdf(NgiDevScene sc, const float* __restrict__ q, const float* __restrict__ wo_in, size_t n,
                                                      int force_degenerated, float* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = q + 16 * i;
    const NgiDevPrim& P = sc.prims[(int)a[0]];
    const int type = (int)a[1];
    NgiGeom g; g.sn = mk3(a[2], a[3], a[4]); g.gn = mk3(a[5], a[6], a[7]);
    ngi_tangent_space(g);
    g.albedo = ngi_constant_albedo(P, type);
    const f3 wi = mk3(a[8], a[9], a[10]);
    f3 wo = mk3(0.0f); bool valid = true;
    if (a[14] != 0.0f) wo = mk3(wo_in[3 * i], wo_in[3 * i + 1], wo_in[3 * i + 2]);
    else valid = ngi_sample_bsdf(P, type, g, wi, a[11], a[12], a[13], wo);
    float pdf = 0.0f; f3 fs = mk3(0.0f);
    if (valid) fs = ngi_eval_bsdf(P, type, g, wi, wo, force_degenerated != 0, pdf);
    float* o = out + 8 * i;
    o[0] = wo.x; o[1] = wo.y; o[2] = wo.z; o[3] = fs.x; o[4] = fs.y; o[5] = fs.z; o[6] = pdf; o[7] = valid ? 1.0f : 0.0f;
}

// ================================================================================================
// scene handle
// ================================================================================================
// The path-state buffer (176 B + queues per slot, 369 MB at the default 2 Mi slots) is scene independent, and
// cudaMalloc / cudaFree of that size cost 0.1-0.3 s: a released buffer is parked per device and handed to the
// next scene handle that asks for the same capacity (the e2e path creates one handle per render).
// Scene arrays and build temporaries come from the device's stream-ordered memory pool (cudaMallocAsync) with the
// release threshold lifted, so that a host application that creates and destroys a scene per render (the e2e path:
// ~45 allocations and ~30 frees per scene) re-uses the pool's memory instead of paying cudaMalloc / cudaFree (each
// cudaFree synchronises the device). NGI_POOL_ALLOC=0 restores plain cudaMalloc / cudaFree.
bool pool_alloc_enabled() {
    static const bool on = [] { const char* e = getenv("NGI_POOL_ALLOC"); return e ? atoi(e) != 0 : true; }();
    return on;
}
cudaError_t ngi_dmalloc(void** p, size_t bytes, cudaStream_t st) {
    if (!pool_alloc_enabled()) return cudaMalloc(p, bytes);
    static std::mutex m;
    static bool tuned[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(m);
        if (dev >= 0 && dev < 64 && !tuned[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            tuned[dev] = true;
        }
    }
    return cudaMallocAsync(p, bytes, st);
}
void ngi_dfree(void* p, cudaStream_t st) {
    if (!p) return;
    if (pool_alloc_enabled()) cudaFreeAsync(p, st); else cudaFree(p);
}

struct WaveCacheEntry { void* mem = nullptr; unsigned capacity = 0; };
constexpr int kWaveCacheSlots = 4;
std::mutex g_wave_mutex;
WaveCacheEntry g_wave_cache[64][kWaveCacheSlots];

void* wave_cache_take(int device, unsigned P) {
    std::lock_guard<std::mutex> lock(g_wave_mutex);
    if (device < 0 || device >= 64) return nullptr;
    for (WaveCacheEntry& e : g_wave_cache[device])
        if (e.mem && e.capacity == P) { void* m = e.mem; e.mem = nullptr; e.capacity = 0; return m; }
    return nullptr;
}
void wave_cache_put(int device, void* mem, unsigned P) {
    std::lock_guard<std::mutex> lock(g_wave_mutex);
    if (device >= 0 && device < 64)
        for (WaveCacheEntry& e : g_wave_cache[device])
            if (!e.mem) { e.mem = mem; e.capacity = P; return; }
    cudaFree(mem);
}

// One wavefront pipeline: path-state buffer, counters, the captured graph of a batch of iterations and the streams it
// runs on. A render runs several lanes concurrently on disjoint sample ranges (same film): the kernels of one lane are
// serially dependent and every persistent trace launch ends in a tail of a few long rays, so a second lane's kernels
// fill the SMs the first lane leaves idle (and vice versa).
struct Lane {
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
    unsigned wave_capacity = 0;
    void* wave_mem = nullptr;
    NgiRenderCounters* counters = nullptr;
    NgiRenderCounters* counters_host = nullptr;  // pinned
    cudaGraphExec_t graph_exec = nullptr;
    NgiWaveParams graph_wp{};
    int graph_iters = 0;
    double* fold_acc = nullptr;                  // lane 0 only: fp64 film accumulator folded into after every iteration (k_film_fold)
    size_t fold_n = 0;
    double* graph_acc = nullptr;                 // what the captured graph folds into
    unsigned* iter_log = nullptr;                // NGI_ITER_LOG=<file>: per-iteration ray counts (profiling sessions: rays of an ncu-captured launch)
    std::vector<cudaEvent_t> events;             // per-kernel timing (NGI_RENDER_TIME_KERNELS)
    NgiWaveParams wp{};
    bool running = false;
};

struct Scene {
    int device = 0;
    cudaStream_t stream = nullptr;
    NgiDevScene dev{};
    NgiSceneInfo info{};
    std::vector<void*> allocs;
    std::vector<size_t> alloc_bytes;             // parallel to `allocs` (a built scene is cloned to other devices array by array)
    std::vector<void*> build_temps;              // temporaries of build_scene: released by its guard on every exit path
    unsigned bvh2_depth = 0;                     // height of the binary BVH (accel = 1 needs it <= its traversal stack)
    std::vector<Lane> lanes;
    int num_lanes = 2;
    // persistent trace kernels: grid = SM count x resident CTAs per SM (queried once per kernel)
    NgiTraceTuning tune{4, 12, 0x3F800000u, 64u, 1u};  // refill_min / tri_min: sweeps in profiles/r01_sweep_trace.txt, profiles/r02_sweep_trace_tq.txt
    unsigned grid_extend = 0, grid_shadow = 0, grid_trace[2] = {0, 0};
    unsigned* trace_cursor = nullptr;
    // the extend and shadow kernels of one iteration are independent: the shadow kernel is forked onto the lane's second
    // stream so that its CTAs fill the SMs the extend kernel's tail leaves idle
    bool overlap_trace = true;
    cudaStream_t bd_streams[4] = {nullptr, nullptr, nullptr, nullptr};   // wavefront bdpt: batches in flight
    double* film_acc = nullptr;                  // fp64 film accumulator (k_film_fold), kept across renders of the same size
    size_t film_acc_n = 0;
    unsigned grid_bdw_extend = 0, grid_bdw_shadow = 0;

    ~Scene() {
        cudaSetDevice(device);
        for (cudaStream_t b : bd_streams) if (b) cudaStreamDestroy(b);
        for (Lane& l : lanes) {
            if (l.graph_exec) cudaGraphExecDestroy(l.graph_exec);
            for (auto e : l.events) cudaEventDestroy(e);
            if (l.wave_mem) wave_cache_put(device, l.wave_mem, l.wave_capacity);
            if (l.iter_log) cudaFree(l.iter_log);
            if (l.counters) cudaFree(l.counters);
            if (l.counters_host) cudaFreeHost(l.counters_host);
            if (l.ev_fork) cudaEventDestroy(l.ev_fork);
            if (l.ev_join) cudaEventDestroy(l.ev_join);
            if (l.ev_done) cudaEventDestroy(l.ev_done);
            if (l.stream2) cudaStreamDestroy(l.stream2);
            if (l.stream) cudaStreamDestroy(l.stream);
        }
        for (void* p : allocs) ngi_dfree(p, stream);
        if (trace_cursor) ngi_dfree(trace_cursor, stream);
        if (film_acc) ngi_dfree(film_acc, stream);
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    }
};

template <class T>
int dev_alloc(Scene* s, T** out, size_t count, bool keep) {
    void* p = nullptr;
    NGI_CUDA(ngi_dmalloc(&p, std::max<size_t>(count * sizeof(T), 16), s->stream));
    if (keep) { s->allocs.push_back(p); s->alloc_bytes.push_back(std::max<size_t>(count * sizeof(T), 16)); s->info.device_bytes += count * sizeof(T); }
    else s->build_temps.push_back(p);
    *out = (T*)p;
    return NGI_OK;
}

template <class K>
int persistent_grid(K kernel, unsigned* out) {
    int dev = 0, sms = 0, per_sm = 0;
    NGI_CUDA(cudaGetDevice(&dev));
    NGI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NGI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceBlock, 0));
    *out = (unsigned)(sms * (per_sm > 0 ? per_sm : 1));
    return NGI_OK;
}

int init_trace_launch(Scene* s) {
    int rc;
    if ((rc = persistent_grid(k_extend, &s->grid_extend))) return rc;
    if ((rc = persistent_grid(k_shadow, &s->grid_shadow))) return rc;
    if ((rc = persistent_grid(k_trace8<false>, &s->grid_trace[0]))) return rc;
    if ((rc = persistent_grid(k_trace8<true>, &s->grid_trace[1]))) return rc;
    NGI_CUDA(ngi_dmalloc((void**)&s->trace_cursor, sizeof(unsigned), s->stream));
    NGI_CUDA(cudaStreamSynchronize(s->stream));
    if (const char* e = getenv("NGI_TRACE_CARVEOUT")) {      // shared-memory carve-out (percent) of the trace kernels: they use none, L1 holds the BVH
        const int pct = atoi(e);
        cudaFuncSetAttribute(k_extend, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(k_shadow, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    if (const char* e = getenv("NGI_TRACE_GRID_PCT")) {      // fewer resident trace CTAs leave registers for the other lane's logic kernels
        const int pct = std::min(100, std::max(10, atoi(e)));
        s->grid_extend = std::max(148u, s->grid_extend * (unsigned)pct / 100u / 148u * 148u);
        s->grid_shadow = std::max(148u, s->grid_shadow * (unsigned)pct / 100u / 148u * 148u);
    }
    if (const char* e = getenv("NGI_TRACE_REFILL_MIN")) s->tune.refill_min = atoi(e);
    if (const char* e = getenv("NGI_TRACE_TRI_MIN")) s->tune.tri_min = atoi(e);
    if (const char* e = getenv("NGI_TRACE_OVERLAP")) s->overlap_trace = atoi(e) != 0;
    if (const char* e = getenv("NGI_LANES")) s->num_lanes = std::min(4, std::max(1, atoi(e)));
    if (const char* e = getenv("NGI_TRACE_CHUNK")) s->tune.chunk = (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("NGI_TRACE_SPREAD")) s->tune.spread = (unsigned)std::max(0, atoi(e));
    return NGI_OK;
}

int build_scene(Scene* s, const NgiSceneDesc* desc) {
    NgiHostArrays ha;
    if (!ngi_prepare_scene(desc, ha)) return set_err(ha.error.find("not supported") != std::string::npos ? NGI_ERR_UNSUPPORTED : NGI_ERR_INVALID_ARGUMENT, ha.error);
    cudaStream_t st = s->stream;
    const unsigned nr = ha.n_real;
    const unsigned n = nr < 2 ? 2 : nr;

    // events and every keep = false allocation below are released when this function returns, whichever way
    struct BuildGuard {
        Scene* s; cudaStream_t st; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        ~BuildGuard() {
            for (void* p : s->build_temps) ngi_dfree(p, st);
            s->build_temps.clear();
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
        }
    } guard{s, st};
    cudaEvent_t& ev0 = guard.ev0;
    cudaEvent_t& ev1 = guard.ev1;
    NGI_CUDA(cudaEventCreate(&ev0));
    NGI_CUDA(cudaEventCreate(&ev1));

    // ---- upload ----
    float* d_pos = nullptr;
    int rc;
    if ((rc = dev_alloc(s, &d_pos, (size_t)std::max(nr, 1u) * 9, false))) return rc;
    if (nr) NGI_CUDA(cudaMemcpyAsync(d_pos, desc->positions, (size_t)nr * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    float4* d_shade = nullptr; NgiDevPrim* d_prims = nullptr; unsigned* d_lights = nullptr; float* d_cdf = nullptr;
    // shading records are assembled on the device from positions + normals + the per-triangle primitive index (k_shade_setup): the
    // host loop and the 80 B / triangle upload it replaces were a third of scene_create on the 1 M-triangle scene
    float* d_nrm = nullptr; int* d_triprim = nullptr;
    if ((rc = dev_alloc(s, &d_shade, (size_t)nr * 5, true))) return rc;
    if ((rc = dev_alloc(s, &d_nrm, (size_t)std::max(nr, 1u) * 9, false))) return rc;
    if ((rc = dev_alloc(s, &d_triprim, std::max(nr, 1u), false))) return rc;
    if (nr) {
        NGI_CUDA(cudaMemcpyAsync(d_nrm, desc->normals, (size_t)nr * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
        NGI_CUDA(cudaMemcpyAsync(d_triprim, ha.tri_prim.data(), (size_t)nr * sizeof(int), cudaMemcpyHostToDevice, st));
        k_shade_setup<<<grid_for(nr), kBlock, 0, st>>>(d_pos, d_nrm, d_triprim, nr, d_shade);
    }
    if ((rc = dev_alloc(s, &d_prims, ha.prims.size(), true))) return rc;
    if ((rc = dev_alloc(s, &d_lights, ha.light_prims.size(), true))) return rc;
    if ((rc = dev_alloc(s, &d_cdf, ha.cdf.size(), true))) return rc;
    NGI_CUDA(cudaMemcpyAsync(d_prims, ha.prims.data(), ha.prims.size() * sizeof(NgiDevPrim), cudaMemcpyHostToDevice, st));
    if (!ha.light_prims.empty()) NGI_CUDA(cudaMemcpyAsync(d_lights, ha.light_prims.data(), ha.light_prims.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    if (!ha.cdf.empty()) NGI_CUDA(cudaMemcpyAsync(d_cdf, ha.cdf.data(), ha.cdf.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    float *d_uv = nullptr, *d_texdata = nullptr; NgiDevTex* d_tex = nullptr;
    if (!ha.shade_uv.empty()) {
        if ((rc = dev_alloc(s, &d_uv, ha.shade_uv.size(), true))) return rc;
        NGI_CUDA(cudaMemcpyAsync(d_uv, ha.shade_uv.data(), ha.shade_uv.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (!ha.textures.empty()) {
        if ((rc = dev_alloc(s, &d_tex, ha.textures.size(), true))) return rc;
        if ((rc = dev_alloc(s, &d_texdata, ha.tex_data.size(), true))) return rc;
        NGI_CUDA(cudaMemcpyAsync(d_tex, ha.textures.data(), ha.textures.size() * sizeof(NgiDevTex), cudaMemcpyHostToDevice, st));
        NGI_CUDA(cudaMemcpyAsync(d_texdata, ha.tex_data.data(), ha.tex_data.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    }

    NGI_CUDA(cudaEventRecord(ev0, st));
    // ---- 0. bounds ----
    int* d_bounds = nullptr;
    if ((rc = dev_alloc(s, &d_bounds, 6, false))) return rc;
    float smin[3] = {0, 0, 0}, smax[3] = {0, 0, 0};
    if (nr) {
        const int init[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        NGI_CUDA(cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
        k_bounds<<<std::min(grid_for((size_t)nr * 3), 148u * 8u), kBlock, 0, st>>>(d_pos, (size_t)nr * 3, d_bounds);
        int hb[6];
        NGI_CUDA(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
        NGI_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 3; k++) { smin[k] = ordered_to_float(hb[k]); smax[k] = ordered_to_float(hb[3 + k]); }
    }
    const float pad = ngi_box_pad(smin, smax, ha.sensor);
    const f3 anchor = mk3(smin[0], smin[1], smin[2]);

    // ---- 1. triangle records + boxes ----
    float4 *d_rec = nullptr, *d_tlo = nullptr, *d_thi = nullptr;
    if ((rc = dev_alloc(s, &d_rec, (size_t)n * 3, false))) return rc;
    if ((rc = dev_alloc(s, &d_tlo, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_thi, n, false))) return rc;
    k_tri_setup<<<grid_for(n), kBlock, 0, st>>>(d_pos, n, nr, pad, anchor, d_rec, d_tlo, d_thi);

    // ---- 2. Morton + sort ----
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr; unsigned *d_vals = nullptr, *d_vals2 = nullptr;
    if ((rc = dev_alloc(s, &d_keys, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_keys2, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_vals, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_vals2, n, false))) return rc;
    const f3 mmin = mk3(smin[0] - pad, smin[1] - pad, smin[2] - pad);
    f3 sinv;
    sinv.x = 1.0f / fmaxf(smax[0] - smin[0] + 2 * pad, 1e-30f);
    sinv.y = 1.0f / fmaxf(smax[1] - smin[1] + 2 * pad, 1e-30f);
    sinv.z = 1.0f / fmaxf(smax[2] - smin[2] + 2 * pad, 1e-30f);
    k_morton<<<grid_for(n), kBlock, 0, st>>>(d_tlo, d_thi, n, mmin, sinv, d_keys, d_vals);
    size_t tmp_bytes = 0;
    NGI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));
    unsigned char* d_tmp = nullptr;
    if ((rc = dev_alloc(s, &d_tmp, tmp_bytes, false))) return rc;
    NGI_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));

    // ---- 3-5. binary BVH by PLOC over the Morton order ----
    float4 *d_tris2 = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_nodes2 = nullptr;
    int *d_left = nullptr, *d_right = nullptr; unsigned* d_ncnt = nullptr;
    if ((rc = dev_alloc(s, &d_tris2, (size_t)n * 3, true))) return rc;
    if ((rc = dev_alloc(s, &d_lo, 2 * (size_t)n - 1, false))) return rc;
    if ((rc = dev_alloc(s, &d_hi, 2 * (size_t)n - 1, false))) return rc;
    if ((rc = dev_alloc(s, &d_nodes2, (size_t)(n - 1) * 4, true))) return rc;
    if ((rc = dev_alloc(s, &d_left, n - 1, false))) return rc;
    if ((rc = dev_alloc(s, &d_right, n - 1, false))) return rc;
    if ((rc = dev_alloc(s, &d_ncnt, n - 1, false))) return rc;
    unsigned* d_depth2 = nullptr;
    if ((rc = dev_alloc(s, &d_depth2, n - 1, false))) return rc;
    k_gather_sorted<<<grid_for(n), kBlock, 0, st>>>(d_vals2, n, d_rec, d_tlo, d_thi, d_tris2, d_lo, d_hi);
    // the sort buffers are dead now; cluster arrays ping-pong
    int *d_cid[2] = {nullptr, nullptr}, *d_nn = nullptr; float4 *d_clo[2] = {nullptr, nullptr}, *d_chi[2] = {nullptr, nullptr};
    unsigned *d_keep = nullptr, *d_spos = nullptr, *d_newc = nullptr; unsigned char* d_scan_tmp = nullptr;
    for (int k = 0; k < 2; k++) {
        if ((rc = dev_alloc(s, &d_cid[k], n, false))) return rc;
        if ((rc = dev_alloc(s, &d_clo[k], n, false))) return rc;
        if ((rc = dev_alloc(s, &d_chi[k], n, false))) return rc;
    }
    if ((rc = dev_alloc(s, &d_nn, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_keep, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_spos, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_newc, 1, false))) return rc;
    // SAH-optimal collapse decisions, one row per binary node, filled by the PLOC merges (ngi_dp_node); NGI_COLLAPSE_GREEDY=1 keeps
    // round 1's greedy collapse for A/B runs
    NgiDpRow* d_dp = nullptr;
    const bool greedy_collapse = getenv("NGI_COLLAPSE_GREEDY") && atoi(getenv("NGI_COLLAPSE_GREEDY"));
    const float sah_c_prim = getenv("NGI_SAH_CPRIM") ? (float)atof(getenv("NGI_SAH_CPRIM")) : NGI_SAH_C_PRIM;
    if (!greedy_collapse && (rc = dev_alloc(s, &d_dp, n - 1, false))) return rc;
    size_t scan_bytes = 0;
    NGI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_keep, d_spos, (int)n, st));
    if ((rc = dev_alloc(s, &d_scan_tmp, scan_bytes, false))) return rc;
    k_ploc_init<<<grid_for(n), kBlock, 0, st>>>((int)n, d_lo, d_hi, d_cid[0], d_clo[0], d_chi[0]);
    {
        unsigned C = n, merges_done = 0, rounds = 0;
        int cur = 0;
        while (C > 1) {
            k_ploc_nearest<<<grid_for(C), kBlock, 0, st>>>(d_clo[cur], d_chi[cur], (int)C, d_nn);
            k_ploc_flag<<<grid_for(C), kBlock, 0, st>>>(d_nn, (int)C, d_keep);
            NGI_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_bytes, d_keep, d_spos, (int)C, st));
            NgiPlocCtx pc;
            pc.nn = d_nn; pc.pos = d_spos; pc.cid_in = d_cid[cur]; pc.clo_in = d_clo[cur]; pc.chi_in = d_chi[cur];
            pc.cid_out = d_cid[cur ^ 1]; pc.clo_out = d_clo[cur ^ 1]; pc.chi_out = d_chi[cur ^ 1];
            pc.lo = d_lo; pc.hi = d_hi; pc.left = d_left; pc.right = d_right; pc.cnt = d_ncnt; pc.n = (int)n;
            pc.next_id = (int)(n - 2) - (int)merges_done;
            pc.dp = d_dp; pc.c_node = 1.0f; pc.c_prim = sah_c_prim; pc.depth = d_depth2;
            k_ploc_merge<<<grid_for(C), kBlock, 0, st>>>(pc, (int)C, d_keep, d_newc);
            unsigned newC = 0;
            NGI_CUDA(cudaMemcpyAsync(&newC, d_newc, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            NGI_CUDA(cudaStreamSynchronize(st));
            if (newC == 0 || newC >= C) return set_err(NGI_ERR_CUDA, "PLOC made no progress");
            merges_done += C - newC;
            C = newC; cur ^= 1;
            if (++rounds > 100000) return set_err(NGI_ERR_CUDA, "PLOC did not terminate");
        }
        if (merges_done != n - 1) return set_err(NGI_ERR_CUDA, "PLOC merge count mismatch");
        NGI_CUDA(cudaMemcpyAsync(&s->bvh2_depth, d_depth2, sizeof(unsigned), cudaMemcpyDeviceToHost, st));     // node 0 = root
        NGI_CUDA(cudaStreamSynchronize(st));
    }
    k_pack2<<<grid_for(n - 1), kBlock, 0, st>>>(d_lo, d_hi, d_left, d_right, (int)n, d_nodes2);

    // ---- 6. collapse to BVH8, one launch per level ----
    uint4* d_nodes8_tmp = nullptr; float4* d_tris8_compact = nullptr; unsigned* d_cnt = nullptr; NgiBuildTask *d_q0 = nullptr, *d_q1 = nullptr;
    if ((rc = dev_alloc(s, &d_nodes8_tmp, (size_t)n * 5, false))) return rc;
    if ((rc = dev_alloc(s, &d_tris8_compact, (size_t)n * 3, false))) return rc;
    if ((rc = dev_alloc(s, &d_cnt, 4, false))) return rc;
    if ((rc = dev_alloc(s, &d_q0, n, false))) return rc;
    if ((rc = dev_alloc(s, &d_q1, n, false))) return rc;
    {
        const unsigned init[4] = {1u, 0u, 0u, 0u};
        NGI_CUDA(cudaMemcpyAsync(d_cnt, init, sizeof(init), cudaMemcpyHostToDevice, st));
        const NgiBuildTask root = {0, 0u};
        NGI_CUDA(cudaMemcpyAsync(d_q0, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    }
    NgiCollapseCtx ctx;
    ctx.lo = d_lo; ctx.hi = d_hi; ctx.left = d_left; ctx.right = d_right; ctx.cnt = d_ncnt; ctx.tris2 = d_tris2; ctx.n = (int)n;
    ctx.nodes8 = d_nodes8_tmp; ctx.tris8 = d_tris8_compact; ctx.counters = d_cnt; ctx.dp = d_dp;
    unsigned n_tasks = 1, depth = 0;
    unsigned hc[4] = {1, 0, 0, 0};
    NgiBuildTask *qin = d_q0, *qout = d_q1;
    while (n_tasks > 0) {
        depth++;
        ctx.out_tasks = qout;
        NGI_CUDA(cudaMemsetAsync(d_cnt + 2, 0, sizeof(unsigned), st));
        k_collapse8<<<grid_for(n_tasks, 128), 128, 0, st>>>(ctx, qin, n_tasks);
        NGI_CUDA(cudaMemcpyAsync(hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
        NGI_CUDA(cudaStreamSynchronize(st));
        n_tasks = hc[2];
        std::swap(qin, qout);
        if (depth > 4096) return set_err(NGI_ERR_CUDA, "BVH8 collapse did not terminate");
    }
    NGI_CUDA(cudaGetLastError());
    if (hc[1] != n) return set_err(NGI_ERR_CUDA, "BVH8 collapse lost triangles: " + std::to_string(hc[1]) + " of " + std::to_string(n));
    if (depth > NGI_BVH8_MAX_DEPTH) return set_err(NGI_ERR_UNSUPPORTED, "BVH8 depth " + std::to_string(depth) + " exceeds the traversal stack");
    const unsigned n_nodes8 = hc[0];
    if ((unsigned long long)n_nodes8 * 24ull >= 0xFFFFFFFFull) return set_err(NGI_ERR_UNSUPPORTED, "more than 2^32 / 24 BVH8 nodes");
    // traversal form: exact-size node array with valid24 masks, triangles at their fixed places 24 * node + 3 * slot + j (ngi_bvh.h)
    uint4* d_nodes8 = nullptr; float4* d_tris8 = nullptr;
    if ((rc = dev_alloc(s, &d_nodes8, (size_t)n_nodes8 * 5, true))) return rc;
    if ((rc = dev_alloc(s, &d_tris8, (size_t)n_nodes8 * 24 * 3, true))) return rc;
    NGI_CUDA(cudaMemsetAsync(d_tris8, 0, (size_t)n_nodes8 * 24 * 3 * sizeof(float4), st));
    k_expand8<<<grid_for(n_nodes8), kBlock, 0, st>>>(d_nodes8_tmp, d_tris8_compact, n_nodes8, d_nodes8, d_tris8);
    if ((rc = bvh_tables_init())) return rc;
    NGI_CUDA(cudaEventRecord(ev1, st));
    NGI_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    NGI_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    NgiDevScene& d = s->dev;
    d.nodes8 = d_nodes8; d.tris8 = d_tris8; d.nodes2 = d_nodes2; d.tris2 = d_tris2;
    d.shade_tris = d_shade; d.prims = d_prims; d.light_prims = d_lights; d.cdf = d_cdf;
    d.shade_uv = d_uv; d.textures = d_tex; d.tex_data = d_texdata;
    d.n_tris = n; d.n_lights = (unsigned)ha.light_prims.size(); d.sensor = ha.sensor;
    s->info.num_tris = nr; s->info.bvh8_nodes = n_nodes8; s->info.bvh2_nodes = n - 1;
    s->info.build_gpu_seconds = ms * 1e-3;
    for (int k = 0; k < 3; k++) { s->info.scene_min[k] = smin[k]; s->info.scene_max[k] = smax[k]; }
    s->info.num_lights = d.n_lights; s->info.bvh8_max_depth = depth; s->info.bvh2_max_depth = s->bvh2_depth;
    return init_trace_launch(s);
}

// ---- wavefront state ---------------------------------------------------------------------------
int ensure_lane(Scene* s, Lane& l, unsigned P) {
    if (!l.stream) {
        NGI_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        NGI_CUDA(cudaStreamCreateWithFlags(&l.stream2, cudaStreamNonBlocking));
        NGI_CUDA(cudaEventCreateWithFlags(&l.ev_fork, cudaEventDisableTiming));
        NGI_CUDA(cudaEventCreateWithFlags(&l.ev_join, cudaEventDisableTiming));
        NGI_CUDA(cudaEventCreateWithFlags(&l.ev_done, cudaEventDisableTiming));
        NGI_CUDA(cudaMalloc((void**)&l.counters, sizeof(NgiRenderCounters)));
        NGI_CUDA(cudaMallocHost((void**)&l.counters_host, sizeof(NgiRenderCounters)));
        if (getenv("NGI_ITER_LOG")) NGI_CUDA(cudaMalloc((void**)&l.iter_log, 2 * kIterLogCap * sizeof(unsigned)));
    }
    if (l.wave_capacity == P && l.wave_mem) return NGI_OK;
    if (l.graph_exec) { cudaGraphExecDestroy(l.graph_exec); l.graph_exec = nullptr; }
    if (l.wave_mem) { wave_cache_put(s->device, l.wave_mem, l.wave_capacity); l.wave_mem = nullptr; }
    // per slot: record A 32 (sample, p) + record B 32 (thr_pix, dir_info) + hit 16 = 80 B; shadow queue 2 entries x 48 B; 3 slot queues
    const size_t bytes = (size_t)P * (8 + 16 + 24 + 16 + 16 + 96 + 4 + 4 + 4);
    l.wave_mem = wave_cache_take(s->device, P);
    if (!l.wave_mem) NGI_CUDA(cudaMalloc(&l.wave_mem, bytes));
    l.wave_capacity = P;
    return NGI_OK;
}

void carve_wave(Lane& l, NgiWaveParams& wp) {
    const size_t P = l.wave_capacity;
    unsigned char* p = (unsigned char*)l.wave_mem;
    wp.sa = (NgiSlotA*)p; p += P * 32;          // cudaMalloc'ed memory is 256-byte aligned: every record starts on a sector
    wp.sb = (NgiSlotB*)p; p += P * 32;
    wp.hit = (float4*)p; p += P * 16;
    wp.shadow_q = (float4*)p; p += P * 96;
    wp.extend_q = (unsigned*)p; p += P * 4;
    wp.surface_q = (unsigned*)p; p += P * 4;
    wp.regen_q = (unsigned*)p; p += P * 4;
    wp.stage_counters = l.counters->stage;
    wp.iter_counters = l.counters->iter;
    wp.fetch_cursors = l.counters->fetch;
    wp.next_sample = &l.counters->next_sample;
    wp.capacity = (unsigned)P;
}

bool same_wp(const NgiWaveParams& a, const NgiWaveParams& b) { return memcmp(&a, &b, sizeof(a)) == 0; }

int launch_iteration(Scene* s, Lane& l, bool timed, size_t& ev_used, bool per_ray = false) {
    const NgiWaveParams& wp = l.wp;
    cudaStream_t st = l.stream;
    const unsigned P = wp.capacity;
    const bool direct = wp.renderer == NGI_RENDERER_PTDIRECT || wp.renderer == NGI_RENDERER_LTDIRECT;   // renderers with a shadow queue
    const bool lt = wp.renderer >= NGI_RENDERER_LT || s->dev.sensor.kind == NGI_ET_AREA;   // generic kernels
    k_iter_begin<<<1, 1, 0, st>>>(l.counters, wp.sample_end, l.iter_log);
    if (timed) {
        while (l.events.size() < ev_used + 4) { cudaEvent_t e; NGI_CUDA(cudaEventCreate(&e)); l.events.push_back(e); }
        NGI_CUDA(cudaEventRecord(l.events[ev_used], st));
    }
    const unsigned sg = std::min(grid_for(P), kStageGrid);
    if (lt) {
        k_classify<true><<<grid_for(P), kBlock, 0, st>>>(s->dev, wp);
        k_surface<true><<<sg, kBlock, 0, st>>>(s->dev, wp);
        k_eye<true><<<sg, kBlock, 0, st>>>(s->dev, wp);
    } else {
        k_classify<false><<<grid_for(P), kBlock, 0, st>>>(s->dev, wp);
        k_surface<false><<<sg, kBlock, 0, st>>>(s->dev, wp);
        k_eye<false><<<sg, kBlock, 0, st>>>(s->dev, wp);
    }
    if (timed) NGI_CUDA(cudaEventRecord(l.events[ev_used + 1], st));
    if (!timed && direct && s->overlap_trace) {
        // fork: shadow on stream2, extend on the lane's main stream, join
        NGI_CUDA(cudaEventRecord(l.ev_fork, st));
        NGI_CUDA(cudaStreamWaitEvent(l.stream2, l.ev_fork, 0));
        k_extend<<<s->grid_extend, kTraceBlock, 0, st>>>(s->dev, wp, s->tune);
        k_shadow<<<s->grid_shadow, kTraceBlock, 0, l.stream2>>>(s->dev, wp, s->tune);
        NGI_CUDA(cudaEventRecord(l.ev_join, l.stream2));
        NGI_CUDA(cudaStreamWaitEvent(st, l.ev_join, 0));
        if (l.fold_acc) k_film_fold<<<kFilmGrid, kBlock, 0, st>>>(wp.film, l.fold_acc, l.fold_n);
        return NGI_OK;
    }
    if (per_ray) k_extend_per_ray<<<std::min(grid_for(P), 148u * 16u), kBlock, 0, st>>>(s->dev, wp);
    else k_extend<<<s->grid_extend, kTraceBlock, 0, st>>>(s->dev, wp, s->tune);
    if (timed) NGI_CUDA(cudaEventRecord(l.events[ev_used + 2], st));
    if (direct && per_ray) k_shadow_per_ray<<<std::min(grid_for((size_t)P * 2), 148u * 16u), kBlock, 0, st>>>(s->dev, wp);
    else if (direct) k_shadow<<<s->grid_shadow, kTraceBlock, 0, st>>>(s->dev, wp, s->tune);
    if (timed) { NGI_CUDA(cudaEventRecord(l.events[ev_used + 3], st)); ev_used += 4; }
    if (l.fold_acc) k_film_fold<<<kFilmGrid, kBlock, 0, st>>>(wp.film, l.fold_acc, l.fold_n);
    return NGI_OK;
}

// ---- wavefront bdpt: host side -------------------------------------------------------------------------------------------
// One batch context = the buffers of NgiBdWave + a stream. Two contexts alternate: while the host waits for the strategy totals of
// batch b (the only host sync of a batch), the kernels of batch b + 1 are already queued on the other stream, and the tails of one
// batch's persistent trace launches are filled by the other's kernels.
struct BdwCtx {
    cudaStream_t stream = nullptr;
    NgiBdWave wv{};
    unsigned* ctl = nullptr;                  // 64 words: counts[32] | cursors[28] | strategy totals (u64) | - | shadow cursor
    unsigned* ctl_host = nullptr;             // pinned copy
    size_t item_cap = 0;
    void* sort_tmp = nullptr; size_t sort_bytes = 0;
    bool pending = false;
    cudaEvent_t done = nullptr;
    // NGI_RENDER_TIME_KERNELS: event pairs around the trace launches ([0] extend, [1] shadow) and around each phase
    bool timed = false;
    std::vector<cudaEvent_t> ev[3];
    double seconds[3] = {0, 0, 0};
    uint64_t timed_launches[3] = {0, 0, 0};
    cudaError_t mark(int which, cudaStream_t st) {
        if (!timed) return cudaSuccess;
        cudaEvent_t e;
        cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) return rc;
        ev[which].push_back(e);
        return cudaEventRecord(e, st);
    }
    void collect() {                           // after a stream synchronise: pairs -> seconds
        for (int w = 0; w < 3; w++) {
            for (size_t i = 0; i + 1 < ev[w].size(); i += 2) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, ev[w][i], ev[w][i + 1]) == cudaSuccess) { seconds[w] += ms * 1e-3; timed_launches[w]++; }
            }
            for (cudaEvent_t e : ev[w]) cudaEventDestroy(e);
            ev[w].clear();
        }
    }
    void release(cudaStream_t st) {
        collect();           // stream-ordered frees: safe once `st` is ordered after the batch streams (or on an error path)
        ngi_dfree(wv.V, st); ngi_dfree(wv.C, st); ngi_dfree(wv.nverts, st); ngi_dfree(wv.rays[0], st); ngi_dfree(wv.rays[1], st);
        ngi_dfree(wv.hits, st); ngi_dfree(wv.offsets, st); ngi_dfree(ctl, st);
        ngi_dfree(wv.items, st); ngi_dfree(wv.items_sorted, st); ngi_dfree(sort_tmp, st);
        if (ctl_host) cudaFreeHost(ctl_host);
        if (done) cudaEventDestroy(done);
        const double sec[3] = {seconds[0], seconds[1], seconds[2]};
        const uint64_t tl[3] = {timed_launches[0], timed_launches[1], timed_launches[2]};
        *this = BdwCtx();
        for (int w = 0; w < 3; w++) { seconds[w] = sec[w]; timed_launches[w] = tl[w]; }    // the totals survive the release
    }
};
// frees whatever a bdpt render holds when it leaves render_bdpt_wave, on success and on every error return
struct BdwRender {
    BdwCtx ctx[4];
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~BdwRender() {
        for (BdwCtx& c : ctx) { if (c.stream) cudaStreamSynchronize(c.stream); c.release(st); }     // (all released already on the success path)
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

int bdw_phase1(Scene* s, BdwCtx& c, const NgiBdParams& bp, const int cap, uint64_t& launches) {
    cudaStream_t st = c.stream;
    NgiBdWave& wv = c.wv;
    NGI_CUDA(c.mark(2, st));
    NGI_CUDA(cudaMemsetAsync(c.ctl, 0, 64 * sizeof(unsigned), st));
    const unsigned walkers = 2u * wv.batch;
    k_bdw_start<<<(walkers + kBlock - 1) / kBlock, kBlock, 0, st>>>(s->dev, bp, wv, cap);
    launches++;
    for (int step = 1; step < cap; step++) {
        // the queue shrinks by about 2x per step (Russian roulette); the kernels loop grid-stride over the device-side count
        const unsigned expect = std::max(walkers >> (step - 1), 1u);
        NGI_CUDA(c.mark(0, st));
        k_bdw_extend<<<s->tune.spread ? s->grid_bdw_extend : std::min(s->grid_bdw_extend, std::max(148u, (expect + kTraceBlock - 1) / kTraceBlock)), kTraceBlock, 0, st>>>(s->dev, wv, step, s->tune);
        NGI_CUDA(c.mark(0, st));
        k_bdw_step<<<std::min(kStageGrid, std::max(148u, (2u * expect + kBlock - 1) / kBlock)), kBlock, 0, st>>>(s->dev, bp, wv, step, cap);
        launches += 2;
    }
    k_bdw_count<<<(wv.batch + kBlock - 1) / kBlock, kBlock, 0, st>>>(s->dev, bp, wv, (unsigned long long*)(c.ctl + kBdwTotals));
    launches++;
    NGI_CUDA(c.mark(2, st));
    NGI_CUDA(cudaMemcpyAsync(c.ctl_host, c.ctl, 64 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    c.pending = true;
    return NGI_OK;
}

int bdw_alloc_items(BdwCtx& c, const size_t cap, cudaStream_t st) {
    NgiBdWave& wv = c.wv;
    ngi_dfree(wv.items, st); ngi_dfree(wv.items_sorted, st); ngi_dfree(c.sort_tmp, st);
    c.item_cap = cap;
    NGI_CUDA(ngi_dmalloc((void**)&wv.items, cap * sizeof(uint2), st));
    NGI_CUDA(ngi_dmalloc((void**)&wv.items_sorted, cap * sizeof(uint2), st));
    c.sort_bytes = 0;
    NGI_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, c.sort_bytes, (const unsigned long long*)wv.items, (unsigned long long*)wv.items_sorted, (int)cap, 32, 48, st));
    NGI_CUDA(ngi_dmalloc(&c.sort_tmp, c.sort_bytes, st));
    return NGI_OK;
}

int bdw_phase2(Scene* s, BdwCtx& c, const NgiBdParams& bp, const int cap, uint64_t& launches, uint64_t& extend_rays, uint64_t& shadow_rays) {
    if (!c.pending) return NGI_OK;
    cudaStream_t st = c.stream;
    NgiBdWave& wv = c.wv;
    NGI_CUDA(cudaStreamSynchronize(st));
    c.pending = false;
    c.collect();
    for (int k = 1; k < cap; k++) extend_rays += c.ctl_host[k];
    wv.n_ray_items = c.ctl_host[kBdwTotals];
    wv.n_rayless = c.ctl_host[kBdwTotals + 1];
    shadow_rays += wv.n_ray_items;
    const size_t need = (size_t)wv.n_ray_items + wv.n_rayless;
    int rc;
    if (need > c.item_cap && (rc = bdw_alloc_items(c, need + need / 4, st))) return rc;
    if (need == 0) return NGI_OK;
    NGI_CUDA(c.mark(2, st));
    k_bdw_expand<<<(wv.batch + kBlock - 1) / kBlock, kBlock, 0, st>>>(s->dev, bp, wv);
    launches++;
    if (wv.n_ray_items) {
        NGI_CUDA(c.mark(1, st));
        k_bdw_shadow<<<s->tune.spread ? s->grid_bdw_shadow : std::min(s->grid_bdw_shadow, std::max(148u, (wv.n_ray_items + kTraceBlock - 1) / kTraceBlock)), kTraceBlock, 0, st>>>(s->dev, wv, s->tune);
        NGI_CUDA(c.mark(1, st));
        launches++;
    }
    // order by y = n | s << 8 (bits 32..47 of the item read as one 64-bit key); dead items (0xFFFF) end up last
    NGI_CUDA(cub::DeviceRadixSort::SortKeys(c.sort_tmp, c.sort_bytes, (const unsigned long long*)wv.items, (unsigned long long*)wv.items_sorted, (int)need, 32, 48, st));
    launches += 3;
    k_bdw_contrib<<<(unsigned)std::min<size_t>(148u * 16u, (need + 127) / 128), 128, 0, st>>>(s->dev, bp, wv);
    k_film_fold<<<kFilmGrid, kBlock, 0, st>>>(bp.film, s->film_acc, s->film_acc_n);
    launches += 2;
    NGI_CUDA(c.mark(2, st));
    return NGI_OK;
}

int render_bdpt_wave(Scene* s, const NgiRenderParams* rp, const NgiBdParams& bp, cudaStream_t st, NgiRenderStats* stats) {
    const int cap = ngi_bd_vertex_cap(bp);
    int rc;
    if (!s->grid_bdw_extend) {
        if ((rc = persistent_grid(k_bdw_extend, &s->grid_bdw_extend))) return rc;
        if ((rc = persistent_grid(k_bdw_shadow, &s->grid_bdw_shadow))) return rc;
    }
    for (cudaStream_t& b : s->bd_streams) if (!b) NGI_CUDA(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    // batch size: every stage of a batch ends in a tail and the batches' ~50 launches are serially dependent, so throughput grows
    // with the batch (profiles/r01_sweep_bdpt_batch.txt: 136 / 241 / 305 / 323 Mpaths/s at 2^17 / 2^19 / 2^21 / 2^22 samples);
    // vertex + cache storage is cap x 2 B x 144 bytes (14.5 GB at 2^21 samples and 24 vertices) per batch in flight
    unsigned B = rp->wave_capacity;
    if (!B) {
        B = 1u << 21;
        while (B > (1u << 16) && (long long)B * 4 > rp->num_samples) B >>= 1;     // at least four batches, so that they overlap
    }
    if (const char* e = getenv("NGI_BDPT_BATCH")) B = (unsigned)std::max(1, atoi(e));
    B = (unsigned)std::min<long long>(B, rp->num_samples);
    int K = 2;
    if (const char* e = getenv("NGI_BDPT_STREAMS")) K = std::min(4, std::max(1, atoi(e)));
    const bool timed = (rp->flags & NGI_RENDER_TIME_KERNELS) != 0;      // per-kernel CUDA events: one batch at a time
    if (timed) K = 1;
    {   // fit the batches in flight into half of the memory that is free (or parked in the stream-ordered pool)
        size_t free_b = 0, total_b = 0;
        NGI_CUDA(cudaMemGetInfo(&free_b, &total_b));
        cudaMemPool_t pool;
        if (pool_alloc_enabled() && cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) {
            unsigned long long reserved = 0, used = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used) free_b += (size_t)(reserved - used);
        }
        const size_t per_sample = (size_t)cap * 2 * (sizeof(NgiBdVertex) + sizeof(NgiBdCache)) + 2 * (4 * sizeof(float4) + sizeof(float4) + sizeof(unsigned)) +
                                  sizeof(unsigned long long) + 3 * 16 * sizeof(uint2);
        while (B > (1u << 12) && (size_t)B * per_sample * (size_t)K > free_b / 2) B >>= 1;
    }
    const long long n_batches = (rp->num_samples + B - 1) / B;
    K = (int)std::min<long long>(K, n_batches);
    BdwRender R;
    R.st = st;
    BdwCtx* ctx = R.ctx;
    cudaEvent_t& ev0 = R.ev0;
    cudaEvent_t& ev1 = R.ev1;
    NGI_CUDA(cudaEventCreate(&ev0)); NGI_CUDA(cudaEventCreate(&ev1));
    for (int k = 0; k < K; k++) {
        BdwCtx& c = ctx[k];
        c.stream = s->bd_streams[k];
        c.timed = timed;
        NgiBdWave& wv = c.wv;
        wv.walkers = 2u * B;
        NGI_CUDA(ngi_dmalloc((void**)&wv.V, (size_t)cap * wv.walkers * sizeof(NgiBdVertex), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.C, (size_t)cap * wv.walkers * sizeof(NgiBdCache), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.nverts, (size_t)wv.walkers * sizeof(unsigned), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.rays[0], (size_t)wv.walkers * 2 * sizeof(float4), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.rays[1], (size_t)wv.walkers * 2 * sizeof(float4), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.hits, (size_t)wv.walkers * sizeof(float4), st));
        NGI_CUDA(ngi_dmalloc((void**)&wv.offsets, (size_t)B * sizeof(unsigned long long), st));
        NGI_CUDA(ngi_dmalloc((void**)&c.ctl, 64 * sizeof(unsigned), st));
        wv.counts = c.ctl; wv.cursors = c.ctl + 32;
        NGI_CUDA(cudaMallocHost((void**)&c.ctl_host, 64 * sizeof(unsigned)));
        if ((rc = bdw_alloc_items(c, (size_t)B * 16 + 1024, st))) return rc;     // expected <= 16 strategies per sample; grown on demand
        NGI_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    }
    NGI_CUDA(cudaEventRecord(ev0, st));          // fork: the film memset and the allocations precede every batch
    for (int k = 0; k < K; k++) NGI_CUDA(cudaStreamWaitEvent(ctx[k].stream, ev0, 0));
    uint64_t launches = 0, extend_rays = 0, shadow_rays = 0;
    for (long long b = 0; b < n_batches; b++) {
        BdwCtx& c = ctx[b % K];
        if ((rc = bdw_phase2(s, c, bp, cap, launches, extend_rays, shadow_rays))) return rc;      // batch b - K (no-op at the start)
        c.wv.first = (unsigned long long)(rp->sample_offset + b * (long long)B);
        c.wv.batch = (unsigned)std::min<long long>(B, rp->num_samples - b * (long long)B);
        if ((rc = bdw_phase1(s, c, bp, cap, launches))) return rc;
    }
    for (long long b = std::max<long long>(0, n_batches - K); b < n_batches; b++)
        if ((rc = bdw_phase2(s, ctx[b % K], bp, cap, launches, extend_rays, shadow_rays))) return rc;
    for (int k = 0; k < K; k++) {                // join
        NGI_CUDA(cudaEventRecord(ctx[k].done, ctx[k].stream));
        NGI_CUDA(cudaStreamWaitEvent(st, ctx[k].done, 0));
    }
    k_film_finish<<<kFilmGrid, kBlock, 0, st>>>(bp.film, s->film_acc, s->film_acc_n);
    launches += 2;                                // k_film_begin (render_impl) + k_film_finish
    NGI_CUDA(cudaEventRecord(ev1, st));
    for (int k = 0; k < K; k++) ctx[k].release(st);
    NGI_CUDA(cudaStreamSynchronize(st));
    NGI_CUDA(cudaGetLastError());
    float ms = 0;
    NGI_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (stats) {
        stats->paths = (uint64_t)rp->num_samples; stats->extend_rays = extend_rays; stats->shadow_rays = shadow_rays;
        stats->kernel_launches = launches; stats->wave_iterations = (uint64_t)n_batches; stats->gpu_seconds = ms * 1e-3;
        if (timed) {       // "logic" = everything of a batch that is not a trace launch (start / step / count / expand / sort / contrib)
            const BdwCtx& c = ctx[0];
            stats->extend_kernel_seconds = c.seconds[0]; stats->extend_launches = c.timed_launches[0];
            stats->shadow_kernel_seconds = c.seconds[1]; stats->shadow_launches = c.timed_launches[1];
            stats->logic_kernel_seconds = std::max(0.0, c.seconds[2] - c.seconds[0] - c.seconds[1]); stats->logic_launches = (uint64_t)n_batches;
            stats->trace_kernel_seconds = c.seconds[0] + c.seconds[1];
        }
    }
    return NGI_OK;
}

int render_impl(Scene* s, const NgiRenderParams* rp, float* film_dev, cudaStream_t st, NgiRenderStats* stats) {
    if (rp->struct_size != sizeof(NgiRenderParams)) return set_err(NGI_ERR_INVALID_ARGUMENT, "NgiRenderParams.struct_size mismatch (ABI)");
    if (rp->renderer < NGI_RENDERER_PT || rp->renderer > NGI_RENDERER_BDPT)
        return set_err(NGI_ERR_UNSUPPORTED, "renderer not supported by this build (pt, ptdirect, lt, ltdirect and bdpt are on the GPU path)");
    const bool has_shadow = rp->renderer == NGI_RENDERER_PTDIRECT || rp->renderer == NGI_RENDERER_LTDIRECT;
    if (rp->width <= 0 || rp->height <= 0 || rp->num_samples < 0 || rp->sample_offset < 0) return set_err(NGI_ERR_INVALID_ARGUMENT, "invalid width/height/num_samples");
    const size_t npx = (size_t)rp->width * rp->height;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (rp->num_samples == 0 || (rp->max_num_vertices != -1 && rp->max_num_vertices < 2)) {
        // MaxNumVertices <= 1: the loop exits before the first direction is sampled (src/nanogi.cpp:485)
        if (!rp->accumulate) NGI_CUDA(cudaMemsetAsync(film_dev, 0, npx * 3 * sizeof(float), st));
        NGI_CUDA(cudaStreamSynchronize(st));
        if (stats) stats->paths = (uint64_t)rp->num_samples;
        return NGI_OK;
    }
    // fp64 accumulation behind the fp32 film (k_film_fold): the accumulator is kept with the scene handle
    const size_t nfilm = npx * 3;
    if (s->film_acc_n != nfilm) {
        for (Lane& l : s->lanes) if (l.graph_exec) { cudaGraphExecDestroy(l.graph_exec); l.graph_exec = nullptr; }
        // from the stream-ordered pool like every other scene array: as a cudaMalloc / cudaFree pair per scene handle it cost the e2e
        // path (one handle per render) 0.2 - 0.36 s per step in scene_destroy (profiles/r02_sweep_l2_persist_and_film_pool.txt)
        if (s->film_acc) { ngi_dfree(s->film_acc, s->stream); s->film_acc = nullptr; s->film_acc_n = 0; }
        NGI_CUDA(ngi_dmalloc((void**)&s->film_acc, nfilm * sizeof(double), s->stream));
        NGI_CUDA(cudaStreamSynchronize(s->stream));          // usable from the caller's stream and the lanes' streams from here on
        s->film_acc_n = nfilm;
    }
    k_film_begin<<<kFilmGrid, kBlock, 0, st>>>(film_dev, s->film_acc, nfilm, rp->accumulate);
    if (rp->renderer == NGI_RENDERER_BDPT) {
        NgiBdParams bp;
        bp.film = film_dev; bp.width = rp->width; bp.height = rp->height; bp.max_verts = rp->max_num_vertices;
        bp.seed_lo = (unsigned)rp->seed; bp.seed_hi = (unsigned)(rp->seed >> 32);
        bp.film_scale = rp->film_norm_samples > 0 ? (float)((double)npx / (double)rp->film_norm_samples) : 1.0f;
        if (!(rp->flags & NGI_RENDER_BDPT_PER_THREAD)) return render_bdpt_wave(s, rp, bp, st, stats);   // the product path (ngi_bdpt_wave.h)
        // cross-check: one thread per sample (ngi_bdpt.h), a single launch over the shard
        unsigned long long* d_cnt = nullptr;
        NGI_CUDA(ngi_dmalloc((void**)&d_cnt, 2 * sizeof(unsigned long long), st));
        NGI_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
        cudaEvent_t e0, e1;
        NGI_CUDA(cudaEventCreate(&e0)); NGI_CUDA(cudaEventCreate(&e1));
        NGI_CUDA(cudaEventRecord(e0, st));
        const unsigned grid = (unsigned)std::min<unsigned long long>(((unsigned long long)rp->num_samples + 127ull) / 128ull, 148ull * 64ull);
        k_bdpt<<<grid, 128, 0, st>>>(s->dev, bp, (unsigned long long)rp->sample_offset, (unsigned long long)rp->num_samples, d_cnt);
        k_film_finish<<<kFilmGrid, kBlock, 0, st>>>(film_dev, s->film_acc, nfilm);
        NGI_CUDA(cudaEventRecord(e1, st));
        unsigned long long h_cnt[2] = {0, 0};
        NGI_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        NGI_CUDA(cudaStreamSynchronize(st));
        NGI_CUDA(cudaGetLastError());
        float ms = 0;
        NGI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        ngi_dfree(d_cnt, st);
        if (stats) {
            stats->paths = (uint64_t)rp->num_samples; stats->extend_rays = h_cnt[0]; stats->shadow_rays = h_cnt[1];
            stats->kernel_launches = 3; stats->wave_iterations = 1; stats->gpu_seconds = ms * 1e-3;
        }
        return NGI_OK;
    }
    const bool per_ray = (rp->flags & NGI_RENDER_PER_RAY_TRACE) != 0;
    const bool timed = (rp->flags & NGI_RENDER_TIME_KERNELS) != 0 || per_ray;
    // default slot count per lane: 2 Mi, 4 Mi for long renders (profiles/r01_sweep_wave.txt: larger waves amortise the
    // fixed tail of every launch; small renders prefer the shorter ramp-up and drain)
    // (8 Mi from 2^30 samples on: C3 +2.2 %, C2 +0.6 % over 4 Mi, profiles/r01_sweep_wave.txt)
    // (4 Mi already from 2^27 samples: the 8-GPU shard of C3, 265 M samples, is just below 2^28)
    unsigned P = rp->wave_capacity ? rp->wave_capacity : (rp->num_samples >= (1ll << 30) ? (1u << 23) : rp->num_samples >= (1ll << 27) ? (1u << 22) : (1u << 21));
    P = std::max(P, 1024u);
    if ((unsigned long long)rp->num_samples < P) P = std::max(1024u, (unsigned)((rp->num_samples + 255) / 256 * 256));
    // lanes: concurrent pipelines on disjoint sample ranges; only worth it when every lane gets several waves of work
    int K = timed ? 1 : s->num_lanes;
    while (K > 1 && (unsigned long long)rp->num_samples < 4ull * P * (unsigned)K) K--;
    if ((int)s->lanes.size() < K) s->lanes.resize(K);
    int rc;

    struct EventPair { cudaEvent_t a = nullptr, b = nullptr; ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evp;
    cudaEvent_t& ev0 = evp.a;
    cudaEvent_t& ev1 = evp.b;
    NGI_CUDA(cudaEventCreate(&ev0));
    NGI_CUDA(cudaEventCreate(&ev1));
    NGI_CUDA(cudaEventRecord(ev0, st));          // also the fork point: the film memset above precedes every lane

    const int kItersPerBatch = 8;
    const int kernels_per_iter = has_shadow ? 6 : 5;   // iter_begin, classify, surface, eye, extend (, shadow)
    for (int k = 0; k < K; k++) {
        Lane& l = s->lanes[k];
        if ((rc = ensure_lane(s, l, P))) return rc;
        NgiWaveParams& wp = l.wp;
        memset(&wp, 0, sizeof(wp));
        carve_wave(l, wp);
        wp.film = film_dev;
        wp.renderer = rp->renderer; wp.max_verts = rp->max_num_vertices; wp.width = rp->width; wp.height = rp->height;
        const long long lo = rp->num_samples / K * k + std::min<long long>(k, rp->num_samples % K);
        const long long cnt = rp->num_samples / K + (k < rp->num_samples % K ? 1 : 0);
        wp.sample_end = (unsigned long long)(rp->sample_offset + lo + cnt);
        wp.seed_lo = (unsigned)rp->seed; wp.seed_hi = (unsigned)(rp->seed >> 32);
        wp.film_scale = rp->film_norm_samples > 0 ? (float)((double)npx / (double)rp->film_norm_samples) : 1.0f;
        l.fold_acc = k == 0 ? s->film_acc : nullptr;          // one fold per iteration of lane 0 covers every lane's splats
        l.fold_n = nfilm;
        NgiRenderCounters init;
        memset(&init, 0, sizeof(init));
        init.next_sample = (unsigned long long)(rp->sample_offset + lo);
        *l.counters_host = init;
        NGI_CUDA(cudaStreamWaitEvent(l.stream, ev0, 0));
        NGI_CUDA(cudaMemcpyAsync(l.counters, l.counters_host, sizeof(init), cudaMemcpyHostToDevice, l.stream));
        NGI_CUDA(cudaMemsetAsync(wp.sb, 0, (size_t)P * 32, l.stream));         // every slot starts idle (dir_info.w = 0)
        if (!timed && (!l.graph_exec || !same_wp(l.graph_wp, wp) || l.graph_iters != kItersPerBatch || l.graph_acc != l.fold_acc)) {
            // the batch of kItersPerBatch iterations is captured once into a CUDA graph and replayed
            if (l.graph_exec) { cudaGraphExecDestroy(l.graph_exec); l.graph_exec = nullptr; }
            cudaGraph_t graph;
            NGI_CUDA(cudaStreamSynchronize(l.stream));
            NGI_CUDA(cudaStreamBeginCapture(l.stream, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < kItersPerBatch; i++) { size_t dummy = 0; rc = launch_iteration(s, l, false, dummy); if (rc) { cudaStreamEndCapture(l.stream, &graph); return rc; } }
            NGI_CUDA(cudaStreamEndCapture(l.stream, &graph));
            NGI_CUDA(cudaGraphInstantiate(&l.graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            l.graph_wp = wp; l.graph_iters = kItersPerBatch; l.graph_acc = l.fold_acc;
        }
        l.running = true;
    }

    size_t ev_used = 0;
    uint64_t launches = 0;
    double logic_ms = 0.0, extend_ms = 0.0, shadow_ms = 0.0;
    uint64_t timed_iters = 0;
    // expected number of iterations ~ (rays per path) * N / P; every lane's counters are polled once per batch
    int running = K;
    while (running > 0) {
        for (int k = 0; k < K; k++) {
            Lane& l = s->lanes[k];
            if (!l.running) continue;
            if (timed) {
                for (int i = 0; i < kItersPerBatch; i++) { rc = launch_iteration(s, l, true, ev_used, per_ray); if (rc) return rc; }
            } else {
                NGI_CUDA(cudaGraphLaunch(l.graph_exec, l.stream));
            }
            launches += (uint64_t)kItersPerBatch * (kernels_per_iter + (l.fold_acc ? 1 : 0));
            NGI_CUDA(cudaMemcpyAsync(l.counters_host, l.counters, sizeof(NgiRenderCounters), cudaMemcpyDeviceToHost, l.stream));
        }
        for (int k = 0; k < K; k++) {
            Lane& l = s->lanes[k];
            if (!l.running) continue;
            NGI_CUDA(cudaStreamSynchronize(l.stream));
            if (timed) {
                for (size_t i = 0; i + 3 < ev_used; i += 4) {
                    float ms = 0;
                    NGI_CUDA(cudaEventElapsedTime(&ms, l.events[i], l.events[i + 1])); logic_ms += ms;
                    NGI_CUDA(cudaEventElapsedTime(&ms, l.events[i + 1], l.events[i + 2])); extend_ms += ms;
                    NGI_CUDA(cudaEventElapsedTime(&ms, l.events[i + 2], l.events[i + 3])); shadow_ms += ms;
                    timed_iters++;
                }
                ev_used = 0;
            }
            const NgiRenderCounters& c = *l.counters_host;
            if (c.next_sample >= l.wp.sample_end && c.iter[0] == 0 && c.iter[1] == 0) {
                l.running = false; running--;
                NGI_CUDA(cudaEventRecord(l.ev_done, l.stream));
                NGI_CUDA(cudaStreamWaitEvent(st, l.ev_done, 0));     // join
            }
        }
    }
    if (const char* path = getenv("NGI_ITER_LOG")) {        // entry i = rays of iteration i - 1 (the counters roll at the start of an iteration)
        if (FILE* f = fopen(path, "a")) {
            std::vector<unsigned> h(2 * kIterLogCap);
            for (int k = 0; k < K; k++) {
                Lane& l = s->lanes[k];
                if (!l.iter_log) continue;
                cudaMemcpy(h.data(), l.iter_log, h.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
                const unsigned long long its = std::min<unsigned long long>(l.counters_host->iterations, kIterLogCap);
                fprintf(f, "# render: renderer %d samples %lld slots %u lane %d of %d iterations %llu\n", rp->renderer, (long long)rp->num_samples, P, k, K, its);
                for (unsigned long long i = 1; i < its; i++) fprintf(f, "%d %llu %u %u\n", k, i - 1, h[2 * i], h[2 * i + 1]);
            }
            fclose(f);
        }
    }
    k_film_finish<<<kFilmGrid, kBlock, 0, st>>>(film_dev, s->film_acc, nfilm);      // every lane has joined `st`
    launches += 2;                                                                  // k_film_begin + k_film_finish
    NGI_CUDA(cudaEventRecord(ev1, st));
    NGI_CUDA(cudaStreamSynchronize(st));
    NGI_CUDA(cudaGetLastError());
    float ms = 0;
    NGI_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (stats) {
        stats->paths = (uint64_t)rp->num_samples;
        for (int k = 0; k < K; k++) {
            const NgiRenderCounters& c = *s->lanes[k].counters_host;
            stats->extend_rays += c.total_extend + c.iter[1];
            stats->shadow_rays += c.total_shadow + c.iter[0];
            stats->wave_iterations += c.iterations;
        }
        stats->kernel_launches = launches;
        stats->gpu_seconds = ms * 1e-3;
        stats->trace_kernel_seconds = (extend_ms + shadow_ms) * 1e-3;
        stats->logic_kernel_seconds = logic_ms * 1e-3;
        stats->extend_kernel_seconds = extend_ms * 1e-3;
        stats->shadow_kernel_seconds = has_shadow ? shadow_ms * 1e-3 : 0.0;
        stats->logic_launches = timed_iters;
        stats->extend_launches = timed_iters;
        stats->shadow_launches = has_shadow ? timed_iters : 0;
    }
    return NGI_OK;
}

template <int ACCEL>
void launch_trace(Scene* s, const NgiRay* rays, size_t n, NgiHit* hits, int any_hit, cudaStream_t st) {
    if (any_hit) k_trace<ACCEL, true><<<grid_for(n), kBlock, 0, st>>>(s->dev, rays, n, hits);
    else k_trace<ACCEL, false><<<grid_for(n), kBlock, 0, st>>>(s->dev, rays, n, hits);
}

template <bool ANY_HIT>
void launch_trace8(Scene* s, const NgiRay* rays, size_t n, NgiHit* hits, cudaStream_t st) {
    QuerySource<ANY_HIT> src;
    src.rays = reinterpret_cast<const float4*>(rays); src.hits = reinterpret_cast<float4*>(hits); src.n = (unsigned)n; src.cur = s->trace_cursor;
    const unsigned need = grid_for(n, kTraceBlock);
    const unsigned grid = std::min(s->grid_trace[ANY_HIT ? 1 : 0], need);
    k_trace8<ANY_HIT><<<grid, kTraceBlock, 0, st>>>(s->dev, src, s->tune);
}

int trace_impl(Scene* s, const NgiRay* rays_dev, size_t n, NgiHit* hits_dev, int any_hit, int accel, double* seconds) {
    if (accel < 0 || accel > 2) return set_err(NGI_ERR_INVALID_ARGUMENT, "accel must be 0 (BVH8), 1 (BVH2) or 2 (brute force)");
    if (accel == 1 && s->bvh2_depth >= 64u)      // ngi_trace_bvh2 keeps one stack entry per level in int stack[64]
        return set_err(NGI_ERR_UNSUPPORTED, "binary BVH of height " + std::to_string(s->bvh2_depth) + " exceeds the cross-check traversal's stack (accel = 1)");
    cudaStream_t st = s->stream;
    struct EventPair { cudaEvent_t a = nullptr, b = nullptr; ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evp;
    cudaEvent_t& ev0 = evp.a;
    cudaEvent_t& ev1 = evp.b;
    NGI_CUDA(cudaEventCreate(&ev0));
    NGI_CUDA(cudaEventCreate(&ev1));
    if (n >= 0xFFFFFF00ull) return set_err(NGI_ERR_INVALID_ARGUMENT, "at most 2^32 - 256 rays per call");
    if (accel == 0) NGI_CUDA(cudaMemsetAsync(s->trace_cursor, 0, sizeof(unsigned), st));
    NGI_CUDA(cudaEventRecord(ev0, st));
    if (n) {
        if (accel == 0) { if (any_hit) launch_trace8<true>(s, rays_dev, n, hits_dev, st); else launch_trace8<false>(s, rays_dev, n, hits_dev, st); }
        else if (accel == 1) launch_trace<1>(s, rays_dev, n, hits_dev, any_hit, st);
        else launch_trace<2>(s, rays_dev, n, hits_dev, any_hit, st);
    }
    NGI_CUDA(cudaEventRecord(ev1, st));
    NGI_CUDA(cudaStreamSynchronize(st));
    NGI_CUDA(cudaGetLastError());
    float ms = 0;
    NGI_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (seconds) *seconds = ms * 1e-3;
    return NGI_OK;
}

// ================================================================================================
// multi-GPU: samples sharded by index, one NCCL reduce of the per-GPU films (ngi_comm.h)
// ================================================================================================
#define NGI_NCCL(api, call)                                                                             \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) return set_err(NGI_ERR_NCCL, std::string(#call) + ": " + (api)->GetErrorString(r__)); \
    } while (0)

int nccl_or_error(NgiNccl** out) {
    NgiNccl* api = ngi_nccl();
    if (!api->handle) return set_err(NGI_ERR_NCCL, api->error);
    *out = api;
    return NGI_OK;
}

void log_nccl_once(NgiNccl* api, const char* what, int ranks) {
    if (const char* q = getenv("NGI_QUIET")) if (atoi(q)) return;
    int v = 0;
    api->GetVersion(&v);
    fprintf(stderr, "[nanogi_gpu] NCCL %d.%d.%d: %s over %d GPU(s)\n", v / 10000, (v / 100) % 100, v % 100, what, ranks);
}

// one rank of a communicator that spans processes (one process per GPU)
struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

// every device of one process: scene built on the first, cloned to the others; per-device film; one communicator per device
struct Group {
    std::vector<int> devices;
    std::vector<Scene*> scenes;
    std::vector<ncclComm_t> comms;
    std::vector<float*> films;
    size_t film_floats = 0;
    ~Group() {
        NgiNccl* api = ngi_nccl();
        for (size_t g = 0; g < devices.size(); g++) {
            cudaSetDevice(devices[g]);
            if (g < films.size() && films[g]) cudaFree(films[g]);
            if (g < comms.size() && comms[g] && api->handle) api->CommDestroy(comms[g]);
            if (g < scenes.size()) delete scenes[g];
        }
    }
};

// A second scene handle on `device` with the arrays of `src` (built on another device): same sizes, contents broadcast by the caller
int clone_scene_layout(const Scene* src, int device, Scene** out) {
    NGI_CUDA(cudaSetDevice(device));
    Scene* c = new Scene;
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return set_err(NGI_ERR_CUDA, cudaGetErrorString(e)); }
    for (size_t i = 0; i < src->allocs.size(); i++) {
        void* p = nullptr;
        e = ngi_dmalloc(&p, src->alloc_bytes[i], c->stream);
        if (e != cudaSuccess) { delete c; return set_err(e == cudaErrorMemoryAllocation ? NGI_ERR_OUT_OF_MEMORY : NGI_ERR_CUDA, cudaGetErrorString(e)); }
        c->allocs.push_back(p); c->alloc_bytes.push_back(src->alloc_bytes[i]);
    }
    auto map = [&](const void* p) -> void* {
        if (!p) return nullptr;
        for (size_t i = 0; i < src->allocs.size(); i++) if (src->allocs[i] == p) return c->allocs[i];
        return nullptr;
    };
    const NgiDevScene& a = src->dev;
    NgiDevScene& d = c->dev;
    d = a;
    d.nodes8 = (const uint4*)map(a.nodes8); d.tris8 = (const float4*)map(a.tris8); d.nodes2 = (const float4*)map(a.nodes2); d.tris2 = (const float4*)map(a.tris2);
    d.shade_tris = (const float4*)map(a.shade_tris); d.prims = (const NgiDevPrim*)map(a.prims); d.light_prims = (const unsigned*)map(a.light_prims);
    d.cdf = (const float*)map(a.cdf); d.shade_uv = (const float*)map(a.shade_uv); d.textures = (const NgiDevTex*)map(a.textures); d.tex_data = (const float*)map(a.tex_data);
    c->info = src->info;
    *out = c;
    return NGI_OK;
}

int group_create(const NgiSceneDesc* desc, const int* devices, int n, Group** out) {
    NgiNccl* api = nullptr;
    int rc;
    if (n > 1 && (rc = nccl_or_error(&api))) return rc;
    std::unique_ptr<Group> g(new Group);
    g->devices.assign(devices, devices + n);
    // the scene is uploaded and its BVH built ONCE, on the first device
    NGI_CUDA(cudaSetDevice(devices[0]));
    {
        Scene* s0 = new Scene;
        s0->device = devices[0];
        cudaError_t e = cudaStreamCreateWithFlags(&s0->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete s0; return set_err(NGI_ERR_CUDA, cudaGetErrorString(e)); }
        rc = build_scene(s0, desc);
        if (rc != NGI_OK) { delete s0; return rc; }
        g->scenes.push_back(s0);
    }
    if (n > 1) {
        g->comms.assign(n, nullptr);
        NGI_NCCL(api, api->CommInitAll(g->comms.data(), n, devices));
        log_nccl_once(api, "ncclCommInitAll; scene broadcast from the first device + one film reduce per render", n);
        for (int k = 1; k < n; k++) {
            Scene* c = nullptr;
            if ((rc = clone_scene_layout(g->scenes[0], devices[k], &c))) return rc;
            g->scenes.push_back(c);
        }
        // every array of the built scene (BVH8 nodes, traversal + shading triangles, primitives, CDFs, textures) goes over NVLink
        const Scene* s0 = g->scenes[0];
        NGI_CUDA(cudaSetDevice(devices[0]));
        NGI_CUDA(cudaStreamSynchronize(s0->stream));
        NGI_NCCL(api, api->GroupStart());
        for (size_t i = 0; i < s0->allocs.size(); i++)
            for (int k = 0; k < n; k++)
                NGI_NCCL(api, api->Broadcast(s0->allocs[i], g->scenes[k]->allocs[i], s0->alloc_bytes[i], ncclChar, 0, g->comms[k], g->scenes[k]->stream));
        NGI_NCCL(api, api->GroupEnd());
        for (int k = 0; k < n; k++) {
            NGI_CUDA(cudaSetDevice(devices[k]));
            NGI_CUDA(cudaStreamSynchronize(g->scenes[k]->stream));
            if (k > 0 && ((rc = bvh_tables_init()) || (rc = init_trace_launch(g->scenes[k])))) return rc;
        }
    }
    *out = g.release();
    return NGI_OK;
}

int group_render(Group* g, const NgiRenderParams* rp, float* film_host, NgiRenderStats* stats) {
    const int n = (int)g->devices.size();
    if (rp->width <= 0 || rp->height <= 0 || rp->num_samples < 0) return set_err(NGI_ERR_INVALID_ARGUMENT, "invalid width/height/num_samples");
    const size_t floats = (size_t)rp->width * rp->height * 3;
    if (g->film_floats != floats) {
        for (int k = 0; k < (int)g->films.size(); k++) { cudaSetDevice(g->devices[k]); cudaFree(g->films[k]); }
        g->films.assign(n, nullptr);
        g->film_floats = 0;
        for (int k = 0; k < n; k++) { NGI_CUDA(cudaSetDevice(g->devices[k])); NGI_CUDA(cudaMalloc((void**)&g->films[k], floats * sizeof(float))); }
        g->film_floats = floats;
    }
    // samples sharded by index (ngi_gpu_shard_range); every device renders its shard concurrently, one host thread each
    std::vector<NgiRenderStats> st(n);
    std::vector<int> rcs(n, NGI_OK);
    std::vector<std::string> errs(n);
    auto work = [&](int k) {
        cudaSetDevice(g->devices[k]);
        NgiRenderParams p = *rp;
        int64_t off = 0, cnt = 0;
        ngi_gpu_shard_range(rp->num_samples, k, n, &off, &cnt);
        p.sample_offset = rp->sample_offset + off; p.num_samples = cnt; p.accumulate = 0;
        rcs[k] = render_impl(g->scenes[k], &p, g->films[k], g->scenes[k]->stream, &st[k]);
        if (rcs[k] != NGI_OK) errs[k] = g_err;
    };
    {
        std::vector<std::thread> th;
        for (int k = 1; k < n; k++) th.emplace_back(work, k);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int k = 0; k < n; k++) if (rcs[k] != NGI_OK) return set_err(rcs[k], "GPU " + std::to_string(g->devices[k]) + ": " + errs[k]);
    // the gather of src/nanogi.cpp:429-437: one reduce(SUM) onto the first device
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    NGI_CUDA(cudaSetDevice(g->devices[0]));
    NGI_CUDA(cudaEventCreate(&e0)); NGI_CUDA(cudaEventCreate(&e1));
    NGI_CUDA(cudaEventRecord(e0, g->scenes[0]->stream));
    if (n > 1) {
        NgiNccl* api = nullptr;
        int rc;
        if ((rc = nccl_or_error(&api))) return rc;
        NGI_NCCL(api, api->GroupStart());
        for (int k = 0; k < n; k++)
            NGI_NCCL(api, api->Reduce(g->films[k], g->films[k], floats, ncclFloat, ncclSum, 0, g->comms[k], g->scenes[k]->stream));
        NGI_NCCL(api, api->GroupEnd());
    }
    NGI_CUDA(cudaEventRecord(e1, g->scenes[0]->stream));
    NGI_CUDA(cudaMemcpyAsync(film_host, g->films[0], floats * sizeof(float), cudaMemcpyDeviceToHost, g->scenes[0]->stream));
    for (int k = 0; k < n; k++) { NGI_CUDA(cudaSetDevice(g->devices[k])); NGI_CUDA(cudaStreamSynchronize(g->scenes[k]->stream)); }
    NGI_CUDA(cudaSetDevice(g->devices[0]));
    float reduce_ms = 0;
    NGI_CUDA(cudaEventElapsedTime(&reduce_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int k = 0; k < n; k++) {
            stats->paths += st[k].paths; stats->extend_rays += st[k].extend_rays; stats->shadow_rays += st[k].shadow_rays;
            stats->wave_iterations += st[k].wave_iterations; stats->kernel_launches += st[k].kernel_launches;
            stats->gpu_seconds = std::max(stats->gpu_seconds, st[k].gpu_seconds);
            stats->trace_kernel_seconds += st[k].trace_kernel_seconds; stats->logic_kernel_seconds += st[k].logic_kernel_seconds;
            stats->extend_kernel_seconds += st[k].extend_kernel_seconds; stats->shadow_kernel_seconds += st[k].shadow_kernel_seconds;
            stats->logic_launches += st[k].logic_launches; stats->extend_launches += st[k].extend_launches; stats->shadow_launches += st[k].shadow_launches;
        }
        stats->reduce_seconds = reduce_ms * 1e-3;
    }
    return NGI_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int ngi_gpu_abi_version(void) { return 2; }   // 2: NgiRenderStats.reduce_seconds, multi-GPU entry points

const char* ngi_gpu_last_error(void) { return g_err.c_str(); }

int ngi_gpu_device_count(void) {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return set_err(NGI_ERR_NO_DEVICE, "no CUDA device available (this module has no CPU fallback)"); }
    return n;
}

int ngi_gpu_scene_create(const NgiSceneDesc* desc, int device, void** out_scene) {
    if (!desc || !out_scene) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    *out_scene = nullptr;
    const int nd = ngi_gpu_device_count();
    if (nd < 0) return nd;
    if (device < 0 || device >= nd) return set_err(NGI_ERR_INVALID_ARGUMENT, "device index out of range");
    NGI_CUDA(cudaSetDevice(device));
    Scene* s = new Scene;
    s->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; return set_err(NGI_ERR_CUDA, cudaGetErrorString(e)); }
    const int rc = build_scene(s, desc);
    if (rc != NGI_OK) { delete s; return rc; }
    *out_scene = s;
    return NGI_OK;
}

int ngi_gpu_scene_info(void* scene, NgiSceneInfo* out) {
    if (!scene || !out) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    *out = ((Scene*)scene)->info;
    return NGI_OK;
}

void ngi_gpu_scene_destroy(void* scene) { delete (Scene*)scene; }

int ngi_gpu_render_device(void* scene, const NgiRenderParams* params, void* film_rgb_device, void* cuda_stream, NgiRenderStats* out_stats) {
    if (!scene || !params || !film_rgb_device) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Scene* s = (Scene*)scene;
    NGI_CUDA(cudaSetDevice(s->device));
    return render_impl(s, params, (float*)film_rgb_device, cuda_stream ? (cudaStream_t)cuda_stream : s->stream, out_stats);
}

int ngi_gpu_render(void* scene, const NgiRenderParams* params, float* film_rgb_host, NgiRenderStats* out_stats) {
    if (!scene || !params || !film_rgb_host) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Scene* s = (Scene*)scene;
    NGI_CUDA(cudaSetDevice(s->device));
    if (params->width <= 0 || params->height <= 0) return set_err(NGI_ERR_INVALID_ARGUMENT, "invalid width/height");
    const size_t bytes = (size_t)params->width * params->height * 3 * sizeof(float);
    float* d_film = nullptr;
    NGI_CUDA(ngi_dmalloc((void**)&d_film, bytes, s->stream));
    NgiRenderParams p = *params;
    p.accumulate = 0;
    int rc = render_impl(s, &p, d_film, s->stream, out_stats);
    if (rc == NGI_OK) {
        cudaError_t e = cudaMemcpyAsync(film_rgb_host, d_film, bytes, cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = set_err(NGI_ERR_CUDA, cudaGetErrorString(e));
    }
    ngi_dfree(d_film, s->stream);
    return rc;
}

int ngi_gpu_trace_device(void* scene, const void* rays_device, uint64_t n, void* hits_device, int any_hit, int accel, double* out_seconds) {
    if (!scene || (n && (!rays_device || !hits_device))) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Scene* s = (Scene*)scene;
    NGI_CUDA(cudaSetDevice(s->device));
    return trace_impl(s, (const NgiRay*)rays_device, (size_t)n, (NgiHit*)hits_device, any_hit, accel, out_seconds);
}

int ngi_gpu_trace(void* scene, const NgiRay* rays_host, uint64_t n, NgiHit* hits_host, int any_hit, int accel) {
    if (!scene || (n && (!rays_host || !hits_host))) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Scene* s = (Scene*)scene;
    NGI_CUDA(cudaSetDevice(s->device));
    if (n == 0) return NGI_OK;
    NgiRay* d_rays = nullptr; NgiHit* d_hits = nullptr;
    NGI_CUDA(cudaMalloc((void**)&d_rays, n * sizeof(NgiRay)));
    cudaError_t e = cudaMalloc((void**)&d_hits, n * sizeof(NgiHit));
    if (e != cudaSuccess) { cudaFree(d_rays); return set_err(NGI_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    int rc = NGI_OK;
    e = cudaMemcpyAsync(d_rays, rays_host, n * sizeof(NgiRay), cudaMemcpyHostToDevice, s->stream);
    if (e != cudaSuccess) rc = set_err(NGI_ERR_CUDA, cudaGetErrorString(e));
    if (rc == NGI_OK) rc = trace_impl(s, d_rays, (size_t)n, d_hits, any_hit, accel, nullptr);
    if (rc == NGI_OK) {
        e = cudaMemcpy(hits_host, d_hits, n * sizeof(NgiHit), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_err(NGI_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_rays); cudaFree(d_hits);
    return rc;
}

int ngi_gpu_eval_bsdf(void* scene, const float* queries_host, const float* wo_in_host, uint64_t n, int force_degenerated, float* out_host) {
    if (!scene || (n && (!queries_host || !wo_in_host || !out_host))) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Scene* s = (Scene*)scene;
    NGI_CUDA(cudaSetDevice(s->device));
    if (n == 0) return NGI_OK;
    float *d_q = nullptr, *d_wo = nullptr, *d_out = nullptr;
    NGI_CUDA(cudaMalloc((void**)&d_q, n * 16 * sizeof(float)));
    NGI_CUDA(cudaMalloc((void**)&d_wo, n * 3 * sizeof(float)));
    NGI_CUDA(cudaMalloc((void**)&d_out, n * 8 * sizeof(float)));
    NGI_CUDA(cudaMemcpyAsync(d_q, queries_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    NGI_CUDA(cudaMemcpyAsync(d_wo, wo_in_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    k_eval_bsdf<<<grid_for(n), kBlock, 0, s->stream>>>(s->dev, d_q, d_wo, (size_t)n, force_degenerated, d_out);
    NGI_CUDA(cudaMemcpyAsync(out_host, d_out, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    NGI_CUDA(cudaStreamSynchronize(s->stream));
    NGI_CUDA(cudaGetLastError());
    cudaFree(d_q); cudaFree(d_wo); cudaFree(d_out);
    return NGI_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------------------------------
void ngi_gpu_shard_range(int64_t num_samples, int rank, int world_size, int64_t* out_offset, int64_t* out_count) {
    // rank r of G takes the contiguous index range [r N / G, (r + 1) N / G): sizes differ by at most one (128-bit product: N can
    // exceed 2^60 / G in principle)
    const __int128 n = num_samples;
    const int64_t lo = (int64_t)(n * rank / world_size), hi = (int64_t)(n * (rank + 1) / world_size);
    if (out_offset) *out_offset = lo;
    if (out_count) *out_count = hi - lo;
}

int ngi_gpu_comm_get_id(NgiCommId* out_id) {
    if (!out_id) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(NgiCommId) >= sizeof(ncclUniqueId), "NgiCommId too small");
    NgiNccl* api = nullptr;
    int rc;
    if ((rc = nccl_or_error(&api))) return rc;
    ncclUniqueId id;
    NGI_NCCL(api, api->GetUniqueId(&id));
    memset(out_id, 0, sizeof(*out_id));
    memcpy(out_id, &id, sizeof(id));
    return NGI_OK;
}

int ngi_gpu_comm_create(const NgiCommId* id, int rank, int world_size, int device, void** out_comm) {
    if (!id || !out_comm || world_size < 1 || rank < 0 || rank >= world_size) return set_err(NGI_ERR_INVALID_ARGUMENT, "bad id / rank / world_size");
    *out_comm = nullptr;
    NgiNccl* api = nullptr;
    int rc;
    if ((rc = nccl_or_error(&api))) return rc;
    const int nd = ngi_gpu_device_count();
    if (nd < 0) return nd;
    if (device < 0 || device >= nd) return set_err(NGI_ERR_INVALID_ARGUMENT, "device index out of range");
    NGI_CUDA(cudaSetDevice(device));
    ncclUniqueId nid;
    memcpy(&nid, id, sizeof(nid));
    Comm* c = new Comm;
    c->rank = rank; c->world = world_size; c->device = device;
    ncclResult_t r = api->CommInitRank(&c->comm, world_size, nid, rank);
    if (r != ncclSuccess) { delete c; return set_err(NGI_ERR_NCCL, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
    if (rank == 0) log_nccl_once(api, "ncclCommInitRank (one process per GPU); one film reduce per render", world_size);
    *out_comm = c;
    return NGI_OK;
}

int ngi_gpu_comm_reduce_film(void* comm, void* film_rgb_device, uint64_t num_floats, int root, void* cuda_stream) {
    if (!comm || !film_rgb_device) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    Comm* c = (Comm*)comm;
    NgiNccl* api = nullptr;
    int rc;
    if ((rc = nccl_or_error(&api))) return rc;
    NGI_CUDA(cudaSetDevice(c->device));
    NGI_NCCL(api, api->Reduce(film_rgb_device, film_rgb_device, (size_t)num_floats, ncclFloat, ncclSum, root, c->comm, (cudaStream_t)cuda_stream));
    // NULL = the legacy default stream, which does NOT order itself against the non-blocking streams this module renders on: return
    // only when the reduce is done, so that a following render cannot reset a film the reduce is still reading / writing
    if (!cuda_stream) NGI_CUDA(cudaStreamSynchronize(nullptr));
    return NGI_OK;
}

void ngi_gpu_comm_destroy(void* comm) {
    Comm* c = (Comm*)comm;
    if (!c) return;
    NgiNccl* api = ngi_nccl();
    if (c->comm && api->handle) { cudaSetDevice(c->device); api->CommDestroy(c->comm); }
    delete c;
}

int ngi_gpu_group_create(const NgiSceneDesc* desc, const int* devices, int num_devices, void** out_group) {
    if (!desc || !out_group || num_devices < 1) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument / no devices");
    *out_group = nullptr;
    const int nd = ngi_gpu_device_count();
    if (nd < 0) return nd;
    std::vector<int> dv(num_devices);
    for (int k = 0; k < num_devices; k++) {
        dv[k] = devices ? devices[k] : k;
        if (dv[k] < 0 || dv[k] >= nd) return set_err(NGI_ERR_INVALID_ARGUMENT, "device index out of range");
        for (int j = 0; j < k; j++) if (dv[j] == dv[k]) return set_err(NGI_ERR_INVALID_ARGUMENT, "duplicate device index");
    }
    Group* g = nullptr;
    const int rc = group_create(desc, dv.data(), num_devices, &g);
    if (rc != NGI_OK) return rc;
    *out_group = g;
    return NGI_OK;
}

int ngi_gpu_group_render(void* group, const NgiRenderParams* params, float* film_rgb_host, NgiRenderStats* out_stats) {
    if (!group || !params || !film_rgb_host) return set_err(NGI_ERR_INVALID_ARGUMENT, "null argument");
    return group_render((Group*)group, params, film_rgb_host, out_stats);
}

int ngi_gpu_group_scene(void* group, int index, void** out_scene) {
    Group* g = (Group*)group;
    if (!g || !out_scene || index < 0 || index >= (int)g->scenes.size()) return set_err(NGI_ERR_INVALID_ARGUMENT, "bad group / index");
    *out_scene = g->scenes[index];
    return NGI_OK;
}

void ngi_gpu_group_destroy(void* group) { delete (Group*)group; }

}  // extern "C"

// ngi_math.h — small vector math + exactly-rounded float ops shared by every kernel.
//
// All device algorithms of this module are written as `NGI_HD` (host+device) inline functions over these
// types so that the very same code can be stepped through on a CPU by the test-only simulator in
// tests/hostsim/ (this container has no GPU). The product (libnanogi_gpu.so) only ever runs them on the
// device; nothing here is a CPU fallback.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NGI_HD __host__ __device__ __forceinline__
#define NGI_HD_NOINLINE __host__ __device__
#else
#define NGI_HD inline
#define NGI_HD_NOINLINE inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned int x, y; };
struct uint4 { unsigned int x, y, z, w; };
struct double2 { double x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r{x, y, z, w}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r{x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r{x, y, z, w}; return r; }
#endif

// ---- exactly-rounded single operations ----------------------------------------------------------
// The triangle test must evaluate ONE expression tree on the GPU and in the CPU oracle:
// every operation is an explicitly rounded IEEE op; fused multiply-adds appear only where written.
#if defined(__CUDA_ARCH__)
#define NGI_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define NGI_MUL(a, b) __fmul_rn((a), (b))
#define NGI_ADD(a, b) __fadd_rn((a), (b))
#define NGI_SUB(a, b) __fsub_rn((a), (b))
#define NGI_RCP(a) __frcp_rn((a))
#else
// host build of these headers (tests/hostsim) is compiled with -ffp-contract=off
#define NGI_FMA(a, b, c) fmaf((a), (b), (c))
#define NGI_MUL(a, b) ((a) * (b))
#define NGI_ADD(a, b) ((a) + (b))
#define NGI_SUB(a, b) ((a) - (b))
#define NGI_RCP(a) (1.0f / (a))
#endif

#define NGI_PI_F 3.14159265358979323846f
#define NGI_INV_PI_F 0.31830988618379067154f
#define NGI_EPS_F 1e-4f                      /* EpsF, reference include/nanogi/basic.hpp:85 */
#define NGI_INF_F 3.402823466e+38f           /* InfF, reference include/nanogi/basic.hpp:84 */

struct f3 { float x, y, z; };
struct d3v { double x, y, z; };

// ---- division, square root, sin / cos of the SHADING code -----------------------------------------
// ngi_shade.h / ngi_wave.h / ngi_bdpt*.h compute in fp32 what the reference computes in fp64, so their last bits are this
// module's choice. With NGI_FAST_SHADE the device takes them from the special-function unit (MUFU.RCP / RSQ / SQRT: <= 2 ulp;
// MUFU.SIN / COS: 2^-21 absolute on [-pi, pi]) — one or two instructions where the correctly rounded sequences are ~10
// (division: MUFU.RCP + FCHK + 5 FFMA + a slow-path call), ~8 (sqrtf) and ~40 (sincosf). Not used by the texel lookup
// (fp64 like the reference, ngi_wave.h) nor by the ray queries (the explicitly rounded NGI_* macros above). The host build of these
// headers (tests/hostsim) keeps the IEEE operations.
#ifndef NGI_FAST_SHADE
#define NGI_FAST_SHADE 1
#endif
#if defined(__CUDA_ARCH__) && NGI_FAST_SHADE
#define NGI_SFU1(op, x) float r_; asm(op " %0, %1;" : "=f"(r_) : "f"(x)); return r_
NGI_HD float ngi_rcpf(const float x) { NGI_SFU1("rcp.approx.ftz.f32", x); }
NGI_HD float ngi_sqrtf(const float x) { NGI_SFU1("sqrt.approx.ftz.f32", x); }
NGI_HD float ngi_rsqrtf(const float x) { NGI_SFU1("rsqrt.approx.ftz.f32", x); }
NGI_HD float ngi_divf(const float a, const float b) { return a * ngi_rcpf(b); }
// theta in [-pi, 2 pi]: brought into [-pi, pi], where the SFU's error bound holds
NGI_HD void ngi_sincosf(float theta, float* s, float* c) {
    if (theta > NGI_PI_F) theta -= 2.0f * NGI_PI_F;
    __sincosf(theta, s, c);
}
#else
NGI_HD float ngi_rcpf(const float x) { return 1.0f / x; }
NGI_HD float ngi_sqrtf(const float x) { return sqrtf(x); }
NGI_HD float ngi_rsqrtf(const float x) { return 1.0f / sqrtf(x); }
NGI_HD float ngi_divf(const float a, const float b) { return a / b; }
NGI_HD void ngi_sincosf(const float theta, float* s, float* c) { sincosf(theta, s, c); }
#endif

NGI_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
NGI_HD f3 mk3(float a) { return mk3(a, a, a); }
NGI_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
NGI_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
NGI_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
NGI_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
NGI_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
NGI_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
NGI_HD f3 operator/(f3 a, f3 b) { return mk3(ngi_divf(a.x, b.x), ngi_divf(a.y, b.y), ngi_divf(a.z, b.z)); }
NGI_HD f3 operator/(f3 a, float s) { return mk3(ngi_divf(a.x, s), ngi_divf(a.y, s), ngi_divf(a.z, s)); }   // (one MUFU.RCP: the asm is CSE'd)
NGI_HD f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
NGI_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NGI_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
NGI_HD float length(f3 a) { return ngi_sqrtf(dot(a, a)); }
// glm::normalize(v) = v * inversesqrt(dot(v, v)); a zero vector yields NaN (relied upon by the sn fallback; 0 * MUFU.RSQ(0) = 0 * inf too)
NGI_HD f3 normalize(f3 a) { const float s = ngi_rsqrtf(dot(a, a)); return a * s; }
NGI_HD bool is_zero(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
NGI_HD float fmin2(float a, float b) { return fminf(a, b); }
NGI_HD float fmax2(float a, float b) { return fmaxf(a, b); }
NGI_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

NGI_HD unsigned f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}
NGI_HD float u2f(unsigned u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
NGI_HD int ngi_popc(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of the highest set bit (x != 0)
NGI_HD int ngi_bfind(unsigned x) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
NGI_HD int ngi_clz64(unsigned long long x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
NGI_HD int ngi_clz32(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// read-only 16-byte loads (LDG.E.128.CONSTANT on the device)
NGI_HD float4 ngi_ldg(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
NGI_HD uint4 ngi_ldg(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
NGI_HD float ngi_ldg(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
NGI_HD unsigned ngi_ldg(const unsigned* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, so a sample's uniforms depend only on
// (seed, sample index, vertex, block) — never on GPU count, wave capacity or queue order. -------------
NGI_HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#if defined(__CUDA_ARCH__)
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
#endif
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 24-bit uniform in [0, 1)
NGI_HD float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

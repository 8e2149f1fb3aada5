// oracle.cpp — CPU restatement of nanogi's `pt` / `ptdirect` hot path.
//
// *** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
// load this library. The product (libnanogi_gpu.so, the `nanogi` CLI) never links or calls it.
//
// PARITY: PINNED AGAINST THE REFERENCE'S OWN CODE, with one declared substitution (Embree, below).
// The reference ships no tests, golden vectors or numeric images and its third-party
// libraries (Boost, glm, Embree 2.5.1, TBB, Assimp, FreeImage, yaml-cpp, ctemplate, Eigen) are absent here, so it
// cannot be built as shipped. Its own sources CAN be compiled where they lie against small stand-in headers for
// those libraries (oracle/refshim/, oracle/build_ref.sh -> oracle/_ref/): Scene::Load, Primitive::*,
// Scene::Intersect's reconstruction, Visible, Random, RenderProcess and ProcessSample_PT / _PTDirect / _LT /
// _LTDirect then run here unmodified. With one thread and the release seed std::time(nullptr) interposed, that
// run is deterministic, and this file's mt19937 mode reproduces its films BIT FOR BIT (float64) on five of six
// scenes and to 4e-15 on the sixth, for all four renderers; the per-function tables agree to the last bit
// (tests/test_reference_pin.py; committed reference outputs: tests/golden/reference_films.npz).
// This file stays a line-by-line restatement, each function citing the lines it follows (paths relative to
// /root/reference); the analytic known-answer tests of tests/test_oracle.py remain as a second anchor.
// What the pin does NOT cover: Embree's own kernel (next paragraph); glm's default constructors are assumed to
// zero-initialise like glm 0.9.5 (the reference relies on it); the draw order inside
// `SampleDirection(rng.Next2D(), rng.Next(), ...)` is unspecified C++ and follows what g++ generates.
//
// Third-party arithmetic that is NOT in the reference tree: Embree v2.5.1 (Dockerfile:16-17),
// called at include/nanogi/rt.hpp:2092-2142 (build) and :2182 (rtcIntersect, the only query).
// It is substituted by a float32 Moeller-Trumbore closest-hit with an explicit operation order
// (tri_test below; the same expression tree is evaluated by the CUDA kernels) over a SAH BVH that is
// a pure conservative filter; ties in t are broken by the lowest global triangle id. That
// substitution is the declared meaning of "the reference intersector" in this repository.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off: no implicit FMA contraction, so the
// explicit fmaf() calls are the only fused operations, exactly like the __fmaf_rn device code).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../include/nanogi_gpu.h"  // POD scene description only (no product code is linked)

namespace {

// ---------------------------------------------------------------------------------------------
// constants: include/nanogi/basic.hpp:81-85
// ---------------------------------------------------------------------------------------------
const double Pi = 3.14159265358979323846264338327950288;
const double InvPi = 1.0 / Pi;
const float InfF = std::numeric_limits<float>::max();
const float EpsF = 1e-4f;

struct d2 { double x = 0, y = 0; };
struct d3 {
    double x = 0, y = 0, z = 0;
    d3() {}
    d3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit d3(double a) : x(a), y(a), z(a) {}
    explicit d3(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
};
inline d3 operator+(const d3& a, const d3& b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline d3 operator-(const d3& a, const d3& b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline d3 operator-(const d3& a) { return d3(-a.x, -a.y, -a.z); }
inline d3 operator*(const d3& a, const d3& b) { return d3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline d3 operator*(const d3& a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
inline d3 operator*(double s, const d3& a) { return d3(a.x * s, a.y * s, a.z * s); }
inline d3 operator/(const d3& a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
inline d3 operator/(const d3& a, const d3& b) { return d3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline d3 operator+(const d3& a, double s) { return d3(a.x + s, a.y + s, a.z + s); }
inline double dot(const d3& a, const d3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline d3 cross(const d3& a, const d3& b) {
    return d3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline double length(const d3& a) { return std::sqrt(dot(a, a)); }
// glm::normalize(v) = v * inversesqrt(dot(v,v))
inline d3 normalize(const d3& a) { return a * (1.0 / std::sqrt(dot(a, a))); }
inline bool is_zero(const d3& a) { return a.x == 0 && a.y == 0 && a.z == 0; }
template <class T> inline T clampv(T v, T lo, T hi) { return std::min(std::max(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// include/nanogi/rt.hpp:55-59
// ---------------------------------------------------------------------------------------------
void OrthonormalBasis(const d3& a, d3& b, d3& c) {
    c = std::abs(a.x) > std::abs(a.y) ? normalize(d3(a.z, 0, -a.x)) : normalize(d3(0, a.z, -a.y));
    b = normalize(cross(c, a));
}
// rt.hpp:66-85
inline double LocalCos(const d3& v) { return v.z; }
inline double LocalTan(const d3& v) {
    const double t = 1.0 - v.z * v.z;
    return t <= 0 ? 0 : std::sqrt(t) / v.z;
}
inline d3 LocalReflect(const d3& wi) { return d3(-wi.x, -wi.y, wi.z); }
inline d3 LocalRefract(const d3& wi, double eta, double cosThetaT) { return d3(-eta * wi.x, -eta * wi.y, cosThetaT); }

// rt.hpp:87-103
d2 UniformConcentricDiskSample(const d2& u) {
    d2 v; v.x = 2.0 * u.x - 1.0; v.y = 2.0 * u.y - 1.0;
    if (v.x == 0 && v.y == 0) return d2();
    double r, theta;
    if (v.x > -v.y) {
        if (v.x > v.y) { r = v.x; theta = (Pi * 0.25) * v.y / v.x; }
        else           { r = v.y; theta = (Pi * 0.25) * (2.0 - v.x / v.y); }
    } else {
        if (v.x < v.y) { r = -v.x; theta = (Pi * 0.25) * (4.0 + v.y / v.x); }
        else           { r = -v.y; theta = (Pi * 0.25) * (6.0 - v.x / v.y); }
    }
    d2 o; o.x = r * std::cos(theta); o.y = r * std::sin(theta);
    return o;
}
// rt.hpp:105-109
d3 CosineSampleHemisphere(const d2& u) {
    const d2 s = UniformConcentricDiskSample(u);
    return d3(s.x, s.y, std::sqrt(std::max(0.0, 1.0 - s.x * s.x - s.y * s.y)));
}
// rt.hpp:116-122
d3 UniformSampleSphere(const d2& u) {
    const double z = 1.0 - 2.0 * u.x;
    const double r = std::sqrt(std::max(0.0, 1.0 - z * z));
    const double phi = 2.0 * Pi * u.y;
    return d3(r * std::cos(phi), r * std::sin(phi), z);
}
// rt.hpp:129-133
d2 UniformSampleTriangle(const d2& u) {
    const double s = std::sqrt(std::max(0.0, u.x));
    d2 o; o.x = 1.0 - s; o.y = u.y * s;
    return o;
}
// rt.hpp:135-140
int PixelIndex(const d2& rasterPos, int w, int h) {
    const int pX = clampv((int)(rasterPos.x * w), 0, w - 1);
    const int pY = clampv((int)(rasterPos.y * h), 0, h - 1);
    return pY * w + pX;
}

// ---------------------------------------------------------------------------------------------
// include/nanogi/basic.hpp:440-497
// ---------------------------------------------------------------------------------------------
struct Distribution1D {
    std::vector<double> cdf{0.0};
    void Add(double v) { cdf.push_back(cdf.back() + v); }
    void Normalize() {
        const double sum = cdf.back();
        const double invSum = 1.0 / sum;
        for (auto& v : cdf) v *= invSum;
    }
    int SampleReuse(double u, double& u2) const {
        int v = static_cast<int>(std::upper_bound(cdf.begin(), cdf.end(), u) - cdf.begin()) - 1;
        int i = clampv<int>(v, 0, static_cast<int>(cdf.size()) - 2);
        u2 = (u - cdf[i]) / (cdf[i + 1] - cdf[i]);
        return i;
    }
};

// rt.hpp:282-302
struct SurfaceGeometry {
    bool degenerated = false;
    d3 p, sn, gn, dpdu, dpdv;
    d2 uv;
    void ComputeTangentSpace() { OrthonormalBasis(sn, dpdu, dpdv); }
    // ToWorld = [dpdu dpdv sn], ToLocal = ToWorld^T
    d3 ToLocal(const d3& w) const { return d3(dot(dpdu, w), dot(dpdv, w), dot(sn, w)); }
    d3 ToWorld(const d3& l) const { return dpdu * l.x + dpdv * l.y + sn * l.z; }
};

enum TransportDirection { LE, EL };

struct Scene;

// rt.hpp:157-270 (Evaluate :262-268)
struct Texture {
    int Width = 0, Height = 0;
    std::vector<float> Data;
    d3 Evaluate(const d2& uv) const {
        auto fract = [](double x) { return x - std::floor(x); };
        const int x = clampv((int)(fract(uv.x) * Width), 0, Width - 1);
        const int y = clampv((int)(fract(uv.y) * Height), 0, Height - 1);
        const int i = Width * y + x;
        return d3(Data[3 * i], Data[3 * i + 1], Data[3 * i + 2]);
    }
};

// rt.hpp:379-1470
struct Primitive {
    int Type = 0;
    int firstTri = -1, numTris = 0;  // MeshRef (de-indexed range)
    const Scene* scene = nullptr;
    int LType = 0, EType = 0, SType = 0;
    d3 L_Le, L_vec, L_Center;
    double L_InvArea = 0, L_Radius = 0;
    Distribution1D Dist;  // L.Area.Dist or E.Area.Dist
    d3 E_Position, E_Vx, E_Vy, E_Vz, E_We;
    double E_Fov = 0, E_Aspect = 1, E_InvArea = 0;
    d3 D_R; const Texture* D_TexR = nullptr;
    d3 G_R, G_Eta, G_K; double G_Roughness = 0; const Texture* G_TexR = nullptr;
    d3 S_R; double S_Eta1 = 1, S_Eta2 = 1;

    void SamplePosition(const d2& u, SurfaceGeometry& geom) const;
    d3 EvaluatePosition(const SurfaceGeometry& geom, bool forceDegenerated) const;
    double EvaluatePositionPDF(const SurfaceGeometry& geom, bool forceDegenerated) const;
    bool SampleDirection(const d2& u, double uComp, int queryType, const SurfaceGeometry& geom, const d3& wi, d3& wo) const;
    d3 EvaluateDirection(const SurfaceGeometry& geom, int queryType, const d3& wi, const d3& wo, TransportDirection transDir, bool forceDegenerated) const;
    double EvaluateDirectionPDF(const SurfaceGeometry& geom, int queryType, const d3& wi, const d3& wo, bool forceDegenerated) const;
    bool RasterPosition(const d3& wo, const SurfaceGeometry& geom, d2& rasterPos) const;
    double EvaluateBechmannDist(const d3& H) const;
    double EvalauteShadowMaskingFunc(const d3& wi, const d3& wo, const d3& H) const;
    d3 EvaluateFrConductor(double cosThetaI) const;
    double EvaluateFresnelTerm(const d3& localWi, double etaI, double etaT) const;
};

// ---------------------------------------------------------------------------------------------
// The substituted intersector (see header). Float32, explicit operation order.
//   e1 = v1 - v0, e2 = v2 - v0 (IEEE float subtraction, precomputed)
//   P = d x e2 ; det = e1 . P ; inv = 1/det ; T = o - v0 ; u = (T . P) * inv
//   Q = T x e1 ; v = (d . Q) * inv ; t = (e2 . Q) * inv
//   cross(a,b).x = fma(a.y, b.z, -(a.z*b.y)) ; dot(a,b) = fma(a.x,b.x, fma(a.y,b.y, a.z*b.z))
// accept iff det != 0, 0 <= u <= 1, v >= 0, u+v <= 1, tmin < t < tmax (NaN fails every test).
// No back-face culling (rt.hpp:2174-2178: no culling flags).
// ---------------------------------------------------------------------------------------------
struct TriF { float v0[3], e1[3], e2[3]; };

inline float crossx(float ay, float az, float by, float bz) { return fmaf(ay, bz, -(az * by)); }
inline float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(ax, bx, fmaf(ay, by, az * bz)); }

inline bool tri_test(const TriF& tr, const float o[3], const float d[3], float tmin, float tmax, float& t, float& u, float& v) {
    const float px = crossx(d[1], d[2], tr.e2[1], tr.e2[2]);
    const float py = crossx(d[2], d[0], tr.e2[2], tr.e2[0]);
    const float pz = crossx(d[0], d[1], tr.e2[0], tr.e2[1]);
    const float det = dot3(tr.e1[0], tr.e1[1], tr.e1[2], px, py, pz);
    if (!(det != 0.0f)) return false;
    const float inv = 1.0f / det;
    const float tx = o[0] - tr.v0[0], ty = o[1] - tr.v0[1], tz = o[2] - tr.v0[2];
    const float uu = dot3(tx, ty, tz, px, py, pz) * inv;
    if (!(uu >= 0.0f && uu <= 1.0f)) return false;
    const float qx = crossx(ty, tz, tr.e1[1], tr.e1[2]);
    const float qy = crossx(tz, tx, tr.e1[2], tr.e1[0]);
    const float qz = crossx(tx, ty, tr.e1[0], tr.e1[1]);
    const float vv = dot3(d[0], d[1], d[2], qx, qy, qz) * inv;
    if (!(vv >= 0.0f && uu + vv <= 1.0f)) return false;
    const float tt = dot3(tr.e2[0], tr.e2[1], tr.e2[2], qx, qy, qz) * inv;
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; u = uu; v = vv;
    return true;
}

struct BNode {
    float bmin[3], bmax[3];
    int32_t left;    // internal: index of left child (right = left+1); leaf: first entry in triIdx
    int32_t count;   // 0 = internal, >0 = leaf triangle count
};

struct Scene {
    // meshes (de-indexed; Assimp stores float, the reference widens to double: rt.hpp:1685-1696)
    std::vector<float> pos, nrm, uv;  // [nTri*9], [nTri*9], [nTri*6] or empty
    std::vector<int> triPrim;
    std::vector<Primitive> Primitives;
    std::vector<Texture> Textures;
    size_t SensorPrimitiveIndex = (size_t)-1;
    std::vector<size_t> LightPrimitiveIndices;
    // intersector
    std::vector<TriF> tris;
    std::vector<BNode> nodes;
    std::vector<uint32_t> triIdx;
    float pad = 0;

    d3 P(int tri, int k) const { const float* p = &pos[(size_t)tri * 9 + 3 * k]; return d3(p[0], p[1], p[2]); }
    d3 Nv(int tri, int k) const { const float* p = &nrm[(size_t)tri * 9 + 3 * k]; return d3(p[0], p[1], p[2]); }
    d2 UV(int tri, int k) const { d2 r; r.x = uv[(size_t)tri * 6 + 2 * k]; r.y = uv[(size_t)tri * 6 + 2 * k + 1]; return r; }

    // rt.hpp:2321-2336
    const Primitive* SampleEmitter(int type, double u) const {
        if ((type & NGI_TYPE_L) > 0) {
            int n = static_cast<int>(LightPrimitiveIndices.size());
            int i = clampv(static_cast<int>(u * n), 0, n - 1);
            return &Primitives.at(LightPrimitiveIndices[i]);
        }
        if ((type & NGI_TYPE_E) > 0) return &Primitives.at(SensorPrimitiveIndex);
        return nullptr;
    }
    // rt.hpp:2338-2352
    double EvaluateEmitterPDF(const Primitive* primitive) const {
        if ((primitive->Type & NGI_TYPE_L) > 0) { int n = static_cast<int>(LightPrimitiveIndices.size()); return 1.0 / n; }
        if ((primitive->Type & NGI_TYPE_E) > 0) return 1;
        return 0;
    }

    void BuildBVH();
    bool TraceClosest(const float o[3], const float d[3], float tmin, float tmax, float& t, float& u, float& v, uint32_t& tri) const;
    bool TraceAny(const float o[3], const float d[3], float tmin, float tmax) const;
    bool TraceBrute(const float o[3], const float d[3], float tmin, float tmax, float& t, float& u, float& v, uint32_t& tri) const;
};

// ---------------------------------------------------------------------------------------------
// BVH (binned SAH), boxes padded so that the BVH is a pure filter for tri_test
// ---------------------------------------------------------------------------------------------
void Scene::BuildBVH() {
    const size_t n = pos.size() / 9;
    tris.resize(n);
    std::vector<float> bmin(n * 3), bmax(n * 3), cen(n * 3);
    float smin[3] = {InfF, InfF, InfF}, smax[3] = {-InfF, -InfF, -InfF};
    for (size_t i = 0; i < n; i++) {
        const float* p = &pos[i * 9];
        for (int k = 0; k < 3; k++) {
            tris[i].v0[k] = p[k];
            tris[i].e1[k] = p[3 + k] - p[k];
            tris[i].e2[k] = p[6 + k] - p[k];
            bmin[i * 3 + k] = std::min(p[k], std::min(p[3 + k], p[6 + k]));
            bmax[i * 3 + k] = std::max(p[k], std::max(p[3 + k], p[6 + k]));
            smin[k] = std::min(smin[k], bmin[i * 3 + k]);
            smax[k] = std::max(smax[k], bmax[i * 3 + k]);
        }
    }
    // conservative padding: rounding of tri_test is ~ulp(|o - v0|); pad by 2^-16 of the largest magnitude
    float mag = 1e-30f;
    for (int k = 0; k < 3; k++) mag = std::max(mag, std::max(std::abs(smin[k]), std::abs(smax[k])));
    if (SensorPrimitiveIndex != (size_t)-1) {
        const Primitive& e = Primitives[SensorPrimitiveIndex];
        mag = std::max(mag, (float)std::max(std::abs(e.E_Position.x), std::max(std::abs(e.E_Position.y), std::abs(e.E_Position.z))));
    }
    pad = mag * (1.0f / 65536.0f);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            bmin[i * 3 + k] -= pad; bmax[i * 3 + k] += pad;
            cen[i * 3 + k] = 0.5f * (bmin[i * 3 + k] + bmax[i * 3 + k]);
        }
    triIdx.resize(n);
    for (size_t i = 0; i < n; i++) triIdx[i] = (uint32_t)i;
    nodes.clear();
    nodes.reserve(2 * n + 1);
    if (n == 0) return;
    struct Task { int node; size_t lo, hi; };
    nodes.push_back(BNode());
    std::vector<Task> stack{{0, 0, n}};
    auto area = [](const float* mn, const float* mx) {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    };
    while (!stack.empty()) {
        Task t = stack.back(); stack.pop_back();
        float mn[3] = {InfF, InfF, InfF}, mx[3] = {-InfF, -InfF, -InfF};
        float cmn[3] = {InfF, InfF, InfF}, cmx[3] = {-InfF, -InfF, -InfF};
        for (size_t i = t.lo; i < t.hi; i++) {
            uint32_t id = triIdx[i];
            for (int k = 0; k < 3; k++) {
                mn[k] = std::min(mn[k], bmin[id * 3 + k]); mx[k] = std::max(mx[k], bmax[id * 3 + k]);
                cmn[k] = std::min(cmn[k], cen[id * 3 + k]); cmx[k] = std::max(cmx[k], cen[id * 3 + k]);
            }
        }
        BNode& nd = nodes[t.node];
        for (int k = 0; k < 3; k++) { nd.bmin[k] = mn[k]; nd.bmax[k] = mx[k]; }
        const size_t cnt = t.hi - t.lo;
        if (cnt <= 2) { nd.left = (int32_t)t.lo; nd.count = (int32_t)cnt; continue; }
        // binned SAH over the widest centroid axis candidates
        const int NB = 16;
        int bestAxis = -1, bestBin = -1; float bestCost = InfF;
        for (int ax = 0; ax < 3; ax++) {
            const float ext = cmx[ax] - cmn[ax];
            if (!(ext > 0)) continue;
            float bmn[NB][3], bmx[NB][3]; int bc[NB];
            for (int b = 0; b < NB; b++) { bc[b] = 0; for (int k = 0; k < 3; k++) { bmn[b][k] = InfF; bmx[b][k] = -InfF; } }
            const float sc = NB / ext;
            for (size_t i = t.lo; i < t.hi; i++) {
                uint32_t id = triIdx[i];
                int b = std::min(NB - 1, (int)((cen[id * 3 + ax] - cmn[ax]) * sc));
                bc[b]++;
                for (int k = 0; k < 3; k++) { bmn[b][k] = std::min(bmn[b][k], bmin[id * 3 + k]); bmx[b][k] = std::max(bmx[b][k], bmax[id * 3 + k]); }
            }
            float rA[NB]; int rC[NB];
            float amn[3] = {InfF, InfF, InfF}, amx[3] = {-InfF, -InfF, -InfF}; int c = 0;
            for (int b = NB - 1; b > 0; b--) {
                for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[b][k]); amx[k] = std::max(amx[k], bmx[b][k]); }
                c += bc[b]; rA[b] = c ? area(amn, amx) : 0; rC[b] = c;
            }
            for (int k = 0; k < 3; k++) { amn[k] = InfF; amx[k] = -InfF; }
            c = 0;
            for (int b = 0; b < NB - 1; b++) {
                for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[b][k]); amx[k] = std::max(amx[k], bmx[b][k]); }
                c += bc[b];
                if (c == 0 || rC[b + 1] == 0) continue;
                float cost = area(amn, amx) * c + rA[b + 1] * rC[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestBin = b; }
            }
        }
        size_t mid;
        if (bestAxis < 0) {
            mid = (t.lo + t.hi) / 2;  // all centroids coincide
        } else {
            const float ext = cmx[bestAxis] - cmn[bestAxis];
            const float sc = NB / ext;
            auto it = std::partition(triIdx.begin() + t.lo, triIdx.begin() + t.hi, [&](uint32_t id) {
                int b = std::min(NB - 1, (int)((cen[id * 3 + bestAxis] - cmn[bestAxis]) * sc));
                return b <= bestBin;
            });
            mid = (size_t)(it - triIdx.begin());
            if (mid == t.lo || mid == t.hi) mid = (t.lo + t.hi) / 2;
        }
        int l = (int)nodes.size();
        nodes.push_back(BNode()); nodes.push_back(BNode());
        nodes[t.node].left = l; nodes[t.node].count = 0;
        stack.push_back({l, t.lo, mid});
        stack.push_back({l + 1, mid, t.hi});
    }
}

// slab test in double on the padded float boxes: conservative w.r.t. tri_test
inline bool box_hit(const BNode& nd, const double o[3], const double id[3], double tmin, double tmax, double& tnear) {
    double t0 = tmin, t1 = tmax;
    for (int k = 0; k < 3; k++) {
        double a = (nd.bmin[k] - o[k]) * id[k], b = (nd.bmax[k] - o[k]) * id[k];
        if (a > b) std::swap(a, b);
        // NaN (0 * inf) must not cull: comparisons with NaN are false, so keep the interval
        if (a > t0) t0 = a;
        if (b < t1) t1 = b;
    }
    tnear = t0;
    return t0 <= t1;
}

bool Scene::TraceClosest(const float o[3], const float d[3], float tmin, float tmax, float& t, float& u, float& v, uint32_t& tri) const {
    if (nodes.empty()) return false;
    double od[3] = {o[0], o[1], o[2]}, id[3] = {1.0 / d[0], 1.0 / d[1], 1.0 / d[2]};
    bool found = false; float bt = tmax; uint32_t bid = 0xFFFFFFFFu; float bu = 0, bv = 0;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode& nd = nodes[stack[--sp]];
        double tn;
        // inclusive culling against the current best t (equal-t lower-id triangles must survive);
        // a hair of slack on the far bound absorbs the float->double mismatch of bt
        if (!box_hit(nd, od, id, (double)tmin * 0.999, (double)bt * 1.000001 + 1e-30, tn)) continue;
        if (nd.count > 0) {
            for (int i = 0; i < nd.count; i++) {
                uint32_t id_ = triIdx[nd.left + i];
                float tt, uu, vv;
                // candidate window is (tmin, tmax) of the ray; the closest-hit reduction is explicit
                if (tri_test(tris[id_], o, d, tmin, tmax, tt, uu, vv)) {
                    if (!found || tt < bt || (tt == bt && id_ < bid)) { found = true; bt = tt; bid = id_; bu = uu; bv = vv; }
                }
            }
        } else {
            double ta, tb;
            bool ha = box_hit(nodes[nd.left], od, id, (double)tmin * 0.999, (double)bt * 1.000001 + 1e-30, ta);
            bool hb = box_hit(nodes[nd.left + 1], od, id, (double)tmin * 0.999, (double)bt * 1.000001 + 1e-30, tb);
            if (ha && hb) {
                if (ta < tb) { stack[sp++] = nd.left + 1; stack[sp++] = nd.left; }
                else { stack[sp++] = nd.left; stack[sp++] = nd.left + 1; }
            } else if (ha) stack[sp++] = nd.left;
            else if (hb) stack[sp++] = nd.left + 1;
        }
    }
    if (found) { t = bt; u = bu; v = bv; tri = bid; }
    return found;
}

bool Scene::TraceAny(const float o[3], const float d[3], float tmin, float tmax) const {
    if (nodes.empty()) return false;
    double od[3] = {o[0], o[1], o[2]}, id[3] = {1.0 / d[0], 1.0 / d[1], 1.0 / d[2]};
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode& nd = nodes[stack[--sp]];
        double tn;
        if (!box_hit(nd, od, id, (double)tmin * 0.999, (double)tmax * 1.000001, tn)) continue;
        if (nd.count > 0) {
            for (int i = 0; i < nd.count; i++) {
                float tt, uu, vv;
                if (tri_test(tris[triIdx[nd.left + i]], o, d, tmin, tmax, tt, uu, vv)) return true;
            }
        } else { stack[sp++] = nd.left; stack[sp++] = nd.left + 1; }
    }
    return false;
}

bool Scene::TraceBrute(const float o[3], const float d[3], float tmin, float tmax, float& t, float& u, float& v, uint32_t& tri) const {
    bool found = false; float bt = tmax; uint32_t bid = 0xFFFFFFFFu; float bu = 0, bv = 0;
    for (uint32_t i = 0; i < tris.size(); i++) {
        float tt, uu, vv;
        if (tri_test(tris[i], o, d, tmin, tmax, tt, uu, vv))
            if (!found || tt < bt || (tt == bt && i < bid)) { found = true; bt = tt; bid = i; bu = uu; bv = vv; }
    }
    if (found) { t = bt; u = bu; v = bv; tri = bid; }
    return found;
}

// ---------------------------------------------------------------------------------------------
// Primitive functions — include/nanogi/rt.hpp:483-1466
// ---------------------------------------------------------------------------------------------

// rt.hpp:488-526 (SampleTriangleMesh lambda)
static void SampleTriangleMesh(const d2& u, const Scene* sc, int firstTri, const Distribution1D& dist, SurfaceGeometry& geom) {
    d2 u2 = u;
    const int i = dist.SampleReuse(u.x, u2.x);
    const d2 b = UniformSampleTriangle(u2);
    const int tri = firstTri + i;
    const d3 p1 = sc->P(tri, 0), p2 = sc->P(tri, 1), p3 = sc->P(tri, 2);
    geom.p = p1 * (1.0 - b.x - b.y) + p2 * b.x + p3 * b.y;
    if (!sc->uv.empty()) {
        const d2 uv1 = sc->UV(tri, 0), uv2 = sc->UV(tri, 1), uv3 = sc->UV(tri, 2);
        geom.uv.x = uv1.x * (1.0 - b.x - b.y) + uv2.x * b.x + uv3.x * b.y;
        geom.uv.y = uv1.y * (1.0 - b.x - b.y) + uv2.y * b.x + uv3.y * b.y;
    }
    geom.degenerated = false;
    geom.gn = normalize(cross(p2 - p1, p3 - p1));
    geom.sn = geom.gn;
    geom.ComputeTangentSpace();
}

// rt.hpp:483-592
void Primitive::SamplePosition(const d2& u, SurfaceGeometry& geom) const {
    if ((Type & NGI_TYPE_L) > 0) {
        if (LType == NGI_L_AREA) { SampleTriangleMesh(u, scene, firstTri, Dist, geom); return; }
        if (LType == NGI_L_POINT) { geom.degenerated = true; geom.p = L_vec; return; }
        if (LType == NGI_L_DIRECTIONAL) {
            const d2 p = UniformConcentricDiskSample(u);
            geom.degenerated = false;
            geom.gn = L_vec; geom.sn = geom.gn;
            geom.ComputeTangentSpace();
            geom.p = L_Center - L_vec * L_Radius + (geom.dpdu * (p.x * L_Radius) + geom.dpdv * (p.y * L_Radius));
            return;
        }
    }
    if ((Type & NGI_TYPE_E) > 0) {
        if (EType == NGI_E_AREA) { SampleTriangleMesh(u, scene, firstTri, Dist, geom); return; }
        if (EType == NGI_E_PINHOLE) { geom.degenerated = true; geom.p = E_Position; return; }
    }
}
// rt.hpp:594-641
d3 Primitive::EvaluatePosition(const SurfaceGeometry&, bool forceDegenerated) const {
    if ((Type & NGI_TYPE_L) > 0) {
        if (LType == NGI_L_AREA) return d3(1);
        if (LType == NGI_L_POINT) return forceDegenerated ? d3(1) : d3();
        if (LType == NGI_L_DIRECTIONAL) return d3(1);
    }
    if ((Type & NGI_TYPE_E) > 0) {
        if (EType == NGI_E_AREA) return d3(1);
        if (EType == NGI_E_PINHOLE) return forceDegenerated ? d3(1) : d3();
    }
    return d3();
}
// rt.hpp:643-690
double Primitive::EvaluatePositionPDF(const SurfaceGeometry&, bool forceDegenerated) const {
    if ((Type & NGI_TYPE_L) > 0) {
        if (LType == NGI_L_AREA) return L_InvArea;
        if (LType == NGI_L_POINT) return forceDegenerated ? 1 : 0;
        if (LType == NGI_L_DIRECTIONAL) return L_InvArea;
    }
    if ((Type & NGI_TYPE_E) > 0) {
        if (EType == NGI_E_AREA) return E_InvArea;
        if (EType == NGI_E_PINHOLE) return forceDegenerated ? 1 : 0;
    }
    return 0;
}

// rt.hpp:692-910. Returns false when the reference returns WITHOUT writing `wo`
// (callers hold a zero-initialised glm::dvec3 -> cos(wo)=0 -> fs=0 -> path ends).
bool Primitive::SampleDirection(const d2& u, double uComp, int queryType, const SurfaceGeometry& geom, const d3& wi, d3& wo) const {
    if ((queryType & NGI_TYPE_L) > 0) {
        if (LType == NGI_L_AREA) { wo = geom.ToWorld(CosineSampleHemisphere(u)); return true; }
        if (LType == NGI_L_POINT) { wo = UniformSampleSphere(u); return true; }
        if (LType == NGI_L_DIRECTIONAL) { wo = L_vec; return true; }
    }
    if ((queryType & NGI_TYPE_E) > 0) {
        if (EType == NGI_E_AREA) { wo = geom.ToWorld(CosineSampleHemisphere(u)); return true; }
        if (EType == NGI_E_PINHOLE) {
            const double rx = 2.0 * u.x - 1.0, ry = 2.0 * u.y - 1.0;
            const double tanFov = std::tan(E_Fov * 0.5);
            const d3 woEye = normalize(d3(E_Aspect * tanFov * rx, tanFov * ry, -1));
            wo = E_Vx * woEye.x + E_Vy * woEye.y + E_Vz * woEye.z;
            return true;
        }
    }
    if ((queryType & NGI_TYPE_D) > 0) {
        const d3 localWi = geom.ToLocal(wi);
        if (LocalCos(localWi) <= 0) return false;
        wo = geom.ToWorld(CosineSampleHemisphere(u));
        return true;
    }
    if ((queryType & NGI_TYPE_G) > 0) {
        const d3 localWi = geom.ToLocal(wi);
        if (LocalCos(localWi) <= 0) return false;
        // SampleBechmannDist, rt.hpp:777-785
        const double tanThetaHSqr = -G_Roughness * G_Roughness * std::log(1.0 - u.x);
        const double cosThetaH = 1.0 / std::sqrt(1.0 + tanThetaHSqr);
        const double cosThetaH2 = cosThetaH * cosThetaH;
        const double sinThetaH = std::sqrt(std::max(0.0, 1.0 - cosThetaH2));
        const double phiH = 2.0 * Pi * u.y;
        const d3 H(sinThetaH * std::cos(phiH), sinThetaH * std::sin(phiH), cosThetaH);
        const d3 localWo = -localWi - 2.0 * dot(-localWi, H) * H;
        if (LocalCos(localWo) <= 0) return false;
        wo = geom.ToWorld(localWo);
        return true;
    }
    if ((queryType & NGI_TYPE_S) > 0) {
        if (SType == NGI_S_REFLECTION) {
            const d3 localWi = geom.ToLocal(wi);
            if (LocalCos(localWi) <= 0) return false;
            wo = geom.ToWorld(LocalReflect(localWi));
            return true;
        }
        if (SType == NGI_S_REFRACTION) {
            const d3 localWi = geom.ToLocal(wi);
            double etaI = S_Eta1, etaT = S_Eta2;
            if (LocalCos(localWi) < 0) std::swap(etaI, etaT);
            const double wiDotN = LocalCos(localWi);
            const double eta = etaI / etaT;
            const double cosThetaTSq = 1.0 - eta * eta * (1.0 - wiDotN * wiDotN);
            if (cosThetaTSq <= 0) { wo = geom.ToWorld(LocalReflect(localWi)); return true; }
            const double cosThetaT = std::sqrt(cosThetaTSq) * (wiDotN > 0 ? -1.0 : 1.0);
            wo = geom.ToWorld(LocalRefract(localWi, eta, cosThetaT));
            return true;
        }
        if (SType == NGI_S_FRESNEL) {
            const d3 localWi = geom.ToLocal(wi);
            double etaI = S_Eta1, etaT = S_Eta2;
            if (LocalCos(localWi) < 0) std::swap(etaI, etaT);
            const double Fr = EvaluateFresnelTerm(localWi, etaI, etaT);
            if (uComp <= Fr) {
                wo = geom.ToWorld(LocalReflect(localWi));
            } else {
                const double wiDotN = LocalCos(localWi);
                const double eta = etaI / etaT;
                const double cosThetaTSq = 1.0 - eta * eta * (1.0 - wiDotN * wiDotN);
                const double cosThetaT = std::sqrt(cosThetaTSq) * (wiDotN > 0 ? -1.0 : 1.0);
                wo = geom.ToWorld(LocalRefract(localWi, eta, cosThetaT));
            }
            return true;
        }
    }
    return false;  // assert(0) in debug; release falls through with wo untouched (rt.hpp:909)
}

// rt.hpp:912-1148
d3 Primitive::EvaluateDirection(const SurfaceGeometry& geom, int queryType, const d3& wi, const d3& wo, TransportDirection transDir, bool forceDegenerated) const {
    if ((queryType & NGI_TYPE_EMITTER) > 0) {
        if ((queryType & NGI_TYPE_L) > 0) {
            if (LType == NGI_L_AREA) {
                const d3 localWo = geom.ToLocal(wo);
                if (LocalCos(localWo) <= 0) return d3();
                return L_Le;
            }
            if (LType == NGI_L_POINT) return L_Le;
            if (LType == NGI_L_DIRECTIONAL) return forceDegenerated ? L_Le : d3();
        }
        if ((queryType & NGI_TYPE_E) > 0) {
            if (EType == NGI_E_AREA) {
                const d3 localWo = geom.ToLocal(wo);
                if (LocalCos(localWo) <= 0) return d3();
                return E_We;
            }
            if (EType == NGI_E_PINHOLE) {
                d2 rasterPos;
                if (!RasterPosition(wo, geom, rasterPos)) return d3();
                const d3 woEye(dot(E_Vx, wo), dot(E_Vy, wo), dot(E_Vz, wo));
                const double tanFov = std::tan(E_Fov * 0.5);
                const double cosTheta = -LocalCos(woEye);
                const double invCosTheta = 1.0 / cosTheta;
                const double A = tanFov * tanFov * E_Aspect * 4.0;
                return d3(invCosTheta * invCosTheta * invCosTheta / A);
            }
        }
    }
    if ((queryType & NGI_TYPE_BSDF) > 0) {
        // shadingNormalCorrection, rt.hpp:994-1005
        double shadingNormalCorrection;
        {
            const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
            const double wiDotNg = dot(wi, geom.gn), woDotNg = dot(wo, geom.gn);
            const double wiDotNs = LocalCos(localWi), woDotNs = LocalCos(localWo);
            if (wiDotNg * wiDotNs <= 0 || woDotNg * woDotNs <= 0) shadingNormalCorrection = 0;
            else if (transDir == LE) shadingNormalCorrection = wiDotNs * woDotNg / (woDotNs * wiDotNg);
            else shadingNormalCorrection = 1;
        }
        if ((queryType & NGI_TYPE_D) > 0) {
            const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
            if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return d3();
            const d3 R = D_TexR ? D_TexR->Evaluate(geom.uv) : D_R;
            return R * InvPi * shadingNormalCorrection;
        }
        if ((queryType & NGI_TYPE_G) > 0) {
            const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
            if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return d3();
            const d3 H = normalize(localWi + localWo);
            const double D = EvaluateBechmannDist(H);
            const double G = EvalauteShadowMaskingFunc(localWi, localWo, H);
            const d3 F = EvaluateFrConductor(dot(localWi, H));
            const d3 R = G_TexR ? G_TexR->Evaluate(geom.uv) : G_R;
            return R * D * G * F / (4.0 * LocalCos(localWi)) / LocalCos(localWo) * shadingNormalCorrection;
        }
        if ((queryType & NGI_TYPE_S) > 0) {
            if (!forceDegenerated) return d3();
            if (SType == NGI_S_REFLECTION) {
                const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
                if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return d3();
                return S_R * shadingNormalCorrection;
            }
            if (SType == NGI_S_REFRACTION) {
                const d3 localWi = geom.ToLocal(wi);
                double etaI = S_Eta1, etaT = S_Eta2;
                if (LocalCos(localWi) < 0) std::swap(etaI, etaT);
                const double eta = etaI / etaT;
                const double refrCorrection = transDir == EL ? eta : 1.0;
                return S_R * shadingNormalCorrection * refrCorrection * refrCorrection;
            }
            if (SType == NGI_S_FRESNEL) {
                const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
                double etaI = S_Eta1, etaT = S_Eta2;
                if (LocalCos(localWi) < 0) std::swap(etaI, etaT);
                const double Fr = EvaluateFresnelTerm(localWi, etaI, etaT);
                if (LocalCos(localWi) * LocalCos(localWo) >= 0) {
                    return S_R * Fr * shadingNormalCorrection;
                } else {
                    const double eta = etaI / etaT;
                    const double refrCorrection = transDir == EL ? eta : 1.0;
                    return S_R * (1.0 - Fr) * shadingNormalCorrection * refrCorrection * refrCorrection;
                }
            }
        }
    }
    return d3();  // assert(0), rt.hpp:1146 (e.g. a pure [L] primitive after `type & ~Emitter`)
}

// rt.hpp:1150-1336
double Primitive::EvaluateDirectionPDF(const SurfaceGeometry& geom, int queryType, const d3& wi, const d3& wo, bool forceDegenerated) const {
    if ((queryType & NGI_TYPE_L) > 0) {
        if (LType == NGI_L_AREA) {
            const d3 localWo = geom.ToLocal(wo);
            if (LocalCos(localWo) <= 0) return 0;
            return InvPi;
        }
        if (LType == NGI_L_POINT) return InvPi * 0.25;
        if (LType == NGI_L_DIRECTIONAL) return forceDegenerated ? 1 : 0;
    }
    if ((queryType & NGI_TYPE_E) > 0) {
        if (EType == NGI_E_AREA) {
            const d3 localWo = geom.ToLocal(wo);
            if (LocalCos(localWo) <= 0) return 0;
            return InvPi;
        }
        if (EType == NGI_E_PINHOLE) {
            d2 rasterPos;
            if (!RasterPosition(wo, geom, rasterPos)) return 0;
            const d3 woEye(dot(E_Vx, wo), dot(E_Vy, wo), dot(E_Vz, wo));
            const double tanFov = std::tan(E_Fov * 0.5);
            const double cosTheta = -LocalCos(woEye);
            const double invCosTheta = 1.0 / cosTheta;
            const double A = tanFov * tanFov * E_Aspect * 4.0;
            return invCosTheta * invCosTheta * invCosTheta / A;
        }
    }
    if ((queryType & NGI_TYPE_D) > 0) {
        const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
        if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return 0;
        return InvPi;
    }
    if ((queryType & NGI_TYPE_G) > 0) {
        const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
        if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return 0;
        const d3 H = normalize(localWi + localWo);
        const double D = EvaluateBechmannDist(H);
        return D * LocalCos(H) / (4.0 * dot(localWo, H)) / LocalCos(localWo);
    }
    if ((queryType & NGI_TYPE_S) > 0) {
        if (!forceDegenerated) return 0;
        if (SType == NGI_S_REFLECTION) {
            const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
            if (LocalCos(localWi) <= 0 || LocalCos(localWo) <= 0) return 0;
            return 1;
        }
        if (SType == NGI_S_REFRACTION) return 1;
        if (SType == NGI_S_FRESNEL) {
            const d3 localWi = geom.ToLocal(wi), localWo = geom.ToLocal(wo);
            double etaI = S_Eta1, etaT = S_Eta2;
            if (LocalCos(localWi) < 0) std::swap(etaI, etaT);
            const double Fr = EvaluateFresnelTerm(localWi, etaI, etaT);
            if (LocalCos(localWi) * LocalCos(localWo) >= 0) return Fr;
            return 1.0 - Fr;
        }
    }
    return 0;
}

// rt.hpp:1344-1399
bool Primitive::RasterPosition(const d3& wo, const SurfaceGeometry& geom, d2& rasterPos) const {
    if (EType == NGI_E_PINHOLE) {
        const d3 woEye(dot(E_Vx, wo), dot(E_Vy, wo), dot(E_Vz, wo));
        if (LocalCos(woEye) >= 0) return false;
        const double tanFov = std::tan(E_Fov * 0.5);
        rasterPos.x = (-woEye.x / woEye.z / tanFov / E_Aspect + 1.0) * 0.5;
        rasterPos.y = (-woEye.y / woEye.z / tanFov + 1.0) * 0.5;
        if (rasterPos.x < 0 || rasterPos.x > 1 || rasterPos.y < 0 || rasterPos.y > 1) return false;
        return true;
    }
    if (EType == NGI_E_AREA) { rasterPos = geom.uv; return true; }
    return false;
}
// rt.hpp:1407-1414
double Primitive::EvaluateBechmannDist(const d3& H) const {
    if (LocalCos(H) <= 0) return 0.0;
    const double ex = LocalTan(H) / G_Roughness;
    const double t1 = std::exp(-(ex * ex));
    const double t2 = (Pi * G_Roughness * G_Roughness * std::pow(LocalCos(H), 4.0));
    return t1 / t2;
}
// rt.hpp:1423-1431 — NOTE the reference computes wi_dot_H from `wo` (:1428-1429); replicated.
double Primitive::EvalauteShadowMaskingFunc(const d3& wi, const d3& wo, const d3& H) const {
    const double n_dot_H = LocalCos(H);
    const double n_dot_wo = LocalCos(wo);
    const double n_dot_wi = LocalCos(wi);
    const double wo_dot_H = std::abs(dot(wo, H));
    const double wi_dot_H = std::abs(dot(wo, H));
    return std::min(1.0, std::min(2.0 * n_dot_H * n_dot_wo / wo_dot_H, 2.0 * n_dot_H * n_dot_wi / wi_dot_H));
}
// rt.hpp:1433-1442
d3 Primitive::EvaluateFrConductor(double cosThetaI) const {
    const d3 eta = G_Eta, k = G_K;
    const d3 tmp = (eta * eta + k * k) * (cosThetaI * cosThetaI);
    const d3 rParl2 = (tmp - (eta * (2.0 * cosThetaI)) + 1.0) / (tmp + (eta * (2.0 * cosThetaI)) + 1.0);
    const d3 tmpF = eta * eta + k * k;
    const d3 rPerp2 = (tmpF - (eta * (2.0 * cosThetaI)) + cosThetaI * cosThetaI) / (tmpF + (eta * (2.0 * cosThetaI)) + cosThetaI * cosThetaI);
    return (rParl2 + rPerp2) * 0.5;
}
// rt.hpp:1450-1466 (the `1.0f` literal at :1454 is harmless: 1.0f == 1.0)
double Primitive::EvaluateFresnelTerm(const d3& localWi, double etaI, double etaT) const {
    const double wiDotN = LocalCos(localWi);
    const double eta = etaI / etaT;
    const double cosThetaTSq = 1.0 - eta * eta * (1.0f - wiDotN * wiDotN);
    if (cosThetaTSq <= 0) return 1;
    const double absCosThetaI = std::abs(wiDotN);
    const double absCosThetaT = std::sqrt(cosThetaTSq);
    const double rhoS = (etaI * absCosThetaI - etaT * absCosThetaT) / (etaI * absCosThetaI + etaT * absCosThetaT);
    const double rhoT = (etaI * absCosThetaT - etaT * absCosThetaI) / (etaI * absCosThetaT + etaT * absCosThetaI);
    return (rhoS * rhoS + rhoT * rhoT) * 0.5;
}

// ---------------------------------------------------------------------------------------------
// Scene::Intersect / Visible — include/nanogi/rt.hpp:2162-2261
// ---------------------------------------------------------------------------------------------
struct Ray { d3 o, d; };
struct Intersection { SurfaceGeometry geom; const Primitive* Prim = nullptr; uint32_t tri = 0; };

struct Counters { uint64_t extend = 0, shadow = 0; };

bool Intersect(const Scene& sc, const Ray& ray, Intersection& isect, float minT, float maxT) {
    // rt.hpp:2165-2178: double -> float ray
    const float o[3] = {(float)ray.o.x, (float)ray.o.y, (float)ray.o.z};
    const float d[3] = {(float)ray.d.x, (float)ray.d.y, (float)ray.d.z};
    float tfar, u, v; uint32_t tri;
    if (!sc.TraceClosest(o, d, minT, maxT, tfar, u, v, tri)) return false;  // rtcIntersect, :2182
    const Primitive* prim = &sc.Primitives[sc.triPrim[tri]];                // :2190-2194
    isect.Prim = prim; isect.tri = tri;
    isect.geom.p = ray.o + ray.d * (double)tfar;                             // :2197
    const d3 p1 = sc.P(tri, 0), p2 = sc.P(tri, 1), p3 = sc.P(tri, 2);        // :2200-2206
    isect.geom.gn = normalize(cross(p2 - p1, p3 - p1));
    const d3 n1 = sc.Nv(tri, 0), n2 = sc.Nv(tri, 1), n3 = sc.Nv(tri, 2);     // :2209-2218
    isect.geom.sn = normalize(n1 * (double)(1.0f - u - v) + n2 * (double)u + n3 * (double)v);
    if (std::isnan(isect.geom.sn.x) || std::isnan(isect.geom.sn.y) || std::isnan(isect.geom.sn.z)) isect.geom.sn = isect.geom.gn;
    if (!sc.uv.empty()) {                                                    // :2221-2227
        const d2 uv1 = sc.UV(tri, 0), uv2 = sc.UV(tri, 1), uv3 = sc.UV(tri, 2);
        const double w = (double)(1.0f - u - v);
        isect.geom.uv.x = uv1.x * w + uv2.x * (double)u + uv3.x * (double)v;
        isect.geom.uv.y = uv1.y * w + uv2.y * (double)u + uv3.y * (double)v;
    }
    isect.geom.degenerated = false;
    isect.geom.ComputeTangentSpace();                                        // :2233
    // dndu/dndv (:2236-2241) are consumed only by ptmnee — not on this path.
    return true;
}
bool Intersect(const Scene& sc, const Ray& ray, Intersection& isect) { return Intersect(sc, ray, isect, EpsF, InfF); }  // :2246-2249

// rt.hpp:2251-2261 (closest-hit query used as an occlusion test; the hit record is discarded,
// so an any-hit traversal returns the identical boolean)
bool Visible(const Scene& sc, const d3& p1, const d3& p2) {
    const d3 p1p2 = p2 - p1;
    const double p1p2L = length(p1p2);
    const d3 dd = p1p2 / p1p2L;
    const float o[3] = {(float)p1.x, (float)p1.y, (float)p1.z};
    const float d[3] = {(float)dd.x, (float)dd.y, (float)dd.z};
    return !sc.TraceAny(o, d, EpsF, (float)(p1p2L) * (1.0f - EpsF));
}

// rt.hpp:2364-2374
double GeometryTerm(const SurfaceGeometry& geom1, const SurfaceGeometry& geom2) {
    d3 p1p2 = geom2.p - geom1.p;
    const double p1p2L2 = dot(p1p2, p1p2);
    const double p1p2L = std::sqrt(p1p2L2);
    p1p2 = p1p2 / p1p2L;
    double t = 1.0;
    if (!geom1.degenerated) t *= std::abs(dot(geom1.sn, p1p2));
    if (!geom2.degenerated) t *= std::abs(dot(geom2.sn, -p1p2));
    return t / p1p2L2;
}

// ---------------------------------------------------------------------------------------------
// Samplers. MtSampler = include/nanogi/basic.hpp:419-434 (mt19937 + uniform_real_distribution),
// draws in the reference's call order. PhiloxSampler = the GPU module's counter-based streams
// (key = seed, counter = (sample index, vertex, block)) so that oracle and GPU consume the SAME
// uniforms for the same (sample, vertex, slot): used for sample-exact replay parity tests.
// ---------------------------------------------------------------------------------------------
struct MtSampler {
    std::mt19937 engine;
    std::uniform_real_distribution<double> distDouble;
    std::uniform_int_distribution<unsigned int> distUInt;
    void SetSeed(unsigned int seed) { engine.seed(seed); distDouble.reset(); distUInt.reset(); }
    double Next() { return distDouble(engine); }
    unsigned int NextUInt() { return distUInt(engine); }
    void begin_sample(int64_t) {}
    // Random::Next2D() is `glm::dvec2(Next(), Next())` (basic.hpp:424) and the renderers call
    // `SampleDirection(rng.Next2D(), rng.Next(), ...)` (src/nanogi.cpp:495): the order in which those draws happen is
    // unspecified by C++. g++ evaluates call arguments right to left, so in the reference AS BUILT the second component is
    // drawn before the first and uComp before the direction pair. The oracle follows that order — verified film-exactly
    // against the reference's own code (oracle/_ref, tests/test_reference_pin.py); statistically every order is the same.
    d2 Next2D() { d2 r; r.y = Next(); r.x = Next(); return r; }
    double sensor_pick(int = 0) { return Next(); }
    d2 sensor_pos(int = 0) { return Next2D(); }
    double light_pick(int) { return Next(); }
    d2 light_pos(int) { return Next2D(); }
    void dir_ucomp(int, d2& u, double& uc) { uc = Next(); u = Next2D(); }
    double rr(int) { return Next(); }
    // bdpt subpaths (kind 0 = light subpath, 1 = eye subpath): plain sequential draws in the reference's call order
    double bd_emitter_pick(int) { return Next(); }
    d2 bd_emitter_pos(int) { return Next2D(); }
    void bd_dir_ucomp(int, int, d2& u, double& uc) { uc = Next(); u = Next2D(); }
    double bd_rr(int, int) { return Next(); }
};

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline double u01_24(uint32_t x) { return (double)((float)(x >> 8) * (1.0f / 16777216.0f)); }

struct PhiloxSampler {
    uint32_t key[2] = {0, 0};
    int64_t sample = 0;
    int cachedVertexA = -1, cachedVertexB = -1, cachedVertexC = -1;
    uint32_t a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c4[4] = {0, 0, 0, 0};
    void begin_sample(int64_t s) { sample = s; cachedVertexA = cachedVertexB = cachedVertexC = -1; }
    void blockC(int v) {
        if (cachedVertexC == v) return;
        uint32_t c[4] = {(uint32_t)sample, (uint32_t)((uint64_t)sample >> 32), (uint32_t)v, 2u};
        philox4x32_10(c, key, c4); cachedVertexC = v;
    }
    void blockA(int v) {
        if (cachedVertexA == v) return;
        uint32_t c[4] = {(uint32_t)sample, (uint32_t)((uint64_t)sample >> 32), (uint32_t)v, 0u};
        philox4x32_10(c, key, a); cachedVertexA = v;
    }
    void blockB(int v) {
        if (cachedVertexB == v) return;
        uint32_t c[4] = {(uint32_t)sample, (uint32_t)((uint64_t)sample >> 32), (uint32_t)v, 1u};
        philox4x32_10(c, key, b); cachedVertexB = v;
    }
    // block 2 = {sensor pick, sensor position u0, u1, -}: values only matter for an E.area sensor (rt.hpp:573-577);
    // pt / ptdirect draw them once per sample (vertex 0), ltdirect once per path vertex
    double sensor_pick(int v = 0) { blockC(v); return u01_24(c4[0]); }
    d2 sensor_pos(int v = 0) { blockC(v); d2 r; r.x = u01_24(c4[1]); r.y = u01_24(c4[2]); return r; }
    double light_pick(int v) { blockB(v); return u01_24(b[0]); }
    d2 light_pos(int v) { blockB(v); d2 r; r.x = u01_24(b[1]); r.y = u01_24(b[2]); return r; }
    void dir_ucomp(int v, d2& u, double& uc) { blockA(v); u.x = u01_24(a[0]); u.y = u01_24(a[1]); uc = u01_24(a[2]); }
    // bdpt: the light subpath uses block 1 (emitter) and block 0 (directions, RR) like lt; the eye subpath block 2 (emitter)
    // and block 3 (directions, RR) — the two subpaths of one sample index must not share uniforms
    uint32_t d4[4] = {0, 0, 0, 0};
    void blockD(int v) { uint32_t c[4] = {(uint32_t)sample, (uint32_t)((uint64_t)sample >> 32), (uint32_t)v, 3u}; philox4x32_10(c, key, d4); }
    double bd_emitter_pick(int kind) { return kind == 0 ? light_pick(0) : sensor_pick(0); }
    d2 bd_emitter_pos(int kind) { return kind == 0 ? light_pos(0) : sensor_pos(0); }
    void bd_dir_ucomp(int kind, int v, d2& u, double& uc) {
        if (kind == 0) { dir_ucomp(v, u, uc); return; }
        blockD(v); u.x = u01_24(d4[0]); u.y = u01_24(d4[1]); uc = u01_24(d4[2]);
    }
    double bd_rr(int kind, int v) { if (kind == 0) return rr(v); blockD(v); return u01_24(d4[3]); }
    double rr(int v) { blockA(v); return u01_24(a[3]); }
};

struct RenderParams { int Width, Height, MaxNumVertices; };

// ---------------------------------------------------------------------------------------------
// ProcessSample_PT — src/nanogi.cpp:446-607
// ---------------------------------------------------------------------------------------------
template <class S>
void ProcessSample_PT(const Scene& scene, const RenderParams& Params, S& rng, std::vector<d3>& film, Counters& cnt) {
    const Primitive* E = scene.SampleEmitter(NGI_TYPE_E, rng.sensor_pick());        // :450
    const double pdfE = scene.EvaluateEmitterPDF(E);
    SurfaceGeometry geomE;
    E->SamplePosition(rng.sensor_pos(), geomE);                                      // :461
    const double pdfPE = E->EvaluatePositionPDF(geomE, true);
    d3 throughput = E->EvaluatePosition(geomE, true) / pdfPE / pdfE;                 // :471
    const Primitive* prim = E;
    int type = NGI_TYPE_E;
    SurfaceGeometry geom = geomE;
    d3 wi;
    int pixelIndex = -1;
    int numVertices = 1;
    while (true) {
        if (Params.MaxNumVertices != -1 && numVertices >= Params.MaxNumVertices) break;      // :485
        const int vtx = numVertices - 1;
        d3 wo;                                                                                // zero-initialised (glm < 0.9.9)
        d2 u; double uc; rng.dir_ucomp(vtx, u, uc);
        prim->SampleDirection(u, uc, type, geom, wi, wo);                                     // :495
        const double pdfD = prim->EvaluateDirectionPDF(geom, type, wi, wo, true);             // :496
        if (type == NGI_TYPE_E) {                                                             // :504-523
            d2 rasterPos;
            if (!prim->RasterPosition(wo, geom, rasterPos)) break;
            pixelIndex = PixelIndex(rasterPos, Params.Width, Params.Height);
        }
        const d3 fs = prim->EvaluateDirection(geom, type, wi, wo, EL, true);                  // :531
        if (is_zero(fs)) break;
        throughput = throughput * (fs / pdfD);                                                // :544
        Ray ray{geom.p, wo};                                                                  // :553
        Intersection isect;
        cnt.extend++;
        if (!Intersect(scene, ray, isect)) break;                                             // :557
        if ((isect.Prim->Type & NGI_TYPE_L) > 0) {                                            // :568-575
            const d3 c = throughput
                * isect.Prim->EvaluateDirection(isect.geom, NGI_TYPE_L, d3(), -ray.d, EL, false)
                * isect.Prim->EvaluatePosition(isect.geom, false);
            film[pixelIndex] = film[pixelIndex] + c;
        }
        const double rrProb = 0.5;                                                            // :583-591
        if (rng.rr(vtx) > rrProb) break;
        throughput = throughput / rrProb;
        geom = isect.geom;                                                                    // :599-603
        prim = isect.Prim;
        type = isect.Prim->Type & ~NGI_TYPE_EMITTER;
        wi = -ray.d;
        numVertices++;
    }
}

// ---------------------------------------------------------------------------------------------
// ProcessSample_PTDirect — src/nanogi.cpp:609-802
// ---------------------------------------------------------------------------------------------
template <class S>
void ProcessSample_PTDirect(const Scene& scene, const RenderParams& Params, S& rng, std::vector<d3>& film, Counters& cnt) {
    const Primitive* E = scene.SampleEmitter(NGI_TYPE_E, rng.sensor_pick());        // :613
    const double pdfE = scene.EvaluateEmitterPDF(E);
    SurfaceGeometry geomE;
    E->SamplePosition(rng.sensor_pos(), geomE);                                      // :624
    const double pdfPE = E->EvaluatePositionPDF(geomE, true);
    d3 throughput = E->EvaluatePosition(geomE, true) / pdfPE / pdfE;                 // :633
    const Primitive* prim = E;
    int type = NGI_TYPE_E;
    SurfaceGeometry geom = geomE;
    d3 wi;
    int pixelIndex = -1;
    int numVertices = 1;
    while (true) {
        if (Params.MaxNumVertices != -1 && numVertices >= Params.MaxNumVertices) break;      // :647
        const int vtx = numVertices - 1;
        // ---- direct light sampling, :654-712 ----
        if (!scene.LightPrimitiveIndices.empty()) {
            const Primitive* L = scene.SampleEmitter(NGI_TYPE_L, rng.light_pick(vtx));        // :659
            const double pdfL = scene.EvaluateEmitterPDF(L);
            SurfaceGeometry geomL;
            L->SamplePosition(rng.light_pos(vtx), geomL);                                     // :670
            const double pdfPL = L->EvaluatePositionPDF(geomL, true);
            const d3 ppL = normalize(geomL.p - geom.p);                                       // :680
            const d3 fsE = prim->EvaluateDirection(geom, type, wi, ppL, EL, false);           // :681
            const d3 fsL = L->EvaluateDirection(geomL, NGI_TYPE_L, d3(), -ppL, LE, false);    // :682
            const double G = GeometryTerm(geom, geomL);                                       // :683
            cnt.shadow++;
            const double V = Visible(scene, geom.p, geomL.p) ? 1.0 : 0.0;                     // :684 (always traced)
            const d3 LeP = L->EvaluatePosition(geomL, true);                                  // :685
            const d3 C = throughput * fsE * G * V * fsL * LeP / pdfL / pdfPL;                 // :686
            if (!is_zero(C)) {                                                                // :694-707
                int index = pixelIndex;
                if (type == NGI_TYPE_E) {
                    d2 rasterPos;
                    prim->RasterPosition(ppL, geom, rasterPos);
                    index = PixelIndex(rasterPos, Params.Width, Params.Height);
                }
                film[index] = film[index] + C;
            }
        }
        // ---- sample next direction, :716-754 ----
        d3 wo;
        d2 u; double uc; rng.dir_ucomp(vtx, u, uc);
        prim->SampleDirection(u, uc, type, geom, wi, wo);                                     // :719
        const double pdfD = prim->EvaluateDirectionPDF(geom, type, wi, wo, true);             // :720
        if (type == NGI_TYPE_E) {                                                             // :728-733
            d2 rasterPos;
            if (!prim->RasterPosition(wo, geom, rasterPos)) break;
            pixelIndex = PixelIndex(rasterPos, Params.Width, Params.Height);
        }
        const d3 fs = prim->EvaluateDirection(geom, type, wi, wo, EL, true);                  // :741
        if (is_zero(fs)) break;
        throughput = throughput * (fs / pdfD);                                                // :754
        Ray ray{geom.p, wo};                                                                  // :763
        Intersection isect;
        cnt.extend++;
        if (!Intersect(scene, ray, isect)) break;                                             // :767
        const double rrProb = 0.5;                                                            // :778-786
        if (rng.rr(vtx) > rrProb) break;
        throughput = throughput / rrProb;
        geom = isect.geom;                                                                    // :794-798
        prim = isect.Prim;
        type = isect.Prim->Type & ~NGI_TYPE_EMITTER;
        wi = -ray.d;
        numVertices++;
    }
}

// ---------------------------------------------------------------------------------------------
// ProcessSample_LT — src/nanogi.cpp:804-943. Light tracing: the film only receives paths that HIT a sensor
// primitive, i.e. nothing with a pinhole (no mesh); meaningful with an E.area sensor.
// ---------------------------------------------------------------------------------------------
template <class S>
void ProcessSample_LT(const Scene& scene, const RenderParams& Params, S& rng, std::vector<d3>& film, Counters& cnt) {
    if (scene.LightPrimitiveIndices.empty()) return;
    const Primitive* L = scene.SampleEmitter(NGI_TYPE_L, rng.light_pick(0));         // :808
    const double pdfL = scene.EvaluateEmitterPDF(L);
    SurfaceGeometry geomL;
    L->SamplePosition(rng.light_pos(0), geomL);                                       // :819
    const double pdfPL = L->EvaluatePositionPDF(geomL, true);
    d3 throughput = L->EvaluatePosition(geomL, true) / pdfPL / pdfL;                  // :830
    const Primitive* prim = L;
    int type = NGI_TYPE_L;
    SurfaceGeometry geom = geomL;
    d3 wi;
    int numVertices = 1;
    while (true) {
        if (Params.MaxNumVertices != -1 && numVertices >= Params.MaxNumVertices) break;      // :842
        const int vtx = numVertices - 1;
        d3 wo;
        d2 u; double uc; rng.dir_ucomp(vtx, u, uc);
        prim->SampleDirection(u, uc, type, geom, wi, wo);                                     // :852
        const double pdfD = prim->EvaluateDirectionPDF(geom, type, wi, wo, true);             // :853
        const d3 fs = prim->EvaluateDirection(geom, type, wi, wo, LE, true);                  // :861
        if (is_zero(fs)) break;
        throughput = throughput * (fs / pdfD);                                                // :874
        Ray ray{geom.p, wo};                                                                  // :883
        Intersection isect;
        cnt.extend++;
        if (!Intersect(scene, ray, isect)) break;                                             // :887
        if ((isect.Prim->Type & NGI_TYPE_E) > 0) {                                            // :899-920
            d2 rasterPos;
            if (!isect.Prim->RasterPosition(-wo, isect.geom, rasterPos)) break;
            const int pixelIndex = PixelIndex(rasterPos, Params.Width, Params.Height);
            const d3 c = throughput
                * isect.Prim->EvaluateDirection(isect.geom, NGI_TYPE_E, d3(), -ray.d, LE, false)
                * isect.Prim->EvaluatePosition(isect.geom, false);
            film[pixelIndex] = film[pixelIndex] + c;
        }
        const double rrProb = 0.5;                                                            // :929-937
        if (rng.rr(vtx) > rrProb) break;
        throughput = throughput / rrProb;
        geom = isect.geom;                                                                    // :945-949
        prim = isect.Prim;
        type = isect.Prim->Type & ~NGI_TYPE_EMITTER;
        wi = -ray.d;
        numVertices++;
    }
}

// ---------------------------------------------------------------------------------------------
// ProcessSample_LTDirect — src/nanogi.cpp:955-1131. Every light-path vertex (the light vertex included) is
// connected to a sampled sensor position. NOTE `pdfPE = L->EvaluatePositionPDF(geomE, true)` (:1017): the
// reference asks the LIGHT primitive, not E — i.e. InvArea of the light for an area light, 1 for a point
// light. Replicated: it is part of the reference's output.
// ---------------------------------------------------------------------------------------------
template <class S>
void ProcessSample_LTDirect(const Scene& scene, const RenderParams& Params, S& rng, std::vector<d3>& film, Counters& cnt) {
    if (scene.LightPrimitiveIndices.empty()) return;
    const Primitive* L = scene.SampleEmitter(NGI_TYPE_L, rng.light_pick(0));         // :959
    const double pdfL = scene.EvaluateEmitterPDF(L);
    SurfaceGeometry geomL;
    L->SamplePosition(rng.light_pos(0), geomL);                                       // :970
    const double pdfPL = L->EvaluatePositionPDF(geomL, true);
    d3 throughput = L->EvaluatePosition(geomL, true) / pdfPL / pdfL;                  // :980
    const Primitive* prim = L;
    int type = NGI_TYPE_L;
    SurfaceGeometry geom = geomL;
    d3 wi;
    int numVertices = 1;
    while (true) {
        if (Params.MaxNumVertices != -1 && numVertices >= Params.MaxNumVertices) break;      // :993
        const int vtx = numVertices - 1;
        {   // ---- direct sensor sampling, :1000-1052 ----
            const Primitive* E = scene.SampleEmitter(NGI_TYPE_E, rng.sensor_pick(vtx));       // :1005
            const double pdfE = scene.EvaluateEmitterPDF(E);
            SurfaceGeometry geomE;
            E->SamplePosition(rng.sensor_pos(vtx), geomE);                                    // :1016
            const double pdfPE = L->EvaluatePositionPDF(geomE, true);                         // :1017 (sic: L)
            const d3 ppE = normalize(geomE.p - geom.p);                                       // :1026
            const d3 fsL = prim->EvaluateDirection(geom, type, wi, ppE, LE, false);           // :1027
            const d3 fsE = E->EvaluateDirection(geomE, NGI_TYPE_E, d3(), -ppE, EL, false);    // :1028
            const double G = GeometryTerm(geom, geomE);                                       // :1029
            cnt.shadow++;
            const double V = Visible(scene, geom.p, geomE.p) ? 1.0 : 0.0;                     // :1030
            const d3 LeP = L->EvaluatePosition(geomE, true);                                  // :1031
            const d3 C = throughput * fsL * G * V * fsE * LeP / pdfE / pdfPE;                 // :1032
            if (!is_zero(C)) {                                                                // :1040-1049
                d2 rasterPos;
                E->RasterPosition(-ppE, geomE, rasterPos);
                const int index = PixelIndex(rasterPos, Params.Width, Params.Height);
                film[index] = film[index] + C;
            }
        }
        d3 wo;
        d2 u; double uc; rng.dir_ucomp(vtx, u, uc);
        prim->SampleDirection(u, uc, type, geom, wi, wo);                                     // :1061
        const double pdfD = prim->EvaluateDirectionPDF(geom, type, wi, wo, true);             // :1062
        const d3 fs = prim->EvaluateDirection(geom, type, wi, wo, LE, true);                  // :1070
        if (is_zero(fs)) break;
        throughput = throughput * (fs / pdfD);                                                // :1083
        Ray ray{geom.p, wo};                                                                  // :1092
        Intersection isect;
        cnt.extend++;
        if (!Intersect(scene, ray, isect)) break;                                             // :1096
        const double rrProb = 0.5;                                                            // :1107-1115
        if (rng.rr(vtx) > rrProb) break;
        throughput = throughput / rrProb;
        geom = isect.geom;                                                                    // :1123-1127
        prim = isect.Prim;
        type = isect.Prim->Type & ~NGI_TYPE_EMITTER;
        wi = -ray.d;
        numVertices++;
    }
}

// ---------------------------------------------------------------------------------------------
// bdpt — include/nanogi/bdpt.hpp:38-539 (PathVertex, Path) and ProcessSample_BDPT, src/nanogi.cpp:1133-1186.
// SURVEY §8f row 4 ("then bdpt"). Restated function by function; the weight actually used by the reference is
// EvaluatePowerHeuristicsMISWeightOpt (bdpt.hpp:184), the O(n) per-strategy EvaluatePDF products.
// ---------------------------------------------------------------------------------------------
struct PathVertex { int type = 0; SurfaceGeometry geom; const Primitive* primitive = nullptr; };   // bdpt.hpp:38-43

struct Path {
    std::vector<PathVertex> vertices;

    // bdpt.hpp:54-123
    template <class S>
    void SampleSubpath(const Scene& scene, S& rng, int kind, TransportDirection transDir, int maxPathVertices, Counters& cnt) {
        PathVertex v;
        vertices.clear();
        for (int step = 0; maxPathVertices == -1 || step < maxPathVertices; step++) {
            if (step == 0) {
                const int type = transDir == LE ? NGI_TYPE_L : NGI_TYPE_E;
                const Primitive* emitter = scene.SampleEmitter(type, rng.bd_emitter_pick(kind));     // :63
                v.primitive = emitter;
                v.type = type;
                emitter->SamplePosition(rng.bd_emitter_pos(kind), v.geom);                            // :66
                vertices.push_back(v);
            } else {
                const PathVertex* pv = &vertices.back();
                const PathVertex* ppv = vertices.size() > 1 ? &vertices[vertices.size() - 2] : nullptr;
                d3 wo;
                const d3 wi = ppv ? normalize(ppv->geom.p - pv->geom.p) : d3();                       // :77
                d2 u; double uc; rng.bd_dir_ucomp(kind, step - 1, u, uc);
                pv->primitive->SampleDirection(u, uc, pv->type, pv->geom, wi, wo);                    // :78
                const d3 f = pv->primitive->EvaluateDirection(pv->geom, pv->type, wi, wo, transDir, true);   // :81
                if (is_zero(f)) break;
                Ray ray{pv->geom.p, wo};
                Intersection isect;
                cnt.extend++;
                if (!Intersect(scene, ray, isect)) break;                                             // :92
                v.geom = isect.geom;                                                                  // :100-102
                v.primitive = isect.Prim;
                v.type = isect.Prim->Type & ~NGI_TYPE_EMITTER;
                const double rrProb = 0.5;                                                            // :108-113
                if (rng.bd_rr(kind, step - 1) > rrProb) { vertices.push_back(v); break; }
                vertices.push_back(v);
            }
        }
    }

    // bdpt.hpp:125-177
    bool Connect(const Scene& scene, int s, int t, const Path& subpathL, const Path& subpathE, Counters& cnt) {
        vertices.clear();
        if (s == 0 && t > 0) {
            if ((subpathE.vertices[t - 1].primitive->Type & NGI_TYPE_L) == 0) return false;
            for (int i = t - 1; i >= 0; i--) vertices.push_back(subpathE.vertices[i]);
            vertices.front().type = NGI_TYPE_L;
        } else if (s > 0 && t == 0) {
            if ((subpathL.vertices[s - 1].primitive->Type & NGI_TYPE_E) == 0) return false;
            for (int i = 0; i < s; i++) vertices.push_back(subpathL.vertices[i]);
            vertices.back().type = NGI_TYPE_E;
        } else {
            cnt.shadow++;
            if (!Visible(scene, subpathL.vertices[s - 1].geom.p, subpathE.vertices[t - 1].geom.p)) return false;
            for (int i = 0; i < s; i++) vertices.push_back(subpathL.vertices[i]);
            for (int i = t - 1; i >= 0; i--) vertices.push_back(subpathE.vertices[i]);
        }
        return true;
    }

    d3 dirTo(int from, int to) const { return normalize(vertices[to].geom.p - vertices[from].geom.p); }

    // bdpt.hpp:187-205
    double SelectionProb(int s) const {
        const double rrProb = 0.5;
        const int n = (int)vertices.size();
        const int t = n - s;
        double selectionProb = 1;
        for (int i = 1; i < s - 1; i++) selectionProb *= rrProb;
        for (int i = t - 2; i >= 1; i--) selectionProb *= rrProb;
        return selectionProb;
    }
    // bdpt.hpp:207-215
    d2 RasterPosition() const {
        const PathVertex& v = vertices[vertices.size() - 1];
        d2 rasterPos;
        v.primitive->RasterPosition(dirTo((int)vertices.size() - 1, (int)vertices.size() - 2), v.geom, rasterPos);
        return rasterPos;
    }
    // bdpt.hpp:217-250
    d3 EvaluateCst(int s) const {
        const int n = (int)vertices.size();
        const int t = n - s;
        d3 cst;
        if (s == 0 && t > 0) {
            const PathVertex& v = vertices[0];
            cst = v.primitive->EvaluatePosition(v.geom, false) * v.primitive->EvaluateDirection(v.geom, v.type, d3(), dirTo(0, 1), EL, false);
        } else if (s > 0 && t == 0) {
            const PathVertex& v = vertices[n - 1];
            cst = v.primitive->EvaluatePosition(v.geom, false) * v.primitive->EvaluateDirection(v.geom, v.type, d3(), dirTo(n - 1, n - 2), LE, false);
        } else if (s > 0 && t > 0) {
            const PathVertex* vL = &vertices[s - 1];
            const PathVertex* vE = &vertices[s];
            const d3 fsL = vL->primitive->EvaluateDirection(vL->geom, vL->type, s - 2 >= 0 ? dirTo(s - 1, s - 2) : d3(), dirTo(s - 1, s), LE, false);
            const d3 fsE = vE->primitive->EvaluateDirection(vE->geom, vE->type, s + 1 < n ? dirTo(s, s + 1) : d3(), dirTo(s, s - 1), EL, false);
            const double G = GeometryTerm(vL->geom, vE->geom);
            cst = fsL * G * fsE;
        }
        return cst;
    }
    static d3 LocalContrb(const d3& f, double p) { return is_zero(f) ? d3() : f / p; }   // bdpt.hpp:258-263
    // bdpt.hpp:252-343
    d3 EvaluateUnweightContribution(const Scene& scene, int s) const {
        const int n = (int)vertices.size();
        const int t = n - s;
        d3 alphaL;
        if (s == 0) alphaL = d3(1);
        else {
            const PathVertex& v0 = vertices[0];
            alphaL = LocalContrb(v0.primitive->EvaluatePosition(v0.geom, true), v0.primitive->EvaluatePositionPDF(v0.geom, true) * scene.EvaluateEmitterPDF(v0.primitive));
            for (int i = 0; i < s - 1; i++) {
                const PathVertex* v = &vertices[i];
                const d3 wi = i >= 1 ? dirTo(i, i - 1) : d3();
                const d3 wo = dirTo(i, i + 1);
                alphaL = alphaL * LocalContrb(v->primitive->EvaluateDirection(v->geom, v->type, wi, wo, LE, true), v->primitive->EvaluateDirectionPDF(v->geom, v->type, wi, wo, true));
            }
        }
        if (is_zero(alphaL)) return d3();
        d3 alphaE;
        if (t == 0) alphaE = d3(1);
        else {
            const PathVertex& vn = vertices[n - 1];
            alphaE = LocalContrb(vn.primitive->EvaluatePosition(vn.geom, true), vn.primitive->EvaluatePositionPDF(vn.geom, true) * scene.EvaluateEmitterPDF(vn.primitive));
            for (int i = n - 1; i > s; i--) {
                const PathVertex* v = &vertices[i];
                const d3 wi = i < n - 1 ? dirTo(i, i + 1) : d3();
                const d3 wo = dirTo(i, i - 1);
                alphaE = alphaE * LocalContrb(v->primitive->EvaluateDirection(v->geom, v->type, wi, wo, EL, true), v->primitive->EvaluateDirectionPDF(v->geom, v->type, wi, wo, true));
            }
        }
        if (is_zero(alphaE)) return d3();
        const d3 cst = EvaluateCst(s);
        if (is_zero(cst)) return d3();
        return alphaL * cst * alphaE;
    }
    // bdpt.hpp:491-535
    double EvaluatePDF(const Scene& scene, int s) const {
        if (is_zero(EvaluateCst(s))) return 0;
        double pdf = 1;
        const int n = (int)vertices.size();
        const int t = n - s;
        if (s > 0) {
            pdf *= vertices[0].primitive->EvaluatePositionPDF(vertices[0].geom, true) * scene.EvaluateEmitterPDF(vertices[0].primitive);
            for (int i = 0; i < s - 1; i++) {
                const PathVertex* vi = &vertices[i];
                pdf *= vi->primitive->EvaluateDirectionPDF(vi->geom, vi->type, i - 1 >= 0 ? dirTo(i, i - 1) : d3(), dirTo(i, i + 1), true);
                pdf *= GeometryTerm(vi->geom, vertices[i + 1].geom);
            }
        }
        if (t > 0) {
            pdf *= vertices[n - 1].primitive->EvaluatePositionPDF(vertices[n - 1].geom, true) * scene.EvaluateEmitterPDF(vertices[n - 1].primitive);
            for (int i = n - 1; i >= s + 1; i--) {
                const PathVertex* vi = &vertices[i];
                pdf *= vi->primitive->EvaluateDirectionPDF(vi->geom, vi->type, i + 1 < n ? dirTo(i, i + 1) : d3(), dirTo(i, i - 1), true);
                pdf *= GeometryTerm(vi->geom, vertices[i - 1].geom);
            }
        }
        return pdf;
    }
    // bdpt.hpp:362-380
    double EvaluatePowerHeuristicsMISWeightOpt(const Scene& scene, int s) const {
        double invWeight = 0;
        const int n = (int)vertices.size();
        const double ps = EvaluatePDF(scene, s);
        for (int i = 0; i <= n; i++) {
            const double pi = EvaluatePDF(scene, i);
            if (pi > 0) { const double r = pi / ps; invWeight += r * r; }
        }
        return 1.0 / invWeight;
    }
    // bdpt.hpp:181-185
    d3 EvaluateContribution(const Scene& scene, int s) const {
        const d3 Cstar = EvaluateUnweightContribution(scene, s);
        return is_zero(Cstar) ? d3() : Cstar * EvaluatePowerHeuristicsMISWeightOpt(scene, s);
    }
};

// ProcessSample_BDPT — src/nanogi.cpp:1133-1186
template <class S>
void ProcessSample_BDPT(const Scene& scene, const RenderParams& Params, S& rng, std::vector<d3>& film, Counters& cnt) {
    if (scene.LightPrimitiveIndices.empty()) return;
    Path subpathL, subpathE, path;
    subpathL.SampleSubpath(scene, rng, 0, LE, Params.MaxNumVertices, cnt);                // :1137
    subpathE.SampleSubpath(scene, rng, 1, EL, Params.MaxNumVertices, cnt);                // :1138
    const int nL = (int)subpathL.vertices.size(), nE = (int)subpathE.vertices.size();
    for (int n = 2; n <= nE + nL; n++) {                                                   // :1148
        if (Params.MaxNumVertices != -1 && n > Params.MaxNumVertices) continue;
        const int minS = std::max(0, n - nE), maxS = std::min(nL, n);
        for (int s = minS; s <= maxS; s++) {
            const int t = n - s;
            if (!path.Connect(scene, s, t, subpathL, subpathE, cnt)) continue;             // :1164
            const d3 C = path.EvaluateContribution(scene, s) / path.SelectionProb(s);      // :1172
            if (is_zero(C)) continue;
            const int px = PixelIndex(path.RasterPosition(), Params.Width, Params.Height); // :1180
            film[px] = film[px] + C;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Scene construction: the loader-side derivations of include/nanogi/rt.hpp:1606-1615 (sensor =
// last E primitive, light list), :1747-1765 (area CDF), :2067-2073 (directional disk)
// ---------------------------------------------------------------------------------------------
Scene* make_scene(const NgiSceneDesc* desc) {
    std::unique_ptr<Scene> sc(new Scene);
    const size_t n = (size_t)desc->num_tris;
    sc->pos.assign(desc->positions, desc->positions + n * 9);
    sc->nrm.assign(desc->normals, desc->normals + n * 9);
    if (desc->texcoords) sc->uv.assign(desc->texcoords, desc->texcoords + n * 6);
    sc->triPrim.assign(n, -1);
    sc->Textures.resize(desc->num_textures);
    for (uint32_t i = 0; i < desc->num_textures; i++) {
        sc->Textures[i].Width = desc->textures[i].width;
        sc->Textures[i].Height = desc->textures[i].height;
        sc->Textures[i].Data.assign(desc->textures[i].rgb, desc->textures[i].rgb + (size_t)3 * desc->textures[i].width * desc->textures[i].height);
    }
    sc->Primitives.resize(desc->num_prims);
    d3 bmin(std::numeric_limits<double>::max()), bmax(-std::numeric_limits<double>::max());
    for (size_t i = 0; i < n * 3; i++) {
        const float* p = &sc->pos[i * 3];
        bmin = d3(std::min(bmin.x, (double)p[0]), std::min(bmin.y, (double)p[1]), std::min(bmin.z, (double)p[2]));
        bmax = d3(std::max(bmax.x, (double)p[0]), std::max(bmax.y, (double)p[1]), std::max(bmax.z, (double)p[2]));
    }
    for (uint32_t i = 0; i < desc->num_prims; i++) {
        const NgiPrimitive& s = desc->prims[i];
        Primitive& p = sc->Primitives[i];
        p.scene = sc.get();
        p.Type = s.type; p.firstTri = s.first_tri; p.numTris = s.first_tri >= 0 ? s.num_tris : 0;
        p.LType = s.l_type; p.EType = s.e_type; p.SType = s.s_type;
        p.D_R = d3(s.d_r); p.G_R = d3(s.g_r); p.G_Eta = d3(s.g_eta); p.G_K = d3(s.g_k); p.G_Roughness = s.g_roughness;
        p.S_R = d3(s.s_r); p.S_Eta1 = s.s_eta1; p.S_Eta2 = s.s_eta2;
        p.L_Le = d3(s.l_le); p.L_vec = d3(s.l_vec);
        p.E_Position = d3(s.e_position); p.E_Vx = d3(s.e_vx); p.E_Vy = d3(s.e_vy); p.E_Vz = d3(s.e_vz);
        p.E_Fov = s.e_fov; p.E_Aspect = s.e_aspect; p.E_We = d3(s.e_we);
        if (s.d_tex >= 0 && (uint32_t)s.d_tex < desc->num_textures) p.D_TexR = &sc->Textures[s.d_tex];
        if (s.g_tex >= 0 && (uint32_t)s.g_tex < desc->num_textures) p.G_TexR = &sc->Textures[s.g_tex];
        for (int t = 0; t < p.numTris; t++) sc->triPrim[p.firstTri + t] = (int)i;
        if ((p.Type & NGI_TYPE_E) > 0) sc->SensorPrimitiveIndex = i;              // rt.hpp:1606-1610
        if ((p.Type & NGI_TYPE_L) > 0) sc->LightPrimitiveIndices.push_back(i);    // rt.hpp:1612-1615
        const bool areaL = (p.Type & NGI_TYPE_L) && p.LType == NGI_L_AREA;
        const bool areaE = (p.Type & NGI_TYPE_E) && p.EType == NGI_E_AREA;
        if ((areaL || areaE) && p.numTris > 0) {                                  // rt.hpp:1747-1765
            double sumArea = 0;
            for (int t = 0; t < p.numTris; t++) {
                const d3 p1 = sc->P(p.firstTri + t, 0), p2 = sc->P(p.firstTri + t, 1), p3 = sc->P(p.firstTri + t, 2);
                const double area = length(cross(p2 - p1, p3 - p1)) * 0.5;
                p.Dist.Add(area);
                sumArea += area;
            }
            p.Dist.Normalize();
            if (areaL) p.L_InvArea = 1.0 / sumArea; else p.E_InvArea = 1.0 / sumArea;
        }
    }
    for (auto& p : sc->Primitives) {                                              // rt.hpp:2059-2079
        if ((p.Type & NGI_TYPE_L) > 0 && p.LType == NGI_L_DIRECTIONAL) {
            p.L_Center = (bmax + bmin) * 0.5;
            p.L_Radius = length(p.L_Center - bmax) * 1.01;
            p.L_InvArea = 1.0 / (2.0 * Pi * p.L_Radius * p.L_Radius);
        }
    }
    sc->BuildBVH();                                                               // rt.hpp:2085-2143
    return sc.release();
}

// ---------------------------------------------------------------------------------------------
// RenderProcess — src/nanogi.cpp:225-440: chunks of GrainSize over [0, NumSamples), lazily
// initialised per-thread Context {rng, film}, final film = sum of thread films * W*H/processed.
// (TBB is replaced by std::thread workers pulling chunks from an atomic counter.)
// ---------------------------------------------------------------------------------------------
template <class S, class SeedFn>
void RenderProcess(const Scene& scene, int renderer, const RenderParams& Params, int64_t NumSamples, int64_t sampleOffset,
                   int64_t normSamples, int numThreads, SeedFn seedSampler, double* filmOut, Counters& total) {
    const int64_t GrainSize = 10000;                                              // src/nanogi.cpp:2014
    const size_t npx = (size_t)Params.Width * Params.Height;
    std::atomic<int64_t> next(0);
    std::mutex contextInitMutex;
    int currentThreadID = 0;
    std::vector<std::vector<d3>> films(numThreads);
    std::vector<Counters> cnts(numThreads);
    auto worker = [&](int slot) {
        S rng; bool init = false;
        std::vector<d3>& film = films[slot];
        while (true) {
            const int64_t begin = next.fetch_add(GrainSize);
            if (begin >= NumSamples) break;
            const int64_t end = std::min(begin + GrainSize, NumSamples);
            if (!init) {                                                          // :292-299
                std::unique_lock<std::mutex> lock(contextInitMutex);
                seedSampler(rng, currentThreadID++);
                film.assign(npx, d3());
                init = true;
            }
            for (int64_t sample = begin; sample != end; sample++) {               // :307-318
                rng.begin_sample(sampleOffset + sample);
                if (renderer == NGI_RENDERER_PT) ProcessSample_PT(scene, Params, rng, film, cnts[slot]);
                else if (renderer == NGI_RENDERER_PTDIRECT) ProcessSample_PTDirect(scene, Params, rng, film, cnts[slot]);
                else if (renderer == NGI_RENDERER_LT) ProcessSample_LT(scene, Params, rng, film, cnts[slot]);
                else if (renderer == NGI_RENDERER_LTDIRECT) ProcessSample_LTDirect(scene, Params, rng, film, cnts[slot]);
                else ProcessSample_BDPT(scene, Params, rng, film, cnts[slot]);
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < numThreads; i++) th.emplace_back(worker, i);
    worker(0);
    for (auto& t : th) t.join();
    // gather, :429-437
    std::vector<d3> film(npx);
    for (auto& f : films) if (!f.empty()) for (size_t i = 0; i < npx; i++) film[i] = film[i] + f[i];
    const double scale = normSamples > 0 ? (double)((int64_t)Params.Width * Params.Height) / (double)normSamples : 1.0;
    for (size_t i = 0; i < npx; i++) { filmOut[3 * i] = film[i].x * scale; filmOut[3 * i + 1] = film[i].y * scale; filmOut[3 * i + 2] = film[i].z * scale; }
    for (auto& c : cnts) { total.extend += c.extend; total.shadow += c.shadow; }
}

thread_local std::string g_err;

}  // namespace

// =============================================================================================
// C API (ctypes-friendly)
// =============================================================================================
extern "C" {

__attribute__((visibility("default"))) const char* oracle_last_error() { return g_err.c_str(); }

__attribute__((visibility("default"))) void* oracle_scene_create(const NgiSceneDesc* desc) {
    if (!desc || desc->struct_size != sizeof(NgiSceneDesc)) { g_err = "bad NgiSceneDesc"; return nullptr; }
    try { return make_scene(desc); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
__attribute__((visibility("default"))) void oracle_scene_destroy(void* s) { delete (Scene*)s; }

// rng_mode 0: mt19937 per thread, master seed = (unsigned)seed (src/nanogi.cpp:186-191, :297)
// rng_mode 1: Philox counter-based streams identical to the GPU module (sample-exact replay)
// film: double[W*H*3], row 0 = bottom. stats: {paths, extend rays, shadow rays, seconds}
__attribute__((visibility("default"))) int oracle_render(void* s, int renderer, int64_t num_samples, int64_t sample_offset,
                                                          int64_t norm_samples, int max_num_vertices, int width, int height,
                                                          uint64_t seed, int rng_mode, int num_threads, double* film, double* stats) {
    Scene* sc = (Scene*)s;
    if (!sc || !film || width <= 0 || height <= 0 || num_samples < 0) { g_err = "invalid argument"; return -1; }
    if (renderer < NGI_RENDERER_PT || renderer > NGI_RENDERER_BDPT) { g_err = "renderer not supported (pt, ptdirect, lt, ltdirect, bdpt)"; return -4; }
    if (sc->SensorPrimitiveIndex == (size_t)-1) { g_err = "scene has no sensor"; return -1; }
    if (num_threads <= 0) num_threads = std::max(1, (int)std::thread::hardware_concurrency() + num_threads);  // src/nanogi.cpp:149-152
    RenderParams P{width, height, max_num_vertices};
    Counters total;
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (rng_mode == 0) {
        MtSampler initRng; initRng.SetSeed((unsigned int)seed);
        RenderProcess<MtSampler>(*sc, renderer, P, num_samples, sample_offset, norm_samples, num_threads,
                                 [&](MtSampler& r, int) { r.SetSeed(initRng.NextUInt()); }, film, total);
    } else {
        RenderProcess<PhiloxSampler>(*sc, renderer, P, num_samples, sample_offset, norm_samples, num_threads,
                                     [&](PhiloxSampler& r, int) { r.key[0] = (uint32_t)seed; r.key[1] = (uint32_t)(seed >> 32); }, film, total);
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (stats) {
        stats[0] = (double)num_samples; stats[1] = (double)total.extend; stats[2] = (double)total.shadow;
        stats[3] = std::chrono::duration<double>(t1 - t0).count();
    }
    return 0;
}

// raw ray queries. mode 0 closest (BVH), 1 any-hit (BVH), 2 closest brute force
__attribute__((visibility("default"))) int oracle_trace(void* s, const NgiRay* rays, uint64_t n, NgiHit* hits, int mode, int num_threads) {
    Scene* sc = (Scene*)s;
    if (!sc || (!rays && n) || (!hits && n)) { g_err = "invalid argument"; return -1; }
    if (num_threads <= 0) num_threads = std::max(1, (int)std::thread::hardware_concurrency());
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        while (true) {
            uint64_t b = next.fetch_add(4096);
            if (b >= n) break;
            uint64_t e = std::min<uint64_t>(b + 4096, n);
            for (uint64_t i = b; i < e; i++) {
                const NgiRay& r = rays[i];
                NgiHit h; h.t = 0; h.u = 0; h.v = 0; h.tri = NGI_NO_HIT;
                if (mode == 1) { if (sc->TraceAny(r.o, r.d, r.tmin, r.tmax)) h.tri = 0; }
                else if (mode == 2) sc->TraceBrute(r.o, r.d, r.tmin, r.tmax, h.t, h.u, h.v, h.tri);
                else sc->TraceClosest(r.o, r.d, r.tmin, r.tmax, h.t, h.u, h.v, h.tri);
                hits[i] = h;
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < num_threads; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}

// Scene::Intersect surface reconstruction for one double ray: out = {hit, tri, p[3], gn[3], sn[3], dpdu[3], dpdv[3], uv[2]}
__attribute__((visibility("default"))) int oracle_intersect(void* s, const double* o, const double* d, double* out) {
    Scene* sc = (Scene*)s;
    Ray r{d3(o), d3(d)}; Intersection is;
    const bool hit = Intersect(*sc, r, is);
    out[0] = hit ? 1 : 0;
    if (!hit) return 0;
    out[1] = is.tri;
    const d3* v[5] = {&is.geom.p, &is.geom.gn, &is.geom.sn, &is.geom.dpdu, &is.geom.dpdv};
    for (int i = 0; i < 5; i++) { out[2 + 3 * i] = v[i]->x; out[3 + 3 * i] = v[i]->y; out[4 + 3 * i] = v[i]->z; }
    out[17] = is.geom.uv.x; out[18] = is.geom.uv.y;
    return 0;
}

__attribute__((visibility("default"))) int oracle_visible(void* s, const double* p1, const double* p2) {
    return Visible(*(Scene*)s, d3(p1), d3(p2)) ? 1 : 0;
}

// Per-function hooks for the table tests of SURVEY §8a rows 8-14.
// geom9 = {sn[3], gn[3], p[3]}; tangent frame is derived with ComputeTangentSpace like every caller does.
static SurfaceGeometry geom_from(const double* g, bool degenerated) {
    SurfaceGeometry geom; geom.sn = d3(g); geom.gn = d3(g + 3); geom.p = d3(g + 6); geom.degenerated = degenerated;
    if (!degenerated) geom.ComputeTangentSpace();
    return geom;
}
// out = {wo[3], wrote}
__attribute__((visibility("default"))) void oracle_sample_direction(void* s, int prim, int query_type, const double* geom9, const double* wi,
                                                                     double u0, double u1, double ucomp, double* out) {
    Scene* sc = (Scene*)s; const Primitive& p = sc->Primitives.at(prim);
    SurfaceGeometry geom = geom_from(geom9, (query_type & NGI_TYPE_E) && p.EType == NGI_E_PINHOLE);
    d3 wo; d2 u; u.x = u0; u.y = u1;
    bool w = p.SampleDirection(u, ucomp, query_type, geom, d3(wi), wo);
    out[0] = wo.x; out[1] = wo.y; out[2] = wo.z; out[3] = w ? 1 : 0;
}
// out = {fs[3], pdf}
__attribute__((visibility("default"))) void oracle_evaluate_direction(void* s, int prim, int query_type, const double* geom9, const double* wi,
                                                                       const double* wo, int trans_dir_el, int force_degenerated, double* out) {
    Scene* sc = (Scene*)s; const Primitive& p = sc->Primitives.at(prim);
    SurfaceGeometry geom = geom_from(geom9, (query_type & NGI_TYPE_E) && p.EType == NGI_E_PINHOLE);
    d3 fs = p.EvaluateDirection(geom, query_type, d3(wi), d3(wo), trans_dir_el ? EL : LE, force_degenerated != 0);
    out[0] = fs.x; out[1] = fs.y; out[2] = fs.z;
    out[3] = p.EvaluateDirectionPDF(geom, query_type, d3(wi), d3(wo), force_degenerated != 0);
}
// out = {p[3], gn[3], sn[3], pdf}
__attribute__((visibility("default"))) void oracle_sample_position(void* s, int prim, double u0, double u1, double* out) {
    Scene* sc = (Scene*)s; const Primitive& p = sc->Primitives.at(prim);
    SurfaceGeometry g; d2 u; u.x = u0; u.y = u1;
    p.SamplePosition(u, g);
    out[0] = g.p.x; out[1] = g.p.y; out[2] = g.p.z; out[3] = g.gn.x; out[4] = g.gn.y; out[5] = g.gn.z;
    out[6] = g.sn.x; out[7] = g.sn.y; out[8] = g.sn.z; out[9] = p.EvaluatePositionPDF(g, true);
}
// out = {ok, x, y, pixelIndex}
__attribute__((visibility("default"))) void oracle_raster_position(void* s, int prim, const double* wo, int w, int h, double* out) {
    Scene* sc = (Scene*)s; const Primitive& p = sc->Primitives.at(prim);
    SurfaceGeometry g; d2 r;
    bool ok = p.RasterPosition(d3(wo), g, r);
    out[0] = ok ? 1 : 0; out[1] = r.x; out[2] = r.y; out[3] = ok ? PixelIndex(r, w, h) : -1;
}
__attribute__((visibility("default"))) double oracle_fresnel(void* s, int prim, double cos_i, double etaI, double etaT) {
    Scene* sc = (Scene*)s;
    return sc->Primitives.at(prim).EvaluateFresnelTerm(d3(std::sqrt(std::max(0.0, 1 - cos_i * cos_i)), 0, cos_i), etaI, etaT);
}
__attribute__((visibility("default"))) double oracle_geometry_term(const double* p1, const double* sn1, int deg1, const double* p2, const double* sn2, int deg2) {
    SurfaceGeometry a, b; a.p = d3(p1); a.sn = d3(sn1); a.degenerated = deg1 != 0; b.p = d3(p2); b.sn = d3(sn2); b.degenerated = deg2 != 0;
    return GeometryTerm(a, b);
}
__attribute__((visibility("default"))) void oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
__attribute__((visibility("default"))) void oracle_orthonormal_basis(const double* a, double* b, double* c) {
    d3 bb, cc; OrthonormalBasis(d3(a), bb, cc);
    b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; c[0] = cc.x; c[1] = cc.y; c[2] = cc.z;
}
// scene info: out = {nTris, nPrims, nLights, sensorIndex, bvhNodes, pad}
__attribute__((visibility("default"))) void oracle_scene_info(void* s, double* out) {
    Scene* sc = (Scene*)s;
    out[0] = (double)sc->tris.size(); out[1] = (double)sc->Primitives.size(); out[2] = (double)sc->LightPrimitiveIndices.size();
    out[3] = (double)(int64_t)sc->SensorPrimitiveIndex; out[4] = (double)sc->nodes.size(); out[5] = sc->pad;
}
// light-primitive CDF and InvArea (rt.hpp:1747-1765): returns count, fills up to cap entries
__attribute__((visibility("default"))) int oracle_light_cdf(void* s, int prim, double* cdf, int cap, double* inv_area) {
    Scene* sc = (Scene*)s; const Primitive& p = sc->Primitives.at(prim);
    int n = (int)p.Dist.cdf.size();
    for (int i = 0; i < n && i < cap; i++) cdf[i] = p.Dist.cdf[i];
    if (inv_area) *inv_area = p.L_InvArea;
    return n;
}

}  // extern "C"

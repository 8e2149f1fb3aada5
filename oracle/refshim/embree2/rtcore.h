// Stand-in for the Embree 2.x API subset the reference uses (rt.hpp:1506-1516, :2085-2143, :2162-2182). Written for this
// repository; see ../README.md.
//
// THIS IS NOT EMBREE'S KERNEL. rtcIntersect here is the float32 Moeller-Trumbore closest hit that DESIGN.md declares as
// "the reference intersector" of this repository — one explicit expression tree (fused multiply-adds written as fmaf, build
// with -ffp-contract=off), strict tnear < t < tfar, no culling, closest hit = lexicographic minimum of (t, geomID, primID) —
// over a binned-SAH BVH whose boxes are padded so that it is a pure filter (identical answers to a loop over all
// triangles). Everything ABOVE the ray query (Scene::Intersect's surface reconstruction, Visible, the renderers) is the
// reference's own code.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)

enum RTCError { RTC_NO_ERROR = 0, RTC_UNKNOWN_ERROR = 1, RTC_INVALID_ARGUMENT = 2, RTC_INVALID_OPERATION = 3, RTC_OUT_OF_MEMORY = 4, RTC_UNSUPPORTED_CPU = 5 };
enum RTCSceneFlags { RTC_SCENE_STATIC = 0, RTC_SCENE_DYNAMIC = 1, RTC_SCENE_COMPACT = 1 << 8, RTC_SCENE_COHERENT = 1 << 9, RTC_SCENE_INCOHERENT = 1 << 10,
                     RTC_SCENE_HIGH_QUALITY = 1 << 11, RTC_SCENE_ROBUST = 1 << 16 };
inline RTCSceneFlags operator|(RTCSceneFlags a, RTCSceneFlags b) { return (RTCSceneFlags)((int)a | (int)b); }
enum RTCAlgorithmFlags { RTC_INTERSECT1 = 1, RTC_INTERSECT4 = 2, RTC_INTERSECT8 = 4, RTC_INTERSECT16 = 8 };
enum RTCGeometryFlags { RTC_GEOMETRY_STATIC = 0, RTC_GEOMETRY_DEFORMABLE = 1, RTC_GEOMETRY_DYNAMIC = 2 };
enum RTCBufferType { RTC_INDEX_BUFFER = 0x01000000, RTC_VERTEX_BUFFER = 0x02000000 };
typedef void (*RTC_ERROR_FUNCTION)(const RTCError code, const char* str);

struct RTCRay;

namespace ngi_embree_shim {
struct Tri { float v0[3], e1[3], e2[3]; unsigned geomID, primID; };
struct Node { float lo[3], hi[3]; int left, count; };   // count > 0: leaf over order[left .. left + count)
struct Geom { size_t numTris = 0, numVerts = 0; std::vector<float> verts; std::vector<int> idx; };
struct Scene {
    std::vector<Geom> geoms;
    std::vector<Tri> tris;          // (geomID, primID) order
    std::vector<unsigned> order;    // BVH leaf order -> index into tris
    std::vector<Node> nodes;
};
inline float crossc(float ay, float az, float by, float bz) { return fmaf(ay, bz, -(az * by)); }
inline float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(ax, bx, fmaf(ay, by, az * bz)); }
inline bool tri_test(const Tri& tr, const float* o, const float* d, float tmin, float tmax, float& t, float& u, float& v) {
    const float px = crossc(d[1], d[2], tr.e2[1], tr.e2[2]), py = crossc(d[2], d[0], tr.e2[2], tr.e2[0]), pz = crossc(d[0], d[1], tr.e2[0], tr.e2[1]);
    const float det = dot3(tr.e1[0], tr.e1[1], tr.e1[2], px, py, pz);
    if (!(det != 0.0f)) return false;
    const float inv = 1.0f / det;
    const float tx = o[0] - tr.v0[0], ty = o[1] - tr.v0[1], tz = o[2] - tr.v0[2];
    const float uu = dot3(tx, ty, tz, px, py, pz) * inv;
    if (!(uu >= 0.0f && uu <= 1.0f)) return false;
    const float qx = crossc(ty, tz, tr.e1[1], tr.e1[2]), qy = crossc(tz, tx, tr.e1[2], tr.e1[0]), qz = crossc(tx, ty, tr.e1[0], tr.e1[1]);
    const float vv = dot3(d[0], d[1], d[2], qx, qy, qz) * inv;
    if (!(vv >= 0.0f && uu + vv <= 1.0f)) return false;
    const float tt = dot3(tr.e2[0], tr.e2[1], tr.e2[2], qx, qy, qz) * inv;
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; u = uu; v = vv;
    return true;
}
inline void build(Scene& s) {
    s.tris.clear();
    float mag = 1e-30f;
    for (unsigned g = 0; g < s.geoms.size(); g++) {
        const Geom& G = s.geoms[g];
        for (size_t p = 0; p < G.numTris; p++) {
            Tri t; t.geomID = g; t.primID = (unsigned)p;
            const float* a = &G.verts[4 * (size_t)G.idx[3 * p]]; const float* b = &G.verts[4 * (size_t)G.idx[3 * p + 1]]; const float* c = &G.verts[4 * (size_t)G.idx[3 * p + 2]];
            for (int k = 0; k < 3; k++) { t.v0[k] = a[k]; t.e1[k] = b[k] - a[k]; t.e2[k] = c[k] - a[k]; mag = std::max(mag, std::max(std::fabs(a[k]), std::max(std::fabs(b[k]), std::fabs(c[k])))); }
            s.tris.push_back(t);
        }
    }
    const size_t n = s.tris.size();
    const float pad = mag * (1.0f / 256.0f);   // generous: rays may start far outside the scene (ulp of |o - v0| matters, not of the scene)
    std::vector<float> lo(n * 3), hi(n * 3), cen(n * 3);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) {
        const Tri& t = s.tris[i];
        const float a = t.v0[k], b = t.v0[k] + t.e1[k], c = t.v0[k] + t.e2[k];
        lo[i * 3 + k] = std::min(a, std::min(b, c)) - pad - std::fabs(a) * 1e-6f; hi[i * 3 + k] = std::max(a, std::max(b, c)) + pad + std::fabs(a) * 1e-6f;
        cen[i * 3 + k] = 0.5f * (lo[i * 3 + k] + hi[i * 3 + k]);
    }
    s.order.resize(n);
    for (size_t i = 0; i < n; i++) s.order[i] = (unsigned)i;
    s.nodes.clear();
    if (n == 0) return;
    struct Task { int node; size_t b, e; };
    s.nodes.push_back(Node());
    std::vector<Task> st{{0, 0, n}};
    while (!st.empty()) {
        const Task t = st.back(); st.pop_back();
        Node nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = 3e38f; nd.hi[k] = -3e38f; }
        float cl[3] = {3e38f, 3e38f, 3e38f}, ch[3] = {-3e38f, -3e38f, -3e38f};
        for (size_t i = t.b; i < t.e; i++) for (int k = 0; k < 3; k++) {
            const unsigned id = s.order[i];
            nd.lo[k] = std::min(nd.lo[k], lo[id * 3 + k]); nd.hi[k] = std::max(nd.hi[k], hi[id * 3 + k]);
            cl[k] = std::min(cl[k], cen[id * 3 + k]); ch[k] = std::max(ch[k], cen[id * 3 + k]);
        }
        if (t.e - t.b <= 4) { nd.left = (int)t.b; nd.count = (int)(t.e - t.b); s.nodes[t.node] = nd; continue; }
        // binned surface-area heuristic (16 bins per axis); falls back to the median of the widest axis
        const int NB = 16;
        int bestAxis = -1, bestBin = -1; float bestCost = 3e38f;
        auto half_area = [](const float* mn, const float* mx) { const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2]; return dx * dy + dy * dz + dz * dx; };
        for (int ax = 0; ax < 3; ax++) {
            const float ext = ch[ax] - cl[ax];
            if (!(ext > 0)) continue;
            float bmn[NB][3], bmx[NB][3]; int bc[NB];
            for (int b = 0; b < NB; b++) { bc[b] = 0; for (int k = 0; k < 3; k++) { bmn[b][k] = 3e38f; bmx[b][k] = -3e38f; } }
            const float sc = NB / ext;
            for (size_t i = t.b; i < t.e; i++) {
                const unsigned id = s.order[i];
                const int b = std::min(NB - 1, (int)((cen[id * 3 + ax] - cl[ax]) * sc));
                bc[b]++;
                for (int k = 0; k < 3; k++) { bmn[b][k] = std::min(bmn[b][k], lo[id * 3 + k]); bmx[b][k] = std::max(bmx[b][k], hi[id * 3 + k]); }
            }
            float rA[NB]; int rC[NB]; float amn[3] = {3e38f, 3e38f, 3e38f}, amx[3] = {-3e38f, -3e38f, -3e38f}; int c = 0;
            for (int b = NB - 1; b > 0; b--) { for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[b][k]); amx[k] = std::max(amx[k], bmx[b][k]); } c += bc[b]; rA[b] = c ? half_area(amn, amx) : 0; rC[b] = c; }
            for (int k = 0; k < 3; k++) { amn[k] = 3e38f; amx[k] = -3e38f; }
            c = 0;
            for (int b = 0; b < NB - 1; b++) {
                for (int k = 0; k < 3; k++) { amn[k] = std::min(amn[k], bmn[b][k]); amx[k] = std::max(amx[k], bmx[b][k]); }
                c += bc[b];
                if (c == 0 || rC[b + 1] == 0) continue;
                const float cost = half_area(amn, amx) * c + rA[b + 1] * rC[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestBin = b; }
            }
        }
        size_t mid = (t.b + t.e) / 2;
        if (bestAxis >= 0) {
            const float sc = NB / (ch[bestAxis] - cl[bestAxis]);
            auto it = std::partition(s.order.begin() + t.b, s.order.begin() + t.e,
                                     [&](unsigned id) { return std::min(NB - 1, (int)((cen[id * 3 + bestAxis] - cl[bestAxis]) * sc)) <= bestBin; });
            const size_t m2 = (size_t)(it - s.order.begin());
            if (m2 != t.b && m2 != t.e) mid = m2;
        }
        nd.left = (int)s.nodes.size(); nd.count = 0;
        s.nodes[t.node] = nd;
        s.nodes.push_back(Node()); s.nodes.push_back(Node());
        st.push_back({nd.left, t.b, mid});
        st.push_back({nd.left + 1, mid, t.e});
    }
}
inline bool box_hit(const Node& nd, const double* o, const double* id, double t0, double t1, double& tnear) {
    for (int k = 0; k < 3; k++) {
        double a = (nd.lo[k] - o[k]) * id[k], b = (nd.hi[k] - o[k]) * id[k];
        if (a > b) std::swap(a, b);
        if (a > t0) t0 = a;       // NaN (0 * inf) compares false: keeps the interval, never culls
        if (b < t1) t1 = b;
    }
    tnear = t0;
    return t0 <= t1;
}
inline RTC_ERROR_FUNCTION& error_function() { static RTC_ERROR_FUNCTION f = nullptr; return f; }
}  // namespace ngi_embree_shim

typedef ngi_embree_shim::Scene* RTCScene;

inline void rtcInit(const char* = nullptr) {}
inline void rtcExit() {}
inline void rtcSetErrorFunction(RTC_ERROR_FUNCTION f) { ngi_embree_shim::error_function() = f; }
inline RTCScene rtcNewScene(RTCSceneFlags, RTCAlgorithmFlags) { return new ngi_embree_shim::Scene; }
inline void rtcDeleteScene(RTCScene s) { delete s; }
inline unsigned rtcNewTriangleMesh(RTCScene s, RTCGeometryFlags, size_t numTriangles, size_t numVertices, size_t = 1) {
    ngi_embree_shim::Geom g; g.numTris = numTriangles; g.numVerts = numVertices;
    g.verts.assign(numVertices * 4, 0.0f); g.idx.assign(numTriangles * 3, 0);   // Embree 2: 16-byte vertex stride, int32 indices
    s->geoms.push_back(std::move(g));
    return (unsigned)(s->geoms.size() - 1);
}
inline void* rtcMapBuffer(RTCScene s, unsigned geomID, RTCBufferType type) {
    return type == RTC_VERTEX_BUFFER ? (void*)s->geoms[geomID].verts.data() : (void*)s->geoms[geomID].idx.data();
}
inline void rtcUnmapBuffer(RTCScene, unsigned, RTCBufferType) {}
inline void rtcCommit(RTCScene s) { ngi_embree_shim::build(*s); }

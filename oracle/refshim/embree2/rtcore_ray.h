// Stand-in for embree2/rtcore_ray.h: the single-ray structure and rtcIntersect. See rtcore.h and ../README.md.
#pragma once
#include "rtcore.h"

struct RTCRay {
    float org[3]; float align0;
    float dir[3]; float align1;
    float tnear, tfar, time; int mask;
    float Ng[3]; float align2;
    float u, v;
    int geomID, primID, instID;
};

inline void rtcIntersect(RTCScene s, RTCRay& ray) {
    using namespace ngi_embree_shim;
    if (s->nodes.empty()) return;
    const double o[3] = {ray.org[0], ray.org[1], ray.org[2]}, id[3] = {1.0 / ray.dir[0], 1.0 / ray.dir[1], 1.0 / ray.dir[2]};
    bool found = false; float bt = ray.tfar, bu = 0, bv = 0; unsigned best = 0xFFFFFFFFu;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    const double t0 = (double)ray.tnear * 0.999;
    while (sp) {
        const Node& nd = s->nodes[stack[--sp]];
        double tn;
        // inclusive culling against the current best t, with slack for the float -> double mismatch
        if (!box_hit(nd, o, id, t0, (double)bt * 1.000001 + 1e-30, tn)) continue;
        if (nd.count > 0) {
            for (int i = 0; i < nd.count; i++) {
                const unsigned ti = s->order[nd.left + i];
                float t, u, v;
                if (tri_test(s->tris[ti], ray.org, ray.dir, ray.tnear, ray.tfar, t, u, v))
                    if (!found || t < bt || (t == bt && ti < best)) { found = true; bt = t; bu = u; bv = v; best = ti; }
            }
        } else {
            // near child first (the visiting order cannot change the answer: the reduction above is order-free)
            double ta, tb;
            const double t1 = (double)bt * 1.000001 + 1e-30;
            const bool ha = box_hit(s->nodes[nd.left], o, id, t0, t1, ta), hb = box_hit(s->nodes[nd.left + 1], o, id, t0, t1, tb);
            if (ha && hb) { if (ta < tb) { stack[sp++] = nd.left + 1; stack[sp++] = nd.left; } else { stack[sp++] = nd.left; stack[sp++] = nd.left + 1; } }
            else if (ha) stack[sp++] = nd.left;
            else if (hb) stack[sp++] = nd.left + 1;
        }
    }
    if (!found) return;
    const Tri& tr = s->tris[best];
    ray.tfar = bt; ray.u = bu; ray.v = bv; ray.geomID = (int)tr.geomID; ray.primID = (int)tr.primID;
    ray.Ng[0] = tr.e1[1] * tr.e2[2] - tr.e1[2] * tr.e2[1]; ray.Ng[1] = tr.e1[2] * tr.e2[0] - tr.e1[0] * tr.e2[2]; ray.Ng[2] = tr.e1[0] * tr.e2[1] - tr.e1[1] * tr.e2[0];
}

// ref_harness.cpp — C interface around the reference's OWN translation unit (TEST INFRASTRUCTURE).
//
// `#include NGI_REF_SRC` pulls in /root/reference/src/nanogi.cpp (and through it include/nanogi/{basic,rt,bdpt}.hpp) exactly
// as oracle/build_ref.sh hands it over — compiled against the stand-in third-party headers of oracle/refshim/. Nothing of the
// reference is restated here: the functions below only CALL the reference's Scene::Load, Primitive::*, Scene::Intersect /
// Visible, GeometryTerm and Renderer::Render (-> RenderProcess -> ProcessSample_PT / _PTDirect / _LT / _LTDirect) and copy
// their results out, with the same signatures as the oracle's hooks (oracle/oracle.cpp) so that tests can put the two side
// by side. `main` is renamed and `private` opened so that the renderer's fields can be set without going through argv.
//
// Determinism: the release build seeds from std::time(nullptr) (src/nanogi.cpp:190). This harness interposes time() (the
// library is linked -Bsymbolic-functions) so that ref_render's `seed` becomes that value; with one thread the whole run is
// then a deterministic function of the seed, which is what lets the oracle's mt19937 mode be compared with it FILM-EXACTLY.
#include <ctime>
#include <cstring>

static long long g_fake_time = -1;
extern "C" time_t time(time_t* t) noexcept {
    time_t v;
    if (g_fake_time >= 0) v = (time_t)g_fake_time;
    else { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); v = ts.tv_sec; }
    if (t) *t = v;
    return v;
}

// every header src/nanogi.cpp includes, first and untouched (include guards keep them from being read again below), so that
// `#define private public` only reaches the text of src/nanogi.cpp itself — i.e. struct Renderer's ProcessSample_* section
#include <nanogi/macros.hpp>
#include <nanogi/basic.hpp>
#include <nanogi/rt.hpp>
#include <nanogi/bdpt.hpp>
#include <boost/program_options.hpp>
#include <ctemplate/template.h>
#include <eigen3/Eigen/Dense>

#define main nanogi_reference_main
#define private public
#include NGI_REF_SRC
#undef private
#undef main

namespace {
thread_local std::string g_err;
struct RefScene { nanogi::Scene scene; };
bool g_logger_running = false;
void ensure_logger() {
    if (g_logger_running) return;
    // the reference logs every loader step to std::cout; drop it unless NGI_REF_LOG is set (nothing else in a Python host uses cout)
    if (!getenv("NGI_REF_LOG")) std::cout.rdbuf(nullptr);
    NGI_LOG_RUN();
    g_logger_running = true;
}

nanogi::SurfaceGeometry geom_from(const double* g, bool degenerated) {
    nanogi::SurfaceGeometry geom;
    geom.sn = glm::dvec3(g[0], g[1], g[2]); geom.gn = glm::dvec3(g[3], g[4], g[5]); geom.p = glm::dvec3(g[6], g[7], g[8]);
    geom.degenerated = degenerated;
    if (!degenerated) geom.ComputeTangentSpace();
    return geom;
}
bool pinhole_query(const nanogi::Primitive& p, int query_type) {
    return (query_type & nanogi::PrimitiveType::E) > 0 && p.Params.E.Type == nanogi::EType::Pinhole;
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char* ref_last_error() { return g_err.c_str(); }

// Scene::Load (include/nanogi/rt.hpp:1519-2154) on a schema.yml file
REF_API void* ref_scene_load(const char* path, double aspect) {
    ensure_logger();
    RefScene* s = new RefScene;
    try {
        if (!s->scene.Load(path, aspect)) { g_err = "Scene::Load failed (see log)"; delete s; return nullptr; }
    } catch (const std::exception& e) { g_err = e.what(); delete s; return nullptr; }
    return s;
}
REF_API void ref_scene_destroy(void* h) { delete (RefScene*)h; }

// out = {primitives, lights, sensor index, triangles}
REF_API void ref_scene_info(void* h, double* out) {
    const nanogi::Scene& sc = ((RefScene*)h)->scene;
    size_t tris = 0;
    for (const auto& p : sc.Primitives) if (p->MeshRef) tris += p->MeshRef->Faces.size() / 3;
    out[0] = (double)sc.Primitives.size(); out[1] = (double)sc.LightPrimitiveIndices.size(); out[2] = (double)sc.SensorPrimitiveIndex; out[3] = (double)tris;
}

// Renderer::Render (src/nanogi.cpp:182-221) with the fields Renderer::Load would have set (:117-180).
// film: double[W*H*3], row 0 = bottom (the reference's own vector<dvec3> layout)
REF_API int ref_render(void* h, int renderer, long long num_samples, int max_num_vertices, int width, int height, int num_threads,
                       long long seed, double* film_out) {
    ensure_logger();
    const nanogi::Scene& sc = ((RefScene*)h)->scene;
    if (renderer < 0 || renderer > 4) { g_err = "renderer must be pt / ptdirect / lt / ltdirect / bdpt"; return -1; }
    Renderer r;
    r.Type = (RendererType)renderer;
    r.NumThreads = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
    r.init.initialize(r.NumThreads);
    r.GrainSize = 10000;                       // src/nanogi.cpp:2014
    r.ProgressUpdateInterval = 1LL << 60;      // keep the progress log quiet
    r.ProgressImageUpdateInterval = -1;
    r.Params.NumSamples = num_samples; r.Params.RenderTime = -1; r.Params.MaxNumVertices = max_num_vertices;
    r.Params.Width = width; r.Params.Height = height;
    std::vector<glm::dvec3> film;
    g_fake_time = seed;
    try { r.Render(sc, film); } catch (const std::exception& e) { g_fake_time = -1; g_err = e.what(); return -1; }
    g_fake_time = -1;
    for (size_t i = 0; i < film.size(); i++) { film_out[3 * i] = film[i].x; film_out[3 * i + 1] = film[i].y; film_out[3 * i + 2] = film[i].z; }
    return 0;
}

// Scene::Intersect (rt.hpp:2162-2249): out = {hit, primitive index, p[3], gn[3], sn[3], dpdu[3], dpdv[3], uv[2]}
REF_API int ref_intersect(void* h, const double* o, const double* d, double* out) {
    const nanogi::Scene& sc = ((RefScene*)h)->scene;
    nanogi::Ray ray; ray.o = glm::dvec3(o[0], o[1], o[2]); ray.d = glm::dvec3(d[0], d[1], d[2]);
    nanogi::Intersection is;
    const bool hit = sc.Intersect(ray, is);
    out[0] = hit ? 1 : 0;
    if (!hit) return 0;
    size_t pi = 0;
    for (; pi < sc.Primitives.size(); pi++) if (sc.Primitives[pi].get() == is.Prim) break;
    out[1] = (double)pi;
    const glm::dvec3* v[5] = {&is.geom.p, &is.geom.gn, &is.geom.sn, &is.geom.dpdu, &is.geom.dpdv};
    for (int i = 0; i < 5; i++) { out[2 + 3 * i] = v[i]->x; out[3 + 3 * i] = v[i]->y; out[4 + 3 * i] = v[i]->z; }
    out[17] = is.geom.uv.x; out[18] = is.geom.uv.y;
    return 0;
}
// Scene::Visible (rt.hpp:2251-2261)
REF_API int ref_visible(void* h, const double* p1, const double* p2) {
    return ((RefScene*)h)->scene.Visible(glm::dvec3(p1[0], p1[1], p1[2]), glm::dvec3(p2[0], p2[1], p2[2])) ? 1 : 0;
}

// Primitive::SampleDirection (rt.hpp:692-910): out = {wo[3]}; wo starts as a default-constructed dvec3 like in the callers
REF_API void ref_sample_direction(void* h, int prim, int query_type, const double* geom9, const double* wi, double u0, double u1, double ucomp, double* out) {
    const nanogi::Primitive& p = *((RefScene*)h)->scene.Primitives.at(prim);
    const nanogi::SurfaceGeometry geom = geom_from(geom9, pinhole_query(p, query_type));
    glm::dvec3 wo;
    p.SampleDirection(glm::dvec2(u0, u1), ucomp, query_type, geom, glm::dvec3(wi[0], wi[1], wi[2]), wo);
    out[0] = wo.x; out[1] = wo.y; out[2] = wo.z;
}
// Primitive::EvaluateDirection + EvaluateDirectionPDF (rt.hpp:912-1336): out = {fs[3], pdf}
REF_API void ref_evaluate_direction(void* h, int prim, int query_type, const double* geom9, const double* wi, const double* wo, int trans_dir_el,
                                    int force_degenerated, double* out) {
    const nanogi::Primitive& p = *((RefScene*)h)->scene.Primitives.at(prim);
    const nanogi::SurfaceGeometry geom = geom_from(geom9, pinhole_query(p, query_type));
    const glm::dvec3 wi3(wi[0], wi[1], wi[2]), wo3(wo[0], wo[1], wo[2]);
    const glm::dvec3 fs = p.EvaluateDirection(geom, query_type, wi3, wo3, trans_dir_el ? nanogi::TransportDirection::EL : nanogi::TransportDirection::LE, force_degenerated != 0);
    out[0] = fs.x; out[1] = fs.y; out[2] = fs.z;
    out[3] = p.EvaluateDirectionPDF(geom, query_type, wi3, wo3, force_degenerated != 0);
}
// Primitive::SamplePosition + EvaluatePositionPDF (rt.hpp:483-690): out = {p[3], gn[3], sn[3], pdf, uv[2]}
REF_API void ref_sample_position(void* h, int prim, double u0, double u1, double* out) {
    const nanogi::Primitive& p = *((RefScene*)h)->scene.Primitives.at(prim);
    nanogi::SurfaceGeometry g;
    p.SamplePosition(glm::dvec2(u0, u1), g);
    out[0] = g.p.x; out[1] = g.p.y; out[2] = g.p.z; out[3] = g.gn.x; out[4] = g.gn.y; out[5] = g.gn.z;
    out[6] = g.sn.x; out[7] = g.sn.y; out[8] = g.sn.z; out[9] = p.EvaluatePositionPDF(g, true); out[10] = g.uv.x; out[11] = g.uv.y;
}
// Primitive::RasterPosition (rt.hpp:1344-1399) + PixelIndex (rt.hpp:135-140): out = {ok, x, y, pixel index}
REF_API void ref_raster_position(void* h, int prim, const double* wo, int w, int hgt, double* out) {
    const nanogi::Primitive& p = *((RefScene*)h)->scene.Primitives.at(prim);
    nanogi::SurfaceGeometry g; g.degenerated = true; g.p = p.Params.E.Pinhole.Position;
    glm::dvec2 r;
    const bool ok = p.RasterPosition(glm::dvec3(wo[0], wo[1], wo[2]), g, r);
    out[0] = ok ? 1 : 0; out[1] = r.x; out[2] = r.y; out[3] = ok ? (double)nanogi::PixelIndex(r, w, hgt) : -1.0;
}
// GeometryTerm (rt.hpp:2364-2374)
REF_API double ref_geometry_term(const double* p1, const double* sn1, int deg1, const double* p2, const double* sn2, int deg2) {
    nanogi::SurfaceGeometry a, b;
    a.p = glm::dvec3(p1[0], p1[1], p1[2]); a.sn = glm::dvec3(sn1[0], sn1[1], sn1[2]); a.degenerated = deg1 != 0;
    b.p = glm::dvec3(p2[0], p2[1], p2[2]); b.sn = glm::dvec3(sn2[0], sn2[1], sn2[2]); b.degenerated = deg2 != 0;
    return nanogi::GeometryTerm(a, b);
}
// OrthonormalBasis (rt.hpp:55-59)
REF_API void ref_orthonormal_basis(const double* a, double* b, double* c) {
    glm::dvec3 bb, cc;
    nanogi::OrthonormalBasis(glm::dvec3(a[0], a[1], a[2]), bb, cc);
    b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; c[0] = cc.x; c[1] = cc.y; c[2] = cc.z;
}
// Random (basic.hpp:419-434): n doubles of Next() after SetSeed(seed) — pins the oracle's MtSampler
REF_API void ref_random_stream(unsigned int seed, int n, double* out, unsigned int* next_uint) {
    nanogi::Random r; r.SetSeed(seed);
    for (int i = 0; i < n; i++) out[i] = r.Next();
    if (next_uint) *next_uint = r.NextUInt();
}

/*
 * nanogi_gpu.h — C ABI of the B200 (sm_100a) render module for nanogi's `pt` / `ptdirect` path
 * (widened to `lt` / `ltdirect`, E.area sensors and TexR textures: SURVEY.md §8f).
 *
 * This is the drop-in boundary. The reference has no formal plugin API; the
 * operator slot a GPU path can live behind is
 *     void Renderer::Render(const Scene& scene, std::vector<glm::dvec3>& film) const
 *         (reference src/nanogi.cpp:182-221; called once from Run, src/nanogi.cpp:2097-2102)
 * i.e. "loaded scene + renderer params in -> normalised film out". Everything below is plain C:
 * POD structs, raw pointers and sizes, int status codes. No C++/STL/torch types cross it.
 *
 * Conventions
 *  - status: 0 = NGI_OK, negative = error (see NgiStatus); message via ngi_gpu_last_error().
 *  - all host buffers are caller-owned; the module copies what it needs at scene_create.
 *  - film layout: float RGB, W*H*3, row-major, ROW 0 = BOTTOM scanline (reference
 *    include/nanogi/rt.hpp:135-140 PixelIndex; writers flip, include/nanogi/basic.hpp:583-589).
 *  - global triangle id = index into the de-indexed triangle arrays = prim.first_tri + faceIndex
 *    (the (geomID, primID) pair of reference include/nanogi/rt.hpp:2190-2191).
 *  - the handle is thread-compatible, not re-entrant.
 */
#ifndef NANOGI_GPU_H
#define NANOGI_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NGI_API __declspec(dllexport)
#else
#define NGI_API __attribute__((visibility("default")))
#endif

/* ---- status codes ------------------------------------------------------------------------ */
typedef enum NgiStatus {
    NGI_OK = 0,
    NGI_ERR_INVALID_ARGUMENT = -1,
    NGI_ERR_NO_DEVICE = -2,       /* no CUDA device: the GPU path never falls back to the CPU   */
    NGI_ERR_CUDA = -3,
    NGI_ERR_UNSUPPORTED = -4,     /* renderer outside pt / ptdirect / lt / ltdirect / bdpt (ptmnee) */
    NGI_ERR_OUT_OF_MEMORY = -5,
    NGI_ERR_NCCL = -6             /* libnccl.so.2 missing or an NCCL call failed (multi-GPU entry points only) */
} NgiStatus;

/* ---- primitive type bitmask: reference include/nanogi/rt.hpp:338-351 ---------------------- */
enum {
    NGI_TYPE_D = 1 << 0,
    NGI_TYPE_G = 1 << 1,
    NGI_TYPE_S = 1 << 2,
    NGI_TYPE_L = 1 << 3,
    NGI_TYPE_E = 1 << 4,
    NGI_TYPE_BSDF = NGI_TYPE_D | NGI_TYPE_G | NGI_TYPE_S,
    NGI_TYPE_EMITTER = NGI_TYPE_L | NGI_TYPE_E
};
/* reference include/nanogi/rt.hpp:353-371 */
enum { NGI_L_AREA = 0, NGI_L_POINT = 1, NGI_L_DIRECTIONAL = 2 };
enum { NGI_E_AREA = 0, NGI_E_PINHOLE = 1 };
enum { NGI_S_REFLECTION = 0, NGI_S_REFRACTION = 1, NGI_S_FRESNEL = 2 };
/* reference src/nanogi.cpp:51-69. pt / ptdirect are the hot path; lt / ltdirect (src/nanogi.cpp:804-1131) run on the
 * same wavefront machinery; bdpt (src/nanogi.cpp:1133-1186, bdpt.hpp) in batches of samples through
 * its own dense stages on the same trace kernels. ptmnee is not accepted. */
enum { NGI_RENDERER_PT = 0, NGI_RENDERER_PTDIRECT = 1, NGI_RENDERER_LT = 2, NGI_RENDERER_LTDIRECT = 3, NGI_RENDERER_BDPT = 4 };

/*
 * One scene primitive = mesh range + material + emitter parameters.
 * Mirrors `struct Primitive::Params`, reference include/nanogi/rt.hpp:389-477. Parameters are the
 * doubles the YAML loader parsed (rt.hpp:1806-2043); the module narrows them to fp32 on upload.
 */
typedef struct NgiPrimitive {
    int32_t type;            /* bitmask of NGI_TYPE_*                                        */
    int32_t first_tri;       /* first global triangle id, -1 when the primitive has no mesh  */
    int32_t num_tris;
    int32_t l_type;          /* NGI_L_*  (valid when type & L)                               */
    int32_t e_type;          /* NGI_E_*  (valid when type & E)                               */
    int32_t s_type;          /* NGI_S_*  (valid when type & S)                               */
    int32_t d_tex;           /* texture index for D.TexR, -1 = use d_r   (rt.hpp:1022)       */
    int32_t g_tex;           /* texture index for G.TexR, -1 = use g_r   (rt.hpp:1045)       */
    double d_r[3];           /* D.R                                                          */
    double g_r[3];           /* G.R                                                          */
    double g_eta[3];         /* G.Eta                                                        */
    double g_k[3];           /* G.K                                                          */
    double g_roughness;      /* G.Roughness                                                  */
    double s_r[3];           /* S.{Reflection,Refraction,Fresnel}.R                          */
    double s_eta1, s_eta2;   /* S.{Refraction,Fresnel}.Eta1/Eta2                             */
    double l_le[3];          /* L.{Area,Point,Directional}.Le                                */
    double l_vec[3];         /* L.Point.Position or L.Directional.Direction                  */
    double e_position[3];    /* E.Pinhole.Position                                           */
    double e_vx[3], e_vy[3], e_vz[3]; /* E.Pinhole.Vx/Vy/Vz (rt.hpp:1898-1900)               */
    double e_fov;            /* vertical fov in RADIANS (rt.hpp:1897)                        */
    double e_aspect;         /* W/H from the CLI (src/nanogi.cpp:2069)                       */
    double e_we[3];          /* E.Area.We (rt.hpp:947-953); the pinhole's We is parsed but unused
                                by evaluation (rt.hpp:1895 vs :955-978)                        */
} NgiPrimitive;

/* Nearest-neighbour RGB texture, reference include/nanogi/rt.hpp:157-270 (row 0 = top after flip) */
typedef struct NgiTexture {
    int32_t width, height;
    const float* rgb;        /* width*height*3 */
} NgiTexture;

/*
 * Flattened scene, the POD image of the reference's `Scene` members that cross the boundary
 * (rt.hpp:1494-1499). Geometry is DE-INDEXED exactly like the reference hands it to Embree
 * (rt.hpp:2113-2136): three float vertices per triangle, identity index buffer.
 * Area CDFs (rt.hpp:1747-1765), InvArea and the directional-light disk (rt.hpp:2067-2073) are
 * derived inside scene_create from these arrays.
 */
typedef struct NgiSceneDesc {
    uint32_t struct_size;        /* = sizeof(NgiSceneDesc), ABI check                        */
    uint32_t num_prims;
    uint64_t num_tris;
    const float* positions;      /* [num_tris][3 verts][3]                                   */
    const float* normals;        /* [num_tris][3 verts][3]  vertex normals                   */
    const float* texcoords;      /* [num_tris][3 verts][2]  or NULL                          */
    const NgiPrimitive* prims;   /* [num_prims], YAML order                                  */
    uint32_t num_textures;
    uint32_t reserved0;
    const NgiTexture* textures;  /* [num_textures] or NULL                                   */
} NgiSceneDesc;

/* Renderer::Params, reference src/nanogi.cpp:85-97, plus the counter-based-RNG / shard fields */
typedef struct NgiRenderParams {
    uint32_t struct_size;        /* = sizeof(NgiRenderParams)                                */
    int32_t renderer;            /* NGI_RENDERER_PT | _PTDIRECT | _LT | _LTDIRECT | _BDPT     */
    int64_t num_samples;         /* samples THIS call traces (the shard)                      */
    int64_t sample_offset;       /* first sample index of the shard                           */
    int64_t film_norm_samples;   /* N in film *= W*H/N (src/nanogi.cpp:436); 0 = no scaling   */
    int32_t max_num_vertices;    /* -1 = unbounded (src/nanogi.cpp:485)                       */
    int32_t width, height;
    int32_t accumulate;          /* render_device only: 1 = add into film, 0 = overwrite      */
    uint64_t seed;               /* Philox key                                                */
    uint32_t wave_capacity;      /* path slots in flight per lane (bdpt: samples per batch); 0 = module default */
    uint32_t flags;              /* NGI_RENDER_* bits                                          */
} NgiRenderParams;

/* time every trace-kernel launch with CUDA events (fills NgiRenderStats.trace_kernel_seconds);
 * the wavefront loop is then launched kernel by kernel instead of as a CUDA graph */
#define NGI_RENDER_TIME_KERNELS 1u
/* cross-check: run the two trace stages as one-thread-per-ray kernels instead of the warp-cooperative
 * persistent kernels (same building blocks, results identical up to the film's summation order) */
#define NGI_RENDER_PER_RAY_TRACE 2u
/* bdpt cross-check: run the one-sample-per-thread megakernel (ngi_bdpt.h) instead of the wavefront
 * stages (ngi_bdpt_wave.h); same Philox counters, results identical up to the film's summation order */
#define NGI_RENDER_BDPT_PER_THREAD 4u

typedef struct NgiRenderStats {
    uint64_t paths;              /* samples processed                                         */
    uint64_t extend_rays;        /* closest-hit rays traced (exact, counted in-kernel)        */
    uint64_t shadow_rays;        /* occlusion rays traced                                     */
    uint64_t wave_iterations;    /* wavefront iterations executed                             */
    uint64_t kernel_launches;    /* kernels of this module launched by the call               */
    double gpu_seconds;          /* CUDA-event time of the render on its stream               */
    double trace_kernel_seconds; /* CUDA-event time summed over the trace kernel launches, 0 if
                                    per-kernel timing was not requested                       */
    /* per-kernel breakdown, filled only with NGI_RENDER_TIME_KERNELS (CUDA events on the render
     * stream around every launch): summed seconds and launch counts of the three wavefront kernels */
    double logic_kernel_seconds;
    double extend_kernel_seconds;
    double shadow_kernel_seconds;
    uint64_t logic_launches, extend_launches, shadow_launches;
    double reduce_seconds;       /* ngi_gpu_group_render: CUDA-event time of the NCCL film reduce on the root device */
} NgiRenderStats;

typedef struct NgiSceneInfo {
    uint64_t num_tris;
    uint64_t bvh8_nodes;
    uint64_t bvh2_nodes;
    uint64_t device_bytes;       /* bytes of geometry + BVH resident in HBM                   */
    double build_gpu_seconds;    /* LBVH + BVH8 build, CUDA-event timed                       */
    float scene_min[3], scene_max[3];
    uint32_t num_lights;
    uint32_t bvh8_max_depth;
    uint32_t bvh2_max_depth;     /* height of the binary BVH; accel = 1 (cross-check traversal) needs it < 64 */
    uint32_t reserved0;
} NgiSceneInfo;

/* ray / hit records of the geometry-parity and ray-throughput entry point */
typedef struct NgiRay { float o[3]; float tmin; float d[3]; float tmax; } NgiRay;   /* 32 B */
typedef struct NgiHit { float t, u, v; uint32_t tri; } NgiHit;                      /* 16 B */
#define NGI_NO_HIT 0xFFFFFFFFu

/* ---- entry points ------------------------------------------------------------------------ */

/* number of CUDA devices, or NGI_ERR_NO_DEVICE */
NGI_API int ngi_gpu_device_count(void);

/* Replaces Scene's Embree build, reference include/nanogi/rt.hpp:2085-2143 (rtcNewScene,
 * rtcNewTriangleMesh, rtcCommit): uploads the flattened scene to `device`, builds the LBVH on the
 * GPU and collapses it to the compressed 8-wide BVH. */
NGI_API int ngi_gpu_scene_create(const NgiSceneDesc* desc, int device, void** out_scene);
NGI_API int ngi_gpu_scene_info(void* scene, NgiSceneInfo* out);
NGI_API void ngi_gpu_scene_destroy(void* scene);

/* Replaces Renderer::Render + RenderProcess + ProcessSample_PT / ProcessSample_PTDirect (and, on the same
 * kernels, ProcessSample_LT / ProcessSample_LTDirect), reference src/nanogi.cpp:182-221, :225-440, :446-607,
 * :609-802, :804-1131. Writes the film (scaled by W*H/film_norm_samples; 0 = raw sums, which is how a host
 * accumulates several passes: --render-time, progress images, resume) into caller HOST memory:
 * float[W*H*3], row 0 = bottom. */
NGI_API int ngi_gpu_render(void* scene, const NgiRenderParams* params, float* film_rgb_host,
                           NgiRenderStats* out_stats);

/* Same, film stays in DEVICE memory (float[W*H*3]) so a multi-GPU caller can reduce it over
 * NCCL before download; `cuda_stream` is a cudaStream_t (NULL = the module's own stream). */
NGI_API int ngi_gpu_render_device(void* scene, const NgiRenderParams* params, void* film_rgb_device,
                                  void* cuda_stream, NgiRenderStats* out_stats);

/* Replaces Scene::Intersect / Scene::Visible as raw ray queries, reference
 * include/nanogi/rt.hpp:2162-2261 (rtcIntersect at :2182). any_hit = 0: closest hit
 * (t, u, v, global tri id; ties -> lowest id); any_hit = 1: occlusion (tri = 0 if any triangle
 * has tmin < t < tmax, else NGI_NO_HIT). Host buffers. `accel`: 0 = BVH8 (product), 1 = BVH2
 * (LBVH, cross-check), 2 = brute force (test aid, O(n_tris) per ray). */
NGI_API int ngi_gpu_trace(void* scene, const NgiRay* rays_host, uint64_t n, NgiHit* hits_host,
                          int any_hit, int accel);

/* Same on DEVICE buffers; returns the CUDA-event time of the trace kernel in *out_seconds. */
NGI_API int ngi_gpu_trace_device(void* scene, const void* rays_device, uint64_t n, void* hits_device,
                                 int any_hit, int accel, double* out_seconds);

/* Parity hook for SURVEY §8a rows 8-14: evaluates the device-side fp32 restatement of
 * Primitive::{SampleDirection, EvaluateDirection, EvaluateDirectionPDF} (reference
 * include/nanogi/rt.hpp:692-1336) for `n` queries on the GPU.
 * in  : per query 16 floats  {prim, type, sn[3], gn[3], wi[3], u0, u1, uComp, wo_given(0/1), pad}
 *        followed (when wo_given) by wo taken from  wo_in[3*i..]
 * out : per query 8 floats   {wo[3], fs[3], pdf, wo_valid}                                     */
NGI_API int ngi_gpu_eval_bsdf(void* scene, const float* queries_host, const float* wo_in_host,
                              uint64_t n, int force_degenerated, float* out_host);

/* ---- multi-GPU: samples sharded by index, ONE NCCL reduce of the per-GPU films ---------------
 * Replaces, across GPUs, the reference's split of [0, NumSamples) into independent chunks with private films
 * (tbb::parallel_for, reference src/nanogi.cpp:281-337) and its final gather `film += ctx.film * (W*H/N)`
 * (src/nanogi.cpp:429-437). Counter-based Philox keyed by the sample index makes the sample SET independent of the
 * GPU count; every GPU pre-scales its splats by W*H/film_norm_samples, so the gather is one ncclReduce(SUM) over NVLink.
 * NCCL is opened with dlopen("libnccl.so.2") on first use: without it these entry points return NGI_ERR_NCCL and the
 * single-GPU entry points are unaffected. */

/* rank r of world_size takes sample indices [offset, offset + count) = [r N / G, (r+1) N / G) */
NGI_API void ngi_gpu_shard_range(int64_t num_samples, int rank, int world_size, int64_t* out_offset, int64_t* out_count);

/* (a) ONE process, several devices: ncclCommInitAll. The scene is uploaded and its BVH built once, on devices[0], and
 * every built array is handed to the other devices with ncclBroadcast. devices = NULL means 0 .. num_devices-1. */
NGI_API int ngi_gpu_group_create(const NgiSceneDesc* desc, const int* devices, int num_devices, void** out_group);
/* Renderer::Render for the whole group: shards params->num_samples (from params->sample_offset) over the devices, renders
 * the shards concurrently, reduces the films onto devices[0] (ncclReduce, SUM) and copies the result to host memory. */
NGI_API int ngi_gpu_group_render(void* group, const NgiRenderParams* params, float* film_rgb_host, NgiRenderStats* out_stats);
/* the scene handle of the group's index-th device (borrowed: valid until ngi_gpu_group_destroy) */
NGI_API int ngi_gpu_group_scene(void* group, int index, void** out_scene);
NGI_API void ngi_gpu_group_destroy(void* group);

/* (b) one process PER GPU (torchrun, MPI): ncclCommInitRank. Rank 0 calls ngi_gpu_comm_get_id and hands the 128 bytes
 * to the other ranks by any host channel; every rank then creates its communicator, renders its shard with
 * ngi_gpu_render_device and calls ngi_gpu_comm_reduce_film (in place; the sum lands on `root`) on the same stream. */
typedef struct NgiCommId { char bytes[128]; } NgiCommId;
NGI_API int ngi_gpu_comm_get_id(NgiCommId* out_id);
NGI_API int ngi_gpu_comm_create(const NgiCommId* id, int rank, int world_size, int device, void** out_comm);
/* `cuda_stream`: pass the stream the film was rendered on (the one given to ngi_gpu_render_device), so that the reduce is ordered
 * after that render and before the next one. NULL = the legacy default stream; the call then returns only after the reduce has
 * completed (the module's own streams are non-blocking and would not wait for it). */
NGI_API int ngi_gpu_comm_reduce_film(void* comm, void* film_rgb_device, uint64_t num_floats, int root, void* cuda_stream);
NGI_API void ngi_gpu_comm_destroy(void* comm);

/* thread-local message of the last failing call */
NGI_API const char* ngi_gpu_last_error(void);

/* version of this ABI */
NGI_API int ngi_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NANOGI_GPU_H */

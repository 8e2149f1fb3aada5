/*
 * nanogi_host.h — C ABI of the host-side front end (libnanogi_host.so): the scene-file loader, the film
 * writers and the command-line parser that sit on the CALLER side of the render boundary. Pure C++17 +
 * zlib, no CUDA. It exists so that tests (ctypes) and other hosts can load reference scene files into
 * the POD `NgiSceneDesc` that include/nanogi_gpu.h consumes.
 *
 *   ngi_host_scene_load   <- Scene::Load            reference include/nanogi/rt.hpp:1519-2154
 *   ngi_host_save_image   <- SaveImage              reference include/nanogi/basic.hpp:506-672
 *   ngi_host_parse_cli    <- Run's option table     reference src/nanogi.cpp:2000-2048
 */
#ifndef NANOGI_HOST_H
#define NANOGI_HOST_H

#include <stdint.h>
#include "nanogi_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct NgiCliOptions {
    int32_t help;
    int32_t has_scene, has_renderer, has_num_threads, has_seed;
    char scene[1024];
    char result[1024];
    char renderer[64];
    int64_t num_samples;
    int32_t max_num_vertices;
    int32_t width, height;
    int32_t num_threads;
    int64_t grain_size;
    int64_t progress_update_interval;
    double render_time;
    double progress_image_update_interval;
    char progress_image_update_format[1024];
    int32_t gpus;
    uint32_t wave_capacity;
    uint64_t seed;
    char device[16];
    int64_t sample_offset;        /* [b200] --sample-offset */
    char resume_from[1024];       /* [b200] --resume-from   */
} NgiCliOptions;

/* Loads a schema.yml scene (YAML + OBJ meshes). aspect = width/height (src/nanogi.cpp:2069).
 * Returns 0 and an opaque handle, or -1 (message via ngi_host_last_error). */
NGI_API int ngi_host_scene_load(const char* path, double aspect, void** out_scene);
/* Fills `out` with pointers INTO the handle (valid until ngi_host_scene_free). */
NGI_API int ngi_host_scene_desc(void* scene, NgiSceneDesc* out);
/* index of the sensor primitive (last E primitive, rt.hpp:1606-1610) or -1; number of L primitives */
NGI_API int ngi_host_scene_sensor(void* scene);
NGI_API int ngi_host_scene_num_lights(void* scene);
NGI_API void ngi_host_scene_free(void* scene);

/* film: float RGB [height][width][3], row 0 = bottom. Format from the extension (.hdr/.exr/.png, and .pfm: lossless float, additive). */
NGI_API int ngi_host_save_image(const char* path, const float* film_rgb, int width, int height);

/* Reads a TexR texture image like Texture::Load (reference include/nanogi/rt.hpp:168-258): float RGB, row 0 = TOP.
 * PNG (8-bit), Radiance .hdr, binary PPM / PFM. Call with rgb_out = NULL to query width / height, then with a buffer
 * of width*height*3 floats. Returns 0 or -1. */
NGI_API int ngi_host_load_image(const char* path, int* width, int* height, float* rgb_out, uint64_t capacity_floats);

/* Parses a nanogi command line. Returns 0, or -1 on a usage error (message via ngi_host_last_error). */
NGI_API int ngi_host_parse_cli(int argc, const char* const* argv, NgiCliOptions* out);
NGI_API const char* ngi_host_usage(void);

NGI_API const char* ngi_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

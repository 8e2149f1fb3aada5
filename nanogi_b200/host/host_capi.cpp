// host_capi.cpp — C ABI over the host front end (see include/nanogi_host.h). No CUDA here.
#include <cstring>
#include <string>

#include "../../include/nanogi_host.h"
#include "cli.hpp"
#include "image_io.hpp"
#include "scene_loader.hpp"

namespace { thread_local std::string g_err; }

static void copy_str(char* dst, size_t cap, const std::string& s) {
    std::strncpy(dst, s.c_str(), cap - 1);
    dst[cap - 1] = 0;
}

extern "C" {

int ngi_host_scene_load(const char* path, double aspect, void** out_scene) {
    if (!path || !out_scene) { g_err = "invalid argument"; return -1; }
    auto* hs = new ngi::HostScene;
    ngi::Logger::get().quiet = true;
    if (!hs->Load(path, aspect)) { g_err = hs->error.empty() ? "scene load failed" : hs->error; delete hs; return -1; }
    *out_scene = hs;
    return 0;
}
int ngi_host_scene_desc(void* scene, NgiSceneDesc* out) {
    if (!scene || !out) { g_err = "invalid argument"; return -1; }
    *out = ((ngi::HostScene*)scene)->desc();
    return 0;
}
int ngi_host_scene_sensor(void* scene) { return scene ? ((ngi::HostScene*)scene)->sensor_prim : -1; }
int ngi_host_scene_num_lights(void* scene) { return scene ? (int)((ngi::HostScene*)scene)->light_prims.size() : 0; }
void ngi_host_scene_free(void* scene) { delete (ngi::HostScene*)scene; }

int ngi_host_save_image(const char* path, const float* film_rgb, int width, int height) {
    if (!path || !film_rgb || width <= 0 || height <= 0) { g_err = "invalid argument"; return -1; }
    ngi::Logger::get().quiet = true;
    if (!ngi::SaveImage(path, film_rgb, width, height)) { g_err = std::string("failed to save ") + path; return -1; }
    return 0;
}

int ngi_host_load_image(const char* path, int* width, int* height, float* rgb_out, uint64_t capacity_floats) {
    if (!path || !width || !height) { g_err = "invalid argument"; return -1; }
    std::vector<float> rgb;
    std::string err;
    if (!ngi::LoadImageRGB(path, *width, *height, rgb, err)) { g_err = err; return -1; }
    if (rgb_out) {
        if (capacity_floats < rgb.size()) { g_err = "buffer too small"; return -1; }
        std::memcpy(rgb_out, rgb.data(), rgb.size() * sizeof(float));
    }
    return 0;
}

int ngi_host_parse_cli(int argc, const char* const* argv, NgiCliOptions* out) {
    if (!out) { g_err = "invalid argument"; return -1; }
    try {
        const ngi::CliOptions o = ngi::ParseCli(argc, argv);
        std::memset(out, 0, sizeof(*out));
        out->help = o.help; out->has_scene = o.has_scene; out->has_renderer = o.has_renderer;
        out->has_num_threads = o.has_num_threads; out->has_seed = o.has_seed;
        copy_str(out->scene, sizeof(out->scene), o.scene);
        copy_str(out->result, sizeof(out->result), o.result);
        copy_str(out->renderer, sizeof(out->renderer), o.renderer);
        out->num_samples = o.num_samples; out->max_num_vertices = o.max_num_vertices;
        out->width = o.width; out->height = o.height; out->num_threads = o.num_threads;
        out->grain_size = o.grain_size; out->progress_update_interval = o.progress_update_interval;
        out->render_time = o.render_time; out->progress_image_update_interval = o.progress_image_update_interval;
        copy_str(out->progress_image_update_format, sizeof(out->progress_image_update_format), o.progress_image_update_format);
        out->gpus = o.gpus; out->wave_capacity = o.wave_capacity; out->seed = o.seed;
        copy_str(out->device, sizeof(out->device), o.device);
        out->sample_offset = o.sample_offset;
        copy_str(out->resume_from, sizeof(out->resume_from), o.resume_from);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
const char* ngi_host_usage(void) { static std::string u = ngi::CliUsage(); return u.c_str(); }
const char* ngi_host_last_error(void) { return g_err.c_str(); }

}

// nanogi_main.cpp — the `nanogi` command line with the B200 render module behind it.
//
// Host-side mirror of the reference application (src/nanogi.cpp): Run (:1993-2121), Renderer::Load
// (:117-180) and Renderer::Render (:182-221). The CLI, scene YAML, renderer names and film outputs are
// the reference's; the per-sample loop is replaced by ONE call across the C ABI (include/nanogi_gpu.h):
//     Scene's Embree commit           ->   ngi_gpu_group_create   (BVH built once on the first GPU, broadcast over NCCL to the others)
//     Renderer::Render(scene, film)   ->   ngi_gpu_group_render   (samples sharded by index, one NCCL reduce of the per-GPU films)
// `pt` and `ptdirect` (the hot path) and `lt` / `ltdirect` / `bdpt` are on the GPU path; ptmnee is reported
// as unsupported instead of silently differing. --render-time and progress images follow RenderProcess's pass loop.
// There is no CPU fallback.
#include <chrono>
#include <functional>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/nanogi_gpu.h"
#include "cli.hpp"
#include "image_io.hpp"
#include "logger.hpp"
#include "scene_loader.hpp"

using namespace ngi;

namespace {

// RendererType + string table, src/nanogi.cpp:51-71
const char* const RendererType_String[] = {"pt", "ptdirect", "lt", "ltdirect", "bdpt", "ptmnee"};

struct Renderer {
    int Type = -1;
    struct { long long NumSamples; double RenderTime; int MaxNumVertices; int Width; int Height; } Params;
    double ProgressImageUpdateInterval = -1;
    std::string ProgressImageUpdateFormat;
    int NumGpus = 1;
    unsigned long long Seed = 0;
    unsigned WaveCapacity = 0;
    long long SampleOffset = 0;          // [b200] first sample index (resume: continue the Philox sample sequence)
    std::string ResumeFrom;              // [b200] film of an earlier run (.pfm) rendered with SampleOffset samples
    void* Group = nullptr;               // [b200] ngi_gpu_group handle: the scene on every GPU + the NCCL communicators
    ~Renderer() { if (Group) ngi_gpu_group_destroy(Group); }

    // Renderer::Load, src/nanogi.cpp:117-180
    bool Load(const CliOptions& vm) {
        Type = -1;
        for (int i = 0; i < 6; i++) if (vm.renderer == RendererType_String[i]) Type = i;
        if (Type < 0) { NGI_LOG_ERROR("Invalid renderer type: " + vm.renderer); return false; }
        if (Type > NGI_RENDERER_BDPT) { NGI_LOG_ERROR("Renderer '" + vm.renderer + "' is not supported by this build (GPU path: pt, ptdirect, lt, ltdirect, bdpt)"); return false; }
        Params.NumSamples = vm.num_samples;
        Params.RenderTime = vm.render_time;
        Params.MaxNumVertices = vm.max_num_vertices;
        Params.Width = vm.width;
        Params.Height = vm.height;
        ProgressImageUpdateInterval = vm.progress_image_update_interval;       // :163-167
        ProgressImageUpdateFormat = vm.progress_image_update_format;
        if (vm.device != "gpu") { NGI_LOG_ERROR("--device " + vm.device + ": only 'gpu' is built in (no CPU fallback)"); return false; }
        const int nd = ngi_gpu_device_count();
        if (nd <= 0) { NGI_LOG_ERROR(std::string("No CUDA device: ") + ngi_gpu_last_error()); return false; }
        NumGpus = vm.gpus;
        if (NumGpus < 1 || NumGpus > nd) { NGI_LOG_ERROR("--gpus " + std::to_string(vm.gpus) + " but " + std::to_string(nd) + " device(s) present"); return false; }
        NGI_LOG_INFO("Number of GPUs: " + std::to_string(NumGpus));
        // src/nanogi.cpp:186-191: release builds seed from the clock
        Seed = vm.has_seed ? vm.seed : (unsigned long long)std::time(nullptr);
        WaveCapacity = vm.wave_capacity;
        SampleOffset = vm.sample_offset;
        ResumeFrom = vm.resume_from;
        if (!ResumeFrom.empty() && SampleOffset <= 0) { NGI_LOG_ERROR("--resume-from needs --sample-offset = the number of samples in that film"); return false; }
        return true;
    }

    // The accelerator build of the reference lives in Scene::Load (rtcCommit, rt.hpp:2085-2143) and is therefore outside its
    // "Rendering" timer; here it is the upload of the flattened scene + the GPU BVH build on the first device, the NCCL
    // communicator (--gpus > 1) and the broadcast of the built scene to the other devices.
    bool Build(const HostScene& scene) {
        const auto start = std::chrono::high_resolution_clock::now();
        const NgiSceneDesc desc = scene.desc();
        if (ngi_gpu_group_create(&desc, nullptr, NumGpus, &Group) != NGI_OK) { NGI_LOG_ERROR(ngi_gpu_last_error()); return false; }
        void* s0 = nullptr; NgiSceneInfo info{};
        if (ngi_gpu_group_scene(Group, 0, &s0) == NGI_OK && ngi_gpu_scene_info(s0, &info) == NGI_OK) {
            char buf[256];
            std::snprintf(buf, sizeof(buf), "BVH8: %llu triangles, %llu nodes, depth %u, %.1f MB on each of %d GPU(s), built in %.1f ms (GPU) / %.1f ms (total)",
                          (unsigned long long)info.num_tris, (unsigned long long)info.bvh8_nodes, info.bvh8_max_depth, info.device_bytes / 1e6, NumGpus,
                          info.build_gpu_seconds * 1e3,
                          (double)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - start).count() / 1e3);
            NGI_LOG_INFO(buf);
        }
        return true;
    }

    // {{count}} expansion of the progress image path, src/nanogi.cpp:374-393 (ctemplate there)
    static std::string ProgressPath(const std::string& format, long long count) {
        char num[32];
        std::snprintf(num, sizeof(num), "%010lld", count);
        std::string out = format;
        const std::string key = "{{count}}";
        for (size_t pos = out.find(key); pos != std::string::npos; pos = out.find(key, pos + 10)) out.replace(pos, key.size(), num);
        return out;
    }

    // Renderer::Render + RenderProcess, src/nanogi.cpp:182-221, :225-440. The reference's outer `while (true)` of
    // parallel_for passes (:243-414) is kept: a pass is ONE call across the C ABI per GPU over a contiguous range of
    // sample indices (counter-based RNG: the sample set does not depend on pass sizes or GPU count). --num-samples
    // mode is a single pass unless progress images are requested; --render-time mode runs passes until the time is up
    // (:336-346), sized to ~0.25 s of GPU work. Raw (unscaled) pass films are accumulated in fp64 and normalised by
    // W*H / processedSamples like :429-437.
    bool Render(std::vector<float>& film) const {
        const auto start = std::chrono::high_resolution_clock::now();
        const size_t npx = (size_t)Params.Width * Params.Height;
        std::vector<double> acc(npx * 3, 0.0);                 // raw sum of sample contributions
        long long processed = 0;                               // samples behind `acc` (incl. a resumed film's)
        if (!ResumeFrom.empty()) {
            int w = 0, h = 0; std::vector<float> rgb; std::string err;
            if (!LoadImageRGB(ResumeFrom, w, h, rgb, err)) { NGI_LOG_ERROR(err); return false; }
            if (w != Params.Width || h != Params.Height) { NGI_LOG_ERROR("--resume-from: image size differs from --width/--height"); return false; }
            // the file holds film * W*H / N0 with row 0 = top; undo both
            const double unscale = (double)SampleOffset / (double)npx;
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++)
                    for (int k = 0; k < 3; k++) acc[((size_t)(h - 1 - y) * w + x) * 3 + k] = (double)rgb[((size_t)y * w + x) * 3 + k] * unscale;
            processed = SampleOffset;
        }
        const long long first = SampleOffset;
        std::vector<float> pass_film(npx * 3);
        NgiRenderStats stats{};
        unsigned long long ext = 0, sh = 0; double gpu_s = 0, reduce_s = 0;
        auto gather = [&](std::vector<float>& out) {
            const double scale = processed > 0 ? (double)npx / (double)processed : 0.0;
            out.resize(npx * 3);
            for (size_t i = 0; i < npx * 3; i++) out[i] = (float)(acc[i] * scale);
        };
        const bool timed = Params.RenderTime > 0;
        const bool passes = timed || ProgressImageUpdateInterval > 0;
        long long pass_size = passes ? (1ll << 24) : Params.NumSamples;
        long long done = 0, progressImageCount = 0;
        auto prevImageUpdateTime = start;
        auto seconds_since = [](std::chrono::high_resolution_clock::time_point t) {
            return (double)(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t).count()) / 1000.0;
        };
        while (true) {
            long long n = pass_size;
            if (!timed) n = std::min(n, Params.NumSamples - done);
            if (n > 0) {
                const long long base = first + done;
                const auto pass_start = std::chrono::high_resolution_clock::now();
                // one call for all GPUs: the module shards [base, base + n) by index, renders the shards concurrently and sums the
                // films with one NCCL reduce; raw sums (film_norm_samples = 0) so that passes add up in fp64 here
                NgiRenderParams p{};
                p.struct_size = sizeof(p);
                p.renderer = Type;
                p.num_samples = n; p.sample_offset = base; p.film_norm_samples = 0;
                p.max_num_vertices = Params.MaxNumVertices; p.width = Params.Width; p.height = Params.Height;
                p.seed = Seed; p.wave_capacity = WaveCapacity;
                if (ngi_gpu_group_render(Group, &p, pass_film.data(), &stats) != NGI_OK) { NGI_LOG_ERROR(ngi_gpu_last_error()); return false; }
                for (size_t i = 0; i < npx * 3; i++) acc[i] += (double)pass_film[i];
                ext += stats.extend_rays; sh += stats.shadow_rays; reduce_s += stats.reduce_seconds;
                const double pass_gpu = stats.gpu_seconds + stats.reduce_seconds;
                gpu_s += pass_gpu;
                done += n; processed += n;
                if (passes && seconds_since(pass_start) < 0.25 && pass_size < (1ll << 34)) pass_size *= 2;
            }
            const double elapsed = seconds_since(start);
            if (!timed) {
                char buf[64]; std::snprintf(buf, sizeof(buf), "Progress: %.1f%%", Params.NumSamples > 0 ? (double)done / Params.NumSamples * 100.0 : 100.0);
                if (passes) NGI_LOG_INFO(buf);
            } else {
                char buf[96]; std::snprintf(buf, sizeof(buf), "Progress: %.1f%% (%.1fs / %.1fs)", elapsed / Params.RenderTime * 100.0, elapsed, Params.RenderTime);
                NGI_LOG_INFO(buf);
            }
            const bool finished = timed ? elapsed > Params.RenderTime : done >= Params.NumSamples;
            if (ProgressImageUpdateInterval > 0 && !finished && seconds_since(prevImageUpdateTime) > ProgressImageUpdateInterval) {   // :356-404
                std::vector<float> snapshot;
                gather(snapshot);
                progressImageCount++;
                NGI_LOG_INFO("Saving progress: ");
                NGI_LOG_INDENTER();
                SaveImage(ProgressPath(ProgressImageUpdateFormat, progressImageCount), snapshot.data(), Params.Width, Params.Height);
                prevImageUpdateTime = std::chrono::high_resolution_clock::now();
            }
            if (finished) break;
        }
        gather(film);                                           // :429-437
        const double elapsed = seconds_since(start);
        NGI_LOG_INFO("Progress: 100.0%");
        NGI_LOG_INFO("# of samples: " + std::to_string(done));
        NGI_LOG_INFO("Elapesed time: " + std::to_string(elapsed));
        char buf[256];
        std::snprintf(buf, sizeof(buf), "GPU render: %.3f s, %.1f Mpaths/s, %.1f Mrays/s (extend %llu, shadow %llu)", gpu_s,
                      gpu_s > 0 ? done / gpu_s / 1e6 : 0.0, gpu_s > 0 ? (ext + sh) / gpu_s / 1e6 : 0.0, ext, sh);
        NGI_LOG_INFO(buf);
        if (NumGpus > 1) {
            std::snprintf(buf, sizeof(buf), "NCCL film reduce over %d GPUs: %.3f ms", NumGpus, reduce_s * 1e3);
            NGI_LOG_INFO(buf);
        }
        return true;
    }
};

// Run, src/nanogi.cpp:1993-2121
bool Run(int argc, char** argv) {
    CliOptions vm;
    try {
        vm = ParseCli(argc, argv);
        if (vm.help) { std::cout << CliUsage() << std::endl; return true; }  // :2030-2035 (returns 1)
    } catch (const CliError& e) {
        std::cerr << "ERROR : " << e.what() << std::endl;
        return false;
    }
    Logger::get().quiet = vm.quiet;
    NGI_LOG_INFO("nanogi");
    NGI_LOG_INFO("B200 (sm_100a) render module behind the reference CLI");

    HostScene scene;
    {
        NGI_LOG_INFO("Loading scene");
        NGI_LOG_INDENTER();
        if (!vm.has_scene) { NGI_LOG_ERROR("the option '--scene' is required"); return false; }
        if (!scene.Load(vm.scene, (double)vm.width / vm.height)) return false;   // :2069
    }
    Renderer renderer;
    {
        NGI_LOG_INFO("Initializing renderer");
        NGI_LOG_INDENTER();
        if (!renderer.Load(vm)) return false;
    }
    {
        NGI_LOG_INFO("Building GPU scene");
        NGI_LOG_INDENTER();
        if (!renderer.Build(scene)) return false;
    }
    std::vector<float> film;
    {
        NGI_LOG_INFO("Rendering");
        NGI_LOG_INDENTER();
        if (!renderer.Render(film)) return false;
    }
    {
        NGI_LOG_INFO("Saving rendered image");
        NGI_LOG_INDENTER();
        SaveImage(vm.result, film.data(), vm.width, vm.height);   // :2113 (result ignored by the reference too)
    }
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    int result = EXIT_SUCCESS;
    try {
        if (!Run(argc, argv)) result = EXIT_FAILURE;
    } catch (const std::exception& e) {
        NGI_LOG_ERROR("EXCEPTION | " + std::string(e.what()));
        result = EXIT_FAILURE;
    }
    return result;
}

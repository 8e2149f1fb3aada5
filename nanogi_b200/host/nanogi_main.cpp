// nanogi_main.cpp — the `nanogi` command line with the B200 render module behind it.
//
// Host-side mirror of the reference application (src/nanogi.cpp): Run (:1993-2121), Renderer::Load
// (:117-180) and Renderer::Render (:182-221). The CLI, scene YAML, renderer names and film outputs are
// the reference's; the per-sample loop is replaced by ONE call across the C ABI (include/nanogi_gpu.h):
//     Renderer::Render(scene, film)   ->   ngi_gpu_scene_create + ngi_gpu_render
// Only `pt` and `ptdirect` are on the GPU path; the other reference renderers (lt, ltdirect, bdpt, ptmnee)
// are reported as unsupported instead of silently differing. There is no CPU fallback.
#include <chrono>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <string>
#include <thread>
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
    int NumGpus = 1;
    unsigned long long Seed = 0;
    unsigned WaveCapacity = 0;

    // Renderer::Load, src/nanogi.cpp:117-180
    bool Load(const CliOptions& vm) {
        Type = -1;
        for (int i = 0; i < 6; i++) if (vm.renderer == RendererType_String[i]) Type = i;
        if (Type < 0) { NGI_LOG_ERROR("Invalid renderer type: " + vm.renderer); return false; }
        if (Type > 1) { NGI_LOG_ERROR("Renderer '" + vm.renderer + "' is not supported by this build (GPU path: pt, ptdirect)"); return false; }
        Params.NumSamples = vm.num_samples;
        Params.RenderTime = vm.render_time;
        Params.MaxNumVertices = vm.max_num_vertices;
        Params.Width = vm.width;
        Params.Height = vm.height;
        if (vm.device != "gpu") { NGI_LOG_ERROR("--device " + vm.device + ": only 'gpu' is built in (no CPU fallback)"); return false; }
        if (Params.RenderTime > 0) { NGI_LOG_ERROR("--render-time mode is not supported by this build; use --num-samples"); return false; }
        const int nd = ngi_gpu_device_count();
        if (nd <= 0) { NGI_LOG_ERROR(std::string("No CUDA device: ") + ngi_gpu_last_error()); return false; }
        NumGpus = vm.gpus;
        if (NumGpus < 1 || NumGpus > nd) { NGI_LOG_ERROR("--gpus " + std::to_string(vm.gpus) + " but " + std::to_string(nd) + " device(s) present"); return false; }
        NGI_LOG_INFO("Number of GPUs: " + std::to_string(NumGpus));
        // src/nanogi.cpp:186-191: release builds seed from the clock
        Seed = vm.has_seed ? vm.seed : (unsigned long long)std::time(nullptr);
        WaveCapacity = vm.wave_capacity;
        return true;
    }

    // Renderer::Render, src/nanogi.cpp:182-221
    bool Render(const HostScene& scene, std::vector<float>& film) const {
        const auto start = std::chrono::high_resolution_clock::now();
        const size_t npx = (size_t)Params.Width * Params.Height;
        const NgiSceneDesc desc = scene.desc();
        std::vector<std::vector<float>> films(NumGpus, std::vector<float>(npx * 3));
        std::vector<NgiRenderStats> stats(NumGpus);
        std::vector<std::string> errors(NumGpus);
        auto work = [&](int g) {
            void* h = nullptr;
            if (ngi_gpu_scene_create(&desc, g, &h) != NGI_OK) { errors[g] = ngi_gpu_last_error(); return; }
            NgiRenderParams p{};
            p.struct_size = sizeof(p);
            p.renderer = Type;
            // samples sharded by index: GPU g takes the contiguous range [g N/G, (g+1) N/G)
            const long long lo = Params.NumSamples * g / NumGpus, hi = Params.NumSamples * (g + 1) / NumGpus;
            p.num_samples = hi - lo; p.sample_offset = lo; p.film_norm_samples = Params.NumSamples;
            p.max_num_vertices = Params.MaxNumVertices; p.width = Params.Width; p.height = Params.Height;
            p.seed = Seed; p.wave_capacity = WaveCapacity;
            if (ngi_gpu_render(h, &p, films[g].data(), &stats[g]) != NGI_OK) errors[g] = ngi_gpu_last_error();
            ngi_gpu_scene_destroy(h);
        };
        std::vector<std::thread> th;
        for (int g = 1; g < NumGpus; g++) th.emplace_back(work, g);
        work(0);
        for (auto& t : th) t.join();
        for (int g = 0; g < NumGpus; g++) if (!errors[g].empty()) { NGI_LOG_ERROR("GPU " + std::to_string(g) + ": " + errors[g]); return false; }
        film = films[0];
        for (int g = 1; g < NumGpus; g++) for (size_t i = 0; i < npx * 3; i++) film[i] += films[g][i];
        const auto end = std::chrono::high_resolution_clock::now();
        const double elapsed = (double)(std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count()) / 1000.0;
        unsigned long long ext = 0, sh = 0; double gs = 0;
        for (auto& s : stats) { ext += s.extend_rays; sh += s.shadow_rays; gs = std::max(gs, s.gpu_seconds); }
        NGI_LOG_INFO("Progress: 100.0%");
        NGI_LOG_INFO("# of samples: " + std::to_string(Params.NumSamples));
        NGI_LOG_INFO("Elapesed time: " + std::to_string(elapsed));
        char buf[256];
        std::snprintf(buf, sizeof(buf), "GPU render: %.3f s, %.1f Mpaths/s, %.1f Mrays/s (extend %llu, shadow %llu)", gs,
                      Params.NumSamples / gs / 1e6, (ext + sh) / gs / 1e6, ext, sh);
        NGI_LOG_INFO(buf);
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
    std::vector<float> film;
    {
        NGI_LOG_INFO("Rendering");
        NGI_LOG_INDENTER();
        if (!renderer.Render(scene, film)) return false;
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

// cli.hpp — nanogi's command line, reproduced name for name (reference src/nanogi.cpp:2000-2023):
//   nanogi [options] <renderer> <scene> <result> <width> <height>
// Same option names, short names, defaults and positional order; `-h` is --height, help is --help only.
// Additive flags of this build: --gpus N, --seed S, --wave-capacity P, --device gpu.
// boost::program_options is replaced by this ~150-line parser (long `--k v`, `--k=v`, short `-k v`, `-kv`).
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ngi {

struct CliOptions {
    bool help = false;
    std::string scene;
    bool has_scene = false;
    std::string result = "render.hdr";                 // :2004
    std::string renderer;                              // :2005 required
    bool has_renderer = false;
    long long num_samples = 10000000LL;                // :2006
    int max_num_vertices = -1;                         // :2007
    int width = 1280;                                  // :2008
    int height = 720;                                  // :2009
    int num_threads = 0; bool has_num_threads = false; // :2010 (accepted, unused by the GPU path)
    long long grain_size = 10000;                      // :2014 (accepted, unused by the GPU path)
    long long progress_update_interval = 100000;       // :2016
    double render_time = -1;                           // :2017
    double progress_image_update_interval = -1;        // :2018
    std::string progress_image_update_format = "progress/{{count}}.png";  // :2019
    // additive
    int gpus = 1;
    unsigned long long seed = 0; bool has_seed = false;
    unsigned wave_capacity = 0;
    long long sample_offset = 0;
    std::string resume_from;
    std::string device = "gpu";
    bool quiet = false;
};

struct CliError : std::runtime_error { using std::runtime_error::runtime_error; };

inline std::string CliUsage() {
    std::ostringstream o;
    o << "Usage: nanogi [options] <renderer> <scene> <result> <width> <height>\n"
      << "Allowed options:\n"
      << "  --help                                Display help message\n"
      << "  -i [ --scene ] arg                    Scene file\n"
      << "  -o [ --result ] arg (=render.hdr)     Rendered result\n"
      << "  -r [ --renderer ] arg                 Rendering technique\n"
      << "  -n [ --num-samples ] arg (=10000000)  Number of samples\n"
      << "  -m [ --max-num-vertices ] arg (=-1)   Maximum number of vertices\n"
      << "  -w [ --width ] arg (=1280)            Width of the rendered image\n"
      << "  -h [ --height ] arg (=720)            Height of the rendered image\n"
      << "  -j [ --num-threads ] arg              Number of threads\n"
      << "  --grain-size arg (=10000)             Grain size\n"
      << "  --progress-update-interval arg (=100000)\n"
      << "                                        Progress update interval\n"
      << "  -t [ --render-time ] arg (=-1)        Render time in seconds (-1 to use # of \n"
      << "                                        samples)\n"
      << "  --progress-image-update-interval arg (=-1)\n"
      << "                                        Progress image update interval (-1: \n"
      << "                                        disable)\n"
      << "  --progress-image-update-format arg (=progress/{{count}}.png)\n"
      << "                                        Progress image update format string \n"
      << "                                         - {{count}}: image count\n"
      << "  --gpus arg (=1)                       [b200] number of GPUs (samples sharded by index)\n"
      << "  --seed arg                            [b200] Philox seed (default: time)\n"
      << "  --wave-capacity arg (=0)              [b200] path slots in flight (0 = default)\n"
      << "  --sample-offset arg (=0)              [b200] first sample index (continue an earlier run's sample sequence)\n"
      << "  --resume-from arg                     [b200] .pfm film of an earlier run with --sample-offset samples; the new\n"
      << "                                        samples are added to it\n"
      << "  --device arg (=gpu)                   [b200] only 'gpu' is built in; there is no CPU fallback\n";
    return o.str();
}

inline CliOptions ParseCli(int argc, const char* const* argv) {
    CliOptions o;
    struct Opt { const char* lname; char sname; bool takes_value; };
    static const Opt opts[] = {
        {"help", 0, false}, {"scene", 'i', true}, {"result", 'o', true}, {"renderer", 'r', true},
        {"num-samples", 'n', true}, {"max-num-vertices", 'm', true}, {"width", 'w', true}, {"height", 'h', true},
        {"num-threads", 'j', true}, {"grain-size", 0, true}, {"progress-update-interval", 0, true},
        {"render-time", 't', true}, {"progress-image-update-interval", 0, true}, {"progress-image-update-format", 0, true},
        {"gpus", 0, true}, {"seed", 0, true}, {"wave-capacity", 0, true}, {"sample-offset", 0, true}, {"resume-from", 0, true}, {"device", 0, true}, {"quiet", 0, false},
    };
    std::map<std::string, std::string> vm;
    std::vector<std::string> positional;
    auto find_long = [&](const std::string& n) -> const Opt* { for (auto& p : opts) if (n == p.lname) return &p; return nullptr; };
    auto find_short = [&](char c) -> const Opt* { for (auto& p : opts) if (p.sname && p.sname == c) return &p; return nullptr; };
    auto is_number = [](const std::string& s) { char* e = nullptr; std::strtod(s.c_str(), &e); return !s.empty() && e && *e == 0; };
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string name = a.substr(2), val; bool has_val = false;
            const size_t eq = name.find('=');
            if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
            const Opt* p = find_long(name);
            if (!p) throw CliError("unrecognised option '--" + name + "'");
            if (p->takes_value) {
                if (!has_val) { if (i + 1 >= argc) throw CliError("the required argument for option '--" + name + "' is missing"); val = argv[++i]; }
                vm[p->lname] = val;
            } else vm[p->lname] = "1";
        } else if (a.size() >= 2 && a[0] == '-' && !is_number(a)) {
            const Opt* p = find_short(a[1]);
            if (!p) throw CliError("unrecognised option '" + a + "'");
            std::string val;
            if (p->takes_value) {
                if (a.size() > 2) val = a.substr(2);
                else { if (i + 1 >= argc) throw CliError(std::string("the required argument for option '--") + p->lname + "' is missing"); val = argv[++i]; }
                vm[p->lname] = val;
            } else vm[p->lname] = "1";
        } else positional.push_back(a);
    }
    // positional: renderer scene result width height (:2023)
    static const char* posnames[] = {"renderer", "scene", "result", "width", "height"};
    if (positional.size() > 5) throw CliError("too many positional options have been specified on the command line");
    for (size_t i = 0; i < positional.size(); i++) {
        if (vm.count(posnames[i])) throw CliError(std::string("option '--") + posnames[i] + "' cannot be specified more than once");
        vm[posnames[i]] = positional[i];
    }
    auto to_ll = [](const std::string& k, const std::string& v) { char* e = nullptr; long long r = std::strtoll(v.c_str(), &e, 10); if (v.empty() || *e) throw CliError("the argument ('" + v + "') for option '--" + k + "' is invalid"); return r; };
    auto to_d = [](const std::string& k, const std::string& v) { char* e = nullptr; double r = std::strtod(v.c_str(), &e); if (v.empty() || *e) throw CliError("the argument ('" + v + "') for option '--" + k + "' is invalid"); return r; };
    o.help = vm.count("help") > 0 || argc == 1;  // :2030
    if (vm.count("scene")) { o.scene = vm["scene"]; o.has_scene = true; }
    if (vm.count("result")) o.result = vm["result"];
    if (vm.count("renderer")) { o.renderer = vm["renderer"]; o.has_renderer = true; }
    if (vm.count("num-samples")) o.num_samples = to_ll("num-samples", vm["num-samples"]);
    if (vm.count("max-num-vertices")) o.max_num_vertices = (int)to_ll("max-num-vertices", vm["max-num-vertices"]);
    if (vm.count("width")) o.width = (int)to_ll("width", vm["width"]);
    if (vm.count("height")) o.height = (int)to_ll("height", vm["height"]);
    if (vm.count("num-threads")) { o.num_threads = (int)to_ll("num-threads", vm["num-threads"]); o.has_num_threads = true; }
    if (vm.count("grain-size")) o.grain_size = to_ll("grain-size", vm["grain-size"]);
    if (vm.count("progress-update-interval")) o.progress_update_interval = to_ll("progress-update-interval", vm["progress-update-interval"]);
    if (vm.count("render-time")) o.render_time = to_d("render-time", vm["render-time"]);
    if (vm.count("progress-image-update-interval")) o.progress_image_update_interval = to_d("progress-image-update-interval", vm["progress-image-update-interval"]);
    if (vm.count("progress-image-update-format")) o.progress_image_update_format = vm["progress-image-update-format"];
    if (vm.count("gpus")) o.gpus = (int)to_ll("gpus", vm["gpus"]);
    if (vm.count("seed")) { o.seed = (unsigned long long)to_ll("seed", vm["seed"]); o.has_seed = true; }
    if (vm.count("wave-capacity")) o.wave_capacity = (unsigned)to_ll("wave-capacity", vm["wave-capacity"]);
    if (vm.count("sample-offset")) o.sample_offset = to_ll("sample-offset", vm["sample-offset"]);
    if (vm.count("resume-from")) o.resume_from = vm["resume-from"];
    if (vm.count("device")) o.device = vm["device"];
    o.quiet = vm.count("quiet") > 0;
    if (!o.help && !o.has_renderer) throw CliError("the option '--renderer' is required but missing");  // :2039
    return o;
}

}  // namespace ngi

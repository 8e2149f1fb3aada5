// logger.hpp — plain stderr logger keeping the reference's line shape
//   "| LEVEL t.ttt | @line | #tid | indent msg"   (include/nanogi/basic.hpp:214-220)
// The reference's asio logging thread (basic.hpp:100-317) is control plane, out of scope.
#pragma once
#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>

namespace ngi {

struct Logger {
    static Logger& get() { static Logger l; return l; }
    std::chrono::high_resolution_clock::time_point start = std::chrono::high_resolution_clock::now();
    int indent = 0;
    bool quiet = false;
    std::mutex mu;
    void log(const char* level, int line, const std::string& msg) {
        if (quiet && level[0] != 'E') return;
        std::lock_guard<std::mutex> lock(mu);
        const double t = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
        std::fprintf(stderr, "| %-5s %.3f | @%4d | #%2d | %s%s\n", level, t, line, 0, std::string(indent * 4, '.').c_str(), msg.c_str());
    }
};
struct LogIndenter {
    LogIndenter() { Logger::get().indent++; }
    ~LogIndenter() { Logger::get().indent--; }
};

}  // namespace ngi

#define NGI_LOG_ERROR(msg) ::ngi::Logger::get().log("ERROR", __LINE__, msg)
#define NGI_LOG_WARN(msg)  ::ngi::Logger::get().log("WARN", __LINE__, msg)
#define NGI_LOG_INFO(msg)  ::ngi::Logger::get().log("INFO", __LINE__, msg)
#define NGI_LOG_INDENTER() ::ngi::LogIndenter _ngi_log_indenter_##__LINE__

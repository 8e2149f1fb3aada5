// Stand-in: Assimp::DefaultLogger::create / get()->attachStream (rt.hpp:1540-1547). Streams are kept and never written to.
#pragma once
#include <memory>
#include <vector>
#include "LogStream.hpp"
namespace Assimp {
class Logger {
public:
    enum LogSeverity { NORMAL, VERBOSE };
    enum ErrorSeverity { Debugging = 1, Info = 2, Warn = 4, Err = 8 };
    bool attachStream(LogStream* s, unsigned int = Debugging | Err | Warn | Info) { streams_.emplace_back(s); return true; }
private:
    std::vector<std::unique_ptr<LogStream>> streams_;
};
class DefaultLogger {
public:
    static Logger* create(const char* = "", Logger::LogSeverity = Logger::NORMAL, unsigned int = 0, void* = nullptr) { return get(); }
    static Logger* get() { static Logger l; return &l; }
    static void kill() {}
};
}

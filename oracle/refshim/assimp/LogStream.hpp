// Stand-in: Assimp::LogStream base class (rt.hpp:2288-2310 derives from it). See scene.h.
#pragma once
namespace Assimp { class LogStream { public: virtual ~LogStream() {} virtual void write(const char* message) = 0; }; }

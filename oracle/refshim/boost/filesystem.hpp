// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// boost::filesystem::path (parent_path, extension, string, operator/, == "..."), exists, create_directories.
#pragma once
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
namespace boost { namespace filesystem {
class path {
public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
    const char* c_str() const { return s_.c_str(); }
    bool empty() const { return s_.empty(); }
    path parent_path() const { const size_t p = s_.find_last_of('/'); return p == std::string::npos ? path() : path(p == 0 ? "/" : s_.substr(0, p)); }
    path filename() const { const size_t p = s_.find_last_of('/'); return p == std::string::npos ? *this : path(s_.substr(p + 1)); }
    path extension() const { const std::string f = filename().s_; const size_t p = f.find_last_of('.'); return (p == std::string::npos || p == 0) ? path() : path(f.substr(p)); }
    path operator/(const path& o) const { if (s_.empty()) return o; if (!o.s_.empty() && o.s_[0] == '/') return o; return path(s_.back() == '/' ? s_ + o.s_ : s_ + "/" + o.s_); }
    bool operator==(const path& o) const { return s_ == o.s_; }
    bool operator!=(const path& o) const { return s_ != o.s_; }
private:
    std::string s_;
};
inline bool operator==(const path& a, const char* b) { return a.string() == b; }
inline bool operator!=(const path& a, const char* b) { return a.string() != b; }
inline bool exists(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool create_directories(const path& p) {
    if (p.empty() || exists(p)) return !p.empty();
    const path parent = p.parent_path();
    if (!parent.empty() && parent != p && !exists(parent)) create_directories(parent);
    return ::mkdir(p.c_str(), 0777) == 0;
}
}}

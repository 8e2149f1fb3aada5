// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// boost::format("...%d...%.3f...%s") % a % b ... ; boost::str(f): printf-style directives only (all the reference uses).
#pragma once
#include <atomic>
#include <cctype>
#include <cstdio>
#include <string>
#include <type_traits>
namespace boost {
class format {
public:
    explicit format(const std::string& f) : fmt_(f), pos_(0) { flush_literal(); }
    template <class T> format& operator%(const T& v) { feed(v); return *this; }
    std::string str() const { return out_; }
private:
    // copies literal text up to the next directive (handles "%%")
    void flush_literal() {
        while (pos_ < fmt_.size()) {
            if (fmt_[pos_] == '%') {
                if (pos_ + 1 < fmt_.size() && fmt_[pos_ + 1] == '%') { out_ += '%'; pos_ += 2; continue; }
                return;
            }
            out_ += fmt_[pos_++];
        }
    }
    // returns the directive without its conversion character and the conversion character
    bool next_directive(std::string& spec, char& conv) {
        if (pos_ >= fmt_.size() || fmt_[pos_] != '%') return false;
        size_t e = pos_ + 1;
        while (e < fmt_.size() && !std::isalpha((unsigned char)fmt_[e])) e++;
        while (e < fmt_.size() && (fmt_[e] == 'l' || fmt_[e] == 'h' || fmt_[e] == 'z')) e++;   // length modifiers are replaced below
        if (e >= fmt_.size()) return false;
        spec = fmt_.substr(pos_, e - pos_);
        while (!spec.empty() && (spec.back() == 'l' || spec.back() == 'h' || spec.back() == 'z')) spec.pop_back();
        conv = fmt_[e];
        pos_ = e + 1;
        return true;
    }
    template <class... A> void emit(const std::string& f, A... a) {
        char buf[512];
        const int n = std::snprintf(buf, sizeof(buf), f.c_str(), a...);
        if (n >= (int)sizeof(buf)) { std::string big((size_t)n + 1, '\0'); std::snprintf(&big[0], big.size(), f.c_str(), a...); big.resize((size_t)n); out_ += big; }
        else if (n > 0) out_ += buf;
    }
    template <class T> typename std::enable_if<std::is_integral<T>::value>::type feed(const T& v) {
        std::string spec; char conv;
        if (next_directive(spec, conv)) {
            if (conv == 'f' || conv == 'e' || conv == 'g') emit(spec + conv, (double)v);
            else if (conv == 's') emit(spec + "lld", (long long)v);
            else if (conv == 'x' || conv == 'X' || conv == 'u') emit(spec + "ll" + conv, (unsigned long long)v);
            else emit(spec + "lld", (long long)v);
        }
        flush_literal();
    }
    template <class T> typename std::enable_if<std::is_floating_point<T>::value>::type feed(const T& v) {
        std::string spec; char conv;
        if (next_directive(spec, conv)) {
            if (conv == 'd' || conv == 'i') emit(spec + "lld", (long long)v);
            else if (conv == 's') emit(spec + "g", (double)v);
            else emit(spec + conv, (double)v);
        }
        flush_literal();
    }
    template <class T> void feed(const std::atomic<T>& v) { feed((T)v.load()); }   // "# of samples: %d" % processedSamples (src/nanogi.cpp:421)
    void feed(const std::string& v) { std::string spec; char conv; if (next_directive(spec, conv)) emit(spec + "s", v.c_str()); flush_literal(); }
    void feed(const char* v) { std::string spec; char conv; if (next_directive(spec, conv)) emit(spec + "s", v); flush_literal(); }
    template <class T> void feed(T* const& v) { std::string spec; char conv; if (next_directive(spec, conv)) emit(spec + "llx", (unsigned long long)(size_t)v); flush_literal(); }
    std::string fmt_, out_;
    size_t pos_;
};
inline std::string str(const format& f) { return f.str(); }
}

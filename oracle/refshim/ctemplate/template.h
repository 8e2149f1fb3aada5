// Stand-in for the ctemplate subset used for the progress image path (src/nanogi.cpp:376-392): {{name}} substitution.
#pragma once
#include <map>
#include <string>
namespace ctemplate {
enum Strip { DO_NOT_STRIP, STRIP_BLANK_LINES, STRIP_WHITESPACE };
class TemplateDictionary {
public:
    explicit TemplateDictionary(const std::string&) {}
    std::string& operator[](const std::string& key) { return values[key]; }
    std::map<std::string, std::string> values;
};
class Template {
public:
    static Template* StringToTemplate(const std::string& text, Strip) { Template* t = new Template; t->text_ = text; return t; }
    bool Expand(std::string* out, const TemplateDictionary* dict) const {
        std::string s = text_;
        for (const auto& kv : dict->values) {
            const std::string key = "{{" + kv.first + "}}";
            for (size_t p = s.find(key); p != std::string::npos; p = s.find(key, p + kv.second.size())) s.replace(p, key.size(), kv.second);
        }
        *out = s;
        return true;
    }
private:
    std::string text_;
};
}

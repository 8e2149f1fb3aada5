// yaml_lite.hpp — the YAML subset that schema.yml scene documents use (reference schema.yml:33-335;
// read by yaml-cpp at include/nanogi/rt.hpp:1525): block maps, block sequences ("- "), flow
// sequences "[a, b, c]" (nestable), flow maps "{a: b}", plain / single- / double-quoted scalars,
// comments, "---" document start. Zero dependencies. Errors throw std::runtime_error with a line number,
// which the scene loader maps to the reference's "YAML exception: ..." message (rt.hpp:2147-2151).
#pragma once
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ngi { namespace yaml {

struct Node {
    enum Kind { Null, Scalar, Seq, Map } kind = Null;
    std::string scalar;
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;
    int line = 0;

    bool defined() const { return kind != Null; }
    explicit operator bool() const { return defined(); }
    size_t size() const { return kind == Seq ? seq.size() : kind == Map ? map.size() : 0; }
    const Node& operator[](const std::string& key) const {
        static const Node null_node;
        if (kind != Map) return null_node;
        for (auto& kv : map) if (kv.first == key) return kv.second;
        return null_node;
    }
    const Node& operator[](size_t i) const {
        static const Node null_node;
        if (kind != Seq || i >= seq.size()) return null_node;
        return seq[i];
    }
    const Node& operator[](int i) const { return (*this)[(size_t)i]; }
    bool has(const std::string& key) const { return (*this)[key].defined(); }

    [[noreturn]] void fail(const std::string& what) const {
        throw std::runtime_error("yaml: " + what + " (line " + std::to_string(line) + ")");
    }
    std::string as_string() const {
        if (kind != Scalar) fail("bad conversion: expected a scalar");
        return scalar;
    }
    double as_double() const {
        if (kind != Scalar) fail("bad conversion: expected a number");
        char* end = nullptr;
        const double v = std::strtod(scalar.c_str(), &end);
        if (end == scalar.c_str() || *end != 0) fail("bad conversion: '" + scalar + "' is not a number");
        return v;
    }
    long long as_int() const {
        if (kind != Scalar) fail("bad conversion: expected an integer");
        char* end = nullptr;
        const long long v = std::strtoll(scalar.c_str(), &end, 10);
        if (end == scalar.c_str() || *end != 0) fail("bad conversion: '" + scalar + "' is not an integer");
        return v;
    }
    bool as_bool() const {
        if (kind != Scalar) fail("bad conversion: expected a bool");
        const std::string& s = scalar;
        if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "Yes" || s == "on" || s == "y" || s == "Y") return true;
        if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "No" || s == "off" || s == "n" || s == "N") return false;
        fail("bad conversion: '" + s + "' is not a bool");
    }
};

namespace detail {

struct Line { int indent; std::string text; int no; };

inline std::string rtrim(const std::string& s) {
    size_t e = s.size();
    while (e > 0 && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\r')) e--;
    return s.substr(0, e);
}
inline std::string ltrim(const std::string& s) {
    size_t b = 0;
    while (b < s.size() && (s[b] == ' ' || s[b] == '\t')) b++;
    return s.substr(b);
}
inline std::string trim(const std::string& s) { return ltrim(rtrim(s)); }

// strip a trailing comment that is outside quotes ('#' at start or preceded by whitespace)
inline std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); i++) {
        const char c = s[i];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '\'' || c == '"') { q = c; continue; }
        if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;

    explicit Parser(const std::string& text) {
        size_t b = 0; int no = 0;
        while (b <= text.size()) {
            size_t e = text.find('\n', b);
            if (e == std::string::npos) e = text.size();
            std::string raw = text.substr(b, e - b);
            no++;
            b = e + 1;
            std::string s = rtrim(strip_comment(raw));
            if (s.empty()) continue;
            int ind = 0;
            while ((size_t)ind < s.size() && s[ind] == ' ') ind++;
            if ((size_t)ind < s.size() && s[ind] == '\t') throw std::runtime_error("yaml: tab indentation (line " + std::to_string(no) + ")");
            std::string t = s.substr(ind);
            if (t == "---" || t == "...") continue;
            if (t[0] == '%') continue;
            lines.push_back({ind, t, no});
        }
    }

    // ---- flow syntax (single line) ------------------------------------------------------------
    static void skip_ws(const std::string& s, size_t& i) { while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++; }

    static std::string parse_quoted(const std::string& s, size_t& i, int no) {
        const char q = s[i++];
        std::string out;
        while (i < s.size()) {
            char c = s[i++];
            if (q == '\'' && c == '\'') {
                if (i < s.size() && s[i] == '\'') { out += '\''; i++; continue; }
                return out;
            }
            if (q == '"' && c == '"') return out;
            if (q == '"' && c == '\\' && i < s.size()) {
                char n = s[i++];
                switch (n) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case '\\': out += '\\'; break; case '"': out += '"'; break; default: out += n; }
                continue;
            }
            out += c;
        }
        throw std::runtime_error("yaml: unterminated quoted scalar (line " + std::to_string(no) + ")");
    }

    static Node parse_flow(const std::string& s, size_t& i, int no) {
        skip_ws(s, i);
        Node n; n.line = no;
        if (i >= s.size()) return n;
        if (s[i] == '[') {
            i++; n.kind = Node::Seq;
            skip_ws(s, i);
            if (i < s.size() && s[i] == ']') { i++; return n; }
            while (true) {
                n.seq.push_back(parse_flow(s, i, no));
                skip_ws(s, i);
                if (i >= s.size()) throw std::runtime_error("yaml: unterminated flow sequence (line " + std::to_string(no) + ")");
                if (s[i] == ',') { i++; continue; }
                if (s[i] == ']') { i++; break; }
                throw std::runtime_error("yaml: expected ',' or ']' (line " + std::to_string(no) + ")");
            }
            return n;
        }
        if (s[i] == '{') {
            i++; n.kind = Node::Map;
            skip_ws(s, i);
            if (i < s.size() && s[i] == '}') { i++; return n; }
            while (true) {
                skip_ws(s, i);
                std::string key;
                if (i < s.size() && (s[i] == '"' || s[i] == '\'')) key = parse_quoted(s, i, no);
                else { size_t b = i; while (i < s.size() && s[i] != ':' && s[i] != ',' && s[i] != '}') i++; key = trim(s.substr(b, i - b)); }
                skip_ws(s, i);
                if (i >= s.size() || s[i] != ':') throw std::runtime_error("yaml: expected ':' in flow map (line " + std::to_string(no) + ")");
                i++;
                n.map.emplace_back(key, parse_flow(s, i, no));
                skip_ws(s, i);
                if (i >= s.size()) throw std::runtime_error("yaml: unterminated flow map (line " + std::to_string(no) + ")");
                if (s[i] == ',') { i++; continue; }
                if (s[i] == '}') { i++; break; }
                throw std::runtime_error("yaml: expected ',' or '}' (line " + std::to_string(no) + ")");
            }
            return n;
        }
        if (s[i] == '"' || s[i] == '\'') { n.kind = Node::Scalar; n.scalar = parse_quoted(s, i, no); return n; }
        size_t b = i;
        while (i < s.size() && s[i] != ',' && s[i] != ']' && s[i] != '}') i++;
        std::string v = trim(s.substr(b, i - b));
        if (v.empty() || v == "~" || v == "null") return n;
        n.kind = Node::Scalar; n.scalar = v;
        return n;
    }

    static Node parse_inline_value(const std::string& text, int no) {
        std::string t = trim(text);
        Node n; n.line = no;
        if (t.empty() || t == "~" || t == "null") return n;
        if (t[0] == '[' || t[0] == '{') {
            size_t i = 0;
            n = parse_flow(t, i, no);
            skip_ws(t, i);
            if (i != t.size()) throw std::runtime_error("yaml: trailing characters after flow value (line " + std::to_string(no) + ")");
            return n;
        }
        if (t[0] == '"' || t[0] == '\'') {
            size_t i = 0;
            n.kind = Node::Scalar; n.scalar = parse_quoted(t, i, no);
            return n;
        }
        n.kind = Node::Scalar; n.scalar = t;
        return n;
    }

    // position of the ": " / trailing ':' that splits "key: value" (outside quotes and flow brackets), or npos
    static size_t find_key_colon(const std::string& t) {
        if (t.empty() || t[0] == '[' || t[0] == '{') return std::string::npos;
        char q = 0;
        for (size_t i = 0; i < t.size(); i++) {
            const char c = t[i];
            if (q) { if (c == q) q = 0; continue; }
            if ((c == '"' || c == '\'') && i == 0) { q = c; continue; }
            if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ' || t[i + 1] == '\t')) return i;
        }
        return std::string::npos;
    }

    static std::string unquote_key(const std::string& k, int no) {
        std::string t = trim(k);
        if (!t.empty() && (t[0] == '"' || t[0] == '\'')) { size_t i = 0; return parse_quoted(t, i, no); }
        return t;
    }

    // ---- block syntax --------------------------------------------------------------------------
    Node parse_block(int indent) {
        Node n;
        if (pos >= lines.size() || lines[pos].indent < indent) return n;
        const Line& first = lines[pos];
        n.line = first.no;
        const int ind = first.indent;
        if (first.text[0] == '-' && (first.text.size() == 1 || first.text[1] == ' ')) {
            n.kind = Node::Seq;
            while (pos < lines.size() && lines[pos].indent == ind && lines[pos].text[0] == '-' &&
                   (lines[pos].text.size() == 1 || lines[pos].text[1] == ' ')) {
                Line cur = lines[pos];
                std::string rest = cur.text.substr(1);
                size_t lead = 0;
                while (lead < rest.size() && rest[lead] == ' ') lead++;
                rest = rest.substr(lead);
                if (rest.empty()) {
                    pos++;
                    n.seq.push_back(parse_block(ind + 1));
                } else if (find_key_colon(rest) != std::string::npos || (rest[0] == '-' && (rest.size() == 1 || rest[1] == ' '))) {
                    // "- key: value": the item is a block whose first line starts at the column of `key`
                    lines[pos].indent = ind + 1 + (int)lead;
                    lines[pos].text = rest;
                    n.seq.push_back(parse_block(ind + 1 + (int)lead));
                } else {
                    pos++;
                    n.seq.push_back(parse_inline_value(rest, cur.no));
                }
            }
            if (pos < lines.size() && lines[pos].indent > ind) throw std::runtime_error("yaml: bad indentation (line " + std::to_string(lines[pos].no) + ")");
            return n;
        }
        const size_t colon0 = find_key_colon(first.text);
        if (colon0 == std::string::npos) {
            // a lone scalar / flow value document
            pos++;
            return parse_inline_value(first.text, first.no);
        }
        n.kind = Node::Map;
        while (pos < lines.size() && lines[pos].indent == ind) {
            const Line cur = lines[pos];
            if (cur.text[0] == '-' && (cur.text.size() == 1 || cur.text[1] == ' ')) break;
            const size_t colon = find_key_colon(cur.text);
            if (colon == std::string::npos) throw std::runtime_error("yaml: expected 'key: value' (line " + std::to_string(cur.no) + ")");
            const std::string key = unquote_key(cur.text.substr(0, colon), cur.no);
            const std::string val = trim(cur.text.substr(colon + 1));
            pos++;
            Node child;
            if (val.empty()) {
                // nested block; a sequence may sit at the same indent as its key
                if (pos < lines.size() && (lines[pos].indent > ind ||
                    (lines[pos].indent == ind && lines[pos].text[0] == '-' && (lines[pos].text.size() == 1 || lines[pos].text[1] == ' '))))
                    child = parse_block(lines[pos].indent);
                child.line = child.line ? child.line : cur.no;
            } else {
                child = parse_inline_value(val, cur.no);
            }
            n.map.emplace_back(key, std::move(child));
        }
        if (pos < lines.size() && lines[pos].indent > ind) throw std::runtime_error("yaml: bad indentation (line " + std::to_string(lines[pos].no) + ")");
        return n;
    }
};

}  // namespace detail

inline Node Load(const std::string& text) {
    detail::Parser p(text);
    if (p.lines.empty()) return Node();
    Node n = p.parse_block(p.lines[0].indent);
    if (p.pos != p.lines.size()) throw std::runtime_error("yaml: unexpected content (line " + std::to_string(p.lines[p.pos].no) + ")");
    return n;
}

}}  // namespace ngi::yaml

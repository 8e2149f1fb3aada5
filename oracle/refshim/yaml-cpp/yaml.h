// Stand-in for the yaml-cpp subset the reference's Scene::Load uses (rt.hpp:1519-2050): LoadFile, node["key"], node[i],
// size(), as<T>(), boolean test. Backed by this repository's own YAML-subset reader. See ../README.md.
#pragma once
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include "../../../nanogi_b200/host/yaml_lite.hpp"

namespace YAML {

class Exception : public std::runtime_error { public: using std::runtime_error::runtime_error; };

class Node {
public:
    Node() {}
    Node(std::shared_ptr<ngi::yaml::Node> root, const ngi::yaml::Node* n) : root_(root), n_(n) {}
    bool IsDefined() const { return n_ && n_->defined(); }
    explicit operator bool() const { return IsDefined(); }
    bool operator!() const { return !IsDefined(); }
    size_t size() const { return n_ ? n_->size() : 0; }
    Node operator[](const std::string& key) const { return n_ ? Node(root_, &(*n_)[key]) : Node(); }
    Node operator[](const char* key) const { return (*this)[std::string(key)]; }
    Node operator[](size_t i) const { return n_ ? Node(root_, &(*n_)[i]) : Node(); }
    Node operator[](int i) const { return (*this)[(size_t)i]; }
    template <class T> T as() const;
private:
    void need() const { if (!IsDefined()) throw Exception("bad conversion (node is not defined)"); }
    std::shared_ptr<ngi::yaml::Node> root_;
    const ngi::yaml::Node* n_ = nullptr;
};
template <> inline std::string Node::as<std::string>() const { need(); return n_->as_string(); }
template <> inline double Node::as<double>() const { need(); try { return n_->as_double(); } catch (const std::exception& e) { throw Exception(e.what()); } }
template <> inline int Node::as<int>() const { return (int)as<double>(); }
template <> inline long long Node::as<long long>() const { return (long long)as<double>(); }
template <> inline bool Node::as<bool>() const { need(); try { return n_->as_bool(); } catch (const std::exception& e) { throw Exception(e.what()); } }

inline Node LoadFile(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw Exception("bad file: " + path);
    std::stringstream ss; ss << in.rdbuf();
    try {
        auto root = std::make_shared<ngi::yaml::Node>(ngi::yaml::Load(ss.str()));
        return Node(root, root.get());
    } catch (const std::exception& e) { throw Exception(e.what()); }
}

}  // namespace YAML

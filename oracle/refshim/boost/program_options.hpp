// Stand-in for the boost::program_options subset the reference's Run uses (src/nanogi.cpp:2000-2048): options_description with
// add_options()("long,s", value<T>()->default_value(v)->required(), "help"), positional_options_description::add,
// command_line_parser(argc, argv).options().positional().run(), store, notify, variables_map (count, operator[], as<T>),
// error / required_option. Written for this repository; see the README.md of oracle/refshim.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

class error : public std::logic_error { public: explicit error(const std::string& w) : std::logic_error(w) {} };
class required_option : public error { public: explicit required_option(const std::string& name) : error("the option '--" + name + "' is required but missing") {} };

struct value_base {
    virtual ~value_base() {}
    virtual bool parse(const std::string& text) = 0;      // stores the parsed value
    virtual bool has_default() const = 0;
    virtual std::string default_text() const = 0;
    virtual std::shared_ptr<void> boxed(bool use_default) const = 0;
    bool is_required = false;
};
template <class T> struct typed_value : value_base {
    T parsed{}; T def{}; bool has_def = false;
    typed_value* default_value(const T& v) { def = v; has_def = true; return this; }
    typed_value* required() { is_required = true; return this; }
    bool parse(const std::string& text) override { return convert(text, parsed); }
    bool has_default() const override { return has_def; }
    std::string default_text() const override { std::ostringstream o; o << def; return o.str(); }
    std::shared_ptr<void> boxed(bool use_default) const override { return std::make_shared<T>(use_default ? def : parsed); }
    static bool convert(const std::string& s, std::string& out) { out = s; return true; }
    template <class U> static bool convert(const std::string& s, U& out) { std::istringstream i(s); i >> out; return !i.fail() && i.eof(); }
};
template <class T> inline typed_value<T>* value() { return new typed_value<T>(); }

struct option { std::string lname; char sname = 0; std::shared_ptr<value_base> val; std::string help; };

class options_description {
public:
    explicit options_description(const std::string& caption = "") : caption_(caption) {}
    class init {
    public:
        explicit init(options_description* o) : o_(o) {}
        init& operator()(const char* name, const char* help) { o_->add(name, nullptr, help); return *this; }
        init& operator()(const char* name, value_base* v, const char* help) { o_->add(name, v, help); return *this; }
    private:
        options_description* o_;
    };
    init add_options() { return init(this); }
    void add(const std::string& name, value_base* v, const std::string& help) {
        option op; const size_t c = name.find(',');
        op.lname = name.substr(0, c); if (c != std::string::npos && c + 1 < name.size()) op.sname = name[c + 1];
        op.val.reset(v); op.help = help;
        opts.push_back(op);
    }
    const option* find_long(const std::string& n) const { for (auto& o : opts) if (o.lname == n) return &o; return nullptr; }
    const option* find_short(char ch) const { for (auto& o : opts) if (o.sname && o.sname == ch) return &o; return nullptr; }
    std::vector<option> opts;
    std::string caption_;
};
inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
    os << d.caption_ << ":\n";
    for (auto& o : d.opts) {
        std::string left = "  ";
        if (o.sname) left += std::string("-") + o.sname + " [ --" + o.lname + " ]"; else left += "--" + o.lname;
        if (o.val) { left += " arg"; if (o.val->has_default()) left += " (=" + o.val->default_text() + ")"; }
        os << left << "  " << o.help << "\n";
    }
    return os;
}

class positional_options_description {
public:
    positional_options_description& add(const char* name, int count) { for (int i = 0; i < count; i++) names.push_back(name); return *this; }
    std::vector<std::string> names;
};

class variable_value {
public:
    variable_value() {}
    explicit variable_value(std::shared_ptr<void> v) : v_(v) {}
    template <class T> const T& as() const { if (!v_) throw error("boost::bad_any_cast: failed conversion using boost::any_cast"); return *static_cast<const T*>(v_.get()); }
    bool empty() const { return !v_; }
private:
    std::shared_ptr<void> v_;
};
class variables_map {
public:
    size_t count(const std::string& k) const { return m_.count(k); }
    const variable_value& operator[](const std::string& k) const { static const variable_value none; auto it = m_.find(k); return it == m_.end() ? none : it->second; }
    std::map<std::string, variable_value> m_;
    std::vector<std::string> required_;
};

struct parsed_options { std::vector<std::pair<const option*, std::string>> items; const options_description* desc = nullptr; };

class command_line_parser {
public:
    command_line_parser(int argc, const char* const* argv) { for (int i = 1; i < argc; i++) args_.push_back(argv[i]); }
    command_line_parser(int argc, char** argv) { for (int i = 1; i < argc; i++) args_.push_back(argv[i]); }
    command_line_parser& options(const options_description& d) { desc_ = &d; return *this; }
    command_line_parser& positional(const positional_options_description& p) { pos_ = &p; return *this; }
    parsed_options run() const {
        parsed_options out; out.desc = desc_;
        size_t npos = 0;
        auto is_number = [](const std::string& s) { char* e = nullptr; std::strtod(s.c_str(), &e); return !s.empty() && e && *e == 0; };
        for (size_t i = 0; i < args_.size(); i++) {
            const std::string& a = args_[i];
            const option* op = nullptr; std::string val; bool has_val = false;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                std::string name = a.substr(2); const size_t eq = name.find('=');
                if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
                op = desc_->find_long(name);
                if (!op) throw error("unrecognised option '--" + name + "'");
            } else if (a.size() >= 2 && a[0] == '-' && !is_number(a)) {
                op = desc_->find_short(a[1]);
                if (!op) throw error("unrecognised option '" + a + "'");
                if (a.size() > 2) { val = a.substr(2); has_val = true; }
            } else {
                if (!pos_ || npos >= pos_->names.size()) throw error("too many positional options have been specified on the command line");
                op = desc_->find_long(pos_->names[npos++]); val = a; has_val = true;
                if (!op) throw error("unknown positional option");
            }
            if (op->val && !has_val) {
                if (i + 1 >= args_.size()) throw error("the required argument for option '--" + op->lname + "' is missing");
                val = args_[++i];
            }
            out.items.emplace_back(op, val);
        }
        return out;
    }
private:
    std::vector<std::string> args_;
    const options_description* desc_ = nullptr;
    const positional_options_description* pos_ = nullptr;
};

inline void store(const parsed_options& p, variables_map& vm) {
    for (auto& it : p.items) {
        const option* op = it.first;
        if (vm.m_.count(op->lname)) throw error("option '--" + op->lname + "' cannot be specified more than once");
        if (op->val) {
            if (!op->val->parse(it.second)) throw error("the argument ('" + it.second + "') for option '--" + op->lname + "' is invalid");
            vm.m_[op->lname] = variable_value(op->val->boxed(false));
        } else vm.m_[op->lname] = variable_value(std::make_shared<bool>(true));
    }
    for (auto& o : p.desc->opts) {
        if (!vm.m_.count(o.lname) && o.val && o.val->has_default()) vm.m_[o.lname] = variable_value(o.val->boxed(true));
        if (o.val && o.val->is_required) vm.required_.push_back(o.lname);
    }
}
inline void notify(variables_map& vm) { for (auto& r : vm.required_) if (!vm.count(r)) throw required_option(r); }

}}

// Minimal stand-in for boost/program_options.hpp so that the reference's
// src/parser.cpp compiles unmodified in the oracle build (test infrastructure only).
// Implements only what parser.cpp:19-515 uses: options_description, value<T>()
// with bound storage and default_value, variables_map (count / operator[] / as<T>),
// parse_command_line (-i/-h/-v and long forms), parse_config_file (key=value lines,
// '#' comments), store (first explicit value wins, defaults filled in), notify.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

inline std::string po_trim(std::string const& s) {
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

template <class T> inline T po_cast(std::string const& s) {
    std::istringstream is(s);
    T v;
    is >> v;
    if (is.fail()) throw std::runtime_error("program_options stub: bad value '" + s + "'");
    return v;
}
template <> inline std::string po_cast<std::string>(std::string const& s) { return s; }
template <> inline bool po_cast<bool>(std::string const& s0) {
    std::string s;
    for (char c: s0) s.push_back(static_cast<char>(std::tolower(c)));
    if (s == "true" || s == "1" || s == "yes" || s == "on" || s == "") return true;
    if (s == "false" || s == "0" || s == "no" || s == "off") return false;
    throw std::runtime_error("program_options stub: bad bool '" + s0 + "'");
}

class value_semantic {
  public:
    virtual ~value_semantic() {}
    virtual void assign(std::string const& raw) = 0;
    virtual bool has_default() const = 0;
    virtual std::string default_raw() const = 0;
};

template <class T> class typed_value: public value_semantic {
  public:
    explicit typed_value(T* store): m_store {store} {}
    typed_value* default_value(T const& v) {
        m_has_default = true;
        std::ostringstream os;
        os.precision(17);
        os << std::boolalpha << v;
        m_default_raw = os.str();
        return this;
    }
    void assign(std::string const& raw) override {
        T v = po_cast<T>(raw);
        if (m_store != nullptr) *m_store = v;
    }
    bool has_default() const override { return m_has_default; }
    std::string default_raw() const override { return m_default_raw; }

  private:
    T* m_store;
    bool m_has_default {false};
    std::string m_default_raw {};
};
template <> inline typed_value<std::string>* typed_value<std::string>::default_value(std::string const& v) {
    m_has_default = true;
    m_default_raw = v;
    return this;
}

template <class T> inline typed_value<T>* value() { return new typed_value<T>(nullptr); }
template <class T> inline typed_value<T>* value(T* store) { return new typed_value<T>(store); }

struct option_entry {
    std::string long_name;
    char short_name {0};
    std::shared_ptr<value_semantic> sem; // null => flag
    std::string help;
};

class options_description;
class options_adder {
  public:
    explicit options_adder(options_description& d): m_d {d} {}
    options_adder& operator()(const char* name, const char* help);
    options_adder& operator()(const char* name, value_semantic* sem, const char* help);

  private:
    options_description& m_d;
};

class options_description {
  public:
    explicit options_description(std::string const& title = ""): m_title {title} {}
    options_adder add_options() { return options_adder {*this}; }
    options_description& add(options_description const& o) {
        for (auto const& e: o.m_entries) m_entries.push_back(e);
        return *this;
    }
    void push(const char* name, value_semantic* sem, const char* help) {
        option_entry e;
        std::string n {name};
        auto comma = n.find(',');
        if (comma != std::string::npos) {
            e.long_name = n.substr(0, comma);
            e.short_name = n[comma + 1];
        }
        else {
            e.long_name = n;
        }
        e.sem.reset(sem);
        e.help = help;
        m_entries.push_back(e);
    }
    option_entry const* find(std::string const& name) const {
        for (auto const& e: m_entries)
            if (e.long_name == name) return &e;
        return nullptr;
    }
    option_entry const* find_short(char c) const {
        for (auto const& e: m_entries)
            if (e.short_name == c) return &e;
        return nullptr;
    }
    std::vector<option_entry> m_entries;
    std::string m_title;
};
inline options_adder& options_adder::operator()(const char* name, const char* help) {
    m_d.push(name, nullptr, help);
    return *this;
}
inline options_adder& options_adder::operator()(const char* name, value_semantic* sem, const char* help) {
    m_d.push(name, sem, help);
    return *this;
}
inline std::ostream& operator<<(std::ostream& os, options_description const& d) {
    os << d.m_title << ":\n";
    for (auto const& e: d.m_entries) os << "  --" << e.long_name << "  " << e.help << "\n";
    return os;
}

struct parsed_options {
    options_description const* desc;
    std::vector<std::pair<std::string, std::string>> values;
};

class variable_value {
  public:
    variable_value() {}
    variable_value(std::string raw, bool defaulted): m_raw {raw}, m_defaulted {defaulted} {}
    template <class T> T as() const { return po_cast<T>(m_raw); }
    bool defaulted() const { return m_defaulted; }
    std::string m_raw;
    bool m_defaulted {false};
    std::shared_ptr<value_semantic> m_sem;
};

class variables_map {
  public:
    size_t count(std::string const& name) const { return m_map.count(name); }
    variable_value const& operator[](std::string const& name) const {
        static const variable_value empty {};
        auto it = m_map.find(name);
        return it == m_map.end() ? empty : it->second;
    }
    std::map<std::string, variable_value> m_map;
};

inline parsed_options parse_command_line(int argc, char* argv[], options_description const& desc) {
    parsed_options p {&desc, {}};
    for (int i {1}; i < argc; i++) {
        std::string a {argv[i]};
        option_entry const* e {nullptr};
        std::string val;
        bool have_val {false};
        if (a.rfind("--", 0) == 0) {
            std::string n {a.substr(2)};
            auto eq = n.find('=');
            if (eq != std::string::npos) {
                val = n.substr(eq + 1);
                n = n.substr(0, eq);
                have_val = true;
            }
            e = desc.find(n);
        }
        else if (a.size() >= 2 && a[0] == '-') {
            e = desc.find_short(a[1]);
            if (a.size() > 2) {
                val = a.substr(2);
                have_val = true;
            }
        }
        if (e == nullptr) throw std::runtime_error("program_options stub: unknown option " + a);
        if (e->sem) {
            if (!have_val) {
                if (i + 1 >= argc) throw std::runtime_error("program_options stub: missing value for " + a);
                val = argv[++i];
            }
        }
        p.values.push_back({e->long_name, val});
    }
    return p;
}

inline parsed_options parse_config_file(std::istream& is, options_description const& desc) {
    parsed_options p {&desc, {}};
    std::string line;
    while (std::getline(is, line)) {
        auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = po_trim(line);
        if (line.empty()) continue;
        auto eq = line.find('=');
        if (eq == std::string::npos) throw std::runtime_error("program_options stub: bad line '" + line + "'");
        std::string key {po_trim(line.substr(0, eq))};
        std::string val {po_trim(line.substr(eq + 1))};
        if (desc.find(key) == nullptr) throw std::runtime_error("program_options stub: unrecognised option '" + key + "'");
        p.values.push_back({key, val});
    }
    return p;
}

inline void store(parsed_options const& p, variables_map& vm) {
    for (auto const& kv: p.values) {
        auto it = vm.m_map.find(kv.first);
        if (it == vm.m_map.end() || it->second.defaulted()) {
            variable_value v {kv.second, false};
            v.m_sem = p.desc->find(kv.first)->sem;
            vm.m_map[kv.first] = v;
        }
    }
    for (auto const& e: p.desc->m_entries) {
        if (e.sem && e.sem->has_default() && vm.m_map.count(e.long_name) == 0) {
            variable_value v {e.sem->default_raw(), true};
            v.m_sem = e.sem;
            vm.m_map[e.long_name] = v;
        }
    }
}

inline void notify(variables_map& vm) {
    for (auto& kv: vm.m_map) {
        if (kv.second.m_sem) kv.second.m_sem->assign(kv.second.m_raw);
    }
}

} } // namespace boost::program_options

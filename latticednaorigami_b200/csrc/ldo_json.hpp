// ldo_json.hpp — minimal JSON reader for the host side (system / moveset / order-parameter / bias
// files). The reference vendors jsoncpp 1.7.4 (src/jsoncpp.cpp, out of scope); only the subset those
// files use is needed: objects, arrays, strings, numbers, booleans, null, // and /* */ comments.
#pragma once

#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ldohost {

class Json {
  public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Type type {Null};
    bool b {false};
    double num {0};
    std::string str {};
    std::vector<Json> arr {};
    std::vector<std::pair<std::string, Json>> obj {};

    bool is_null() const { return type == Null; }
    size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
    bool has(std::string const& key) const {
        if (type != Object) return false;
        for (auto const& kv: obj)
            if (kv.first == key) return true;
        return false;
    }
    // Missing members read as null (jsoncpp semantics for const access)
    Json const& operator[](std::string const& key) const {
        static const Json null_value {};
        if (type != Object) return null_value;
        for (auto const& kv: obj)
            if (kv.first == key) return kv.second;
        return null_value;
    }
    Json const& operator[](size_t i) const {
        static const Json null_value {};
        if (type != Array || i >= arr.size()) return null_value;
        return arr[i];
    }
    int as_int() const {
        if (type == Number) return static_cast<int>(num);
        if (type == Bool) return b ? 1 : 0;
        if (type == Null) return 0;
        throw std::runtime_error("JSON value is not convertible to int");
    }
    double as_double() const {
        if (type == Number) return num;
        if (type == Bool) return b ? 1 : 0;
        if (type == Null) return 0;
        throw std::runtime_error("JSON value is not convertible to double");
    }
    bool as_bool() const {
        if (type == Bool) return b;
        if (type == Number) return num != 0;
        if (type == Null) return false;
        throw std::runtime_error("JSON value is not convertible to bool");
    }
    std::string as_string() const {
        if (type == String) return str;
        if (type == Null) return "";
        if (type == Bool) return b ? "true" : "false";
        if (type == Number) {
            char buf[64];
            snprintf(buf, sizeof(buf), "%.17g", num);
            return buf;
        }
        throw std::runtime_error("JSON value is not convertible to string");
    }

    static Json parse(std::string const& text) {
        Parser p {text, 0};
        Json v = p.value();
        p.skip();
        if (p.pos != text.size()) p.error("trailing characters");
        return v;
    }

  private:
    struct Parser {
        std::string const& t;
        size_t pos;
        [[noreturn]] void error(std::string const& what) {
            throw std::runtime_error("JSON parse error at offset " + std::to_string(pos) + ": " + what);
        }
        void skip() {
            for (;;) {
                while (pos < t.size() && (t[pos] == ' ' || t[pos] == '\t' || t[pos] == '\n' || t[pos] == '\r')) pos++;
                if (pos + 1 < t.size() && t[pos] == '/' && t[pos + 1] == '/') {
                    while (pos < t.size() && t[pos] != '\n') pos++;
                }
                else if (pos + 1 < t.size() && t[pos] == '/' && t[pos + 1] == '*') {
                    pos += 2;
                    while (pos + 1 < t.size() && !(t[pos] == '*' && t[pos + 1] == '/')) pos++;
                    pos += 2;
                }
                else {
                    return;
                }
            }
        }
        Json value() {
            skip();
            if (pos >= t.size()) error("unexpected end");
            char c = t[pos];
            if (c == '{') return object();
            if (c == '[') return array();
            if (c == '"') {
                Json v;
                v.type = String;
                v.str = string();
                return v;
            }
            if (t.compare(pos, 4, "true") == 0) {
                pos += 4;
                Json v;
                v.type = Bool;
                v.b = true;
                return v;
            }
            if (t.compare(pos, 5, "false") == 0) {
                pos += 5;
                Json v;
                v.type = Bool;
                v.b = false;
                return v;
            }
            if (t.compare(pos, 4, "null") == 0) {
                pos += 4;
                return Json {};
            }
            return number();
        }
        Json number() {
            const char* start = t.c_str() + pos;
            char* end = nullptr;
            double d = std::strtod(start, &end);
            if (end == start) error("bad number");
            pos += static_cast<size_t>(end - start);
            Json v;
            v.type = Number;
            v.num = d;
            return v;
        }
        std::string string() {
            std::string out;
            pos++; // opening quote
            while (pos < t.size() && t[pos] != '"') {
                char c = t[pos++];
                if (c == '\\') {
                    if (pos >= t.size()) error("bad escape");
                    char e = t[pos++];
                    switch (e) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u': {
                        if (pos + 4 > t.size()) error("bad unicode escape");
                        unsigned code = static_cast<unsigned>(std::strtoul(t.substr(pos, 4).c_str(), nullptr, 16));
                        pos += 4;
                        if (code < 0x80) {
                            out.push_back(static_cast<char>(code));
                        }
                        else if (code < 0x800) {
                            out.push_back(static_cast<char>(0xC0 | (code >> 6)));
                            out.push_back(static_cast<char>(0x80 | (code & 0x3F)));
                        }
                        else {
                            out.push_back(static_cast<char>(0xE0 | (code >> 12)));
                            out.push_back(static_cast<char>(0x80 | ((code >> 6) & 0x3F)));
                            out.push_back(static_cast<char>(0x80 | (code & 0x3F)));
                        }
                        break;
                    }
                    default: out.push_back(e);
                    }
                }
                else {
                    out.push_back(c);
                }
            }
            if (pos >= t.size()) error("unterminated string");
            pos++; // closing quote
            return out;
        }
        Json array() {
            Json v;
            v.type = Array;
            pos++;
            skip();
            if (pos < t.size() && t[pos] == ']') {
                pos++;
                return v;
            }
            for (;;) {
                v.arr.push_back(value());
                skip();
                if (pos >= t.size()) error("unterminated array");
                if (t[pos] == ',') {
                    pos++;
                    continue;
                }
                if (t[pos] == ']') {
                    pos++;
                    return v;
                }
                error("expected , or ]");
            }
        }
        Json object() {
            Json v;
            v.type = Object;
            pos++;
            skip();
            if (pos < t.size() && t[pos] == '}') {
                pos++;
                return v;
            }
            for (;;) {
                skip();
                if (pos >= t.size() || t[pos] != '"') error("expected member name");
                std::string key = string();
                skip();
                if (pos >= t.size() || t[pos] != ':') error("expected :");
                pos++;
                v.obj.push_back({key, value()});
                skip();
                if (pos >= t.size()) error("unterminated object");
                if (t[pos] == ',') {
                    pos++;
                    continue;
                }
                if (t[pos] == '}') {
                    pos++;
                    return v;
                }
                error("expected , or }");
            }
        }
    };
};

} // namespace ldohost

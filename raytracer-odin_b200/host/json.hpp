// json.hpp — minimal JSON DOM for the glTF loader (the reference parses through vendor:cgltf,
// input.odin:28; only what read_gltf touches is needed: objects, arrays, numbers, strings, bools).
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace orh {

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj; // insertion order kept

    const Json* find(const char* key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return find(key) != nullptr; }
    const Json& at(const char* key) const {
        static const Json null_json;
        const Json* j = find(key);
        return j ? *j : null_json;
    }
    const Json& at(size_t i) const {
        static const Json null_json;
        return (kind == Array && i < arr.size()) ? arr[i] : null_json;
    }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    bool is_null() const { return kind == Null; }
    double number(double dflt) const { return kind == Number ? num : dflt; }
    int64_t integer(int64_t dflt) const { return kind == Number ? (int64_t)num : dflt; }
};

class JsonParser {
public:
    // Returns false and sets `err` on malformed input.
    static bool parse(const std::string& text, Json* out, std::string* err) {
        JsonParser p(text);
        p.skip_ws();
        if (!p.value(out, 0)) { *err = p.err_ + " at byte " + std::to_string(p.i_); return false; }
        p.skip_ws();
        if (p.i_ != p.s_.size()) { *err = "trailing characters at byte " + std::to_string(p.i_); return false; }
        return true;
    }

private:
    explicit JsonParser(const std::string& s) : s_(s) {}
    const std::string& s_;
    size_t i_ = 0;
    std::string err_;

    bool fail(const char* m) { err_ = m; return false; }
    void skip_ws() {
        while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) i_++;
    }
    bool literal(const char* lit) {
        const size_t n = std::strlen(lit);
        if (s_.compare(i_, n, lit) != 0) return fail("bad literal");
        i_ += n;
        return true;
    }
    static void utf8(uint32_t cp, std::string* o) {
        if (cp < 0x80) o->push_back((char)cp);
        else if (cp < 0x800) { o->push_back((char)(0xC0 | (cp >> 6))); o->push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            o->push_back((char)(0xE0 | (cp >> 12))); o->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            o->push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            o->push_back((char)(0xF0 | (cp >> 18))); o->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            o->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o->push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    bool hex4(uint32_t* v) {
        if (i_ + 4 > s_.size()) return fail("short \\u escape");
        uint32_t x = 0;
        for (int k = 0; k < 4; k++) {
            const char c = s_[i_++];
            x <<= 4;
            if (c >= '0' && c <= '9') x |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') x |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') x |= (uint32_t)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        *v = x;
        return true;
    }
    bool string(std::string* o) {
        if (s_[i_] != '"') return fail("expected string");
        i_++;
        while (i_ < s_.size()) {
            const char c = s_[i_++];
            if (c == '"') return true;
            if (c != '\\') { o->push_back(c); continue; }
            if (i_ >= s_.size()) break;
            const char e = s_[i_++];
            switch (e) {
            case '"': o->push_back('"'); break;
            case '\\': o->push_back('\\'); break;
            case '/': o->push_back('/'); break;
            case 'b': o->push_back('\b'); break;
            case 'f': o->push_back('\f'); break;
            case 'n': o->push_back('\n'); break;
            case 'r': o->push_back('\r'); break;
            case 't': o->push_back('\t'); break;
            case 'u': {
                uint32_t cp;
                if (!hex4(&cp)) return false;
                if (cp >= 0xD800 && cp < 0xDC00 && i_ + 1 < s_.size() && s_[i_] == '\\' && s_[i_ + 1] == 'u') {
                    i_ += 2;
                    uint32_t lo;
                    if (!hex4(&lo)) return false;
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                utf8(cp, o);
                break;
            }
            default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool value(Json* out, int depth) {
        if (depth > 256) return fail("nesting too deep");
        if (i_ >= s_.size()) return fail("unexpected end");
        const char c = s_[i_];
        if (c == '{') {
            out->kind = Json::Object;
            i_++;
            skip_ws();
            if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
            for (;;) {
                skip_ws();
                std::string key;
                if (i_ >= s_.size() || !string(&key)) return err_.empty() ? fail("expected key") : false;
                skip_ws();
                if (i_ >= s_.size() || s_[i_] != ':') return fail("expected ':'");
                i_++;
                skip_ws();
                out->obj.emplace_back(std::move(key), Json());
                if (!value(&out->obj.back().second, depth + 1)) return false;
                skip_ws();
                if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
                if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            out->kind = Json::Array;
            i_++;
            skip_ws();
            if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
            for (;;) {
                skip_ws();
                out->arr.emplace_back();
                if (!value(&out->arr.back(), depth + 1)) return false;
                skip_ws();
                if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
                if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { out->kind = Json::String; return string(&out->str); }
        if (c == 't') { out->kind = Json::Bool; out->b = true; return literal("true"); }
        if (c == 'f') { out->kind = Json::Bool; out->b = false; return literal("false"); }
        if (c == 'n') { out->kind = Json::Null; return literal("null"); }
        if (c == '-' || (c >= '0' && c <= '9')) {
            const char* begin = s_.c_str() + i_;
            char* end = nullptr;
            out->num = std::strtod(begin, &end);
            if (end == begin) return fail("bad number");
            out->kind = Json::Number;
            i_ += (size_t)(end - begin);
            return true;
        }
        return fail("unexpected character");
    }
};

} // namespace orh

// Minimal JSON document model for the glTF importer (kjb_asset.cpp).  RFC 8259 values; numbers kept as double
// (serde_json, which the reference's `gltf` crate parses with, does the same before narrowing to the field type).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace kjb_asset_detail {

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order kept; lookups are linear (glTF objects are small)

    bool is(Kind k) const { return kind == k; }
    const Json* get(const char* key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Json& operator[](size_t i) const { static const Json null_json; return (kind == Array && i < arr.size()) ? arr[i] : null_json; }
    size_t size() const { return kind == Array ? arr.size() : 0; }
    double number_or(double d) const { return kind == Number ? num : d; }
    // typed field readers with glTF defaults
    double f(const char* key, double d) const { const Json* j = get(key); return j && j->kind == Number ? j->num : d; }
    // integer field: NaN / out-of-range doubles would make the cast undefined; they saturate (callers range-check the result)
    int64_t i(const char* key, int64_t d) const {
        const Json* j = get(key); if (!j || j->kind != Number) return d;
        const double v = j->num;
        if (!(v == v)) return -1;
        if (v >= 9007199254740992.0) return INT64_C(9007199254740992);
        if (v <= -9007199254740992.0) return -INT64_C(9007199254740992);
        return (int64_t)v;
    }
    // a size / offset / count: non-negative, integral, below 2^53 — anything else is reported through `ok`
    size_t size(const char* key, size_t d, bool& ok) const {
        const Json* j = get(key); if (!j) return d;
        if (j->kind != Number) { ok = false; return 0; }
        const double v = j->num;
        if (!(v >= 0.0) || v >= 9007199254740992.0 || v != (double)(uint64_t)v) { ok = false; return 0; }
        return (size_t)v;
    }
    bool has(const char* key) const { return get(key) != nullptr; }
    std::string s(const char* key, const char* d = "") const { const Json* j = get(key); return j && j->kind == String ? j->str : std::string(d); }
};

class JsonParser {
public:
    JsonParser(const char* p, size_t n) : p_(p), e_(p + n) {}
    bool parse(Json& out, std::string& err) {
        // UTF-8 byte order mark is tolerated (GLB chunks never carry one, exported .gltf files sometimes do)
        if (e_ - p_ >= 3 && (unsigned char)p_[0] == 0xEF && (unsigned char)p_[1] == 0xBB && (unsigned char)p_[2] == 0xBF) p_ += 3;
        if (!value(out, 0)) { err = err_.empty() ? "malformed JSON" : err_; return false; }
        ws();
        if (p_ != e_) { err = "trailing characters after JSON document"; return false; }
        return true;
    }

private:
    const char *p_, *e_;
    std::string err_;
    void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool fail(const char* m) { if (err_.empty()) err_ = m; return false; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(e_ - p_) < n || memcmp(p_, s, n) != 0) return false; p_ += n; return true; }

    static void utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o += char(c);
        else if (c < 0x800) { o += char(0xC0 | (c >> 6)); o += char(0x80 | (c & 0x3F)); }
        else if (c < 0x10000) { o += char(0xE0 | (c >> 12)); o += char(0x80 | ((c >> 6) & 0x3F)); o += char(0x80 | (c & 0x3F)); }
        else { o += char(0xF0 | (c >> 18)); o += char(0x80 | ((c >> 12) & 0x3F)); o += char(0x80 | ((c >> 6) & 0x3F)); o += char(0x80 | (c & 0x3F)); }
    }
    bool hex4(uint32_t& v) {
        if (e_ - p_ < 4) return false;
        v = 0;
        for (int k = 0; k < 4; ++k) {
            const char c = *p_++; v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10; else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else return false;
        }
        return true;
    }
    bool string(std::string& o) {
        if (p_ >= e_ || *p_ != '"') return fail("expected string");
        ++p_;
        while (p_ < e_) {
            const char c = *p_++;
            if (c == '"') return true;
            if ((unsigned char)c < 0x20) return fail("control character in string");
            if (c != '\\') { o += c; continue; }
            if (p_ >= e_) break;
            const char x = *p_++;
            switch (x) {
                case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
                case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break; case 'r': o += '\r'; break; case 't': o += '\t'; break;
                case 'u': {
                    uint32_t u; if (!hex4(u)) return fail("bad \\u escape");
                    if (u >= 0xD800 && u < 0xDC00 && e_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2; uint32_t lo; if (!hex4(lo)) return fail("bad \\u escape");
                        if (lo >= 0xDC00 && lo < 0xE000) u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00); else return fail("unpaired surrogate");
                    }
                    utf8(o, u); break;
                }
                default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool number(Json& out) {
        const char* s = p_;
        if (p_ < e_ && *p_ == '-') ++p_;
        if (p_ >= e_ || !(*p_ >= '0' && *p_ <= '9')) return fail("bad number");
        while (p_ < e_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) ++p_;
        std::string tmp(s, p_);
        char* end = nullptr;
        out.kind = Json::Number; out.num = strtod(tmp.c_str(), &end);
        if (end != tmp.c_str() + tmp.size()) return fail("bad number");
        return true;
    }
    bool value(Json& out, int depth) {
        if (depth > 256) return fail("JSON nested too deeply");
        ws();
        if (p_ >= e_) return fail("unexpected end of JSON");
        const char c = *p_;
        if (c == '{') {
            ++p_; out.kind = Json::Object; ws();
            if (p_ < e_ && *p_ == '}') { ++p_; return true; }
            for (;;) {
                ws(); std::string k; if (!string(k)) return false;
                ws(); if (p_ >= e_ || *p_ != ':') return fail("expected ':'"); ++p_;
                out.obj.emplace_back(std::move(k), Json());
                if (!value(out.obj.back().second, depth + 1)) return false;
                ws(); if (p_ >= e_) return fail("unterminated object");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_; out.kind = Json::Array; ws();
            if (p_ < e_ && *p_ == ']') { ++p_; return true; }
            for (;;) {
                out.arr.emplace_back();
                if (!value(out.arr.back(), depth + 1)) return false;
                ws(); if (p_ >= e_) return fail("unterminated array");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { out.kind = Json::String; return string(out.str); }
        if (c == 't') { if (!lit("true")) return fail("bad literal"); out.kind = Json::Bool; out.b = true; return true; }
        if (c == 'f') { if (!lit("false")) return fail("bad literal"); out.kind = Json::Bool; out.b = false; return true; }
        if (c == 'n') { if (!lit("null")) return fail("bad literal"); out.kind = Json::Null; return true; }
        return number(out);
    }
};

}  // namespace kjb_asset_detail

// json_min.hpp — a small recursive-descent JSON reader, enough for the ujson-written metadata
// files of the reference (Table.scala:27-35, Column.scala:21-29, Segment.scala:41-45).
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace imm3 {

struct JsonValue {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    double num = 0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;  // insertion order kept

    const JsonValue* get(const char* key) const {
        if (type != Obj) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class JsonParser {
  public:
    explicit JsonParser(const std::string& s) : p_(s.c_str()), end_(s.c_str() + s.size()) {}
    bool parse(JsonValue* out) {
        if (!value(out)) return false;
        ws();
        return p_ == end_;
    }

  private:
    const char* p_;
    const char* end_;
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    bool lit(const char* s) {
        size_t n = std::strlen(s);
        if ((size_t)(end_ - p_) < n || std::strncmp(p_, s, n)) return false;
        p_ += n;
        return true;
    }
    bool string(std::string* out) {
        if (p_ >= end_ || *p_ != '"') return false;
        ++p_;
        out->clear();
        while (p_ < end_ && *p_ != '"') {
            char c = *p_++;
            if (c == '\\') {
                if (p_ >= end_) return false;
                c = *p_++;
                switch (c) {
                    case 'n': c = '\n'; break;
                    case 't': c = '\t'; break;
                    case 'r': c = '\r'; break;
                    case 'b': c = '\b'; break;
                    case 'f': c = '\f'; break;
                    case 'u': {
                        if (end_ - p_ < 4) return false;
                        unsigned x = 0;
                        for (int i = 0; i < 4; i++) {
                            char h = p_[i];
                            unsigned d = (h >= '0' && h <= '9') ? h - '0' : (h | 32) >= 'a' && (h | 32) <= 'f' ? (h | 32) - 'a' + 10 : 99;
                            if (d == 99) return false;
                            x = x * 16 + d;
                        }
                        p_ += 4;
                        // UTF-8 encode the BMP code point (identifiers are ASCII in practice)
                        if (x < 0x80) { out->push_back((char)x); }
                        else if (x < 0x800) { out->push_back((char)(0xC0 | (x >> 6))); out->push_back((char)(0x80 | (x & 0x3F))); }
                        else { out->push_back((char)(0xE0 | (x >> 12))); out->push_back((char)(0x80 | ((x >> 6) & 0x3F))); out->push_back((char)(0x80 | (x & 0x3F))); }
                        continue;
                    }
                    default: break;  // \" \\ \/
                }
            }
            out->push_back(c);
        }
        if (p_ >= end_) return false;
        ++p_;
        return true;
    }
    bool value(JsonValue* out) {
        ws();
        if (p_ >= end_) return false;
        char c = *p_;
        if (c == '{') {
            ++p_;
            out->type = JsonValue::Obj;
            ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!string(&k)) return false;
                ws();
                if (p_ >= end_ || *p_ != ':') return false;
                ++p_;
                out->obj.emplace_back(std::move(k), JsonValue());
                if (!value(&out->obj.back().second)) return false;
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return true; }
                return false;
            }
        }
        if (c == '[') {
            ++p_;
            out->type = JsonValue::Arr;
            ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return true; }
            for (;;) {
                out->arr.emplace_back();
                if (!value(&out->arr.back())) return false;
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return true; }
                return false;
            }
        }
        if (c == '"') { out->type = JsonValue::Str; return string(&out->str); }
        if (lit("true")) { out->type = JsonValue::Bool; out->num = 1; return true; }
        if (lit("false")) { out->type = JsonValue::Bool; out->num = 0; return true; }
        if (lit("null")) { out->type = JsonValue::Null; return true; }
        char* e = nullptr;
        out->num = std::strtod(p_, &e);
        if (e == p_) return false;
        out->type = JsonValue::Num;
        p_ = e;
        return true;
    }
};

inline std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\t': o += "\\t"; break;
            case '\r': o += "\\r"; break;
            default: o.push_back(c);
        }
    }
    return o;
}

}  // namespace imm3

// Minimal JSON DOM (objects keep insertion order: ASTUtils.toBinaryClauseFromFilterJsonNode folds the members of an
// n-ary filter node in document order, core/src/main/scala/com/cardinal/utils/ast/ASTUtils.scala:379-404).
#pragma once
#include <cmath>
#include <cstdlib>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "lk_common.h"

namespace lk {

struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  int64_t i64 = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  bool is_null() const { return kind == Null; }
  bool is_str() const { return kind == Str; }
  bool is_obj() const { return kind == Obj; }
  bool is_arr() const { return kind == Arr; }
  const Json* get(const std::string& k) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  // Jackson JsonNode.textValue(): the string for textual nodes, otherwise "null"
  const std::string* text() const { return kind == Str ? &str : nullptr; }
  int64_t as_i64(int64_t dflt = 0) const { return kind == Num ? (is_int ? i64 : (int64_t)num) : dflt; }
  bool as_bool(bool dflt = false) const { return kind == Bool ? b : dflt; }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string& s) : s_(s) {}
  Json parse() {
    Json v = value();
    ws();
    LK_CHECK(p_ == s_.size(), LK_ERR_INVALID, "json: trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  void ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) p_++;
  }
  char peek() {
    ws();
    LK_CHECK(p_ < s_.size(), LK_ERR_INVALID, "json: unexpected end");
    return s_[p_];
  }
  void expect(char c) {
    LK_CHECK(peek() == c, LK_ERR_INVALID, strf("json: expected '%c' at %zu", c, p_));
    p_++;
  }
  static void utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
  }
  unsigned hex4() {
    LK_CHECK(p_ + 4 <= s_.size(), LK_ERR_INVALID, "json: bad \\u escape");
    unsigned v = 0;
    for (int i = 0; i < 4; i++) {
      char c = s_[p_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail(LK_ERR_INVALID, "json: bad \\u escape");
    }
    return v;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (true) {
      LK_CHECK(p_ < s_.size(), LK_ERR_INVALID, "json: unterminated string");
      char c = s_[p_++];
      if (c == '"') break;
      if (c != '\\') { out += c; continue; }
      LK_CHECK(p_ < s_.size(), LK_ERR_INVALID, "json: unterminated escape");
      char e = s_[p_++];
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          unsigned cp = hex4();
          if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
            p_ += 2;
            unsigned lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(out, cp);
          break;
        }
        default: fail(LK_ERR_INVALID, "json: bad escape");
      }
    }
    return out;
  }
  Json value() {
    char c = peek();
    Json v;
    if (c == '{') {
      p_++;
      v.kind = Json::Obj;
      if (peek() == '}') { p_++; return v; }
      while (true) {
        std::string k = string();
        expect(':');
        v.obj.emplace_back(std::move(k), value());
        char d = peek();
        p_++;
        if (d == '}') break;
        LK_CHECK(d == ',', LK_ERR_INVALID, "json: expected ',' or '}'");
      }
      return v;
    }
    if (c == '[') {
      p_++;
      v.kind = Json::Arr;
      if (peek() == ']') { p_++; return v; }
      while (true) {
        v.arr.push_back(value());
        char d = peek();
        p_++;
        if (d == ']') break;
        LK_CHECK(d == ',', LK_ERR_INVALID, "json: expected ',' or ']'");
      }
      return v;
    }
    if (c == '"') { v.kind = Json::Str; v.str = string(); return v; }
    if (s_.compare(p_, 4, "true") == 0) { p_ += 4; v.kind = Json::Bool; v.b = true; return v; }
    if (s_.compare(p_, 5, "false") == 0) { p_ += 5; v.kind = Json::Bool; v.b = false; return v; }
    if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return v; }
    size_t q = p_;
    bool isint = true;
    if (q < s_.size() && (s_[q] == '-' || s_[q] == '+')) q++;
    while (q < s_.size() && ((s_[q] >= '0' && s_[q] <= '9') || s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E' || s_[q] == '-' || s_[q] == '+')) {
      if (s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E') isint = false;
      q++;
    }
    LK_CHECK(q > p_, LK_ERR_INVALID, strf("json: unexpected character '%c' at %zu", c, p_));
    std::string t = s_.substr(p_, q - p_);
    v.kind = Json::Num;
    v.num = strtod(t.c_str(), nullptr);
    v.is_int = isint;
    if (isint) v.i64 = strtoll(t.c_str(), nullptr, 10);
    p_ = q;
    return v;
  }
};

inline Json parse_json(const std::string& s) { return JsonParser(s).parse(); }

inline void json_escape(std::string& out, const std::string& s) {
  out += '"';
  for (unsigned char c : s) {
    if (c == '"') out += "\\\"";
    else if (c == '\\') out += "\\\\";
    else if (c == '\n') out += "\\n";
    else if (c == '\r') out += "\\r";
    else if (c == '\t') out += "\\t";
    else if (c < 0x20) out += strf("\\u%04x", c);
    else out += (char)c;
  }
  out += '"';
}

}  // namespace lk

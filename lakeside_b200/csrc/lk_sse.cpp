// Native serialisation of result rows into the reference's per-segment stream elements (host code, no CUDA).
//
// Replaces, per row, DataPoint -> SketchInput(map sketch) -> mutable.HashMap -> Jackson -> "data: ...\r\n\r\n":
//   PushDownAggregatorStage.scala:95-106   SketchInput(ts, SketchTags(tags, "map", Right(Map(globalAgg -> value))))
//   Commons.scala:474-502                  dataPointResponseToSSE: message = {timestamp, tags, type: "sketch", sketchType, sketch}
//   SSEMessage.scala:23-34                 GenericSSEPayload(id = "_", type = "data", message) ; "data: " + json + "\r\n\r\n"
// and is read back by SegmentSequencer.decode (SegmentSequencer.scala:35-101): tags must be JSON strings; a sketch value
// may be a JSON number or one of the strings "NaN", "Infinity", "-Infinity" (what Jackson's default
// QUOTE_NON_NUMERIC_NUMBERS writes for non-finite doubles).  Key order inside an object is irrelevant to that decoder
// (the reference's comes out of a HashMap); doubles are written in the shortest form that parses back to the same bits.
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>

#include "lk_common.h"
#include "lk_query.h"

namespace lk {

namespace {

struct Out {
  char* p;
  int64_t cap, n = 0;  // n keeps counting past cap: the caller learns the size it needs
  void put(const char* s, size_t len) {
    if (n + (int64_t)len <= cap) memcpy(p + n, s, len);
    n += (int64_t)len;
  }
  void lit(const char* s) { put(s, strlen(s)); }
  void ch(char c) { put(&c, 1); }
  // JSON string, escaped like Jackson's default: \" \\ \b \f \n \r \t, other control characters as \u00XX,
  // everything else (UTF-8 included) verbatim
  void str(const char* s, size_t len) {
    ch('"');
    size_t run = 0;
    for (size_t i = 0; i < len; i++) {
      const unsigned char c = (unsigned char)s[i];
      const char* esc = nullptr;
      char ubuf[8];
      switch (c) {
        case '"': esc = "\\\""; break;
        case '\\': esc = "\\\\"; break;
        case '\b': esc = "\\b"; break;
        case '\f': esc = "\\f"; break;
        case '\n': esc = "\\n"; break;
        case '\r': esc = "\\r"; break;
        case '\t': esc = "\\t"; break;
        default:
          if (c < 0x20) {
            static const char hex[] = "0123456789ABCDEF";
            ubuf[0] = '\\'; ubuf[1] = 'u'; ubuf[2] = '0'; ubuf[3] = '0'; ubuf[4] = hex[c >> 4]; ubuf[5] = hex[c & 15]; ubuf[6] = 0;
            esc = ubuf;
          }
      }
      if (esc) {
        put(s + run, i - run);
        lit(esc);
        run = i + 1;
      }
    }
    put(s + run, len - run);
    ch('"');
  }
  void i64(int64_t v) {
    char b[24];
    auto r = std::to_chars(b, b + sizeof b, v);
    put(b, (size_t)(r.ptr - b));
  }
  void f64(double v) {
    if (std::isnan(v)) { lit("\"NaN\""); return; }
    if (std::isinf(v)) { lit(v > 0 ? "\"Infinity\"" : "\"-Infinity\""); return; }
    char b[40];
    auto r = std::to_chars(b, b + sizeof b, v);  // shortest round-trip
    put(b, (size_t)(r.ptr - b));
  }
};

}  // namespace

// Rows [row0, row1) of `res`; value column v is published under sketch key keys[v] (v < n_keys <= n_values).
// fallback_tags: n_fallback (key, value) pairs used when a row has no tag left (Commons.scala:448-451 queryTags).
int64_t result_to_sse(const HostResult& res, int64_t row0, int64_t row1, const char* const* keys, int n_keys, const char* const* fallback_tags,
                      int n_fallback, char* buf, int64_t cap) {
  LK_CHECK(row0 >= 0 && row1 >= row0 && row1 <= res.n, LK_ERR_INVALID, "lk_result_to_sse: row range out of bounds");
  LK_CHECK(n_keys >= 1 && n_keys <= res.n_values && keys, LK_ERR_INVALID, "lk_result_to_sse: one sketch key per published value column");
  LK_CHECK(n_fallback >= 0 && (n_fallback == 0 || fallback_tags), LK_ERR_INVALID, "lk_result_to_sse: bad fallback tags");
  Out o{buf, buf ? cap : 0};
  for (int64_t i = row0; i < row1; i++) {
    o.lit("data: {\"id\":\"_\",\"type\":\"data\",\"message\":{\"timestamp\":");
    o.i64(res.ts[i]);
    o.lit(",\"tags\":{");
    int ntags = 0;
    for (int t = 0; t < res.n_tags; t++) {
      const int32_t c = res.codes[t][i];
      if (c < 0) continue;  // SQL NULL
      const std::string& v = res.dicts[t][(size_t)c];
      if (v.empty() || v == "null") continue;  // Commons.scala:430-434
      if (ntags++) o.ch(',');
      const std::string& name = res.col_names[(size_t)(1 + res.n_values + t)];
      o.str(name.data(), name.size());
      o.ch(':');
      o.str(v.data(), v.size());
    }
    if (ntags == 0)
      for (int f = 0; f < n_fallback; f++) {
        if (f) o.ch(',');
        o.str(fallback_tags[2 * f], strlen(fallback_tags[2 * f]));
        o.ch(':');
        o.str(fallback_tags[2 * f + 1], strlen(fallback_tags[2 * f + 1]));
      }
    o.lit("},\"type\":\"sketch\",\"sketchType\":\"map\",\"sketch\":{");
    for (int v = 0; v < n_keys; v++) {
      if (v) o.ch(',');
      o.str(keys[v], strlen(keys[v]));
      o.ch(':');
      o.f64(res.values[(size_t)v][i]);  // a SQL-NULL aggregate reads 0.0 through getDouble (Commons.scala:426): stored as 0.0
    }
    o.lit("}}}\r\n\r\n");
  }
  return o.n;
}

}  // namespace lk

// TEST INFRASTRUCTURE: sequential host emulation of the fused scan kernel (lakeside_b200/csrc/lk_scan.cuh).
// It runs the SAME host planner (lk_plan.cpp) and the SAME per-value decode primitives (lk_device.h) the kernels
// use, one tile after another on the CPU, so that index/cursor/decode bugs are caught in the CPU test-suite where no
// GPU exists.  It is compiled only by tests/ (never linked into liblakeside_b200.so) and is not a product path.
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "lk_query.h"

namespace lk {
struct Query::Device {};
Query::Query() = default;
Query::~Query() = default;
HostResult::~HostResult() = default;
}  // namespace lk

using namespace lk;

struct EmulResult {
  std::vector<int64_t> ts;
  std::vector<std::vector<double>> values;
  std::vector<std::vector<uint8_t>> nulls;
  std::vector<std::vector<int32_t>> codes;
  std::vector<std::vector<std::string>> dicts;
  std::vector<std::string> col_names;
  std::string info;
  int64_t survivors = 0;
};

static thread_local std::string g_err;

struct Cell {
  uint64_t rows = 0;
  uint64_t acc[LK_MAX_AGGS] = {0, 0, 0, 0, 0, 0, 0};
};

static bool col_pos(const ColCursor& c, const uint32_t* bits, const uint16_t* pref, uint32_t r, uint32_t& vidx) {
  if (c.flags & CUR_ALL_VALID) { vidx = c.vidx0 + r; return true; }
  if (c.flags & CUR_ALL_NULL) return false;
  uint32_t w = bits[r >> 5], b = r & 31;
  vidx = c.vidx0 + pref[r >> 5] + (uint32_t)__builtin_popcount(w & ((1u << b) - 1));
  return (w >> b) & 1;
}

static void run(Query& q, EmulResult& out) {
  plan_query(q);
  out.info = q.info_json;
  std::vector<uint64_t> arena((q.arena_bytes + 7) / 8 + 2, 0);
  uint8_t* A = (uint8_t*)arena.data();
  for (auto& u : q.uploads) memcpy(A + u.arena_off, u.src ? u.src : q.segs[u.seg].data + u.file_off, u.len);
  const ScanParams& P = q.params;
  const int np = (int)P.npcols;
  // definition bitmaps the way def_expand_kernel builds them: one run at a time, every chunk of q.def_chunks
  std::vector<uint32_t> defbm(q.defbm_words + 2, 0);
  for (const DefChunk& dc : q.def_chunks) {
    uint32_t* w = defbm.data() + dc.word0;
    for (uint32_t k = 0; k < dc.run_n; k++) {
      const Run r = q.runs[dc.run_lo + k];
      const uint32_t next = k + 1 < dc.run_n ? q.runs[dc.run_lo + k + 1].start : dc.num_rows;
      lk_def_expand_run(A, dc.base_off, r, next, [&](uint32_t word, uint32_t m) { w[word] |= m; },
                        [&](uint32_t a, uint32_t b) { for (uint32_t i = a; i < b; i++) { LK_CHECK(w[i] == 0, LK_ERR_INVALID, "def bitmap: whole word written twice"); w[i] = 0xffffffffu; } });
    }
  }
  std::map<uint64_t, Cell> table;  // ordered by cell = bucket * n_groups + gid => sorted by timestamp
  uint32_t phase_min = 0xffffffffu, phase_max = 0;
  for (uint32_t t = 0; t < P.ntiles; t++) {
    const TileDesc& td = q.tiles[t];
    const ColCursor* cur = &q.cursors[td.cursor0];
    const ChunkInfo* ci = &q.chunk_infos[(size_t)td.rg * np];
    LK_CHECK(td.nrows <= (uint32_t)LK_TILE_ROWS_MAX, LK_ERR_INVALID, "tile too large");
    const uint32_t nwords = (td.nrows + 31) / 32;
    std::vector<std::vector<uint32_t>> bits(np, std::vector<uint32_t>(nwords + 1, 0));
    std::vector<std::vector<uint16_t>> pref(np, std::vector<uint16_t>(nwords + 1, 0));
    for (int p = 0; p < np; p++) {
      if (cur[p].flags & (CUR_ALL_VALID | CUR_ALL_NULL)) continue;
      uint32_t running = 0;
      for (uint32_t w = 0; w < nwords; w++) {
        uint32_t nb = std::min(32u, td.nrows - 32 * w);
        bits[p][w] = lk_def_word(A, q.runs.data(), cur[p], ci[p], td.row0 + 32 * w, nb);
        {  // the scan kernel reads the expanded bitmap instead of walking the runs: both must agree
          const uint32_t bit = td.row0 + 32 * w;
          const uint32_t* bw = defbm.data() + ci[p].defbm_word0 + (bit >> 5);
          uint32_t x = (bit & 31) ? (bw[0] >> (bit & 31)) | (bw[1] << (32 - (bit & 31))) : bw[0];
          if (nb < 32) x &= (1u << nb) - 1;
          LK_CHECK(x == bits[p][w], LK_ERR_INVALID, "expanded definition bitmap differs from the run walk");
        }
        pref[p][w] = (uint16_t)running;
        running += (uint32_t)__builtin_popcount(bits[p][w]);
      }
      LK_CHECK(running == cur[p].nvals, LK_ERR_INVALID, "def bitmap popcount != cursor nvals");
    }
    for (uint32_t r = 0; r < td.nrows; r++) {
      uint32_t idx = 0, vidx;
      for (int f = 0; f < P.n_filter; f++) {
        const FilterCol& fc = P.filter[f];
        int p = fc.pcol;
        uint32_t cls;
        if (!col_pos(cur[p], bits[p].data(), pref[p].data(), r, vidx)) cls = fc.null_cls;
        else if (fc.numeric) {
          uint32_t bad = 0;
          uint64_t b = lk_value_bits(A, q.runs.data(), cur[p], ci[p], vidx, &bad);
          LK_CHECK(!bad, LK_ERR_IO, "bad code");
          cls = lk_numeric_class(fc, lk_bits_to_f64(b, ci[p].phys_type));
        } else {
          uint32_t code = lk_dict_code(A, q.runs.data(), cur[p], ci[p], vidx);
          LK_CHECK(code < ci[p].dict_n, LK_ERR_IO, "bad code");
          cls = q.lut_cls[ci[p].lut_cls + code];
        }
        idx += cls * fc.stride;
      }
      if (!((q.pass_bits[idx >> 5] >> (idx & 31)) & 1)) continue;
      if (P.notnull_pcol >= 0 && !col_pos(cur[P.notnull_pcol], bits[P.notnull_pcol].data(), pref[P.notnull_pcol].data(), r, vidx)) continue;
      if (!col_pos(cur[P.ts_pcol], bits[P.ts_pcol].data(), pref[P.ts_pcol].data(), r, vidx)) continue;
      uint32_t bad = 0;
      int64_t ts = (int64_t)lk_value_bits(A, q.runs.data(), cur[P.ts_pcol], ci[P.ts_pcol], vidx, &bad);
      LK_CHECK(!bad, LK_ERR_IO, "bad code");
      if (ts < P.ts_lo || ts >= P.ts_hi) continue;
      uint64_t rel = (uint64_t)(ts - P.base), bucket = rel / (uint64_t)P.step;
      if (P.is_metrics) {
        uint32_t ph = (uint32_t)(rel - bucket * (uint64_t)P.step);
        phase_min = std::min(phase_min, ph);
        phase_max = std::max(phase_max, ph);
      }
      uint64_t gid = 0;
      for (int k = 0; k < P.n_keys; k++) {
        int p = P.keys[k].pcol;
        uint32_t g = P.keys[k].null_code;
        if (col_pos(cur[p], bits[p].data(), pref[p].data(), r, vidx)) {
          uint32_t code = lk_dict_code(A, q.runs.data(), cur[p], ci[p], vidx);
          LK_CHECK(code < ci[p].dict_n, LK_ERR_IO, "bad code");
          g = q.lut_gcode[ci[p].lut_gcode + code];
        }
        gid += (uint64_t)g * P.keys[k].stride;
      }
      Cell& c = table[bucket * P.n_groups + gid];
      c.rows++;
      out.survivors++;
      for (int a = 0; a < P.n_aggs; a++) {
        int p = P.aggs[a].pcol;
        if (!col_pos(cur[p], bits[p].data(), pref[p].data(), r, vidx)) continue;
        uint64_t raw = lk_value_bits(A, q.runs.data(), cur[p], ci[p], vidx, &bad);
        double x = lk_bits_to_f64(raw, ci[p].phys_type);
        uint64_t xb;
        memcpy(&xb, &x, 8);
        switch (P.aggs[a].op) {
          case AGG_SUM: { double s; memcpy(&s, &c.acc[a], 8); s += x; memcpy(&c.acc[a], &s, 8); break; }
          case AGG_COUNT: c.acc[a]++; break;
          case AGG_MIN: c.acc[a] = std::max(c.acc[a], lk_min_encode(xb)); break;
          default: c.acc[a] = std::max(c.acc[a], lk_max_encode(xb)); break;
        }
      }
    }
  }
  LK_CHECK(!(P.is_metrics && phase_min != 0xffffffffu && phase_min != phase_max), LK_ERR_UNSUPPORTED,
           "metric timestamps are not on one step-aligned grid");
  uint32_t phase = (P.is_metrics && phase_min != 0xffffffffu) ? phase_min : 0;
  const int na = P.n_aggs, nk = P.n_keys;
  out.values.assign(na, {});
  out.nulls.assign(na, {});
  out.codes.assign(nk, {});
  out.dicts = q.key_dicts;
  out.col_names.push_back(q.ts_col_name);
  for (int a = 0; a < na; a++) out.col_names.push_back("value");
  for (auto& k : q.key_names) out.col_names.push_back(k);
  for (auto& kv : table) {
    uint64_t cell = kv.first, bucket = cell / P.n_groups, gid = cell % P.n_groups;
    out.ts.push_back(P.base + (int64_t)bucket * P.step + phase);
    for (int a = 0; a < na; a++) {
      uint64_t w = kv.second.acc[a];
      double v = 0;
      uint8_t isnull = 0;
      uint64_t d;
      switch (P.aggs[a].op) {
        case AGG_SUM: memcpy(&v, &w, 8); break;
        case AGG_COUNT: v = (double)w; break;
        case AGG_MIN: if (!w) isnull = 1; else { d = lk_min_decode(w); memcpy(&v, &d, 8); } break;
        default: if (!w) isnull = 1; else { d = lk_max_decode(w); memcpy(&v, &d, 8); } break;
      }
      if (q.aggs[a].divisor != 1.0 && !isnull) v /= q.aggs[a].divisor;
      out.values[a].push_back(v);
      out.nulls[a].push_back(isnull);
    }
    for (int k = 0; k < nk; k++) {
      uint32_t g = (uint32_t)((gid / P.keys[k].stride) % ((uint64_t)P.keys[k].null_code + 1));
      out.codes[k].push_back(g == P.keys[k].null_code ? -1 : (int32_t)g);
    }
  }
}

extern "C" {
__attribute__((visibility("default"))) const char* lk_emul_last_error() { return g_err.c_str(); }

__attribute__((visibility("default"))) int lk_emul_eval(const char* req_json, const char* aggs_json, const char* path, int tile_rows, int nseg,
                                                         const void* const* bufs, const size_t* lens, EmulResult** out) {
  try {
    Query q;
    q.req = parse_push_down_request(req_json);
    if (path && *path) q.path_opt = path;
    if (tile_rows > 0) global_options().tile_rows = (uint32_t)tile_rows;
    global_options().host_threads = 2;
    if (aggs_json && *aggs_json) {
      Json j = parse_json(aggs_json);
      for (auto& x : j.arr) {
        AggSpec s;
        s.aggregation = x.get("aggregation")->str;
        s.op = s.aggregation == "sum" ? AGG_SUM : s.aggregation == "count" ? AGG_COUNT : s.aggregation == "min" ? AGG_MIN : AGG_MAX;
        const Json* ro = x.get("rollup");
        s.value_column = q.req.expr.dataset == "metrics" ? "rollup_" + (ro ? ro->str : std::string("sum")) : "_cardinalhq.value";
        q.aggs.push_back(s);
      }
    }
    for (int i = 0; i < nseg; i++) {
      SegmentInput s;
      s.data = (const uint8_t*)bufs[i];
      s.len = lens[i];
      q.segs.push_back(s);
    }
    auto r = new EmulResult();
    try { run(q, *r); } catch (...) { delete r; throw; }
    *out = r;
    return 0;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
__attribute__((visibility("default"))) int64_t lk_emul_rows(EmulResult* r) { return (int64_t)r->ts.size(); }
__attribute__((visibility("default"))) int64_t lk_emul_survivors(EmulResult* r) { return r->survivors; }
__attribute__((visibility("default"))) int lk_emul_nvalues(EmulResult* r) { return (int)r->values.size(); }
__attribute__((visibility("default"))) int lk_emul_ntags(EmulResult* r) { return (int)r->codes.size(); }
__attribute__((visibility("default"))) const int64_t* lk_emul_ts(EmulResult* r) { return r->ts.data(); }
__attribute__((visibility("default"))) const double* lk_emul_value(EmulResult* r, int a) { return r->values[a].data(); }
__attribute__((visibility("default"))) const uint8_t* lk_emul_null(EmulResult* r, int a) { return r->nulls[a].data(); }
__attribute__((visibility("default"))) const int32_t* lk_emul_codes(EmulResult* r, int k) { return r->codes[k].data(); }
__attribute__((visibility("default"))) int lk_emul_dict_size(EmulResult* r, int k) { return (int)r->dicts[k].size(); }
__attribute__((visibility("default"))) const char* lk_emul_dict(EmulResult* r, int k, int i) { return r->dicts[k][i].c_str(); }
__attribute__((visibility("default"))) const char* lk_emul_col_name(EmulResult* r, int i) { return r->col_names[i].c_str(); }
__attribute__((visibility("default"))) const char* lk_emul_info(EmulResult* r) { return r->info.c_str(); }
__attribute__((visibility("default"))) void lk_emul_free(EmulResult* r) { delete r; }
}

// regex engine under test (the class is not exported by the product library)
extern "C" __attribute__((visibility("default"))) int lk_emul_regex(const char* pattern, int case_insensitive, const char* s, int len) {
  try {
    Regex re(pattern, case_insensitive != 0);
    return re.search(s, (size_t)len) ? 1 : 0;
  } catch (const Error& e) {
    g_err = e.what();
    return -e.code;
  }
}

// C ABI of liblakeside_b200 (see include/lakeside_b200.h for the contract and the reference seams it replaces).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__GLIBC__)
#include <malloc.h>
#endif

#include "lk_cache.h"
#include "lk_engine.h"
#include "lk_merge.h"

using namespace lk;

struct lk_query { Query q; };
struct lk_result { HostResult* r; };
struct lk_comm { Comm* c; };

static thread_local std::string tl_error;

template <class F>
static int guard(F&& f) {
  try {
    f();
    tl_error.clear();
    return LK_OK;
  } catch (const Error& e) {
    tl_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    tl_error = "out of host memory";
    return LK_ERR_NOMEM;
  } catch (const std::exception& e) {
    tl_error = e.what();
    return LK_ERR_INVALID;
  }
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#pragma GCC visibility push(default)
extern "C" {

int lk_init(const char* options_json) {
  return guard([&] {
    Options& o = global_options();
    if (options_json && *options_json) {
      Json j = parse_json(options_json);
      LK_CHECK(j.is_obj(), LK_ERR_INVALID, "lk_init options must be a JSON object");
      if (const Json* v = j.get("device")) o.device = (int)v->as_i64();
      if (const Json* v = j.get("max_hash_slots")) o.max_hash_slots = (uint64_t)v->as_i64();
      if (const Json* v = j.get("dense_max_cells")) o.dense_max_cells = (uint64_t)v->as_i64();
      if (const Json* v = j.get("segment_cache_bytes")) o.segment_cache_bytes = v->as_i64();
      if (const Json* v = j.get("tile_rows")) o.tile_rows = (uint32_t)v->as_i64();
      if (const Json* v = j.get("host_threads")) o.host_threads = (int)v->as_i64();
      if (const Json* v = j.get("tune_host_malloc")) o.tune_host_malloc = v->as_i64() != 0;
      LK_CHECK(o.max_hash_slots >= 1024 && (o.max_hash_slots & (o.max_hash_slots - 1)) == 0, LK_ERR_INVALID, "max_hash_slots must be a power of two >= 1024");
    }
#if defined(__GLIBC__)
    if (o.tune_host_malloc) {
      // Every query builds ~0.3 GB of transient index memory on the C heap (run lists of the column chunks).  With glibc's
      // defaults the freed top of the heap is unmapped at the end of each query and faulted in again by the next one:
      // 8 ms on the caller's thread per 100-segment query (measured).  Keep it inside the process instead.
      mallopt(M_TRIM_THRESHOLD, 1 << 30);
      mallopt(M_MMAP_THRESHOLD, 32 << 20);
    }
#endif
    device_init();
  });
}

void lk_shutdown(void) { device_shutdown(); }
const char* lk_last_error(void) { return tl_error.c_str(); }
const char* lk_version(void) { return "lakeside_b200 0.1 (sm_100a)"; }
int lk_device_count(void) { return device_count(); }

void* lk_host_alloc(size_t bytes) {
  void* p = nullptr;
  guard([&] { device_init(); p = pinned_alloc(bytes); });
  return p;
}
void lk_host_free(void* p) { pinned_free(p); }

int lk_cache_stats(int64_t* out) {
  if (!out) return LK_ERR_INVALID;
  SegmentCacheStats st = segment_cache().stats();
  out[0] = st.capacity_bytes; out[1] = st.resident_bytes; out[2] = st.segments;
  out[3] = st.column_hits; out[4] = st.column_misses; out[5] = st.evicted_segments;
  return LK_OK;
}
int lk_cache_configure(int64_t capacity_bytes) {
  return guard([&] {
    LK_CHECK(capacity_bytes >= -1, LK_ERR_INVALID, "capacity_bytes must be >= 0, or -1 for the default (a third of the device's memory)");
    global_options().segment_cache_bytes = capacity_bytes;
    if (capacity_bytes < 0) capacity_bytes = device_default_cache_bytes();  // 0 before lk_init: device_init resolves the default then
    segment_cache().set_capacity((size_t)capacity_bytes);
  });
}
void lk_cache_clear(void) { segment_cache().clear(); }

int lk_query_create(const char* pushdown_request_json, const char* options_json, lk_query** out) {
  return guard([&] {
    LK_CHECK(pushdown_request_json && out, LK_ERR_INVALID, "null argument");
    auto h = std::make_unique<lk_query>();
    h->q.req = parse_push_down_request(pushdown_request_json);
    if (options_json && *options_json) {
      Json j = parse_json(options_json);
      LK_CHECK(j.is_obj(), LK_ERR_INVALID, "query options must be a JSON object");
      if (const Json* p = j.get("path"); p && p->text()) {
        LK_CHECK(p->str == "auto" || p->str == "dense" || p->str == "hash" || p->str == "records", LK_ERR_INVALID, "path must be auto|dense|hash|records");
        h->q.path_opt = p->str;
      }
      if (const Json* e = j.get("exact_sums")) h->q.exact_sums = e->as_bool();
      if (const Json* e = j.get("seq_offset")) h->q.seq_offset = (uint64_t)e->as_i64();
      if (const Json* a = j.get("aggregates"); a && a->is_arr()) {
        const BaseExpr& e = h->q.req.expr;
        for (auto& x : a->arr) {
          LK_CHECK(x.is_obj(), LK_ERR_INVALID, "aggregates[] entries must be objects");
          const Json* ag = x.get("aggregation");
          LK_CHECK(ag && ag->text(), LK_ERR_INVALID, "aggregates[].aggregation missing");
          const Json* ro = x.get("rollup");
          // re-uses the planner's resolution through a temporary single-aggregate expression
          AggSpec s;
          s.aggregation = ag->str;
          LK_CHECK(!(ag->str.rfind("p", 0) == 0 || ag->str.find("ces") != std::string::npos), LK_ERR_UNSUPPORTED,
                   "percentile / cardinality-estimate aggregations are outside the GPU path");
          if (ag->str == "sum") s.op = AGG_SUM;
          else if (ag->str == "count") s.op = AGG_COUNT;
          else if (ag->str == "min") s.op = AGG_MIN;
          else if (ag->str == "max") s.op = AGG_MAX;
          else fail(ag->str == "avg" ? LK_ERR_UNSUPPORTED : LK_ERR_QUERY, "aggregate function " + ag->str + " is not available");
          if (e.dataset == "metrics") s.value_column = "rollup_" + ((ro && ro->text()) ? ro->str : std::string("sum"));
          else {
            LK_CHECK(!e.chart.has_field_name || e.chart.field_name == "_cardinalhq.value", LK_ERR_UNSUPPORTED, "chart fieldName in a fused multi-aggregate pass");
            s.value_column = "_cardinalhq.value";
          }
          h->q.aggs.push_back(s);
        }
      }
    }
    *out = h.release();
  });
}

int lk_query_add_segment_buffer(lk_query* q, const void* data, size_t len) {
  return guard([&] {
    LK_CHECK(q && data, LK_ERR_INVALID, "null argument");
    LK_CHECK(!q->q.prepared, LK_ERR_INVALID, "query already prepared");
    SegmentInput s;
    s.data = (const uint8_t*)data;
    s.len = len;
    s.name = "buffer#" + std::to_string(q->q.segs.size());
    q->q.segs.push_back(std::move(s));
  });
}

int lk_query_add_segment_file(lk_query* q, const char* path) {
  return guard([&] {
    LK_CHECK(q && path, LK_ERR_INVALID, "null argument");
    LK_CHECK(!q->q.prepared, LK_ERR_INVALID, "query already prepared");
    struct stat st;
    LK_CHECK(stat(path, &st) == 0 && S_ISREG(st.st_mode), LK_ERR_IO, std::string("IO Error: cannot open ") + path);
    device_init();
    SegmentInput s;
    s.len = (size_t)st.st_size;
    s.name = path;
    s.has_identity = true;
    s.id_mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    s.id_ino = (uint64_t)st.st_ino;
    // a segment the cache knows (same path, size, mtime, inode) brings its footer along and is only read if a column misses
    SegmentIdentity id;
    id.path = s.name; id.size = s.len; id.mtime_ns = s.id_mtime_ns; id.ino = s.id_ino;
    s.cached = segment_cache().lookup(id);
    if (s.cached) {
      s.meta = s.cached->meta;
      s.meta_from_cache = true;
    } else segment_open(s);
    q->q.segs.push_back(std::move(s));
  });
}

int lk_query_plan(lk_query* q) {
  return guard([&] {
    LK_CHECK(q, LK_ERR_INVALID, "null argument");
    LK_CHECK(!q->q.prepared, LK_ERR_INVALID, "query already planned");
    double t0 = now_ms();
    // With a GPU present the column chunks start moving to HBM as soon as the arena layout is known, exactly as in
    // lk_query_prepare; without one (host-logic tests, dictionary agreement on a CPU box) planning is host-only.
    Query* qp = &q->q;
    if (device_count() > 0) {
      q->q.on_layout = [qp] { device_layout(*qp); };
      q->q.device_index = !getenv("LK_HOST_INDEX");  // run headers and cursors are built on the device during prepare
    }
    plan_query(q->q);
    q->q.on_layout = nullptr;
    q->q.t_ms[4] = now_ms() - t0;
  });
}

int lk_query_prepare(lk_query* q) {
  return guard([&] {
    LK_CHECK(q, LK_ERR_INVALID, "null argument");
    LK_CHECK(!device_is_resident(q->q), LK_ERR_INVALID, "query already prepared");
    if (!q->q.prepared) {
      double t0 = now_ms();
      device_init();  // fail before any host work when there is no GPU
      Query* qp = &q->q;
      q->q.on_layout = [qp] { device_layout(*qp); };  // column chunks start moving while the host walks the page headers
      q->q.device_index = !getenv("LK_HOST_INDEX");          // run headers and cursors are built on the device (LK_HOST_INDEX: on the host)
      plan_query(q->q);
      q->q.on_layout = nullptr;
      q->q.t_ms[4] = now_ms() - t0;
    }
    device_upload(q->q);
  });
}

int lk_query_export_dictionaries(lk_query* q, const void** blob, size_t* len) {
  return guard([&] {
    LK_CHECK(q && blob && len, LK_ERR_INVALID, "null argument");
    LK_CHECK(q->q.prepared, LK_ERR_INVALID, "query not prepared");
    q->q.dict_blob = export_dictionaries(q->q);
    *blob = q->q.dict_blob.data();
    *len = q->q.dict_blob.size();
  });
}

int lk_query_import_dictionaries(lk_query* q, const void* blob, size_t len) {
  return guard([&] {
    LK_CHECK(q && blob, LK_ERR_INVALID, "null argument");
    LK_CHECK(q->q.prepared, LK_ERR_INVALID, "query not prepared");
    import_dictionaries(q->q, (const uint8_t*)blob, len);
    device_mark_group_tables_stale(q->q);
  });
}

int lk_query_execute(lk_query* q) {
  return guard([&] { LK_CHECK(q && q->q.prepared, LK_ERR_INVALID, "query not prepared"); device_execute(q->q); });
}

int lk_query_sync(lk_query* q) {
  return guard([&] { LK_CHECK(q, LK_ERR_INVALID, "null argument"); device_sync(q->q); });
}

int lk_query_phase(lk_query* q, uint32_t* phase_min, uint32_t* phase_max) {
  return guard([&] { LK_CHECK(q && phase_min && phase_max, LK_ERR_INVALID, "null argument"); device_phase(q->q, phase_min, phase_max); });
}
int lk_query_set_phase(lk_query* q, uint32_t phase_min, uint32_t phase_max) {
  return guard([&] { LK_CHECK(q, LK_ERR_INVALID, "null argument"); device_set_phase(q->q, phase_min, phase_max); });
}

int lk_query_partial_dense(lk_query* q, int64_t* n_cells, int* n_planes, void** plane_ptrs, int* plane_ops) {
  return guard([&] {
    LK_CHECK(q && n_cells && n_planes && plane_ptrs && plane_ops, LK_ERR_INVALID, "null argument");
    device_partial_dense(q->q, n_cells, n_planes, plane_ptrs, plane_ops);
  });
}

int lk_query_partial_sparse(lk_query* q, int nparts, void** entries, int64_t* counts, int* stride_bytes) {
  return guard([&] {
    LK_CHECK(q && entries && counts && stride_bytes, LK_ERR_INVALID, "null argument");
    device_partial_sparse(q->q, nparts, entries, counts, stride_bytes);
  });
}
int lk_query_merge_sparse(lk_query* q, const void* device_entries, int64_t n) {
  return guard([&] {
    LK_CHECK(q && (device_entries || n == 0), LK_ERR_INVALID, "null argument");
    device_merge_sparse(q->q, device_entries, n);
  });
}

int lk_query_finalize_device(lk_query* q) {
  return guard([&] { LK_CHECK(q && q->q.prepared, LK_ERR_INVALID, "query not prepared"); device_finalize_device(q->q); });
}

int lk_query_finalize(lk_query* q, lk_result** out) {
  return guard([&] {
    LK_CHECK(q && out, LK_ERR_INVALID, "null argument");
    LK_CHECK(q->q.prepared, LK_ERR_INVALID, "query not prepared");
    device_finalize_device(q->q);
    auto r = std::make_unique<lk_result>();
    r->r = device_fetch(q->q);
    *out = r.release();
  });
}

int lk_query_timings(lk_query* q, double* ms) {
  return guard([&] {
    LK_CHECK(q && ms, LK_ERR_INVALID, "null argument");
    if (q->q.dev) { device_sync(q->q); device_timings(q->q); }
    for (int i = 0; i < 8; i++) ms[i] = q->q.t_ms[i];
  });
}

int64_t lk_query_touched_bytes(lk_query* q) { return q ? q->q.touched_bytes : -1; }
int64_t lk_query_total_rows(lk_query* q) { return q ? q->q.total_rows : -1; }
int64_t lk_query_survivors(lk_query* q) {
  int64_t v = -1;
  guard([&] { LK_CHECK(q && q->q.dev, LK_ERR_INVALID, "query not prepared"); v = device_survivors(q->q); });
  return v;
}

int64_t lk_query_eval(lk_query* q, const char* aggregation, const char* chart_type, const char* metric_type, double* out, int64_t cap) {
  int64_t n = -1;
  guard([&] {
    LK_CHECK(q && q->q.dev && aggregation && chart_type && metric_type, LK_ERR_INVALID, "lk_query_eval: bad arguments");
    n = device_eval(q->q, aggregation, chart_type, metric_type, out, cap);
  });
  return n;
}

int lk_formula_eval(lk_query* e1, lk_query* e2, const char* spec_json, int64_t cap, int64_t* out_ts, double* out_value, int32_t* out_side,
                    int64_t* out_row, int64_t* n_out) {
  return guard([&] {
    LK_CHECK(spec_json && n_out && (cap == 0 || (out_ts && out_value && out_side && out_row)), LK_ERR_INVALID, "lk_formula_eval: bad arguments");
    *n_out = device_formula(e1 ? &e1->q : nullptr, e2 ? &e2->q : nullptr, spec_json, cap, out_ts, out_value, out_side, out_row);
  });
}

int lk_comm_create(int rank, int world, int64_t pool_records, int max_aggs, lk_comm** out) {
  return guard([&] {
    LK_CHECK(out, LK_ERR_INVALID, "null argument");
    auto h = std::make_unique<lk_comm>();
    h->c = comm_create(rank, world, pool_records, max_aggs);
    *out = h.release();
  });
}
int lk_comm_handle(lk_comm* c, const void** blob, size_t* len) {
  return guard([&] { LK_CHECK(c && blob && len, LK_ERR_INVALID, "null argument"); comm_handle(c->c, blob, len); });
}
int lk_comm_connect(lk_comm* c, const void* blobs, size_t len_each) {
  return guard([&] { LK_CHECK(c && blobs, LK_ERR_INVALID, "null argument"); comm_connect(c->c, blobs, len_each); });
}
void lk_comm_destroy(lk_comm* c) {
  if (!c) return;
  comm_destroy(c->c);
  delete c;
}
int lk_query_set_comm(lk_query* q, lk_comm* c) {
  return guard([&] {
    LK_CHECK(q, LK_ERR_INVALID, "null argument");
    q->q.comm = c ? c->c : nullptr;
  });
}

int lk_query_stream(lk_query* q, void** cuda_stream) {
  return guard([&] {
    LK_CHECK(q && cuda_stream && q->q.dev, LK_ERR_INVALID, "query not prepared");
    *cuda_stream = device_stream(q->q);
  });
}

int lk_query_info_json(lk_query* q, const char** json) {
  return guard([&] {
    LK_CHECK(q && json, LK_ERR_INVALID, "null argument");
    *json = q->q.info_json.c_str();
  });
}

void lk_query_destroy(lk_query* q) {
  const bool trace = getenv("LK_PLAN_TRACE") != nullptr;
  const double t0 = trace ? now_ms() : 0;
  delete q;
  if (trace) fprintf(stderr, "[lk destroy] %35.2f ms\n", now_ms() - t0);
}

int lk_eval(const char* pushdown_request_json, const char* const* parquet_paths, int n_paths, lk_result** out) {
  lk_query* q = nullptr;
  int rc = lk_query_create(pushdown_request_json, nullptr, &q);
  for (int i = 0; rc == LK_OK && i < n_paths; i++) rc = lk_query_add_segment_file(q, parquet_paths[i]);
  if (rc == LK_OK) rc = lk_query_prepare(q);
  if (rc == LK_OK) rc = lk_query_execute(q);
  if (rc == LK_OK) rc = lk_query_finalize(q, out);
  std::string err = tl_error;
  lk_query_destroy(q);
  tl_error = err;
  return rc;
}

// ---- result accessors ----
int64_t lk_result_num_rows(const lk_result* r) { return r ? r->r->n : 0; }
int lk_result_num_values(const lk_result* r) { return r ? r->r->n_values : 0; }
int lk_result_num_tags(const lk_result* r) { return r ? r->r->n_tags : 0; }
int lk_result_num_cols(const lk_result* r) { return r ? (int)r->r->col_names.size() : 0; }
const char* lk_result_col_name(const lk_result* r, int col) {
  if (!r || col < 0 || col >= (int)r->r->col_names.size()) return nullptr;
  return r->r->col_names[col].c_str();
}
const int64_t* lk_result_ts(const lk_result* r) { return r ? r->r->ts : nullptr; }
const double* lk_result_value(const lk_result* r, int a) { return (r && a >= 0 && a < r->r->n_values) ? r->r->values[a] : nullptr; }
const uint8_t* lk_result_value_null(const lk_result* r, int a) { return (r && a >= 0 && a < r->r->n_values) ? r->r->nulls[a] : nullptr; }
const int32_t* lk_result_tag_codes(const lk_result* r, int t) { return (r && t >= 0 && t < r->r->n_tags) ? r->r->codes[t] : nullptr; }
int lk_result_tag_dict(const lk_result* r, int t, int32_t* n, const char* const** strings) {
  if (!r || t < 0 || t >= r->r->n_tags || !n || !strings) return LK_ERR_INVALID;
  *n = (int32_t)r->r->dicts[t].size();
  *strings = r->r->dict_ptrs[t].data();
  return LK_OK;
}
int64_t lk_result_get_long(const lk_result* r, int64_t row, int col) {
  if (!r || row < 0 || row >= r->r->n) return 0;
  if (r->r->tag_query) return col == 2 ? (int64_t)r->r->values[0][row] : 0;
  if (col == 1) return r->r->ts[row];
  if (col >= 2 && col < 2 + r->r->n_values) return (int64_t)r->r->values[col - 2][row];
  return 0;
}
double lk_result_get_double(const lk_result* r, int64_t row, int col) {
  if (!r || row < 0 || row >= r->r->n) return 0.0;
  if (r->r->tag_query) return col == 2 ? r->r->values[0][row] : 0.0;
  if (col == 1) return (double)r->r->ts[row];
  if (col >= 2 && col < 2 + r->r->n_values) return r->r->values[col - 2][row];
  return 0.0;
}
const char* lk_result_get_string(const lk_result* r, int64_t row, int col) {
  if (!r || row < 0 || row >= r->r->n) return nullptr;
  if (r->r->tag_query) {  // (tag, count): Commons.toDataPoint reads BOTH through getString (Commons.scala:407-416)
    if (col == 1) { int32_t c = r->r->codes[0][row]; return c < 0 ? nullptr : r->r->dict_ptrs[0][c]; }
    if (col != 2) return nullptr;
    static thread_local char buf[32];  // valid until the calling thread's next lk_result_get_string
    snprintf(buf, sizeof buf, "%lld", (long long)r->r->values[0][row]);
    return buf;
  }
  int t = col - 2 - r->r->n_values;
  if (t < 0 || t >= r->r->n_tags) return nullptr;
  int32_t c = r->r->codes[t][row];
  return c < 0 ? nullptr : r->r->dict_ptrs[t][c];
}
int64_t lk_result_to_sse(const lk_result* r, int64_t row0, int64_t row1, const char* const* sketch_keys, int n_keys,
                         const char* const* fallback_tags, int n_fallback, char* buf, int64_t cap) {
  int64_t n = -1;
  guard([&] {
    LK_CHECK(r && r->r, LK_ERR_INVALID, "null result");
    n = result_to_sse(*r->r, row0, row1, sketch_keys, n_keys, fallback_tags, n_fallback, buf, cap);
  });
  return n;
}
void lk_result_free(lk_result* r) {
  if (!r) return;
  delete r->r;
  delete r;
}

// ---- K-way merge ----
int lk_merge_create(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val, const int64_t* lens,
                    int reverse, lk_merge** out) {
  return guard([&] {
    LK_CHECK(out && k >= 0 && (k == 0 || (ts && lens)), LK_ERR_INVALID, "bad argument");
    *out = merge_create(k, ts, gid, val, lens, reverse != 0);
  });
}
int lk_merge_run(lk_merge* m) { return guard([&] { LK_CHECK(m, LK_ERR_INVALID, "null argument"); merge_run(m); }); }
int lk_merge_sync(lk_merge* m) { return guard([&] { LK_CHECK(m, LK_ERR_INVALID, "null argument"); merge_sync(m); }); }
int lk_merge_timings(lk_merge* m, double* ms) { return guard([&] { LK_CHECK(m && ms, LK_ERR_INVALID, "null argument"); merge_timings(m, ms); }); }
int lk_merge_download(lk_merge* m, int64_t* out_ts, int32_t* out_gid, double* out_val, int32_t* out_src) {
  return guard([&] { LK_CHECK(m, LK_ERR_INVALID, "null argument"); merge_download(m, out_ts, out_gid, out_val, out_src); });
}
int lk_merge_reduce(lk_merge* m, int op, int64_t* n_out, int64_t* out_ts, int32_t* out_gid, double* out_val) {
  return guard([&] { LK_CHECK(m && n_out, LK_ERR_INVALID, "null argument"); merge_reduce(m, op, n_out, out_ts, out_gid, out_val); });
}
void lk_merge_destroy(lk_merge* m) { merge_destroy(m); }

int lk_merge_streams(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val, const int64_t* lens,
                     int reverse, int64_t* out_ts, int32_t* out_gid, double* out_val, int32_t* out_src) {
  lk_merge* m = nullptr;
  int rc = lk_merge_create(k, ts, gid, val, lens, reverse, &m);
  if (rc == LK_OK) rc = lk_merge_run(m);
  if (rc == LK_OK) rc = lk_merge_download(m, out_ts, out_gid, out_val, out_src);
  std::string err = tl_error;
  lk_merge_destroy(m);
  tl_error = err;
  return rc;
}

}  // extern "C"
#pragma GCC visibility pop

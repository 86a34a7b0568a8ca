// Host planner: what Commons.toGlobResultSet + BaseExpr.generateSql decide before DuckDB runs
// (core/src/main/scala/com/cardinal/utils/Commons.scala:200-254, utils/ast/BaseExpr.scala:108-513), restated as a
// device plan: touched columns, [startTs,endTs) and step, predicate -> dictionary-code class tables + pass bitmap,
// group-by -> global dictionaries + code remap tables, and the tile/cursor/run seek index over the Parquet pages.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <string_view>

#include "lk_query.h"

namespace lk {

static const char* TIMESTAMP = "_cardinalhq.timestamp";
static const char* NAME = "_cardinalhq.name";
static const char* VALUE = "_cardinalhq.value";

void* (*PoolAlloc::alloc)(size_t) = malloc;
void (*PoolAlloc::release)(void*) = free;

Options& global_options() {
  static Options o;
  return o;
}

void parallel_for(int n, int threads, const std::function<void(int)>& fn) {
  if (threads <= 0) threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
  threads = std::min(threads, n);
  if (threads <= 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::string> errs(threads);
  std::vector<int> codes(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t] {
      try {
        for (int i; (i = next.fetch_add(1)) < n;) fn(i);
      } catch (const Error& e) {
        codes[t] = e.code;
        errs[t] = e.what();
        next.store(n);
      } catch (const std::exception& e) {
        codes[t] = LK_ERR_INVALID;
        errs[t] = e.what();
        next.store(n);
      }
    });
  for (auto& th : pool) th.join();
  for (int t = 0; t < threads; t++)
    if (codes[t]) fail(codes[t], errs[t]);
}

static bool starts_with(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }

// (aggregation, rollup) -> aggregate op + value column.  BaseExpr.getChartSql, BaseExpr.scala:348-369, 376-395.
static AggSpec resolve_agg(const BaseExpr& e, const std::string& aggregation, bool has_rollup, const std::string& rollup,
                           bool allow_field) {
  AggSpec a;
  a.aggregation = aggregation;
  LK_CHECK(!(starts_with(aggregation, "p") || aggregation.find("ces") != std::string::npos), LK_ERR_UNSUPPORTED,
           "percentile / cardinality-estimate aggregations (DDSketch / HLL branches of PushDownAggregatorStage) are outside the GPU path");
  LK_CHECK(aggregation != "avg", LK_ERR_UNSUPPORTED,
           "avg reaches the worker decomposed into sum and count (QueryEngineV2.scala:280-283); request those");
  if (aggregation == "sum") a.op = AGG_SUM;
  else if (aggregation == "count") a.op = AGG_COUNT;
  else if (aggregation == "min") a.op = AGG_MIN;
  else if (aggregation == "max") a.op = AGG_MAX;
  else fail(LK_ERR_QUERY, "Catalog Error: aggregate function " + aggregation + " does not exist");
  if (e.dataset == "metrics") {
    LK_CHECK(!e.chart.has_field_name, LK_ERR_UNSUPPORTED, "chart fieldName on a metrics expression");
    a.value_column = "rollup_" + (has_rollup ? rollup : std::string("sum"));
  } else if (e.chart.has_field_name && e.chart.field_name != VALUE) {
    LK_CHECK(allow_field, LK_ERR_UNSUPPORTED, "chart fieldName in a fused multi-aggregate pass");
    LK_CHECK(e.chart.has_field_type, LK_ERR_INVALID, "Required property: fieldType when chartType = `field`");
    a.value_column = e.chart.field_name + "$" + e.chart.field_type;
    if (e.chart.field_type == "duration") a.divisor = 1000000.0;
    else if (e.chart.field_type == "datasize") a.divisor = 1000.0;
  } else {
    a.value_column = VALUE;
  }
  return a;
}

static int pcol_index(Query& q, const std::string& name) {
  for (size_t i = 0; i < q.pcols.size(); i++)
    if (q.pcols[i].name == name) return (int)i;
  LK_CHECK(q.pcols.size() < (size_t)LK_MAX_PCOLS, LK_ERR_UNSUPPORTED, "query touches too many columns");
  PCol p;
  p.name = name;
  q.pcols.push_back(p);
  return (int)q.pcols.size() - 1;
}

static Truth k_and(Truth a, Truth b) { return (a == F || b == F) ? F : (a == N || b == N) ? N : T; }
static Truth k_or(Truth a, Truth b) { return (a == T || b == T) ? T : (a == N || b == N) ? N : F; }
static Truth k_not(Truth a) { return a == N ? N : (a == T ? F : T); }

static Truth eval_tree(const Clause& c, int& leaf_counter, const std::function<Truth(int)>& leaf_truth) {
  switch (c.kind) {
    case Clause::Leaf: return leaf_truth(leaf_counter++);
    case Clause::And: { Truth a = eval_tree(*c.a, leaf_counter, leaf_truth); Truth b = eval_tree(*c.b, leaf_counter, leaf_truth); return k_and(a, b); }
    case Clause::Or: { Truth a = eval_tree(*c.a, leaf_counter, leaf_truth); Truth b = eval_tree(*c.b, leaf_counter, leaf_truth); return k_or(a, b); }
    case Clause::Not: return k_not(eval_tree(*c.a, leaf_counter, leaf_truth));
  }
  return F;
}

static void build_info_json(Query& q);
void refresh_info_json(Query& q) { build_info_json(q); }

namespace {
struct PlanTrace {
  bool on = getenv("LK_PLAN_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[lk plan] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
}  // namespace

void layout_private_arena(Query& q, const std::vector<uint8_t>& placed) {
  uint64_t arena = 0;
  for (size_t k = 0; k < q.slots.size(); k++) {
    if (k < placed.size() && placed[k]) continue;
    const ChunkSlot& sl = q.slots[k];
    arena = (arena + 255) & ~255ull;
    q.rgs[sl.rgi].arena_base[sl.pcol] = arena;
    q.uploads.push_back({sl.seg, sl.file_off, sl.len, arena});
    arena += sl.len;
    if (sl.reserve) arena = ((arena + 7) & ~7ull) + sl.reserve;
  }
  q.arena_bytes = ((arena + 255) & ~255ull) + 256;  // tail padding: lk_load_u64 may read the next aligned word
}

void plan_query(Query& q) {
  PlanTrace trace;
  const BaseExpr& e = q.req.expr;
  const Options& opt = global_options();
  // ---- shape checks: everything that is not the aggregate push-down is outside the GPU path ----
  // Tag query with a tagDataType (BaseExpr.scala:127-143, the branch for a tag that is a column of the files):
  //   SELECT "tag" as "tag", COUNT(*) AS count FROM T WHERE <filter> AND ts >= S AND ts < E GROUP BY "tag"
  // = this engine with the tag as the only key column, one time bucket and COUNT over the timestamp column (a row that
  // passes `ts >= S` has a timestamp, so that count IS the row count).  Without a tagDataType the reference selects whole
  // rows (SELECT *): not an aggregate, outside the GPU path.
  q.tag_query = q.req.is_tag_query;
  LK_CHECK(!q.tag_query || q.req.has_tag_data_type, LK_ERR_UNSUPPORTED, "tag queries without a tagDataType (SELECT *) are outside the GPU path");
  LK_CHECK(q.tag_query || e.has_chart, LK_ERR_UNSUPPORTED, "exemplar queries (no chart options) are outside the GPU path");
  LK_CHECK(!e.has_extract && !e.has_compute, LK_ERR_UNSUPPORTED, "extract / compute sub-queries are outside the GPU path");
  LK_CHECK(!q.req.segments.empty(), LK_ERR_INVALID, "PushDownRequest has no segmentRequests");
  LK_CHECK(q.segs.size() == q.req.segments.size(), LK_ERR_INVALID,
           strf("%zu segment(s) added but the request lists %zu segmentRequests", q.segs.size(), q.req.segments.size()));
  q.is_metrics = e.dataset == "metrics" && !q.tag_query;  // (a tag query has no timestamp grid to check: one bucket)
  q.ts_col_name = q.is_metrics ? TIMESTAMP : "step_ts";
  // Commons.scala:225-232: startTs = min, endTs = max over the glob, step of the HEAD segment
  q.ts_lo = q.req.segments[0].start_ts;
  q.ts_hi = q.req.segments[0].end_ts;
  for (auto& s : q.req.segments) { q.ts_lo = std::min(q.ts_lo, s.start_ts); q.ts_hi = std::max(q.ts_hi, s.end_ts); }
  q.step = q.req.segments[0].step_ms;
  LK_CHECK(q.step > 0, LK_ERR_INVALID, "stepInMillis must be positive");
  LK_CHECK(q.ts_lo >= 0, LK_ERR_UNSUPPORTED, "negative startTs");
  if (q.tag_query) {
    LK_CHECK(q.aggs.empty(), LK_ERR_INVALID, "a tag query takes no aggregates option");
    LK_CHECK(q.req.tag_type.empty() || q.req.tag_type == "string", LK_ERR_UNSUPPORTED, "tag queries on non-string tags");
    AggSpec a;
    a.op = AGG_COUNT;
    a.aggregation = "count";
    a.value_column = TIMESTAMP;
    q.aggs.push_back(a);
    // one bucket over [startTs, endTs): bucket = (ts - base) / step with base = startTs, step = the whole range
    q.step = std::max<int64_t>(1, q.ts_hi - q.ts_lo);
  }
  if (q.aggs.empty()) q.aggs.push_back(resolve_agg(e, e.chart.aggregation, e.chart.has_rollup, e.chart.rollup, true));
  LK_CHECK(q.aggs.size() <= (size_t)LK_MAX_AGGS, LK_ERR_UNSUPPORTED, "too many aggregates in one pass");
  const bool value_not_null = !q.is_metrics && e.chart.has_field_name && e.chart.field_name != VALUE;  // BaseExpr.scala:407-426

  // ---- footers: DESCRIBE SELECT * FROM read_parquet([...], union_by_name=True) (Commons.scala:213-221) ----
  parallel_for((int)q.segs.size(), opt.host_threads, [&](int i) {
    // (files bring their footer along: from the segment cache, or parsed from their tail when they were added)
    if (!q.segs[i].meta_from_cache && !q.segs[i].sparse) q.segs[i].meta = parse_footer(q.segs[i].data, q.segs[i].len);
  });
  auto exists = [&](const std::string& name) {
    for (auto& s : q.segs)
      if (s.meta.leaf_index(name) >= 0) return true;
    return false;
  };
  std::vector<std::string> fs;
  field_set(e, fs);
  std::vector<std::string> non_existent;  // nonExistentFields = fieldSet() - columnsThatExist (Commons.scala:224)
  for (auto& f : fs)
    if (!exists(f)) non_existent.push_back(f);
  auto is_non_existent = [&](const std::string& n) { return std::find(non_existent.begin(), non_existent.end(), n) != non_existent.end(); };

  trace.mark("footers");
  // ---- touched columns ----
  q.pcols.clear();
  q.ts_pcol = pcol_index(q, TIMESTAMP);
  q.pcols[q.ts_pcol].is_ts = true;
  LK_CHECK(exists(TIMESTAMP), LK_ERR_QUERY, std::string("Binder Error: column ") + TIMESTAMP + " not found");
  LK_CHECK(q.tag_query || exists(NAME), LK_ERR_QUERY, std::string("Binder Error: column ") + NAME + " not found");

  // filter leaves
  q.leaves.clear();
  q.leaf_preds.clear();
  q.leaf_filter_col.clear();
  q.fcols.clear();
  collect_leaves(*e.filter, q.leaves);
  for (size_t li = 0; li < q.leaves.size(); li++) {
    const Clause& l = *q.leaves[li];
    LK_CHECK(!l.extracted && !l.computed, LK_ERR_UNSUPPORTED, "filter on an extracted/computed field '" + l.k + "'");
    q.leaf_preds.push_back(compile_leaf(l));
    if (is_non_existent(l.k)) {  // the leaf is the literal `false` (BaseExpr.scala:462-464)
      q.leaf_filter_col.push_back(-1);
      continue;
    }
    LK_CHECK(exists(l.k), LK_ERR_QUERY, "Binder Error: column " + l.k + " not found");  // only reachable under a NotClause
    int p = pcol_index(q, l.k);
    q.pcols[p].is_filter = true;
    if (q.pcols[p].filter_slot < 0) {
      LK_CHECK(q.fcols.size() < (size_t)LK_MAX_FILTER, LK_ERR_UNSUPPORTED, "too many filter columns");
      q.pcols[p].filter_slot = (int)q.fcols.size();
      FilterColPlan fc;
      fc.pcol = p;
      q.fcols.push_back(fc);
    }
    q.fcols[q.pcols[p].filter_slot].leaves.push_back((int)li);
    q.leaf_filter_col.push_back(q.pcols[p].filter_slot);
  }
  // key columns: name, then the group-bys that exist (BaseExpr.scala:338-346), in chart order
  q.key_pcols.clear();
  q.key_names.clear();
  if (q.tag_query) {  // GROUP BY "tag" alone
    LK_CHECK(exists(q.req.tag_name), LK_ERR_QUERY, "Binder Error: column " + q.req.tag_name + " not found");
    int p = pcol_index(q, q.req.tag_name);
    q.pcols[p].is_key = true;
    q.pcols[p].key_slot = 0;
    q.key_pcols.push_back(p);
    q.key_names.push_back(q.req.tag_name);
  } else {
    int p = pcol_index(q, NAME);
    q.pcols[p].is_key = true;
    q.pcols[p].key_slot = 0;
    q.key_pcols.push_back(p);
    q.key_names.push_back("name");
    for (auto& g : e.chart.group_bys) {
      if (is_non_existent(g)) continue;
      LK_CHECK(q.key_pcols.size() < (size_t)LK_MAX_KEYS, LK_ERR_UNSUPPORTED, "too many group-by columns");
      int gp = pcol_index(q, g);
      q.pcols[gp].is_key = true;
      if (q.pcols[gp].key_slot < 0) q.pcols[gp].key_slot = (int)q.key_pcols.size();
      q.key_pcols.push_back(gp);
      q.key_names.push_back(g);
    }
  }
  q.agg_pcols.clear();
  for (auto& a : q.aggs) {
    LK_CHECK(exists(a.value_column), LK_ERR_QUERY, "Binder Error: column " + a.value_column + " not found");
    int p = pcol_index(q, a.value_column);
    q.pcols[p].is_value = true;
    q.agg_pcols.push_back(p);
  }

  // ---- physical types must agree across the glob (union_by_name would otherwise insert casts) ----
  for (auto& pc : q.pcols) {
    pc.phys_type = -1;
    for (auto& s : q.segs) {
      int li = s.meta.leaf_index(pc.name);
      if (li < 0) continue;
      int t = s.meta.leaves[li].phys_type;
      LK_CHECK(pc.phys_type < 0 || pc.phys_type == t, LK_ERR_UNSUPPORTED, "column '" + pc.name + "' has different physical types across segments");
      pc.phys_type = t;
    }
    pc.string_typed = pc.phys_type == PT_BYTE_ARRAY;
    bool numeric = pc.phys_type == PT_INT32 || pc.phys_type == PT_INT64 || pc.phys_type == PT_FLOAT || pc.phys_type == PT_DOUBLE;
    if (pc.is_ts) LK_CHECK(pc.phys_type == PT_INT64, LK_ERR_UNSUPPORTED, "timestamp column must be INT64");
    if (pc.is_key) LK_CHECK(pc.string_typed, LK_ERR_UNSUPPORTED, "group-by column '" + pc.name + "' is not a string column");
    if (pc.is_value) LK_CHECK(numeric, LK_ERR_UNSUPPORTED, "value column '" + pc.name + "' is not numeric");
    if (pc.is_filter) LK_CHECK(numeric || pc.string_typed, LK_ERR_UNSUPPORTED, "filter column '" + pc.name + "' has an unsupported type");
  }
  for (auto& fc : q.fcols) {
    fc.numeric = !q.pcols[fc.pcol].string_typed;
    fc.n_num = 0;
    fc.num_leaf_slot.clear();
    for (int li : fc.leaves) {
      const LeafPredicate& lp = q.leaf_preds[li];
      if (fc.numeric) {
        LK_CHECK(lp.op == LeafPredicate::Exists || lp.is_numeric_op(), LK_ERR_UNSUPPORTED,
                 "string operator '" + q.leaves[li]->op + "' on numeric column '" + q.pcols[fc.pcol].name + "'");
        if (lp.is_numeric_op()) {
          LK_CHECK(fc.n_num < LK_MAX_NUMLEAF, LK_ERR_UNSUPPORTED, "too many comparisons on column '" + q.pcols[fc.pcol].name + "'");
          fc.num_leaf_slot.push_back(fc.n_num++);
        } else fc.num_leaf_slot.push_back(-1);
      } else {
        LK_CHECK(!lp.is_numeric_op(), LK_ERR_UNSUPPORTED,
                 "numeric operator '" + q.leaves[li]->op + "' on string column '" + q.pcols[fc.pcol].name + "'");
      }
    }
  }

  // ---- index every touched column chunk (page headers, run headers, dictionaries) ----
  q.rgs.clear();
  for (size_t s = 0; s < q.segs.size(); s++)
    for (size_t g = 0; g < q.segs[s].meta.row_groups.size(); g++) {
      if (q.segs[s].meta.row_groups[g].num_rows == 0) continue;
      RowGroupPlan rp;
      rp.seg = (int)s;
      rp.rg = (int)g;
      rp.num_rows = (uint32_t)q.segs[s].meta.row_groups[g].num_rows;
      q.rgs.push_back(std::move(rp));
    }
  const int np = (int)q.pcols.size();
  // ---- arena layout + upload list: needs the footers only, so the H2D copies of the column chunks can start NOW
  //      (q.on_layout) and overlap the page/run indexing below ----
  q.uploads.clear();
  q.slots.clear();
  for (size_t i = 0; i < q.rgs.size(); i++) {
    RowGroupPlan& rp = q.rgs[i];
    const SegmentInput& seg = q.segs[rp.seg];
    rp.arena_base.assign(np, 0);
    rp.from_cache.assign(np, 0);
    rp.chunks.clear();
    rp.chunks.resize(np);
    for (int p = 0; p < np; p++) {
      int li = seg.meta.leaf_index(q.pcols[p].name);
      if (li < 0) continue;
      const ColumnChunkMeta& cm = seg.meta.row_groups[rp.rg].columns[li];
      ChunkSlot sl;
      sl.rgi = (int)i; sl.pcol = p; sl.seg = rp.seg; sl.leaf = li;
      chunk_byte_range(cm, seg.len, q.pcols[p].name, sl.file_off, sl.len);
      // room behind a string chunk for its PLAIN pages re-encoded as dictionary indices (none when the footer rules them out)
      sl.reserve = synth_reserve(cm);
      q.slots.push_back(sl);
    }
  }
  trace.mark("arena layout");
  for (auto& sl : q.slots)
    if (q.segs[sl.seg].data) sl.host = q.segs[sl.seg].data + sl.file_off;
  // the device layer places the chunks (segment cache, private arena), reads what sparse files still owe and starts the
  // copies; without one: all private
  if (q.on_layout) q.on_layout();
  else layout_private_arena(q, std::vector<uint8_t>());
  std::vector<int> slot_of(q.rgs.size() * (size_t)np, -1);
  for (size_t k = 0; k < q.slots.size(); k++) slot_of[(size_t)q.slots[k].rgi * np + q.slots[k].pcol] = (int)k;
  parallel_for((int)q.rgs.size(), opt.host_threads, [&](int i) {
    RowGroupPlan& rp = q.rgs[i];
    const SegmentInput& seg = q.segs[rp.seg];
    for (int p = 0; p < np; p++) {
      int li = seg.meta.leaf_index(q.pcols[p].name);
      if (li < 0) continue;  // union_by_name: the column is NULL for this file
      if (rp.from_cache[p]) continue;  // index (and bytes) came with the cached column
      const PCol& pc = q.pcols[p];
      // the chunk's bytes in host memory, addressed by FILE offset like everything index_chunk reads (only offsets inside
      // the chunk are touched, so a view that holds just this chunk serves as well as the whole file)
      const ChunkSlot& sl = q.slots[slot_of[(size_t)i * np + p]];
      LK_CHECK(sl.host != nullptr, LK_ERR_IO, "segment '" + seg.name + "': column chunk bytes were not read");
      const uint8_t* file_view = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(sl.host) - (uintptr_t)sl.file_off);
      rp.chunks[p] = index_chunk(file_view, seg.len, seg.meta.leaves[li], seg.meta.row_groups[rp.rg].columns[li],
                                 seg.meta.row_groups[rp.rg].num_rows, pc.string_typed, !q.device_index);
      if (pc.string_typed)
        for (auto& pg : rp.chunks[p].pages)
          LK_CHECK(pg.dict_coded || pg.nvals == 0, LK_ERR_UNSUPPORTED, "string column '" + pc.name + "' has PLAIN (non-dictionary) pages");
    }
  });
  for (auto& rp : q.rgs)
    for (int p = 0; p < np; p++) {
      const ChunkIndex& ci = rp.chunks[p];
      if (!ci.present || ci.synth.empty() || rp.from_cache[p]) continue;
      const SegmentInput& seg = q.segs[rp.seg];
      const ColumnChunkMeta& cm = seg.meta.row_groups[rp.rg].columns[seg.meta.leaf_index(q.pcols[p].name)];
      LK_CHECK(ci.synth.size() <= synth_reserve(cm), LK_ERR_UNSUPPORTED, "column '" + q.pcols[p].name + "': PLAIN string pages need more room than the footer announced");
      Query::Upload u{rp.seg, 0, ci.synth.size(), rp.arena_base[p] + ci.synth_base};
      u.src = ci.synth.data();
      q.uploads.push_back(u);
    }
  q.total_rows = 0;
  q.touched_bytes = 0;
  for (auto& rp : q.rgs) {
    rp.seq_base = (uint64_t)q.total_rows;
    q.total_rows += rp.num_rows;
    for (auto& c : rp.chunks)
      if (c.present) q.touched_bytes += c.total_compressed_size;
  }

  trace.mark("index chunks");
  uint32_t def_mask_out = 0;
  // ---- runs pool, tiles, cursors ----
  const uint32_t tile_rows = std::max(64u, std::min(opt.tile_rows, (uint32_t)LK_TILE_ROWS_MAX)) & ~31u;
  std::vector<size_t> run_base(q.rgs.size() + 1, 0), tile_base(q.rgs.size() + 1, 0);
  std::vector<std::vector<uint32_t>> bounds(q.rgs.size());
  std::vector<size_t> nruns_of(q.rgs.size(), 0);
  parallel_for((int)q.rgs.size(), opt.host_threads, [&](int i) {
    const RowGroupPlan& rp = q.rgs[i];
    size_t nr = 0;
    std::vector<uint32_t>& b = bounds[i];
    // tile boundaries: every tile_rows rows and every page boundary of every touched column (both lists ascending: merged)
    std::vector<uint32_t> pg_rows;
    for (auto& c : rp.chunks) {
      nr += c.def_runs.size() + c.val_runs.size();
      for (auto& pg : c.pages) pg_rows.push_back(pg.first_row);
    }
    pg_rows.push_back(rp.num_rows);
    std::sort(pg_rows.begin(), pg_rows.end());
    b.reserve(rp.num_rows / tile_rows + pg_rows.size() + 2);
    size_t k = 0;
    for (uint32_t r = 0; r < rp.num_rows; r += tile_rows) {
      while (k < pg_rows.size() && pg_rows[k] < r) { if (b.empty() || b.back() != pg_rows[k]) b.push_back(pg_rows[k]); k++; }
      if (b.empty() || b.back() != r) b.push_back(r);
    }
    for (; k < pg_rows.size(); k++) if (b.empty() || b.back() != pg_rows[k]) b.push_back(pg_rows[k]);
    nruns_of[i] = nr;
  });
  for (size_t i = 0; i < q.rgs.size(); i++) {
    run_base[i + 1] = run_base[i] + nruns_of[i];
    tile_base[i + 1] = tile_base[i] + (bounds[i].size() - 1);
  }
  LK_CHECK(run_base.back() < 0xffffffffull && tile_base.back() * np < 0xffffffffull, LK_ERR_UNSUPPORTED, "glob too large for 32-bit index pools");
  q.tiles.resize_uninit(tile_base.back());
  q.chunk_infos.assign(q.rgs.size() * np, ChunkInfo{});
  if (q.device_index) {
    // the device walks the run headers and computes the cursors (lk_engine.cu: idx_* kernels): hand it the pages and chunks
    q.idx_pages.clear();
    q.zpages.clear();
    q.idx_chunks.assign(q.rgs.size() * np, IdxChunk{});
    parallel_for((int)q.rgs.size(), opt.host_threads, [&](int i) {
      const std::vector<uint32_t>& b = bounds[i];
      for (size_t t = 0; t + 1 < b.size(); t++) {
        TileDesc& td = q.tiles[tile_base[i] + t];
        td.row0 = b[t];
        td.nrows = b[t + 1] - b[t];
        td.rg = (uint32_t)i;
        td.cursor0 = (uint32_t)((tile_base[i] + t) * np);
      }
    });
    for (size_t i = 0; i < q.rgs.size(); i++) {
      const RowGroupPlan& rp = q.rgs[i];
      for (int p = 0; p < np; p++) {
        const ChunkIndex& ci = rp.chunks[p];
        ChunkInfo& info = q.chunk_infos[i * np + p];
        info.phys_type = (uint32_t)std::max(0, q.pcols[p].phys_type);
        info.seq_base = rp.seq_base;
        IdxChunk& ic = q.idx_chunks[i * np + p];
        ic.present = ci.present;
        ic.page0 = (uint32_t)q.idx_pages.size();
        if (!ci.present) continue;
        auto rebase = [&](uint64_t foff) { return rp.arena_base[p] + (foff - ci.file_start); };
        info.dict_off = ci.has_dict ? rebase(ci.dict_off) : 0;
        info.dict_n = ci.dict_n;
        info.base_off = rp.arena_base[p];
        ic.base_off = rp.arena_base[p];
        ic.max_def = (uint32_t)ci.max_def;
        ic.esz = (ci.phys_type == PT_INT32 || ci.phys_type == PT_FLOAT) ? 4 : 8;
        ic.num_rows = ci.num_rows;
        ic.dict_n = ci.dict_n;
        ic.string_typed = q.pcols[p].string_typed;
        const uint32_t chunk_page0 = (uint32_t)q.idx_pages.size();
        for (auto& zp : ci.zpages) {
          ZPage z;
          z.src_off = rebase(zp.src_off);
          z.dst_off = rebase(zp.dst_virt);
          z.src_len = zp.src_len;
          z.dst_len = zp.dst_len;
          z.page = zp.page < 0 ? 0xffffffffu : chunk_page0 + (uint32_t)zp.page;
          z.flags = rp.from_cache[p] ? (zp.flags & ~(uint32_t)ZP_DECODE) : zp.flags;  // a cached block holds the inflated bytes already
          q.zpages.push_back(z);
        }
        for (auto& pg : ci.pages) {
          IdxPage ip;
          memset(&ip, 0, sizeof ip);
          ip.def_off = pg.def_end > pg.def_off ? rebase(pg.def_off) : 0;
          ip.def_end = pg.def_end > pg.def_off ? rebase(pg.def_end) : 0;
          ip.val_off = pg.synth ? rp.arena_base[p] + ci.synth_base + pg.values_off : rebase(pg.values_off);
          ip.val_end = ip.val_off + pg.values_len;
          ip.first_row = pg.first_row;
          ip.num_rows = pg.num_rows;
          ip.chunk = (uint32_t)(i * np + p);
          ip.dict_coded = pg.dict_coded;
          ip.bit_width = pg.bit_width;
          q.idx_pages.push_back(ip);
        }
        ic.npages = (uint32_t)q.idx_pages.size() - ic.page0;
      }
    }
    LK_CHECK(q.idx_pages.size() < 0x7fffffffull, LK_ERR_UNSUPPORTED, "glob has too many pages");
    q.def_chunks.clear();
    q.defbm_words = 0;
    q.def_blocks_total = 0;
    def_mask_out = 0;
  } else {
  q.runs.resize_uninit(run_base.back());
  q.cursors.resize_uninit(tile_base.back() * np);  // zeroed slice by slice by the workers below
  uint32_t def_mask = 0;
  std::vector<uint32_t> def_masks(q.rgs.size(), 0);
  std::vector<DefChunk> def_tmp(q.rgs.size() * np, DefChunk{});  // run_n > 0 <=> some tile of the chunk mixes NULLs and values
  parallel_for((int)q.rgs.size(), opt.host_threads, [&](int i) {
    const RowGroupPlan& rp = q.rgs[i];
    const uint8_t* file = q.segs[rp.seg].data;
    size_t rpos = run_base[i];
    std::vector<uint32_t> dbase(np, 0), vbase(np, 0);
    for (int p = 0; p < np; p++) {
      const ChunkIndex& ci = rp.chunks[p];
      ChunkInfo& info = q.chunk_infos[(size_t)i * np + p];
      info.phys_type = (uint32_t)std::max(0, q.pcols[p].phys_type);
      info.seq_base = rp.seq_base;
      if (!ci.present) continue;
      auto rebase = [&](uint64_t foff) { return rp.arena_base[p] + (foff - ci.file_start); };
      info.dict_off = ci.has_dict ? rebase(ci.dict_off) : 0;
      info.dict_n = ci.dict_n;
      dbase[p] = (uint32_t)rpos;
      info.base_off = rp.arena_base[p];
      for (auto& r : ci.def_runs) q.runs[rpos++] = r;
      vbase[p] = (uint32_t)rpos;
      for (auto& r : ci.val_runs) q.runs[rpos++] = r;
    }
    // cursors: one linear sweep per column over the (ascending) tile boundaries -- no per-tile binary searches
    const std::vector<uint32_t>& b = bounds[i];
    const size_t nt = b.size() - 1;
    if (nt) memset(&q.cursors[tile_base[i] * np], 0, nt * np * sizeof(ColCursor));
    for (size_t t = 0; t < nt; t++) {
      const size_t ti = tile_base[i] + t;
      TileDesc& td = q.tiles[ti];
      td.row0 = b[t];
      td.nrows = b[t + 1] - b[t];
      td.rg = (uint32_t)i;
      td.cursor0 = (uint32_t)(ti * np);
    }
    std::vector<uint32_t> vb(b.size()), db(b.size());  // value index at / def run holding every boundary row
    for (int p = 0; p < np; p++) {
      const ChunkIndex& ci = rp.chunks[p];
      if (!ci.present) {
        for (size_t t = 0; t < nt; t++) q.cursors[(tile_base[i] + t) * np + p].flags = CUR_ALL_NULL;
        continue;
      }
      const size_t nd = ci.def_runs.size(), nv = ci.val_runs.size(), npg = ci.pages.size();
      size_t di = 0;
      for (size_t k = 0; k < b.size(); k++) {
        const uint32_t r = b[k];
        if (ci.max_def == 0 || nd == 0) { vb[k] = ci.max_def == 0 ? r : 0; db[k] = 0; continue; }
        while (di + 1 < nd && ci.def_runs[di + 1].start <= r) di++;
        vb[k] = ci.vidx_in_run(file, (int)di, r);
        db[k] = (uint32_t)di;
      }
      size_t vi = 0, pi = 0;
      const unsigned esz = (ci.phys_type == PT_INT32 || ci.phys_type == PT_FLOAT) ? 4 : 8;
      for (size_t t = 0; t < nt; t++) {
        const uint32_t r0 = b[t], r1 = b[t + 1], v0 = vb[t], v1 = vb[t + 1];
        ColCursor& c = q.cursors[(tile_base[i] + t) * np + p];
        c.vidx0 = v0;
        c.nvals = v1 - v0;
        if (c.nvals == r1 - r0) c.flags |= CUR_ALL_VALID;
        else if (c.nvals == 0) c.flags |= CUR_ALL_NULL;
        else {
          const uint32_t d0 = db[t];
          uint32_t d1 = db[t + 1];
          if (ci.def_runs[d1].start >= r1) d1--;  // the run holding row r1 - 1
          c.drun_lo = dbase[p] + d0;
          c.drun_n = (uint16_t)(d1 - d0 + 1);
          def_masks[i] |= 1u << p;
          DefChunk& dc = def_tmp[(size_t)i * np + p];
          dc.base_off = rp.arena_base[p];
          dc.run_lo = dbase[p];
          dc.run_n = (uint32_t)nd;
          dc.num_rows = rp.num_rows;
        }
        if (c.nvals > 0) {
          while (pi + 1 < npg && ci.pages[pi + 1].first_row <= r0) pi++;
          const PageInfo& pg = ci.pages[pi];
          if (pg.dict_coded) {
            c.flags |= CUR_DICT;
            c.width = pg.bit_width;
            while (vi + 1 < nv && ci.val_runs[vi + 1].start <= v0) vi++;
            size_t zi = vi;
            while (zi + 1 < nv && ci.val_runs[zi + 1].start <= v1 - 1) zi++;
            c.vrun_lo = vbase[p] + (uint32_t)vi;
            c.vrun_n = (uint16_t)(zi - vi + 1);
          } else {
            c.plain_off = rp.arena_base[p] + (pg.values_off - ci.file_start) + (uint64_t)(v0 - pg.first_vidx) * esz;
          }
        }
      }
    }
    // the per-chunk run lists are in the pool now: give their storage back here, on the worker and under the H2D copies,
    // instead of in the query's destructor (250 MB of large blocks: 8 ms of munmap on the caller's thread)
    for (auto& c : q.rgs[i].chunks) {
      std::vector<Run>().swap(c.def_runs);
      std::vector<Run>().swap(c.val_runs);
      std::vector<uint32_t>().swap(c.def_nn_before);
    }
  });
  for (auto m : def_masks) def_mask |= m;
  layout_def_chunks(q, def_tmp);
  def_mask_out = def_mask;
  q.n_runs = q.runs.size();
  }
  trace.mark("tiles/cursors/runs");
  // ---- predicate: per-column classes over dictionary entries, then the pass bitmap over class combinations ----
  q.lut_cls.clear();
  for (size_t f = 0; f < q.fcols.size(); f++) {
    FilterColPlan& fc = q.fcols[f];
    fc.sig2cls.clear();
    fc.cls_sig.clear();
    if (fc.numeric) {
      fc.ncls = (1u << fc.n_num) + 1;
      for (uint32_t m = 0; m < fc.ncls; m++) {
        std::vector<uint8_t> sig;
        bool is_null = m == (1u << fc.n_num);
        for (size_t j = 0; j < fc.leaves.size(); j++) {
          int slot = fc.num_leaf_slot[j];
          if (slot < 0) sig.push_back(is_null ? F : T);
          else sig.push_back(is_null ? N : ((m >> slot) & 1) ? T : F);
        }
        fc.cls_sig.push_back(sig);
      }
      continue;
    }
    std::vector<uint8_t> nullsig;
    for (int li : fc.leaves) nullsig.push_back(q.leaf_preds[li].eval_string(nullptr));
    fc.sig2cls[nullsig] = 0;
    fc.cls_sig.push_back(nullsig);
    std::unordered_map<std::string, uint32_t> cache;
    for (size_t i = 0; i < q.rgs.size(); i++) {
      const ChunkIndex& ci = q.rgs[i].chunks[fc.pcol];
      if (!ci.present) continue;
      q.chunk_infos[i * np + fc.pcol].lut_cls = (uint32_t)q.lut_cls.size();
      for (auto& s : ci.dict_strings) {
        auto it = cache.find(s);
        uint32_t cls;
        if (it != cache.end()) cls = it->second;
        else {
          std::vector<uint8_t> sig;
          for (int li : fc.leaves) sig.push_back(q.leaf_preds[li].eval_string(&s));
          auto jt = fc.sig2cls.find(sig);
          if (jt == fc.sig2cls.end()) {
            cls = (uint32_t)fc.cls_sig.size();
            fc.sig2cls[sig] = cls;
            fc.cls_sig.push_back(sig);
          } else cls = jt->second;
          cache.emplace(s, cls);
        }
        LK_CHECK(cls < 256, LK_ERR_UNSUPPORTED, "too many predicate classes on one column");
        q.lut_cls.push_back((uint8_t)cls);
      }
    }
    fc.ncls = (uint32_t)fc.cls_sig.size();
  }
  uint64_t combos = 1;
  std::vector<uint32_t> fstride(q.fcols.size(), 1);
  for (size_t f = 0; f < q.fcols.size(); f++) {
    fstride[f] = (uint32_t)combos;
    combos *= q.fcols[f].ncls;
    LK_CHECK(combos <= (1ull << 22), LK_ERR_UNSUPPORTED, "predicate has too many class combinations");
  }
  q.pass_bits.assign((combos + 31) / 32, 0);
  {
    std::vector<int> pos_in_col(q.leaves.size(), 0);
    for (auto& fc : q.fcols)
      for (size_t j = 0; j < fc.leaves.size(); j++) pos_in_col[fc.leaves[j]] = (int)j;
    std::vector<uint32_t> cls(q.fcols.size(), 0);
    for (uint64_t c = 0; c < combos; c++) {
      uint64_t r = c;
      for (size_t f = 0; f < q.fcols.size(); f++) { cls[f] = (uint32_t)(r % q.fcols[f].ncls); r /= q.fcols[f].ncls; }
      int counter = 0;
      Truth t = eval_tree(*e.filter, counter, [&](int li) -> Truth {
        int f = q.leaf_filter_col[li];
        if (f < 0) return F;
        return (Truth)q.fcols[f].cls_sig[cls[f]][pos_in_col[li]];
      });
      if (t == T) q.pass_bits[c >> 5] |= 1u << (c & 31);
    }
  }
  // Rough selectivity of the WHERE clause (only used to choose between the hash table and the record path): dictionary
  // entries taken as equally likely, columns as independent, NULLs ignored; unknown (1.0) with numeric comparisons.
  q.est_selectivity = 1.0;
  {
    bool ok = combos <= 65536;
    for (auto& fc : q.fcols) ok = ok && !fc.numeric;
    if (ok) {
      std::vector<std::vector<double>> pc(q.fcols.size());
      for (size_t f = 0; f < q.fcols.size(); f++) {
        const FilterColPlan& fc = q.fcols[f];
        pc[f].assign(fc.ncls, 0.0);
        double tot = 0;
        for (size_t i = 0; i < q.rgs.size(); i++) {
          const ChunkInfo& ci = q.chunk_infos[i * np + fc.pcol];
          for (uint32_t code = 0; code < ci.dict_n && ci.lut_cls + code < q.lut_cls.size(); code++) { pc[f][q.lut_cls[ci.lut_cls + code]] += 1; tot += 1; }
        }
        if (tot > 0) for (auto& x : pc[f]) x /= tot;
        else pc[f][0] = 1.0;  // no dictionary at all: the column is absent / all NULL
      }
      double sel = 0;
      for (uint64_t c = 0; c < combos; c++) {
        if (!((q.pass_bits[c >> 5] >> (c & 31)) & 1)) continue;
        double pr = 1;
        uint64_t r = c;
        for (size_t f = 0; f < q.fcols.size(); f++) { pr *= pc[f][r % q.fcols[f].ncls]; r /= q.fcols[f].ncls; }
        sel += pr;
      }
      q.est_selectivity = q.fcols.empty() ? 1.0 : sel;
    }
  }

  trace.mark("predicate tables");
  // ---- ScanParams (device pointers are filled in by the device layer) ----
  ScanParams& P = q.params;
  memset(&P, 0, sizeof P);
  P.ntiles = (uint32_t)q.tiles.size();
  P.npcols = (uint32_t)np;
  P.ts_pcol = q.ts_pcol;
  P.n_filter = (int)q.fcols.size();
  for (size_t f = 0; f < q.fcols.size(); f++) {
    const FilterColPlan& fc = q.fcols[f];
    FilterCol& d = P.filter[f];
    d.pcol = (uint8_t)fc.pcol;
    d.numeric = fc.numeric;
    d.n_leaves = (uint8_t)fc.n_num;
    d.null_cls = fc.numeric ? (uint8_t)(1u << fc.n_num) : 0;
    d.stride = fstride[f];
    if (fc.numeric)
      for (size_t j = 0; j < fc.leaves.size(); j++) {
        int slot = fc.num_leaf_slot[j];
        if (slot < 0) continue;
        const LeafPredicate& lp = q.leaf_preds[fc.leaves[j]];
        d.ops[slot] = (uint8_t)(lp.op - LeafPredicate::Gt);
        d.consts[slot] = lp.number;
      }
  }
  P.n_aggs = (int)q.aggs.size();
  for (size_t a = 0; a < q.aggs.size(); a++) { P.aggs[a].op = q.aggs[a].op; P.aggs[a].pcol = (uint8_t)q.agg_pcols[a]; }
  // chart-field filter `field$type IS NOT NULL`: expressed as one more conjunct on the value column being non-null
  P.def_mask = def_mask_out;
  P.ts_lo = q.ts_lo;
  P.ts_hi = q.ts_hi;
  P.step = q.step;
  P.is_metrics = q.is_metrics;
  if (q.ts_hi <= q.ts_lo) { q.nbuckets = 0; q.base = q.ts_lo; }
  else if (q.is_metrics || q.tag_query) {
    q.base = q.ts_lo;
    uint64_t nb = ((uint64_t)(q.ts_hi - q.ts_lo) + q.step - 1) / q.step;
    LK_CHECK(nb < (1ull << 31), LK_ERR_UNSUPPORTED, "too many time buckets");
    q.nbuckets = (uint32_t)nb;
  } else {
    q.base = q.ts_lo - q.ts_lo % q.step;  // step_ts = ts - ts % step (BaseExpr.scala:163-165)
    uint64_t nb = ((uint64_t)(q.ts_hi - q.base) + q.step - 1) / q.step;
    LK_CHECK(nb < (1ull << 31), LK_ERR_UNSUPPORTED, "too many time buckets");
    q.nbuckets = (uint32_t)nb;
  }
  P.base = q.base;
  P.fits32 = q.ts_hi > q.base && (uint64_t)(q.ts_hi - q.base) < (1ull << 32) && (uint64_t)q.step < (1ull << 32);
  P.nbuckets = q.nbuckets;
  P.notnull_pcol = value_not_null ? q.agg_pcols[0] : -1;
  q.params.survivors = nullptr;

  trace.mark("scan params");
  // ---- group-by: global dictionaries + remap tables ----
  q.local_dicts.assign(q.key_pcols.size(), {});
  for (size_t k = 0; k < q.key_pcols.size(); k++) {
    std::vector<std::string>& d = q.local_dicts[k];
    // the same few dozen strings come back from every row group: dedupe first (hash set of views into the chunk indexes),
    // sort only the distinct ones (100 segments x 4 key columns: 14 k strings, 144 distinct)
    std::unordered_set<std::string_view> seen;
    for (auto& rp : q.rgs) {
      const ChunkIndex& ci = rp.chunks[q.key_pcols[k]];
      for (auto& s : ci.dict_strings)
        if (seen.insert(std::string_view(s)).second) d.emplace_back(s);
    }
    std::sort(d.begin(), d.end());
  }
  q.key_dicts = q.local_dicts;
  trace.mark("local dictionaries");
  rebuild_group_tables(q);
  trace.mark("group tables");
  q.prepared = true;
}

void layout_def_chunks(Query& q, const std::vector<DefChunk>& def_tmp) {
  // definition bitmaps: one bit per row for every chunk that has a tile mixing NULLs and values; the device expands
  // the chunk's hybrid RLE/bit-packed definition levels into it (def_expand_kernel) and the scan reads 16 bits per lane
  q.def_chunks.clear();
  q.defbm_words = 0;
  q.def_blocks_total = 0;
  for (size_t k = 0; k < def_tmp.size(); k++) {
    DefChunk dc = def_tmp[k];
    if (!dc.run_n) continue;
    dc.word0 = (uint32_t)q.defbm_words;
    dc.cum = (uint32_t)q.def_blocks_total;
    q.chunk_infos[k].defbm_word0 = dc.word0;
    q.defbm_words += (dc.num_rows + 31) / 32 + 2;  // + 2: the scan's funnel-shifted reads touch one word past the last
    q.def_blocks_total += (dc.run_n + LK_DEF_BLOCK_RUNS - 1) / LK_DEF_BLOCK_RUNS;
    q.def_chunks.push_back(dc);
  }
  // lanes beyond a short last tile still read "their" 16 bits (and mask them away): up to 512 rows past the chunk's end
  if (q.defbm_words) q.defbm_words += 32;
  LK_CHECK(q.defbm_words < 0xffffffffull && q.def_blocks_total < 0x7fffffffull, LK_ERR_UNSUPPORTED, "glob too large for 32-bit definition bitmaps");
}

void rebuild_group_tables(Query& q) {
  const Options& opt = global_options();
  const int np = (int)q.pcols.size();
  ScanParams& P = q.params;
  q.lut_gcode.clear();
  // one table per (row group, distinct key pcol); a column listed twice in groupBys shares it
  std::vector<int> done_pcol;
  for (size_t k = 0; k < q.key_pcols.size(); k++) {
    int p = q.key_pcols[k];
    if (std::find(done_pcol.begin(), done_pcol.end(), p) != done_pcol.end()) {
      // duplicate group-by column: reuse the dictionary of its first occurrence
      for (size_t k2 = 0; k2 < k; k2++)
        if (q.key_pcols[k2] == p) { q.key_dicts[k] = q.key_dicts[k2]; break; }
      continue;
    }
    done_pcol.push_back(p);
    std::unordered_map<std::string, uint32_t> idx;
    idx.reserve(q.key_dicts[k].size() * 2);
    for (size_t i = 0; i < q.key_dicts[k].size(); i++) idx.emplace(q.key_dicts[k][i], (uint32_t)i);
    for (size_t i = 0; i < q.rgs.size(); i++) {
      const ChunkIndex& ci = q.rgs[i].chunks[p];
      if (!ci.present) continue;
      q.chunk_infos[i * np + p].lut_gcode = (uint32_t)q.lut_gcode.size();
      for (auto& s : ci.dict_strings) {
        auto it = idx.find(s);
        LK_CHECK(it != idx.end(), LK_ERR_INVALID, "imported dictionary for '" + q.key_names[k] + "' lacks value '" + s + "'");
        q.lut_gcode.push_back(it->second);
      }
    }
  }
  // group id = mixed radix over (dictionary size + 1 NULL code) of every key column
  P.n_keys = (int)q.key_pcols.size();
  unsigned __int128 g = 1;
  for (int k = P.n_keys - 1; k >= 0; k--) {
    P.keys[k].pcol = (uint8_t)q.key_pcols[k];
    P.keys[k].null_code = (uint32_t)q.key_dicts[k].size();
    P.keys[k].stride = (uint64_t)g;
    g *= (unsigned __int128)q.key_dicts[k].size() + 1;
    LK_CHECK(g < ((unsigned __int128)1 << 62), LK_ERR_UNSUPPORTED, "group space exceeds 2^62 combinations");
  }
  q.n_groups = (uint64_t)g;
  P.n_groups = q.n_groups;
  unsigned __int128 cells = g * q.nbuckets;
  LK_CHECK(cells < ((unsigned __int128)1 << 63), LK_ERR_UNSUPPORTED, "group x bucket space exceeds 2^63 cells");
  q.n_cells = (uint64_t)cells;
  bool dense;
  if (q.path_opt == "dense") dense = true;
  else if (q.path_opt == "hash") dense = false;
  else if (q.path_opt == "records") dense = false;
  else dense = q.n_cells <= opt.dense_max_cells && (q.n_cells <= (1ull << 20) || q.n_cells <= 8ull * (uint64_t)std::max<int64_t>(q.total_rows, 1));
  if (dense) LK_CHECK(q.n_cells <= (1ull << 31), LK_ERR_UNSUPPORTED, "dense table too large; use path=hash");
  q.path = dense ? 0 : 1;
  // Record path instead of the hash table whenever the group space is too large for dense planes: the survivors are
  // appended as records and aggregated by a sort in finalize (no table to probe, nothing to clear).  Measured on B200:
  // C2 (1/16 of the rows survive) 2.16 -> 1.55 ms per step, C4 (47 % survive, 50 M records) 13.6 -> 7.9 ms.  Only while one
  // record per row fits comfortably in HBM.
  // the record key packs (bucket, group id) into 64 bits
  uint32_t gid_bits = 1, bkt_bits = 1;
  while (gid_bits < 63 && (std::max<uint64_t>(q.n_groups, 1) - 1) >> gid_bits) gid_bits++;
  while (bkt_bits < 31 && (std::max<uint32_t>(q.nbuckets, 1) - 1) >> bkt_bits) bkt_bits++;
  P.rec_idx_bits = 0;
  P.rec_gid_bits = gid_bits;
  const bool records_fit = (uint64_t)std::max<int64_t>(q.total_rows, 1) * 8 * (1 + q.aggs.size()) <= (16ull << 30) && q.total_rows < (1ll << 31) &&
                           gid_bits + bkt_bits <= 63;  // (the all-ones key marks a record folded into its owner)
  // (exact_sums: every record also carries its sequence number, one more word per row)
  if (!dense && records_fit && (!q.exact_sums || q.aggs.size() < (size_t)LK_MAX_AGGS) && (q.path_opt == "records" || q.path_opt == "auto")) q.path = 2;
  P.path = q.path;
  // few cells => many rows per cell => warp-level pre-reduction pays
  P.warp_agg = dense && q.n_groups <= 4096;
  { const char* e = getenv("LK_SCAN_STOP_AFTER"); P.stop_after = e ? atoi(e) : 0; }
  q.hash_stride = q.aggs.size() <= 3 ? 32 : 64;
  uint64_t want = 2 * std::min<uint64_t>((uint64_t)std::max<int64_t>(q.total_rows, 1), std::max<uint64_t>(q.n_cells, 1));
  uint64_t slots = 1024;
  while (slots < want && slots < opt.max_hash_slots) slots <<= 1;
  q.hash_slots = slots;
  P.h_stride = q.hash_stride;
  P.h_mask = slots - 1;
  build_info_json(q);
}

static void build_info_json(Query& q) {
  std::string s = "{";
  s += strf("\"path\":\"%s\",\"tiles\":%zu,\"row_groups\":%zu,\"pcols\":%zu,\"total_rows\":%lld,\"touched_bytes\":%lld,", q.path == 0 ? "dense" : q.path == 1 ? "hash" : "records",
            q.tiles.size(), q.rgs.size(), q.pcols.size(), (long long)q.total_rows, (long long)q.touched_bytes);
  s += strf("\"n_groups\":%llu,\"n_buckets\":%u,\"n_cells\":%llu,\"hash_slots\":%llu,\"hash_stride\":%u,\"warp_agg\":%d,", (unsigned long long)q.n_groups,
            q.nbuckets, (unsigned long long)q.n_cells, (unsigned long long)q.hash_slots, q.hash_stride, q.params.warp_agg);
  s += strf("\"ts_lo\":%lld,\"ts_hi\":%lld,\"step\":%lld,\"base\":%lld,\"is_metrics\":%d,\"arena_bytes\":%llu,\"runs\":%zu,\"def_chunks\":%zu,\"def_bitmap_bytes\":%llu,", (long long)q.ts_lo,
            (long long)q.ts_hi, (long long)q.step, (long long)q.base, (int)q.is_metrics, (unsigned long long)q.arena_bytes, (size_t)q.n_runs,
            q.def_chunks.size(), (unsigned long long)q.defbm_words * 4);
  s += "\"columns\":[";
  for (size_t i = 0; i < q.pcols.size(); i++) {
    if (i) s += ",";
    json_escape(s, q.pcols[i].name);
  }
  s += "],\"keys\":[";
  for (size_t i = 0; i < q.key_names.size(); i++) {
    if (i) s += ",";
    s += "{\"name\":";
    json_escape(s, q.key_names[i]);
    s += strf(",\"dict\":%zu}", q.key_dicts[i].size());
  }
  s += "],\"filter_classes\":[";
  for (size_t i = 0; i < q.fcols.size(); i++) s += strf("%s%u", i ? "," : "", q.fcols[i].ncls);
  s += "],\"aggregates\":[";
  for (size_t i = 0; i < q.aggs.size(); i++) {
    if (i) s += ",";
    s += "{\"aggregation\":";
    json_escape(s, q.aggs[i].aggregation);
    s += ",\"column\":";
    json_escape(s, q.aggs[i].value_column);
    s += "}";
  }
  s += "]}";
  q.info_json = s;
}

// blob: u32 n_keys; per key: u32 n; per string: u32 len, bytes
std::string export_dictionaries(const Query& q) {
  std::string b;
  auto put = [&](uint32_t v) { b.append((const char*)&v, 4); };
  put((uint32_t)q.local_dicts.size());
  for (auto& d : q.local_dicts) {
    put((uint32_t)d.size());
    for (auto& s : d) { put((uint32_t)s.size()); b.append(s); }
  }
  return b;
}

void import_dictionaries(Query& q, const uint8_t* blob, size_t len) {
  size_t p = 0;
  auto get = [&]() {
    LK_CHECK(p + 4 <= len, LK_ERR_INVALID, "dictionary blob truncated");
    uint32_t v;
    memcpy(&v, blob + p, 4);
    p += 4;
    return v;
  };
  uint32_t nk = get();
  LK_CHECK(nk == q.key_pcols.size(), LK_ERR_INVALID, "dictionary blob has a different number of key columns");
  std::vector<std::vector<std::string>> dicts(nk);
  for (uint32_t k = 0; k < nk; k++) {
    uint32_t n = get();
    dicts[k].reserve(n);
    for (uint32_t i = 0; i < n; i++) {
      uint32_t l = get();
      LK_CHECK(p + l <= len, LK_ERR_INVALID, "dictionary blob truncated");
      dicts[k].emplace_back((const char*)blob + p, l);
      p += l;
    }
  }
  q.key_dicts = std::move(dicts);
  rebuild_group_tables(q);
}


}  // namespace lk

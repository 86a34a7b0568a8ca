// Host-side state of one glob evaluation (lk_query) and its result (lk_result).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "lk_device.h"
#include "lk_expr.h"
#include "lk_parquet.h"

namespace lk {

// Index pools are hundreds of MB for a 100-segment glob: allocated uninitialised and filled (first-touched) by the
// planner's worker threads, not zero-filled by one thread the way std::vector::assign would.
// The allocator is malloc/free until the device layer swaps in its pinned-memory pool (H2D of the pools then runs at
// PCIe speed instead of through the driver's staging buffer).
struct PoolAlloc {
  static void* (*alloc)(size_t);
  static void (*release)(void*);
};
template <class T>
struct RawVec {
  T* p = nullptr;
  size_t n = 0;
  void (*rel)(void*) = nullptr;
  RawVec() = default;
  RawVec(const RawVec&) = delete;
  RawVec& operator=(const RawVec&) = delete;
  ~RawVec() { if (p) rel(p); }
  void resize_uninit(size_t count) {
    if (p) rel(p);
    p = nullptr;
    n = count;
    if (count) {
      rel = PoolAlloc::release;
      p = static_cast<T*>(PoolAlloc::alloc(count * sizeof(T)));
      if (!p) { n = 0; throw std::bad_alloc(); }
    }
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct Options {
  int device = 0;
  int64_t segment_cache_bytes = -1;  // HBM-resident segment cache: -1 = a third of the device's memory, 0 = off
  uint64_t max_hash_slots = 1ull << 27;
  uint64_t dense_max_cells = 1ull << 25;
  uint32_t tile_rows = LK_TILE_ROWS_MAX;
  int host_threads = 0;  // 0 = hardware concurrency (capped)
  bool tune_host_malloc = true;  // lk_init raises glibc's trim / mmap thresholds (process-wide) so transient index memory is recycled
};
Options& global_options();

struct AggSpec {
  AggOp op;
  std::string aggregation;   // as written in the request ("sum", "count", "min", "max")
  std::string value_column;  // rollup_<x> | _cardinalhq.value | field$type
  double divisor = 1.0;
};

struct CachedSegment;  // lk_cache.h
struct CachedColumn;

struct SegmentInput {
  const uint8_t* data = nullptr;  // null for a file whose footer came from the segment cache, until a cache miss needs its bytes
  size_t len = 0;
  std::string name;
  void* owned_pinned = nullptr;  // file read into pinned memory owned by the query
  FileMeta meta;
  // files only: identity for the HBM-resident segment cache (path + size + mtime + inode) and the entry found at add time
  bool has_identity = false;
  uint64_t id_mtime_ns = 0, id_ino = 0;
  std::shared_ptr<CachedSegment> cached;
  bool meta_from_cache = false;
  // a file is read sparsely: its footer when it is added, then -- once the plan knows them -- only the byte ranges of the
  // touched column chunks that no cached block holds (one pinned block, the ranges back to back); `data` stays null
  bool sparse = false;
  size_t sparse_bytes = 0;
};

// one touched column chunk: where its bytes are in the file and how much room it needs in device memory
struct ChunkSlot {
  int rgi = -1, pcol = -1, seg = -1, leaf = -1;  // row-group slot (Query::rgs), touched column, segment, leaf index in the file
  uint64_t file_off = 0, len = 0, reserve = 0;
  const uint8_t* host = nullptr;  // the chunk's first byte in host memory (inside the whole file, or inside a sparse read)
};

// one physical column touched by the query
struct PCol {
  std::string name;
  bool is_ts = false, is_key = false, is_filter = false, is_value = false;
  bool string_typed = false;
  int phys_type = -1;
  int key_slot = -1, filter_slot = -1;
};

struct FilterColPlan {
  int pcol = -1;
  bool numeric = false;
  std::vector<int> leaves;  // indices into Query::leaves (this column's leaves)
  // string columns: leaf-truth signature -> class id (class 0 = SQL NULL)
  std::map<std::vector<uint8_t>, uint32_t> sig2cls;
  std::vector<std::vector<uint8_t>> cls_sig;
  uint32_t ncls = 0;
  std::vector<int> num_leaf_slot;  // numeric: for each entry of `leaves`, its comparison slot or -1 (exists)
  int n_num = 0;
};

struct RowGroupPlan {
  int seg = -1, rg = -1;
  uint32_t num_rows = 0;
  std::vector<ChunkIndex> chunks;  // per pcol
  std::vector<uint64_t> arena_base;  // per pcol: arena offset of the chunk's first byte
  std::vector<uint8_t> from_cache;   // per pcol: bytes and index of the chunk came from the segment cache (nothing to parse or upload)
  uint64_t seq_base = 0;
};

struct HostResult;
struct Comm;  // sharded evaluation: peer receive pools (lk_engine.cu)

struct Query {
  PushDownRequest req;
  std::vector<AggSpec> aggs;
  std::string path_opt = "auto";
  bool exact_sums = false;
  uint64_t seq_offset = 0;  // exact_sums, sharded: global row sequence number of this shard's first row
  Comm* comm = nullptr;  // attached communicator: the record path exchanges its records through it during the scan
  std::vector<SegmentInput> segs;

  // ---- plan ----
  bool prepared = false;
  bool is_metrics = false;
  bool tag_query = false;  // isTagQuery with a tagDataType: SELECT "tag", COUNT(*) ... GROUP BY "tag" (BaseExpr.scala:127-143)
  int64_t ts_lo = 0, ts_hi = 0, step = 0, base = 0;
  uint32_t nbuckets = 0;
  std::string ts_col_name;
  std::vector<PCol> pcols;
  int ts_pcol = -1;
  std::vector<const Clause*> leaves;
  std::vector<LeafPredicate> leaf_preds;
  std::vector<int> leaf_filter_col;  // per leaf: filter column slot, or -1 => literal FALSE (missing column)
  std::vector<FilterColPlan> fcols;
  std::vector<int> key_pcols;                          // name first, then existing group-bys
  std::vector<std::string> key_names;                  // JDBC names: "name", group-by names
  std::vector<std::vector<std::string>> key_dicts;     // global dictionaries (sorted)
  std::vector<std::vector<std::string>> local_dicts;   // union of this rank's dictionaries (export)
  std::vector<int> agg_pcols;
  std::vector<RowGroupPlan> rgs;
  int64_t touched_bytes = 0, total_rows = 0;
  uint64_t n_groups = 1;
  uint64_t n_cells = 0;
  double est_selectivity = 1.0;  // rough fraction of rows the WHERE clause keeps (path choice only)
  int path = 0;  // 0 dense, 1 hash
  uint64_t hash_slots = 0;
  uint32_t hash_stride = 0;

  // host copies of the device pools
  RawVec<TileDesc> tiles;
  RawVec<ColCursor> cursors;
  RawVec<Run> runs;
  std::vector<ChunkInfo> chunk_infos;
  // device-built seek index (device_index): page / chunk descriptors for the idx_* kernels instead of host-built runs / cursors
  bool device_index = false;
  std::vector<IdxPage> idx_pages;
  std::vector<IdxChunk> idx_chunks;  // [row group][pcol]
  std::vector<ZPage> zpages;         // SNAPPY pages: inflated (unless their block came from the segment cache) and parsed by the device before the index build
  uint64_t n_runs = 0;               // runs in the pool (device mode: known after the count pass)
  std::vector<DefChunk> def_chunks;  // chunks whose definition levels are expanded on the device before every scan
  uint64_t defbm_words = 0;          // size of the bitmap pool (32-bit words)
  uint64_t def_blocks_total = 0;     // CTAs of def_expand_kernel
  std::vector<uint8_t> lut_cls;
  std::vector<uint32_t> lut_gcode;
  std::vector<uint32_t> pass_bits;
  struct Upload { int seg; uint64_t file_off, len, arena_off; const uint8_t* src = nullptr; };  // src: host bytes that are not in a segment (re-encoded pages)
  std::vector<Upload> uploads;
  uint64_t arena_bytes = 0;
  std::vector<ChunkSlot> slots;  // every touched column chunk, in (row group, column) order
  // segment cache: columns this query uses (pinned for its lifetime) and the ones it uploaded itself, published once resident
  std::vector<std::shared_ptr<CachedColumn>> cache_refs;
  struct FreshColumn { int seg, pcol, leaf; std::shared_ptr<CachedColumn> col; };
  std::vector<FreshColumn> cache_fresh;
  // called by plan_query as soon as the arena layout is known (footers only): the device layer starts the H2D copies
  // of the column chunks there so that they overlap the host-side page/run indexing
  std::function<void()> on_layout;
  ScanParams params{};

  // ---- device ----
  struct Device;
  std::unique_ptr<Device> dev;
  double t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::string info_json;
  std::string dict_blob;

  Query();
  ~Query();
};

struct HostResult {
  bool tag_query = false;  // JDBC columns are (tag, "count") instead of (timestamp, value.., name, group-bys..)
  int64_t n = 0;
  int n_values = 0, n_tags = 0;
  std::vector<std::string> col_names;
  uint8_t* pinned = nullptr;  // one pinned block: ts | values | codes | nulls
  int64_t* ts = nullptr;
  std::vector<double*> values;
  std::vector<uint8_t*> nulls;
  std::vector<int32_t*> codes;
  std::vector<std::vector<std::string>> dicts;
  std::vector<std::vector<const char*>> dict_ptrs;
  ~HostResult();
};

// planning (host only, no CUDA): lk_plan.cpp
// places the chunk slots that are not marked in `placed` one after the other in the query's private arena (arena_base,
// uploads, arena_bytes); the default layout when no device layer takes part (on_layout unset)
void layout_private_arena(Query& q, const std::vector<uint8_t>& placed);
void plan_query(Query& q);           // parse + index + compile; fills the host pools and ScanParams (device pointers unset)
void rebuild_group_tables(Query& q); // after key_dicts changed (dictionary import)
// definition bitmaps: which chunks get one (run_n > 0 in def_tmp[row group * npcols + pcol]) and where; fills q.def_chunks,
// chunk_infos[].defbm_word0, defbm_words
void layout_def_chunks(Query& q, const std::vector<DefChunk>& def_tmp);
void refresh_info_json(Query& q);     // after the device filled in what the host plan left open (run / definition-chunk counts)
std::string export_dictionaries(const Query& q);
void import_dictionaries(Query& q, const uint8_t* blob, size_t len);
void parallel_for(int n, int threads, const std::function<void(int)>& fn);
// lk_sse.cpp: rows -> the reference's SSE stream elements (map sketches)
int64_t result_to_sse(const HostResult& res, int64_t row0, int64_t row1, const char* const* keys, int n_keys, const char* const* fallback_tags,
                      int n_fallback, char* buf, int64_t cap);

}  // namespace lk

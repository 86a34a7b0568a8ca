/* lakeside_b200.h -- C ABI of liblakeside_b200.so
 *
 * B200-native (sm_100a) implementation of cardinalhq/lakeside's per-segment DataExpr evaluation path.
 * The library replaces, behind the reference's own seam, what the reference delegates to DuckDB over JDBC:
 *   - Commons.toGlobResultSet       core/src/main/scala/com/cardinal/utils/Commons.scala:200-254
 *       (DESCRIBE + BaseExpr.generateSql + statement.executeQuery over read_parquet([...], union_by_name=True))
 *   - the ResultSet reads of Commons.resultSetToSource / toDataPoint   Commons.scala:280-341, 399-462
 *   - the K-way mergeSorted chains   Commons.scala:391-392, WorkerApi.scala:173, QueryEngineV2.scala:76-97
 *   - the map-sketch merge of TimeGroupedSketchAggregator   core/.../eval/TimeGroupedSketchAggregator.scala:63-93
 *
 * Convention follows the reference's existing FFI (JNA over a C shared object, C strings in, plain data out:
 * query-api/src/main/resources/lib-trigram.h:77-78, core/.../queries/NLPUtils.scala:30-52): plain C types only,
 * no C++ / torch types in any signature.  All functions return LK_OK (0) or an LK_ERR_* code; the message is
 * available from lk_last_error() (thread-local).  There is NO CPU fallback: without a CUDA device every compute
 * entry point returns LK_ERR_CUDA.
 *
 * Threading: every lk_query owns its CUDA stream and buffers; distinct queries may run concurrently from
 * different threads (the reference evaluates all globs of a request at once, Commons.scala:368-392).
 */
#ifndef LAKESIDE_B200_H
#define LAKESIDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LK_OK 0
#define LK_ERR_INVALID 1     /* malformed request JSON / bad argument */
#define LK_ERR_UNSUPPORTED 2 /* query or file shape outside the GPU path (extract/compute, percentile, compressed pages ...) */
#define LK_ERR_IO 3          /* file could not be read / not a Parquet file */
#define LK_ERR_CUDA 4        /* no device / CUDA failure */
#define LK_ERR_QUERY 5       /* the reference's SQL would fail to bind in DuckDB (e.g. value column absent): stream nothing */
#define LK_ERR_NOMEM 6

typedef struct lk_query lk_query;   /* one glob evaluation: request + segments + device state */
typedef struct lk_result lk_result; /* host-resident result rows, sorted by timestamp ascending */
typedef struct lk_merge lk_merge;   /* K-way merge job */
typedef struct lk_comm lk_comm;     /* sharded evaluation: this rank's receive pools + its peers' (one GPU per rank, one NVLink node) */

/* ---- library ------------------------------------------------------------------------------------------ */
/* options_json (may be NULL): {"device": 0, "max_hash_slots": 134217728, "dense_max_cells": 33554432,
 *                              "tile_rows": 512, "host_threads": 8, "tune_host_malloc": 1, "segment_cache_bytes": -1}
 * tune_host_malloc (default 1, glibc only): raises the process's M_TRIM_THRESHOLD / M_MMAP_THRESHOLD so that the
 * ~0.3 GB of transient host index memory a query builds is recycled inside the process instead of being unmapped at
 * the end of every query and faulted in again by the next (8 ms per 100-segment query); pass 0 to leave the host's
 * allocator settings alone. */
int lk_init(const char* options_json);
void lk_shutdown(void);
const char* lk_last_error(void);
const char* lk_version(void);
int lk_device_count(void); /* 0 when no CUDA device is visible; never fails */

/* ---- HBM-resident segment cache ------------------------------------------------------------------------------
 * Column chunks of segment FILES (lk_eval, lk_query_add_segment_file) stay in device memory after the query that uploaded
 * them, keyed by the file's identity (path, size, mtime, inode) and column, together with the footer / page / dictionary
 * index the host built: a later query over the same segments reads no file and moves no segment byte over PCIe.  The
 * analogue of the worker's cache of downloaded segment files (query-worker WorkerApi.scala:53-64), one level further down.
 * Owned by the library, thread-safe, least-recently-used segments evicted first; queries pin what they use.  Capacity:
 * lk_init option "segment_cache_bytes" (default: a third of the device's memory; 0 = off) or lk_cache_configure.
 * Segments handed over as caller buffers (lk_query_add_segment_buffer) have no identity and are never cached.
 * stats: [0] capacity bytes, [1] resident bytes, [2] segments, [3] column hits, [4] column misses, [5] evicted segments. */
int lk_cache_stats(int64_t* stats /*[6]*/);
int lk_cache_configure(int64_t capacity_bytes); /* -1: the default */
void lk_cache_clear(void);

/* Pinned host memory for segment bytes handed to lk_query_add_segment_buffer (fast H2D). */
void* lk_host_alloc(size_t bytes);
void lk_host_free(void* p);

/* ---- one-shot evaluation: drop-in for Commons.toGlobResultSet (Commons.scala:200-254) ------------------- */
/* pushdown_request_json: the PushDownRequest JSON the worker receives (model/SegmentRequest.scala:30-60) whose
 * segmentRequests are exactly the glob's segments; parquet_paths[i] is toParquetFilePath(segmentRequests[i])
 * (Commons.scala:256-278), already resolved by the caller. */
int lk_eval(const char* pushdown_request_json, const char* const* parquet_paths, int n_paths, lk_result** out);

/* ---- staged evaluation (resident segments, fused multi-aggregate pass, sharding) ------------------------ */
/* options_json (may be NULL):
 *   {"aggregates": [{"aggregation":"sum","rollup":"sum"}, ...]   fused pass: several (aggregation, rollup) pairs sharing
 *                                                              filter and grouping (QueryEngineV2.scala:280-296 issues them
 *                                                              as separate requests); default = the request's own chart
 *    "path": "auto"|"dense"|"hash"|"records",                   aggregate layout (records: survivors are appended and
 *                                                               aggregated by a sort in finalize; auto picks it when
 *                                                               the group space is too large for dense planes)
 *    "exact_sums": false,                                       fixed-order (segment order, then row order) double sums, bit-identical
 *                                                               to a sequential evaluator; on every path, and sharded (record path + lk_comm)
 *    "seq_offset": 0}                                           exact_sums, sharded: global row number of this shard's first row (shards hold
 *                                                               contiguous blocks of the request's segments; = rows of the shards before it) */
int lk_query_create(const char* pushdown_request_json, const char* options_json, lk_query** out);
int lk_query_add_segment_file(lk_query* q, const char* path);
/* The buffer is borrowed and must stay valid until lk_query_prepare returns. */
int lk_query_add_segment_buffer(lk_query* q, const void* data, size_t len);
/* Parses footers, page headers and dictionaries, builds the page/run/tile index, compiles the predicate to
 * dictionary-code tables, uploads the touched column chunks to HBM.  After this the query is device-resident. */
int lk_query_prepare(lk_query* q);
/* The host half of prepare: footers, page/run/tile index, predicate and group tables.  Optional; lk_query_prepare
 * runs it when it has not been called.  When a GPU is visible the copies of the touched column chunks to HBM are
 * started here (asynchronously, as soon as the arena layout is known) so that they overlap the index build and the
 * dictionary agreement of the sharded flow; on a host without a GPU it makes no CUDA call. */
int lk_query_plan(lk_query* q);
/* Sharded evaluation: every rank exports the dictionaries of its group-by columns, the host unions them
 * (any order-insensitive union, e.g. sorted) and imports the same blob on every rank so that all ranks index one
 * dense (group x bucket) space.  Blob format: see INTEGRATION.md.  Call between prepare and execute. */
int lk_query_export_dictionaries(lk_query* q, const void** blob, size_t* len);
int lk_query_import_dictionaries(lk_query* q, const void* blob, size_t len);
/* Launches the fused decode+filter+aggregate kernels on the query's stream (asynchronous).  May be called
 * repeatedly: each call first clears the partial aggregate table. */
int lk_query_execute(lk_query* q);
int lk_query_sync(lk_query* q);
/* Device-resident partial aggregates after execute (dense path), for an NCCL reduce by the host:
 *   plane 0            : uint64 row counts per cell (presence)
 *   plane 1 + a        : aggregate a; sum -> float64, count -> uint64, min/max -> order-preserving uint64 keys
 * Layout: cell = bucket * n_groups + group.  op[a]: 0 sum(f64 add) 1 count(u64 add) 2 min(u64 min) 3 max(u64 max). */
int lk_query_partial_dense(lk_query* q, int64_t* n_cells, int* n_planes, void** plane_ptrs /*[8]*/, int* plane_ops /*[8]*/);
/* Sharded dense / hash paths, metrics: GROUP BY "_cardinalhq.timestamp" buckets the raw timestamps, which sit at one offset
 * ("phase") from startTs modulo the step.  A rank learns the phase from its own surviving rows; the rank that finalizes
 * reduced or foreign cells must use the phase ALL ranks saw -- it may have kept no row itself, and ranks that saw different
 * phases must fail instead of merging distinct timestamps.  After execute: lk_query_phase gives this rank's [min, max]
 * (min = 0xffffffff: no surviving row); the host reduces MIN / MAX over the ranks next to the data exchange and hands the
 * global pair to every finalizing rank with lk_query_set_phase before lk_query_finalize*.  (The record path exchanges the
 * phase by itself through its lk_comm.) */
int lk_query_phase(lk_query* q, uint32_t* phase_min, uint32_t* phase_max);
int lk_query_set_phase(lk_query* q, uint32_t phase_min, uint32_t phase_max);
/* Hash path, sharded evaluation: the (group x bucket) cells are hash-partitioned over `nparts` owners (ranks).
 * partial_sparse moves every occupied entry {uint64 key = cell + 1; uint64 acc[..]} (stride_bytes each) out of the
 * table into one device buffer ordered by partition (counts[p] entries for partition p) and leaves the table empty;
 * after the all-to-all each rank merges the entries of ITS partition (own and foreign) with merge_sparse and
 * finalizes: every rank then holds the final rows of its partition. */
int lk_query_partial_sparse(lk_query* q, int nparts, void** entries, int64_t* counts /*[nparts]*/, int* stride_bytes);
int lk_query_merge_sparse(lk_query* q, const void* device_entries, int64_t n);
/* ---- sharded evaluation of the record path: the exchange runs INSIDE the scan, over NVLink / NVSwitch -----------------
 * Replaces, for partial results of one query on several GPUs, the reference's fan-out / fan-in over HTTP + SSE
 * (SegmentSequencer.scala:137-158, WorkerManager.scala:150-156) and the re-aggregation by (timestamp, tags) that follows it
 * (TimeGroupedSketchAggregator.scala:157-177).  One rank per GPU; the (group x bucket) cells are hash-partitioned over the
 * ranks; while a rank scans its segments every survivor record is stored straight into the receive pool of the rank that
 * owns its cell (peer memory mapped with CUDA IPC -- or plain pointers when the ranks are threads of one process); finalize
 * waits ON THE DEVICE for all sources and each rank then holds the final rows of its partition.  No host round trip and no
 * collective call on the data path.  Setup is in two steps because the host owns the out-of-band channel (the reference's
 * is HTTP): every rank creates its half and exports a handle blob, the host all-gathers the blobs, every rank connects.
 * pool_records: capacity of this rank's receive pool (records of all sources that hash to this rank, per query; a query
 * that overflows it fails with LK_ERR_NOMEM); max_aggs: most aggregates of one fused pass.  Every rank must execute the
 * same sequence of queries on a communicator (one epoch per lk_query_execute).  Group-by dictionaries are agreed on
 * beforehand with lk_query_export_dictionaries / lk_query_import_dictionaries, as for the dense path. */
int lk_comm_create(int rank, int world, int64_t pool_records, int max_aggs, lk_comm** out);
int lk_comm_handle(lk_comm* c, const void** blob, size_t* len);
int lk_comm_connect(lk_comm* c, const void* blobs /* world blobs in rank order */, size_t len_each);
void lk_comm_destroy(lk_comm* c);
/* Attach before the first lk_query_execute (NULL detaches).  With a communicator attached the query must take the record
 * path ("path":"records" or what auto picks for a large group space); execute + finalize then include the exchange. */
int lk_query_set_comm(lk_query* q, lk_comm* c);
/* Compacts the non-empty cells into result rows sorted by timestamp, still in HBM.  Idempotent until the next execute.
 * The hash path also returns its table to the clean state.  Dense and hash paths read one scalar back; the record path is
 * fully asynchronous once its scratch has been sized by the query's first finalize: the row count and any error of the
 * scan / finalize kernels are then reported by the next lk_query_sync, lk_query_finalize or lk_query_eval. */
int lk_query_finalize_device(lk_query* q);
/* lk_query_finalize_device + copy of the rows to (pinned) host memory. */
int lk_query_finalize(lk_query* q, lk_result** out);
/* Rows that satisfied the WHERE clause and the [startTs, endTs) range in the last execute (-1 on error). */
int64_t lk_query_survivors(lk_query* q);
/* BaseExpr.eval on the reduced rows of the last finalize (BaseExpr.scala:665-695 `eval`, :47-95 `getFromSketch`;
 * ASTUtils.scala:190-219 `getTransformerFunc`): out[i] = transform(sketch value of row i), where the sketch value is the
 * aggregate named `aggregation` ("sum" | "count" | "min" | "max"; "avg" = sum / count; NaN when the query did not compute
 * it; on pre-rolled metrics the key K is the aggregate over `rollup_K`, so count = sum(rollup_count)) and the transform depends on (dataset, chart type, metric type): metrics count-chart over a rate metric
 * x * (step / 1000), rate chart over a counter x / (step / 1000), events rate chart x / (step / 1000) -- integer
 * division of step first -- else identity.  Computed on the device from the result columns still in HBM; rows are in
 * the order of lk_result_*.  Returns the number of rows written, -1 on error. */
int64_t lk_query_eval(lk_query* q, const char* aggregation, const char* chart_type, const char* metric_type, double* out, int64_t cap);
/* Formula.eval (core/.../utils/ast/Formula.scala:32-69; evaluated per SketchGroup by QueryEngineV2.scala:310-389) over the
 * reduced rows of two finalized queries, on the device: e1 <op> e2 per (timestamp, group key).
 *   spec_json: {"op": "add" | "sub" | "mul" | "div",
 *               "e1": {"aggregation": "sum", "chartType": "line", "metricType": "gauge", "groupBys": [..]} | {"constant": 2.0},
 *               "e2": the same}
 * A BaseExpr side is the query handed in for it (e1 / e2; NULL for a constant side); its per-row value is that of
 * lk_query_eval, its group key the side's sorted group-by values joined by ":" (ASTUtils.scala:87-89; NULL, "" and "null"
 * read as ""; groupBys defaults to the request's chart group-bys), and a later row of the same (timestamp, key) replaces an
 * earlier one, as in the reference's map.  Equal keys are combined; `add` takes a missing side as 0, the other operators
 * give no result; a zero divisor gives no result.  A constant side takes the keys of the other side (ASTUtils.scala:50-64).
 * Output, in timestamp order: timestamp, value, and where the result's tags come from -- side (1 = e1's row, 2 = e2's row:
 * e1 is the constant or the side `add` filled in; 0 = no tags: a constant e1 without group-bys) and the row index in that
 * side's lk_result (for side 0: the e2 row the value was computed from).
 * *n_out = the number of rows (they must fit cap).  LK_ERR_UNSUPPORTED: sides with different numbers of group-bys,
 * group-by values containing ':' with two or more group-bys (the joined keys would be ambiguous), two constants. */
int lk_formula_eval(lk_query* e1, lk_query* e2, const char* spec_json, int64_t cap, int64_t* out_ts, double* out_value,
                    int32_t* out_side, int64_t* out_row, int64_t* n_out);
/* Timings of the last execute/finalize in milliseconds (CUDA events on the query's stream):
 * [0] H2D upload, [1] scan kernel, [2] finalize kernels, [3] D2H, [4] host planning, [5] definition-level expansion
 * (def_expand_kernel + the clear of its bitmaps, at the start of every execute), [6] sharded record path: the device-side wait
 * for the other ranks' records at the start of finalize (included in [2]). */
int lk_query_timings(lk_query* q, double* ms /*[8]*/);
/* Algorithmic bytes: sum of total_compressed_size of the touched column chunks. */
int64_t lk_query_touched_bytes(lk_query* q);
int64_t lk_query_total_rows(lk_query* q);
int lk_query_stream(lk_query* q, void** cuda_stream);
int lk_query_info_json(lk_query* q, const char** json); /* plan description (path, tiles, groups, buckets ...) */
void lk_query_destroy(lk_query* q);

/* ---- result: what Commons.toDataPoint reads from the JDBC ResultSet (Commons.scala:399-462) ------------- */
int64_t lk_result_num_rows(const lk_result* r);
int lk_result_num_values(const lk_result* r); /* 1 unless a fused multi-aggregate pass */
int lk_result_num_tags(const lk_result* r);   /* "name" + existing group-by columns */
/* JDBC column names in order: col 0 = "_cardinalhq.timestamp" (metrics) or "step_ts" (events); then the value
 * column(s); then "name"; then the group-by columns that exist (BaseExpr.scala:338-346, 391-403). */
int lk_result_num_cols(const lk_result* r);
const char* lk_result_col_name(const lk_result* r, int col);
const int64_t* lk_result_ts(const lk_result* r);
const double* lk_result_value(const lk_result* r, int a);       /* SQL NULL reads as 0.0, like ResultSet.getDouble */
const uint8_t* lk_result_value_null(const lk_result* r, int a); /* 1 where the aggregate is SQL NULL */
const int32_t* lk_result_tag_codes(const lk_result* r, int t);  /* -1 = SQL NULL */
int lk_result_tag_dict(const lk_result* r, int t, int32_t* n, const char* const** strings);
/* Row-at-a-time accessors for a java.sql.ResultSet shim (1-based column index like JDBC).
 * Tag queries (PushDownRequest.isTagQuery with a tagDataType; BaseExpr.scala:127-143: SELECT "tag" as "tag", COUNT(*) AS count
 * ... GROUP BY "tag") have the two JDBC columns (tag, "count") instead: get_string(row, 1) = the tag value (NULL for the NULL
 * group), get_string(row, 2) / get_long(row, 2) = the row count (the string is valid until the calling thread's next
 * lk_result_get_string); in the column view the tag is tag column 0 and the count is value column 0. */
int64_t lk_result_get_long(const lk_result* r, int64_t row, int col);
double lk_result_get_double(const lk_result* r, int64_t row, int col);
const char* lk_result_get_string(const lk_result* r, int64_t row, int col); /* NULL for SQL NULL */
/* Rows [row0, row1) as the reference's per-segment stream elements, one Server-Sent Event per row:
 *   data: {"id":"_","type":"data","message":{"timestamp":T,"tags":{..},"type":"sketch","sketchType":"map","sketch":{K:V,..}}}\r\n\r\n
 * i.e. what PushDownAggregatorStage.scala:95-106 (map sketch of globalAgg -> value), Commons.scala:474-502
 * (dataPointResponseToSSE) and SSEMessage.scala:23-34 (GenericSSEPayload.toChunkStreamPart) produce for a DataPoint, and
 * what SegmentSequencer.scala:35-101 decodes.  Value column v is published under sketch_keys[v] (n_keys <= number of
 * value columns); NULL / "" / "null" tags are dropped and a row left without tags takes the n_fallback (key, value)
 * pairs of fallback_tags (the segment's queryTags, Commons.scala:430-451); non-finite doubles are the strings "NaN",
 * "Infinity", "-Infinity" as Jackson writes them.  Returns the number of bytes the rows need; they are written only if
 * they fit `cap` (call with buf = NULL to size the buffer).  -1 on error. */
int64_t lk_result_to_sse(const lk_result* r, int64_t row0, int64_t row1, const char* const* sketch_keys, int n_keys,
                         const char* const* fallback_tags, int n_fallback, char* buf, int64_t cap);
void lk_result_free(lk_result* r);

/* ---- K-way merge of sorted per-segment streams (mergeSorted chains) ------------------ */
/* Streams are SoA (ts int64, gid int32, value f64), each sorted by ts (ascending, or descending if reverse).
 * Output order = the left-deep ``s1.mergeSorted(s2)`` fold: by ts, ties by stream index DESCENDING, then original
 * position.  Host pointers; out_* must hold sum(lens) elements; out_src (optional) receives the source stream. */
int lk_merge_streams(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val,
                     const int64_t* lens, int reverse, int64_t* out_ts, int32_t* out_gid, double* out_val,
                     int32_t* out_src);
/* Resident variant for kernel timing: create uploads, run launches the merge-path kernels, download copies back. */
int lk_merge_create(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val,
                    const int64_t* lens, int reverse, lk_merge** out);
int lk_merge_run(lk_merge* m);  /* asynchronous on the job's stream */
int lk_merge_sync(lk_merge* m);
/* ms: [0] H2D upload, [1] the merge kernels of the last run, [2] of which run detection + ordering of the runs,
 * [3] which kernels ran: 1 = merge by runs (streams whose timestamps repeat: whole runs are ranked and copied), 2 = element-wise merge path. */
int lk_merge_timings(lk_merge* m, double* ms /*[4]*/);
int lk_merge_download(lk_merge* m, int64_t* out_ts, int32_t* out_gid, double* out_val, int32_t* out_src);
/* TimeGroupedSketchAggregator map-sketch merge over the merged stream: combines equal (ts, gid) with
 * op (0 sum/count add, 2 min, 3 max) in merged order; outputs one element per (ts, gid), sorted by ts. */
int lk_merge_reduce(lk_merge* m, int op, int64_t* n_out, int64_t* out_ts, int32_t* out_gid, double* out_val);
void lk_merge_destroy(lk_merge* m);

#ifdef __cplusplus
}
#endif
#endif /* LAKESIDE_B200_H */

// Interface between the C-ABI layer (lk_api.cpp, plain C++) and the CUDA translation units.
#pragma once
#include "lk_query.h"

namespace lk {

int device_count();
void device_init();
void device_shutdown();
int num_sms();
int64_t device_default_cache_bytes();  // a third of the device's memory; 0 before device_init
void* pinned_alloc(size_t bytes);
void pinned_free(void* p);

bool device_is_resident(const Query& q);   // lk_query_prepare has completed
void device_layout(Query& q);              // plan_query's on_layout hook: segment cache / private arena placement + async H2D of what is not resident yet
void device_begin_upload(Query& q);        // async H2D of the uploads listed so far
void segment_load(SegmentInput& s);        // whole file -> pinned host memory owned by the query
void segment_open(SegmentInput& s);        // footer only (parsed from the file's tail); the touched chunks follow in device_layout
void device_upload(Query& q);              // ... plus the index pools; waits for all of it
void device_mark_group_tables_stale(Query& q);
void device_execute(Query& q);             // clear table + fused scan kernel (async)
void device_sync(Query& q);
void* device_stream(Query& q);
void device_phase(Query& q, uint32_t* pmin, uint32_t* pmax);
void device_set_phase(Query& q, uint32_t pmin, uint32_t pmax);
void device_partial_dense(Query& q, int64_t* n_cells, int* n_planes, void** ptrs, int* ops);
void device_partial_sparse(Query& q, int nparts, void** entries_out, int64_t* counts, int* stride_out);
void device_merge_sparse(Query& q, const void* dev_entries, int64_t n);
void device_finalize_device(Query& q);     // compaction into result rows in HBM
void device_resolve(Query& q);             // waits for a pending (asynchronous) finalize: row count, deferred errors
HostResult* device_fetch(Query& q);        // D2H
void device_timings(Query& q);
int64_t device_survivors(Query& q);
// sharded evaluation: communicator over the peers' receive pools (CUDA IPC on one NVLink / NVSwitch node)
Comm* comm_create(int rank, int world, int64_t pool_records, int max_aggs);
void comm_handle(Comm* c, const void** blob, size_t* len);
void comm_connect(Comm* c, const void* blobs, size_t len_each);
void comm_destroy(Comm* c);
int comm_world(const Comm* c);

// Formula.eval (Formula.scala:32-69) over the reduced rows of two finalized queries (either may be null for a constant side)
int64_t device_formula(Query* q1, Query* q2, const std::string& spec_json, int64_t cap, int64_t* out_ts, double* out_val, int32_t* out_side, int64_t* out_row);
int64_t device_eval(Query& q, const std::string& aggregation, const std::string& chart_type, const std::string& metric_type, double* out, int64_t cap);  // BaseExpr.eval on the reduced rows

}  // namespace lk

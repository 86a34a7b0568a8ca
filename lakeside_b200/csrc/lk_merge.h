// K-way merge of sorted per-segment streams on the GPU (merge-path partitioning + in-CTA merge tree).
// Replaces the left-deep `sources.fold(Source.empty)(_ mergeSorted _)` chains
// (core/src/main/scala/com/cardinal/utils/Commons.scala:391-392, query-worker WorkerApi.scala:173,
//  query-api QueryEngineV2.scala:76-97) and, with merge_reduce, the map-sketch merge of
// TimeGroupedSketchAggregator (core/.../eval/TimeGroupedSketchAggregator.scala:63-93).
#pragma once
#include <cstdint>

struct lk_merge;

namespace lk {
lk_merge* merge_create(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val, const int64_t* lens, bool reverse);
void merge_run(lk_merge* m);
void merge_sync(lk_merge* m);
void merge_timings(lk_merge* m, double* ms);
void merge_download(lk_merge* m, int64_t* out_ts, int32_t* out_gid, double* out_val, int32_t* out_src);
void merge_reduce(lk_merge* m, int op, int64_t* n_out, int64_t* out_ts, int32_t* out_gid, double* out_val);
void merge_destroy(lk_merge* m);
}  // namespace lk

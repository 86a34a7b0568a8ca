#include "lk_merge.h"

#include "lk_common.h"

struct lk_merge { int dummy; };

namespace lk {
lk_merge* merge_create(int, const int64_t* const*, const int32_t* const*, const double* const*, const int64_t*, bool) { fail(LK_ERR_UNSUPPORTED, "merge: not built yet"); }
void merge_run(lk_merge*) { fail(LK_ERR_UNSUPPORTED, "merge: not built yet"); }
void merge_sync(lk_merge*) {}
void merge_timings(lk_merge*, double*) {}
void merge_download(lk_merge*, int64_t*, int32_t*, double*, int32_t*) { fail(LK_ERR_UNSUPPORTED, "merge: not built yet"); }
void merge_reduce(lk_merge*, int, int64_t*, int64_t*, int32_t*, double*) { fail(LK_ERR_UNSUPPORTED, "merge: not built yet"); }
void merge_destroy(lk_merge* m) { delete m; }
}  // namespace lk

// GPU K-way merge of sorted per-segment streams (merge-path partitioning + in-CTA pairwise merge tree) and the
// map-sketch re-aggregation of the merged stream.
//
// Order produced = the reference's left-deep fold `sources.fold(Source.empty)((s1, s2) => s1.mergeSorted(s2))`
// (Commons.scala:391-392, WorkerApi.scala:173, QueryEngineV2.scala:96) with akka-stream's MergeSorted emitting the
// left head only when left < right: by timestamp, ties by stream index DESCENDING, then by position in the stream.
// Every element therefore has a unique sort key (ts', K-1-src, pos) and the merge is a deterministic total order.
//
// Two sets of kernels.  Streams whose timestamps repeat (what the reference merges: step-aligned aggregates, hundreds of
// elements per timestamp and stream) are merged by RUNS, see mrun_* below; everything else element-wise:
// one pass over the data: kernel 1 finds, for every output tile boundary (rank p * TILE), the exact per-stream split
// (multi-sequence selection by bisection over the key domain, a thread per stream); kernel 2 gives each CTA its K
// sub-ranges (exactly TILE elements), loads their keys into shared memory, merges the K sorted runs pairwise
// (log2 K levels of rank-by-binary-search, ping-pong buffers) and writes the tile out coalesced, gathering the
// payload.  Algorithmic traffic = 20 B read + 20 B written per element.
#include "lk_merge.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>  // lk_merge_reduce only (stable sorts of the merged order)
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "lk_common.h"
#include "lk_engine.h"

#define CUDA_CHECK(x)                                                                                        \
  do {                                                                                                       \
    cudaError_t err__ = (x);                                                                                 \
    if (err__ != cudaSuccess) ::lk::fail(LK_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(err__)); \
  } while (0)

namespace lk {

constexpr int MG_TILE = 2048;
constexpr int MG_BLOCK = 256;
// element i of a tile buffer lives at slot i + i / 8: a thread's 8 consecutive elements and its neighbours' then spread
// over all shared-memory banks (unpadded, the 128-byte stride between lanes is a 32-way conflict on every 16-byte access)
constexpr int MG_TILE_PADDED = MG_TILE + MG_TILE / 8;
__device__ __forceinline__ uint32_t mg_pad(uint32_t i) { return i + (i >> 3); }
constexpr uint32_t MG_COARSE = 8;  // tiles per coarse split boundary (C5 on B200: 64 -> 0.98 ms of fine splits, 8 -> see profiles)

struct MergeElem {
  unsigned long long key;  // order-mapped timestamp
  uint32_t tag;            // (K - 1 - src) * MG_TILE + local index: unique, realises the tie rule
  uint32_t gidx;           // index into the concatenated input
};

__device__ __forceinline__ bool elem_less(const MergeElem& a, const MergeElem& b) {
  return a.key < b.key || (a.key == b.key && a.tag < b.tag);
}

// order-preserving map int64 -> uint64; reverse streams (descending ts) are merged ascending on the complement
__device__ __forceinline__ unsigned long long ts_key(long long ts, int reverse) {
  unsigned long long u = (unsigned long long)ts ^ 0x8000000000000000ull;
  return reverse ? ~u : u;
}
__device__ __forceinline__ long long key_ts(unsigned long long k, int reverse) {
  if (reverse) k = ~k;
  return (long long)(k ^ 0x8000000000000000ull);
}

// first index in [lo, hi) of stream whose key is >= k (lower) or > k (upper)
__device__ __forceinline__ uint32_t bound(const long long* __restrict__ ts, uint32_t lo, uint32_t hi, unsigned long long k, int reverse, bool upper) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    unsigned long long v = ts_key(ts[mid], reverse);
    if (upper ? v <= k : v < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// kernel 1: per-stream split of output rank r = blockIdx.x * rank_stride.  splits[p * K + j] = elements of stream j before r.
// Two passes: a COARSE one (rank_stride = MG_COARSE tiles, whole streams, the whole key domain; it also records the key
// under every boundary) and the FINE one (every tile boundary), which searches only the windows between its two
// neighbouring coarse boundaries -- a few hundred elements per stream and a key range that is usually a single value --
// instead of bisecting the whole key domain over whole streams for each of the thousands of tiles.
__global__ void __launch_bounds__(MG_BLOCK) merge_split_kernel(const long long* __restrict__ ts, const uint32_t* __restrict__ offs, int K,
                                                               uint64_t total, unsigned long long kmin, unsigned long long kmax, int reverse,
                                                               uint32_t* __restrict__ splits, uint64_t rank_stride,
                                                               const uint32_t* __restrict__ coarse, const unsigned long long* __restrict__ ckey,
                                                               uint32_t per_coarse, unsigned long long* __restrict__ key_out) {
  __shared__ unsigned long long red[MG_BLOCK / 32];
  __shared__ unsigned long long bcast;
  uint64_t r = (uint64_t)blockIdx.x * rank_stride;
  uint32_t* out = splits + (size_t)blockIdx.x * K;
  if (r >= total) {  // the sentinel boundary after the last tile
    for (int j = threadIdx.x; j < K; j += MG_BLOCK) out[j] = offs[j + 1] - offs[j];
    if (key_out && threadIdx.x == 0) key_out[blockIdx.x] = kmax;
    return;
  }
  const uint32_t* c0 = nullptr;  // fine pass: splits of the coarse boundaries either side
  const uint32_t* c1 = nullptr;
  if (coarse) {
    const uint32_t c = blockIdx.x / per_coarse;
    c0 = coarse + (size_t)c * K;
    c1 = c0 + K;
    if (blockIdx.x % per_coarse == 0) {  // this boundary IS a coarse one
      for (int j = threadIdx.x; j < K; j += MG_BLOCK) out[j] = c0[j];
      return;
    }
    r -= (uint64_t)c * per_coarse * rank_stride;  // rank inside the windows
    kmin = ckey[c];
    kmax = ckey[c + 1];
  }
  auto win_lo = [&](int j) -> uint32_t { return offs[j] + (c0 ? c0[j] : 0u); };
  auto win_hi = [&](int j) -> uint32_t { return c1 ? offs[j] + c1[j] : offs[j + 1]; };
  auto block_sum = [&](unsigned long long v) -> unsigned long long {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < MG_BLOCK / 32; w++) t += red[w];
      bcast = t;
    }
    __syncthreads();
    return bcast;
  };
  // smallest key t with count_le(t) > r
  unsigned long long lo = kmin, hi = kmax;
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    unsigned long long c = 0;
    for (int j = threadIdx.x; j < K; j += MG_BLOCK) { const uint32_t wl = win_lo(j); c += bound(ts, wl, win_hi(j), mid, reverse, true) - wl; }
    c = block_sum(c);
    if (c > r) hi = mid; else lo = mid + 1;
  }
  const unsigned long long tstar = lo;
  if (key_out && threadIdx.x == 0) key_out[blockIdx.x] = tstar;
  unsigned long long less = 0;
  for (int j = threadIdx.x; j < K; j += MG_BLOCK) { const uint32_t wl = win_lo(j); less += bound(ts, wl, win_hi(j), tstar, reverse, false) - wl; }
  less = block_sum(less);
  // ties at tstar are consumed from the highest stream index downwards
  unsigned long long rem = r - less;
  __shared__ unsigned long long rem_s;
  if (threadIdx.x == 0) rem_s = rem;
  __syncthreads();
  for (int jb = ((K - 1) / MG_BLOCK) * MG_BLOCK; jb >= 0; jb -= MG_BLOCK) {
    const int j = jb + (MG_BLOCK - 1 - (int)threadIdx.x);  // thread 0 handles the highest stream of the chunk
    uint32_t lb = 0, cnt = 0;
    if (j < K) {
      const uint32_t wh = win_hi(j);
      lb = bound(ts, win_lo(j), wh, tstar, reverse, false);
      cnt = bound(ts, lb, wh, tstar, reverse, true) - lb;
    }
    // inclusive scan of cnt in thread order (descending stream index)
    __shared__ uint32_t warp_tot[MG_BLOCK / 32];
    uint32_t incl = cnt;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; w++) woff += warp_tot[w];
    const unsigned long long before = (unsigned long long)woff + incl - cnt;  // ties consumed by higher streams of this chunk
    const unsigned long long rem0 = rem_s;
    if (j < K) {
      unsigned long long take = rem0 > before ? rem0 - before : 0;
      if (take > cnt) take = cnt;
      out[j] = (lb - offs[j]) + (uint32_t)take;
    }
    __syncthreads();
    if (threadIdx.x == MG_BLOCK - 1) {
      const unsigned long long chunk_total = (unsigned long long)woff + incl;
      rem_s = rem0 > chunk_total ? rem0 - chunk_total : 0;
    }
    __syncthreads();
  }
}

// kernel 2: merge the K sub-ranges of one output tile
__global__ void __launch_bounds__(MG_BLOCK) merge_tile_kernel(const long long* __restrict__ ts, const int* __restrict__ gid, const double* __restrict__ val,
                                                              const uint32_t* __restrict__ offs, int K, uint64_t total, int reverse,
                                                              const uint32_t* __restrict__ splits, long long* __restrict__ out_ts,
                                                              int* __restrict__ out_gid, double* __restrict__ out_val, int* __restrict__ out_src) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MergeElem* bufA = reinterpret_cast<MergeElem*>(smem_raw);
  MergeElem* bufB = bufA + MG_TILE_PADDED;
  uint32_t* rb = reinterpret_cast<uint32_t*>(bufB + MG_TILE_PADDED);  // run boundaries, K + 1 entries (ping)
  uint32_t* rb2 = rb + (K + 1);                                 // (pong)
  uint32_t* rid = rb2 + (K + 1);                                // stream of every non-empty run
  __shared__ int nruns_s;
  const uint32_t* s0 = splits + (size_t)blockIdx.x * K;
  const uint32_t* s1 = s0 + K;
  const uint64_t tile_base = (uint64_t)blockIdx.x * MG_TILE;
  const uint32_t n = (uint32_t)min((uint64_t)MG_TILE, total - tile_base);
  // run boundaries = exclusive prefix of the sub-range lengths (single-warp scan over K); streams that contribute
  // nothing to this tile are dropped here (with heavy timestamp ties a tile draws from a dozen of the K streams, so the
  // merge tree below has 4 levels instead of log2 K)
  if (threadIdx.x < 32) {
    uint32_t carry = 0, nr = 0;
    for (int jb = 0; jb < K; jb += 32) {
      const int j = jb + threadIdx.x;
      uint32_t c = j < K ? s1[j] - s0[j] : 0, incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)threadIdx.x >= d) incl += o; }
      const unsigned ne = __ballot_sync(0xffffffffu, c > 0);
      if (c > 0) {
        const uint32_t slot = nr + __popc(ne & ((1u << threadIdx.x) - 1));
        rb[slot] = carry + incl - c;
        rid[slot] = (uint32_t)j;
      }
      nr += __popc(ne);
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) { rb[nr] = carry; nruns_s = (int)nr; }
  }
  __syncthreads();
  const int nruns0 = nruns_s;
  // load: slot s belongs to the stream j with rb[j] <= s < rb[j + 1]
  for (uint32_t s = threadIdx.x; s < n; s += MG_BLOCK) {
    int lo = 0, hi = nruns0;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (rb[mid] <= s) lo = mid; else hi = mid; }
    const int j = (int)rid[lo];
    const uint32_t gi = offs[j] + s0[j] + (s - rb[lo]);
    MergeElem e;
    e.key = ts_key(ts[gi], reverse);
    e.tag = (uint32_t)(K - 1 - j) * MG_TILE + s;
    e.gidx = gi;
    bufA[mg_pad(s)] = e;
  }
  __syncthreads();
  // pairwise merge tree over runs
  MergeElem* src = bufA;
  MergeElem* dst = bufB;
  int nruns = nruns0;
  while (nruns > 1) {
    const int nnew = (nruns + 1) >> 1;
    // merge path: every thread owns MG_TILE / MG_BLOCK consecutive output slots; one diagonal search finds where its
    // first slot cuts the pair of runs it falls into, then it merges sequentially (and walks on into the next pair when
    // the pair ends inside its slots).  Keys are unique, so the order is total.
    {
      constexpr uint32_t PER = MG_TILE / MG_BLOCK;
      uint32_t s = threadIdx.x * PER;
      const uint32_t s_end = min(n, s + PER);
      if (s < s_end) {
        int lo = 0, hi = nruns;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (rb[mid] <= s) lo = mid; else hi = mid; }
        int pr = lo >> 1;
        bool first = true;
        while (s < s_end) {
          const uint32_t a0 = rb[2 * pr], a1 = rb[min(2 * pr + 1, nruns)], b1 = rb[min(2 * pr + 2, nruns)];
          const uint32_t na = a1 - a0, nb = b1 - a1;
          uint32_t ia = 0;
          if (first) {  // diagonal d = s - a0 of this pair
            const uint32_t d = s - a0;
            uint32_t plo = d > nb ? d - nb : 0u, phi = min(d, na);
            while (plo < phi) {
              const uint32_t mid = (plo + phi) >> 1;
              if (elem_less(src[mg_pad(a0 + mid)], src[mg_pad(a1 + (d - 1 - mid))])) plo = mid + 1; else phi = mid;
            }
            ia = plo;
            first = false;
          }
          uint32_t ib = (s - a0) - ia;
          const uint32_t stop = min(s_end, b1);
          MergeElem ea, eb;
          if (ia < na) ea = src[mg_pad(a0 + ia)];
          if (ib < nb) eb = src[mg_pad(a1 + ib)];
          while (s < stop) {
            const bool take_a = ia < na && (ib >= nb || elem_less(ea, eb));
            if (take_a) { dst[mg_pad(s++)] = ea; if (++ia < na) ea = src[mg_pad(a0 + ia)]; }
            else { dst[mg_pad(s++)] = eb; if (++ib < nb) eb = src[mg_pad(a1 + ib)]; }
          }
          pr++;
        }
      }
    }
    for (int i = threadIdx.x; i <= nnew; i += MG_BLOCK) rb2[i] = i == nnew ? rb[nruns] : rb[2 * i];
    __syncthreads();
    MergeElem* t = src; src = dst; dst = t;
    uint32_t* tr = rb; rb = rb2; rb2 = tr;
    nruns = nnew;
  }
  for (uint32_t s = threadIdx.x; s < n; s += MG_BLOCK) {
    const MergeElem e = src[mg_pad(s)];
    const uint64_t o = tile_base + s;
    out_ts[o] = key_ts(e.key, reverse);
    if (out_gid) out_gid[o] = gid[e.gidx];
    if (out_val) out_val[o] = val[e.gidx];
    if (out_src) out_src[o] = K - 1 - (int)(e.tag / MG_TILE);
  }
}

// ------------------------------------------------------------------------------------------------------------
// merge by RUNS: the fast path for what the reference actually merges -- per-segment streams of step-aligned
// aggregates, where one timestamp covers hundreds of consecutive elements of a stream (C5: 360 distinct timestamps,
// 182 elements per run).  Inside a stream the elements of one timestamp are contiguous, and the tie rule (stream
// index descending, then position) keeps them contiguous in the output: the merged stream is a permutation of whole
// runs.  So instead of ranking 16.8 M elements the path ranks 92 k runs and then moves each run with one coalesced
// copy -- one pass of 8 B read (run detection) + 12 B read + 24 B written per element:
//   mrun_flags     a warp per 32 consecutive elements: head flag = first element of a stream or ts != predecessor's
//                  (ballot -> one bit per element), heads counted per 2048-element block
//   mrun_scan      single block: exclusive prefix of the block counts -> R runs
//   mrun_fill      run r = (key, first element, stream); every key also enters a small hash set -> D distinct keys
//   mrun_distinct  single block: bitonic sort of the D distinct keys in shared memory
//   mrun_matrix    cell (d, K-1-stream) of a D x K matrix <- run index (a stream holds at most one run per key): the
//                  matrix in row-major order IS the merged order of the runs
//   mrun_cells_*   prefix over the cells of (run length, non-empty) -> output offset and rank of every run, compacted
//   mrun_copy      a CTA per 2048 output elements: the runs that overlap it (two binary searches), then every element is
//                  one coalesced gather of (gid, value) from its run's source position; ts and source come from the run
// All counts stay on the device.  The path needs R <= total / 8 + K, D <= 8192 and D x K <= total / 4 + 64 Ki cells; the
// first lk_merge_run of a job reads that verdict back once and otherwise takes the element-wise merge-path kernels.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t MRUN_DMAX = 8192;            // distinct keys the run path sorts in one CTA's shared memory
constexpr uint32_t MRUN_HSET = 4 * MRUN_DMAX;   // slots of the distinct-key hash set
constexpr unsigned long long MRUN_EMPTY = ~0ull;
constexpr uint32_t MRUN_CELLS_PER_CTA = 1024;   // 256 threads x 4 consecutive cells
constexpr int MRUN_CPT = MRUN_CELLS_PER_CTA / 256;

struct MrunCtl {
  uint32_t n_runs;      // R
  uint32_t n_hashed;    // distinct keys claimed in the hash set
  uint32_t n_distinct;  // D (with the all-ones key, which the set cannot hold)
  uint32_t has_max;     // the all-ones key occurs
  uint32_t ok;          // the run path applies (set by mrun_scan, cleared by mrun_distinct)
  uint32_t pad[3];
};

// exclusive prefix over the CTA (blockDim.x a multiple of 32, <= 1024); returns the CTA's total through *total
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* warp_tot /*[32]*/, unsigned long long* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  __syncthreads();  // warp_tot may still be read from the previous call
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  unsigned long long woff = 0, tot = 0;
  for (int w = 0; w < nw; w++) { const unsigned long long t = warp_tot[w]; if (w < wid) woff += t; tot += t; }
  *total = tot;
  return woff + incl - v;
}

__global__ void __launch_bounds__(MG_BLOCK) mrun_flags_kernel(const long long* __restrict__ ts, const uint32_t* __restrict__ offs, int K, uint64_t total,
                                                              uint32_t* __restrict__ flags, uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t sflags[MG_TILE / 32];
  const uint64_t lo = (uint64_t)blockIdx.x * MG_TILE, hi = min(total, lo + MG_TILE);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int IT = MG_TILE / MG_BLOCK;  // 32-element groups per warp
  long long v[IT], first_prev[IT];
#pragma unroll
  for (int it = 0; it < IT; it++) {
    const uint64_t i = lo + (uint64_t)(wid * IT + it) * 32 + lane;
    v[it] = i < hi ? ts[i] : 0;
    first_prev[it] = (lane == 0 && i > 0 && i < hi) ? ts[i - 1] : 0;
  }
#pragma unroll
  for (int it = 0; it < IT; it++) {
    const uint64_t i = lo + (uint64_t)(wid * IT + it) * 32 + lane;
    long long prev = __shfl_up_sync(0xffffffffu, v[it], 1);
    if (lane == 0) prev = first_prev[it];
    const bool head = i < hi && (i == 0 || v[it] != prev);
    const uint32_t w = __ballot_sync(0xffffffffu, head);
    if (lane == 0) sflags[wid * IT + it] = w;
  }
  __syncthreads();
  // streams that start inside this block: their first element is a head whatever its timestamp
  {
    int a = 0, b = K;  // first stream with offs[j] >= lo
    while (a < b) { const int mid = (a + b) >> 1; if (offs[mid] < lo) a = mid + 1; else b = mid; }
    for (int j = a + (int)threadIdx.x; j < K && offs[j] < hi; j += MG_BLOCK)
      if (offs[j + 1] > offs[j]) { const uint32_t r = (uint32_t)(offs[j] - lo); atomicOr(&sflags[r >> 5], 1u << (r & 31)); }
  }
  __syncthreads();
  if (threadIdx.x < MG_TILE / 32) flags[(size_t)blockIdx.x * (MG_TILE / 32) + threadIdx.x] = sflags[threadIdx.x];
  if (threadIdx.x < 32) {
    uint32_t c = __popc(sflags[threadIdx.x]) + __popc(sflags[threadIdx.x + 32]);
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
  }
}

// in-place exclusive prefix of the block counts; R and the verdict "few enough runs" go to ctl
__global__ void __launch_bounds__(1024) mrun_scan_kernel(uint32_t* __restrict__ counts, uint32_t n, uint32_t run_cap, MrunCtl* __restrict__ ctl) {
  __shared__ unsigned long long warp_tot[32];
  constexpr uint32_t PER = 8;
  unsigned long long carry = 0;
  for (uint32_t base = 0; base < n; base += 1024 * PER) {
    const uint32_t i0 = base + threadIdx.x * PER;
    uint32_t c[PER];
    unsigned long long sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; k++) { c[k] = i0 + k < n ? counts[i0 + k] : 0; sum += c[k]; }
    unsigned long long tot;
    unsigned long long ex = carry + block_excl_scan(sum, warp_tot, &tot);
#pragma unroll
    for (uint32_t k = 0; k < PER; k++) { if (i0 + k < n) counts[i0 + k] = (uint32_t)ex; ex += c[k]; }
    carry += tot;
  }
  if (threadIdx.x == 0) {
    ctl->n_runs = (uint32_t)carry;
    ctl->ok = carry <= run_cap ? 1u : 0u;
  }
}

__device__ __forceinline__ uint32_t mrun_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  return (uint32_t)k;
}

__global__ void __launch_bounds__(64) mrun_fill_kernel(const long long* __restrict__ ts, const uint32_t* __restrict__ offs, int K, uint64_t total,
                                                       const uint32_t* __restrict__ flags, const uint32_t* __restrict__ block_prefix, int reverse,
                                                       MrunCtl* __restrict__ ctl, unsigned long long* __restrict__ run_key, uint32_t* __restrict__ run_start,
                                                       uint32_t* __restrict__ run_stream, unsigned long long* __restrict__ hset,
                                                       unsigned long long* __restrict__ dlist) {
  if (!ctl->ok) return;
  __shared__ unsigned long long warp_tot[32];
  if (blockIdx.x == 0 && threadIdx.x == 0) run_start[ctl->n_runs] = (uint32_t)total;
  uint32_t w = flags[(size_t)blockIdx.x * 64 + threadIdx.x];
  unsigned long long tot;
  uint32_t pos = block_prefix[blockIdx.x] + (uint32_t)block_excl_scan(__popc(w), warp_tot, &tot);
  while (w) {
    const int bit = __ffs(w) - 1;
    w &= w - 1;
    const uint32_t i = blockIdx.x * MG_TILE + threadIdx.x * 32 + bit;
    const unsigned long long key = ts_key(ts[i], reverse);
    int a = 0, b = K;  // last stream with offs[j] <= i (empty streams in front of it share the offset and lose)
    while (b - a > 1) { const int mid = (a + b) >> 1; if (offs[mid] <= i) a = mid; else b = mid; }
    run_key[pos] = key;
    run_start[pos] = i;
    run_stream[pos] = (uint32_t)a;
    pos++;
    if (key == MRUN_EMPTY) { ctl->has_max = 1; continue; }
    if (*(volatile uint32_t*)&ctl->n_hashed > MRUN_DMAX) continue;  // too many distinct keys: the verdict is already "no"
    uint32_t h = mrun_hash(key) & (MRUN_HSET - 1);
    for (uint32_t probes = 0; probes < MRUN_HSET; probes++, h = (h + 1) & (MRUN_HSET - 1)) {
      unsigned long long cur = *(volatile unsigned long long*)&hset[h];
      if (cur == MRUN_EMPTY) cur = atomicCAS(&hset[h], MRUN_EMPTY, key);
      if (cur == MRUN_EMPTY) {
        const uint32_t idx = atomicAdd(&ctl->n_hashed, 1u);
        if (idx < MRUN_DMAX) dlist[idx] = key;
        break;
      }
      if (cur == key) break;
    }
  }
}

// sorts the distinct keys (bitonic, shared memory) and closes the verdict
__global__ void __launch_bounds__(256) mrun_distinct_kernel(MrunCtl* __restrict__ ctl, const unsigned long long* __restrict__ dlist,
                                                             unsigned long long* __restrict__ dsorted, int K, uint64_t cell_cap) {
  extern __shared__ unsigned long long skeys[];
  if (!ctl->ok) return;
  const uint32_t nh = ctl->n_hashed;
  const uint32_t D = nh + (ctl->has_max ? 1u : 0u);
  if (nh > MRUN_DMAX || D > MRUN_DMAX || (uint64_t)D * (uint64_t)K > cell_cap) {
    __syncthreads();
    if (threadIdx.x == 0) ctl->ok = 0;
    return;
  }
  uint32_t P = 1;
  while (P < D) P <<= 1;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) skeys[i] = i < nh ? dlist[i] : MRUN_EMPTY;  // the all-ones key sorts last, like the padding
  __syncthreads();
  if (D <= 1024) {  // few keys (the usual case: one per step of the query range): every key counts the smaller ones, no stage barriers
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) {
      const unsigned long long k = skeys[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < D; j++) rank += skeys[j] < k;  // keys are distinct (the padding equals the all-ones key but sits behind it)
      dsorted[i < nh ? rank : D - 1] = k;
    }
    if (threadIdx.x == 0) ctl->n_distinct = D;
    return;
  }
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const unsigned long long a = skeys[i], b = skeys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { skeys[i] = b; skeys[l] = a; }
        }
      }
      __syncthreads();
    }
  for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) dsorted[i] = skeys[i];
  if (threadIdx.x == 0) ctl->n_distinct = D;
}

__global__ void __launch_bounds__(256) mrun_matrix_kernel(const MrunCtl* __restrict__ ctl, const unsigned long long* __restrict__ run_key,
                                                          const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ run_stream,
                                                          const unsigned long long* __restrict__ dsorted, int K, unsigned long long* __restrict__ mat) {
  if (!ctl->ok) return;
  const uint32_t R = ctl->n_runs, D = ctl->n_distinct;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    const unsigned long long key = run_key[r];
    uint32_t a = 0, b = D;
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (dsorted[mid] < key) a = mid + 1; else b = mid; }
    // cell = run length << 32 | run index + 1: the prefix kernels never go back to the run arrays for an empty cell
    mat[(uint64_t)a * K + (uint32_t)(K - 1) - run_stream[r]] = ((unsigned long long)(run_start[r + 1] - run_start[r]) << 32) | (r + 1);
  }
}

// prefix over the cells of (elements << 32 | 1) per non-empty cell: phase 1 = per-CTA sums, phase 2 = one block over the
// sums, phase 3 = the CTA's cells again with its carry; every run then knows its output offset and its rank
__device__ __forceinline__ unsigned long long mrun_cell_weight(unsigned long long cell) {
  return (cell & 0xffffffff00000000ull) | (cell ? 1ull : 0ull);
}

__global__ void __launch_bounds__(256) mrun_cells_sum_kernel(const MrunCtl* __restrict__ ctl, int K, const unsigned long long* __restrict__ mat,
                                                             unsigned long long* __restrict__ partials) {
  __shared__ unsigned long long warp_tot[32];
  if (!ctl->ok) return;
  const uint64_t n = (uint64_t)ctl->n_distinct * K;
  const uint64_t base = (uint64_t)blockIdx.x * MRUN_CELLS_PER_CTA;
  if (base >= n) return;
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < MRUN_CPT; k++) { const uint64_t i = base + k * 256 + threadIdx.x; if (i < n) s += mrun_cell_weight(mat[i]); }
  unsigned long long tot;
  block_excl_scan(s, warp_tot, &tot);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) mrun_cells_scan_kernel(const MrunCtl* __restrict__ ctl, int K, unsigned long long* __restrict__ partials) {
  __shared__ unsigned long long warp_tot[32];
  if (!ctl->ok) return;
  const uint32_t n = (uint32_t)(((uint64_t)ctl->n_distinct * K + MRUN_CELLS_PER_CTA - 1) / MRUN_CELLS_PER_CTA);  // CTAs that hold cells
  unsigned long long carry = 0;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < n ? partials[i] : 0;
    unsigned long long tot;
    const unsigned long long ex = carry + block_excl_scan(v, warp_tot, &tot);
    if (i < n) partials[i] = ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(256) mrun_cells_emit_kernel(const MrunCtl* __restrict__ ctl, int K, uint64_t total, unsigned long long* __restrict__ mat,
                                                              const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ run_stream,
                                                              const unsigned long long* __restrict__ run_key, const unsigned long long* __restrict__ partials,
                                                              uint32_t* __restrict__ m_dst, uint32_t* __restrict__ m_src, uint32_t* __restrict__ m_stream,
                                                              unsigned long long* __restrict__ m_key) {
  __shared__ unsigned long long warp_tot[32];
  if (!ctl->ok) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) m_dst[ctl->n_runs] = (uint32_t)total;
  const uint64_t n = (uint64_t)ctl->n_distinct * K;
  const uint64_t base = (uint64_t)blockIdx.x * MRUN_CELLS_PER_CTA + (uint64_t)threadIdx.x * MRUN_CPT;
  if ((uint64_t)blockIdx.x * MRUN_CELLS_PER_CTA >= n) return;
  unsigned long long cells[MRUN_CPT];
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < MRUN_CPT; k++) {
    cells[k] = base + k < n ? mat[base + k] : 0;
    if (cells[k]) mat[base + k] = 0;  // the matrix goes back to all-empty for the job's next run
    s += mrun_cell_weight(cells[k]);
  }
  unsigned long long tot;
  unsigned long long ex = partials[blockIdx.x] + block_excl_scan(s, warp_tot, &tot);
#pragma unroll
  for (int k = 0; k < MRUN_CPT; k++) {
    const unsigned long long c = cells[k];
    if (!c) continue;
    const uint32_t r = (uint32_t)c - 1, rank = (uint32_t)ex;
    m_dst[rank] = (uint32_t)(ex >> 32);
    m_src[rank] = run_start[r];
    m_stream[rank] = run_stream[r];
    m_key[rank] = run_key[r];
    ex += mrun_cell_weight(c);
  }
}

__global__ void __launch_bounds__(MG_BLOCK) mrun_copy_kernel(const MrunCtl* __restrict__ ctl, const int* __restrict__ gid, const double* __restrict__ val,
                                                             uint64_t total, int reverse, const uint32_t* __restrict__ m_dst, const uint32_t* __restrict__ m_src,
                                                             const uint32_t* __restrict__ m_stream, const unsigned long long* __restrict__ m_key,
                                                             long long* __restrict__ out_ts, int* __restrict__ out_gid, double* __restrict__ out_val,
                                                             int* __restrict__ out_src) {
  __shared__ uint32_t s_dst[MG_TILE + 1];
  __shared__ uint32_t s_src[MG_TILE];
  __shared__ uint32_t s_stream[MG_TILE];
  __shared__ unsigned long long s_key[MG_TILE];
  __shared__ uint32_t s_e[2];
  if (!ctl->ok) return;
  const uint32_t R = ctl->n_runs;
  const uint64_t lo64 = (uint64_t)blockIdx.x * MG_TILE;
  const uint32_t lo = (uint32_t)lo64, hi = (uint32_t)min(total, lo64 + MG_TILE);
  if (threadIdx.x < 64) {  // warps 0 and 1: 32-ary search for the last run whose first output slot is <= x (m_dst[0] = 0 <= x)
    const uint32_t x = threadIdx.x < 32 ? lo : hi - 1, lane = threadIdx.x & 31;
    uint32_t a = 0, b = R;  // answer in [a, b)
    while (b - a > 1) {
      const uint32_t step = (b - a + 31) / 32;              // lane probes a + lane * step
      const uint32_t p = a + lane * step;
      const bool le = p < b && m_dst[p] <= x;               // monotone: true for lanes 0..c-1
      const uint32_t c = __popc(__ballot_sync(0xffffffffu, le));  // >= 1 (lane 0 probes a)
      const uint32_t na = a + (c - 1) * step;
      b = min(b, na + step);
      a = na;
    }
    if (lane == 0) s_e[threadIdx.x >> 5] = a;
  }
  __syncthreads();
  const uint32_t e0 = s_e[0], ne = s_e[1] - e0 + 1;  // <= MG_TILE: every run holds at least one element
  for (uint32_t k = threadIdx.x; k < ne; k += MG_BLOCK) {
    s_dst[k] = m_dst[e0 + k];
    s_src[k] = m_src[e0 + k];
    s_stream[k] = m_stream[e0 + k];
    s_key[k] = m_key[e0 + k];
  }
  __syncthreads();
  constexpr int PER = MG_TILE / MG_BLOCK;
  uint32_t src[PER], e[PER];
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const uint32_t o = lo + k * MG_BLOCK + threadIdx.x;
    uint32_t a = 0, b = ne;
    while (b - a > 1) { const uint32_t mid = (a + b) >> 1; if (s_dst[mid] <= o) a = mid; else b = mid; }
    e[k] = a;
    src[k] = s_src[a] + (o - s_dst[a]);
  }
  int g[PER];
  double v[PER];
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const bool in = lo + k * MG_BLOCK + threadIdx.x < hi;
    g[k] = (in && out_gid) ? gid[src[k]] : 0;
    v[k] = (in && out_val) ? val[src[k]] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const uint32_t o = lo + k * MG_BLOCK + threadIdx.x;
    if (o >= hi) continue;
    out_ts[o] = key_ts(s_key[e[k]], reverse);
    if (out_gid) out_gid[o] = g[k];
    if (out_val) out_val[o] = v[k];
    if (out_src) out_src[o] = (int)s_stream[e[k]];
  }
}

// ---- map-sketch re-aggregation of the merged stream (TimeGroupedSketchAggregator.scala:63-93) ----
// Elements with equal timestamp are contiguous after the merge; equal (ts, gid) pairs are combined with
// op 0: + (sum / count), 2: min, 3: max.  Output: one element per (ts, gid), sorted by ts then gid.
__global__ void mr_iota_kernel(uint32_t* v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void mr_gid_keys_kernel(const int* __restrict__ gid, uint32_t n, unsigned long long* __restrict__ keys) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (unsigned long long)(uint32_t)gid[i] ^ 0x80000000ull;  // signed order
}
__global__ void mr_ts_keys_kernel(const long long* __restrict__ ts, const uint32_t* __restrict__ idx, uint32_t n, int reverse,
                                  unsigned long long* __restrict__ keys) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ts_key(ts[idx[i]], reverse);
}
// head[i] = 1 iff sorted element i starts a new (ts, gid) group
__global__ void mr_heads_kernel(const unsigned long long* __restrict__ sorted_ts, const int* __restrict__ gid, const uint32_t* __restrict__ order,
                                uint32_t n, uint32_t* __restrict__ head) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || sorted_ts[i] != sorted_ts[i - 1] || gid[order[i]] != gid[order[i - 1]]) ? 1u : 0u;
}
__device__ __forceinline__ double java_min(double a, double b) {  // Math.min: NaN wins, -0.0 < +0.0
  if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000ll);
  if (a == 0.0 && b == 0.0) return (__double_as_longlong(a) < 0) ? a : b;
  return a <= b ? a : b;
}
__device__ __forceinline__ double java_max(double a, double b) {
  if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000ll);
  if (a == 0.0 && b == 0.0) return (__double_as_longlong(a) < 0) ? b : a;
  return a >= b ? a : b;
}
// one thread per group: folds the group's values strictly in merged (arrival) order, as SimpleSketchMerger does
__global__ void mr_fold_kernel(const unsigned long long* __restrict__ sorted_ts, const int* __restrict__ gid, const double* __restrict__ val,
                               const uint32_t* __restrict__ order, const uint32_t* __restrict__ head, const uint32_t* __restrict__ pos, uint32_t n,
                               int op, int reverse, long long* __restrict__ out_ts, int* __restrict__ out_gid, double* __restrict__ out_val) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  double acc = val[order[i]];
  for (uint32_t j = i + 1; j < n && !head[j]; j++) {
    const double v = val[order[j]];
    acc = op == 2 ? java_min(acc, v) : op == 3 ? java_max(acc, v) : acc + v;
  }
  const uint32_t o = pos[i];
  out_ts[o] = key_ts(sorted_ts[i], reverse);
  out_gid[o] = gid[order[i]];
  out_val[o] = acc;
}

}  // namespace lk

using namespace lk;

struct lk_merge {
  int K = 0;
  bool reverse = false;
  uint64_t total = 0;
  unsigned long long kmin = 0, kmax = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[4] = {};
  long long* d_ts = nullptr;
  int* d_gid = nullptr;
  double* d_val = nullptr;
  uint32_t* d_offs = nullptr;
  uint32_t* d_splits = nullptr;
  uint32_t* d_coarse = nullptr;         // splits of every MG_COARSE-th tile boundary
  unsigned long long* d_ckey = nullptr;  // and the key under each of them
  long long* o_ts = nullptr;
  int* o_gid = nullptr;
  double* o_val = nullptr;
  int* o_src = nullptr;
  uint32_t ntiles = 0;
  bool ran = false;
  double ms[4] = {0, 0, 0, 0};
  // merge by runs (see mrun_* above): scratch + the verdict of the job's first run
  int mode = 0;  // 0 not decided yet, 1 runs, 2 element-wise merge path
  uint32_t run_cap = 0, cell_parts = 0, known_distinct = 0;
  uint64_t cell_cap = 0;
  lk::MrunCtl* r_ctl = nullptr;
  uint32_t* r_flags = nullptr;
  uint32_t* r_blocks = nullptr;
  unsigned long long* r_hset = nullptr;
  unsigned long long* r_dlist = nullptr;
  unsigned long long* r_dsorted = nullptr;
  unsigned long long* r_key = nullptr;
  uint32_t* r_start = nullptr;
  uint32_t* r_stream = nullptr;
  unsigned long long* r_mat = nullptr;
  unsigned long long* r_parts = nullptr;
  uint32_t* m_dst = nullptr;
  uint32_t* m_src = nullptr;
  uint32_t* m_stream = nullptr;
  unsigned long long* m_key = nullptr;
  cudaEvent_t ev_mid = nullptr;
  // steady state of the run path (verdict known): the two clears and nine small-to-large kernels of a run are one CUDA graph --
  // the ordering kernels take 4-10 us each, about as long as the gaps between separately launched dependent kernels
  cudaGraphExec_t graph = nullptr;
};

namespace lk {

static unsigned long long host_key(long long ts, bool reverse) {
  unsigned long long u = (unsigned long long)ts ^ 0x8000000000000000ull;
  return reverse ? ~u : u;
}

lk_merge* merge_create(int k, const int64_t* const* ts, const int32_t* const* gid, const double* const* val, const int64_t* lens, bool reverse) {
  device_init();
  auto m = new lk_merge();
  try {
    m->K = k;
    m->reverse = reverse;
    std::vector<uint32_t> offs(k + 1, 0);
    uint64_t total = 0;
    for (int j = 0; j < k; j++) {
      LK_CHECK(lens[j] >= 0, LK_ERR_INVALID, "negative stream length");
      total += (uint64_t)lens[j];
      LK_CHECK(total < 0xfffffff0ull, LK_ERR_UNSUPPORTED, "merge: more than 2^32 elements");
      offs[j + 1] = (uint32_t)total;
    }
    LK_CHECK((uint64_t)k * MG_TILE < 0xffffffffull, LK_ERR_UNSUPPORTED, "merge: too many streams");
    m->total = total;
    CUDA_CHECK(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
    for (auto& e : m->ev) CUDA_CHECK(cudaEventCreate(&e));
    bool any = false;
    for (int j = 0; j < k; j++) {
      if (!lens[j]) continue;
      // streams must be sorted; the key domain is bracketed by the heads and tails
      unsigned long long a = host_key(ts[j][0], reverse), b = host_key(ts[j][lens[j] - 1], reverse);
      LK_CHECK(a <= b, LK_ERR_INVALID, "merge: stream " + std::to_string(j) + " is not sorted in the requested direction");
      if (!any) { m->kmin = a; m->kmax = b; any = true; }
      m->kmin = std::min(m->kmin, a);
      m->kmax = std::max(m->kmax, b);
    }
    size_t n = std::max<uint64_t>(total, 1);
    CUDA_CHECK(cudaEventRecord(m->ev[0], m->st));
    CUDA_CHECK(cudaMallocAsync(&m->d_ts, n * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->d_gid, n * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->d_val, n * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->d_offs, (k + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->o_ts, n * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->o_gid, n * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->o_val, n * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->o_src, n * 4, m->st));
    m->ntiles = (uint32_t)((total + MG_TILE - 1) / MG_TILE);
    CUDA_CHECK(cudaMallocAsync(&m->d_splits, ((size_t)m->ntiles + 1) * std::max(k, 1) * 4, m->st));
    const size_t ncoarse = (m->ntiles + MG_COARSE - 1) / MG_COARSE;
    CUDA_CHECK(cudaMallocAsync(&m->d_coarse, (ncoarse + 1) * std::max(k, 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->d_ckey, (ncoarse + 1) * 8, m->st));
    CUDA_CHECK(cudaMemcpyAsync(m->d_offs, offs.data(), (k + 1) * 4, cudaMemcpyHostToDevice, m->st));
    // merge by runs: scratch sized for the shapes the path accepts (see the limits above)
    m->run_cap = (uint32_t)std::min<uint64_t>(total / 8 + (uint64_t)k + 1, 0xfffffff0ull);
    m->cell_cap = total / 4 + 65536;
    m->cell_parts = (uint32_t)((m->cell_cap + MRUN_CELLS_PER_CTA - 1) / MRUN_CELLS_PER_CTA);
    if (const char* e = getenv("LK_MERGE_MODE")) m->mode = !strcmp(e, "elems") ? 2 : 0;  // tests: force the element-wise kernels
    CUDA_CHECK(cudaEventCreate(&m->ev_mid));
    CUDA_CHECK(cudaMallocAsync(&m->r_ctl, sizeof(MrunCtl), m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_flags, ((size_t)m->ntiles + 1) * (MG_TILE / 32) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_blocks, ((size_t)m->ntiles + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_hset, (size_t)MRUN_HSET * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_dlist, (size_t)MRUN_DMAX * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_dsorted, (size_t)MRUN_DMAX * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_key, ((size_t)m->run_cap + 1) * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_start, ((size_t)m->run_cap + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_stream, ((size_t)m->run_cap + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->r_mat, (size_t)m->cell_cap * 8, m->st));
    CUDA_CHECK(cudaMemsetAsync(m->r_mat, 0, (size_t)m->cell_cap * 8, m->st));  // all cells empty; mrun_cells_emit restores that after every run
    CUDA_CHECK(cudaMallocAsync(&m->r_parts, ((size_t)m->cell_parts + 1) * 8, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->m_dst, ((size_t)m->run_cap + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->m_src, ((size_t)m->run_cap + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->m_stream, ((size_t)m->run_cap + 1) * 4, m->st));
    CUDA_CHECK(cudaMallocAsync(&m->m_key, ((size_t)m->run_cap + 1) * 8, m->st));
    for (int j = 0; j < k; j++) {
      if (!lens[j]) continue;
      CUDA_CHECK(cudaMemcpyAsync(m->d_ts + offs[j], ts[j], lens[j] * 8, cudaMemcpyHostToDevice, m->st));
      if (gid && gid[j]) CUDA_CHECK(cudaMemcpyAsync(m->d_gid + offs[j], gid[j], lens[j] * 4, cudaMemcpyHostToDevice, m->st));
      else CUDA_CHECK(cudaMemsetAsync(m->d_gid + offs[j], 0, lens[j] * 4, m->st));
      if (val && val[j]) CUDA_CHECK(cudaMemcpyAsync(m->d_val + offs[j], val[j], lens[j] * 8, cudaMemcpyHostToDevice, m->st));
      else CUDA_CHECK(cudaMemsetAsync(m->d_val + offs[j], 0, lens[j] * 8, m->st));
    }
    CUDA_CHECK(cudaEventRecord(m->ev[1], m->st));
    CUDA_CHECK(cudaStreamSynchronize(m->st));  // the host arrays are borrowed for the call only
    float f = 0;
    CUDA_CHECK(cudaEventElapsedTime(&f, m->ev[0], m->ev[1]));
    m->ms[0] = f;
    static bool attr_set = false;
    size_t smem = 2 * MG_TILE_PADDED * sizeof(MergeElem) + 3 * (size_t)(k + 1) * 4 + 16;
    LK_CHECK(smem <= 200 * 1024, LK_ERR_UNSUPPORTED, "merge: too many streams for one shared-memory tile");
    (void)attr_set;
    CUDA_CHECK(cudaFuncSetAttribute(merge_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK(cudaFuncSetAttribute(mrun_distinct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MRUN_DMAX * 8)));
  } catch (...) {
    merge_destroy(m);
    throw;
  }
  return m;
}

static void merge_run_elems(lk_merge* m) {
  size_t smem = 2 * MG_TILE_PADDED * sizeof(MergeElem) + 3 * (size_t)(m->K + 1) * 4 + 16;
  const uint32_t ncoarse = (m->ntiles + MG_COARSE - 1) / MG_COARSE;  // coarse boundaries 0..ncoarse (the last one is the sentinel)
  merge_split_kernel<<<ncoarse + 1, MG_BLOCK, 0, m->st>>>(m->d_ts, m->d_offs, m->K, m->total, m->kmin, m->kmax, m->reverse ? 1 : 0, m->d_coarse,
                                                          (uint64_t)MG_COARSE * MG_TILE, nullptr, nullptr, 1, m->d_ckey);
  merge_split_kernel<<<m->ntiles + 1, MG_BLOCK, 0, m->st>>>(m->d_ts, m->d_offs, m->K, m->total, m->kmin, m->kmax, m->reverse ? 1 : 0, m->d_splits,
                                                            (uint64_t)MG_TILE, m->d_coarse, m->d_ckey, MG_COARSE, nullptr);
  merge_tile_kernel<<<m->ntiles, MG_BLOCK, smem, m->st>>>(m->d_ts, m->d_gid, m->d_val, m->d_offs, m->K, m->total, m->reverse ? 1 : 0, m->d_splits,
                                                          m->o_ts, m->o_gid, m->o_val, m->o_src);
  CUDA_CHECK(cudaGetLastError());
}

static void merge_enqueue_runs(lk_merge* m, bool first_half, bool second_half, bool record_mid) {
  const int rev = m->reverse ? 1 : 0;
  if (first_half) {
    // run detection and the verdict (R, D, cells) -- see mrun_* above
    CUDA_CHECK(cudaMemsetAsync(m->r_ctl, 0, sizeof(MrunCtl), m->st));
    CUDA_CHECK(cudaMemsetAsync(m->r_hset, 0xff, (size_t)MRUN_HSET * 8, m->st));
    mrun_flags_kernel<<<m->ntiles, MG_BLOCK, 0, m->st>>>(m->d_ts, m->d_offs, m->K, m->total, m->r_flags, m->r_blocks);
    mrun_scan_kernel<<<1, 1024, 0, m->st>>>(m->r_blocks, m->ntiles, m->run_cap, m->r_ctl);
    mrun_fill_kernel<<<m->ntiles, 64, 0, m->st>>>(m->d_ts, m->d_offs, m->K, m->total, m->r_flags, m->r_blocks, rev, m->r_ctl, m->r_key, m->r_start,
                                                  m->r_stream, m->r_hset, m->r_dlist);
    // shared memory for the keys it sorts: all MRUN_DMAX until the first run has told how many there are
    const size_t dsmem = (m->mode == 0 || m->known_distinct > 1024) ? (size_t)MRUN_DMAX * 8 : 1024 * 8;
    mrun_distinct_kernel<<<1, 256, dsmem, m->st>>>(m->r_ctl, m->r_dlist, m->r_dsorted, m->K, m->cell_cap);
    CUDA_CHECK(cudaGetLastError());
  }
  if (second_half) {
    const int wide = num_sms() * 4;
    mrun_matrix_kernel<<<wide, 256, 0, m->st>>>(m->r_ctl, m->r_key, m->r_start, m->r_stream, m->r_dsorted, m->K, m->r_mat);
    mrun_cells_sum_kernel<<<m->cell_parts, 256, 0, m->st>>>(m->r_ctl, m->K, m->r_mat, m->r_parts);
    mrun_cells_scan_kernel<<<1, 1024, 0, m->st>>>(m->r_ctl, m->K, m->r_parts);
    mrun_cells_emit_kernel<<<m->cell_parts, 256, 0, m->st>>>(m->r_ctl, m->K, m->total, m->r_mat, m->r_start, m->r_stream, m->r_key, m->r_parts, m->m_dst,
                                                             m->m_src, m->m_stream, m->m_key);
    if (record_mid) CUDA_CHECK(cudaEventRecord(m->ev_mid, m->st));
    mrun_copy_kernel<<<m->ntiles, MG_BLOCK, 0, m->st>>>(m->r_ctl, m->d_gid, m->d_val, m->total, rev, m->m_dst, m->m_src, m->m_stream, m->m_key, m->o_ts,
                                                        m->o_gid, m->o_val, m->o_src);
    CUDA_CHECK(cudaGetLastError());
  }
}

void merge_run(lk_merge* m) {
  CUDA_CHECK(cudaSetDevice(global_options().device));
  CUDA_CHECK(cudaEventRecord(m->ev[2], m->st));
  static const bool no_graph = getenv("LK_MERGE_NO_GRAPH") != nullptr;  // tuning aid
  if (m->total > 0 && m->mode == 1 && !no_graph) {
    // verdict known (the job's inputs never change): the whole run is one graph launch
    if (!m->graph) {
      cudaGraph_t g = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(m->st, cudaStreamCaptureModeThreadLocal));
      try {
        merge_enqueue_runs(m, true, true, false);
      } catch (...) { cudaStreamEndCapture(m->st, &g); if (g) cudaGraphDestroy(g); throw; }
      CUDA_CHECK(cudaStreamEndCapture(m->st, &g));
      cudaError_t e = cudaGraphInstantiate(&m->graph, g, 0);
      cudaGraphDestroy(g);
      CUDA_CHECK(e);
    }
    CUDA_CHECK(cudaEventRecord(m->ev_mid, m->st));  // (no split inside a graph launch: [2] reads 0)
    CUDA_CHECK(cudaGraphLaunch(m->graph, m->st));
    CUDA_CHECK(cudaEventRecord(m->ev[3], m->st));
    m->ran = true;
    return;
  }
  if (m->total > 0 && m->mode != 2) {
    merge_enqueue_runs(m, true, false, false);
    if (m->mode == 0) {  // first run of this job: one read-back decides which kernels follow (the inputs never change)
      MrunCtl h;
      CUDA_CHECK(cudaMemcpyAsync(&h, m->r_ctl, sizeof h, cudaMemcpyDeviceToHost, m->st));
      CUDA_CHECK(cudaStreamSynchronize(m->st));
      m->mode = h.ok ? 1 : 2;
      m->known_distinct = h.n_distinct;
    }
  }
  if (m->total > 0 && m->mode == 1) merge_enqueue_runs(m, false, true, true);
  else if (m->total > 0) {
    CUDA_CHECK(cudaEventRecord(m->ev_mid, m->st));
    merge_run_elems(m);
  }
  CUDA_CHECK(cudaEventRecord(m->ev[3], m->st));
  m->ran = true;
}

void merge_sync(lk_merge* m) { CUDA_CHECK(cudaStreamSynchronize(m->st)); }

void merge_timings(lk_merge* m, double* ms) {
  merge_sync(m);
  float f = 0;
  if (m->ran && cudaEventElapsedTime(&f, m->ev[2], m->ev[3]) == cudaSuccess) m->ms[1] = f;
  // [2]: run detection + ordering of the runs (or, element-wise path, the wait for the verdict); [3]: 1 = merged by runs, 2 = element-wise
  if (m->ran && m->total > 0 && cudaEventElapsedTime(&f, m->ev[2], m->ev_mid) == cudaSuccess) m->ms[2] = f;
  m->ms[3] = m->mode;
  cudaGetLastError();
  for (int i = 0; i < 4; i++) ms[i] = m->ms[i];
}

void merge_download(lk_merge* m, int64_t* out_ts, int32_t* out_gid, double* out_val, int32_t* out_src) {
  LK_CHECK(m->ran, LK_ERR_INVALID, "lk_merge_download before lk_merge_run");
  size_t n = m->total;
  if (n) {
    if (out_ts) CUDA_CHECK(cudaMemcpyAsync(out_ts, m->o_ts, n * 8, cudaMemcpyDeviceToHost, m->st));
    if (out_gid) CUDA_CHECK(cudaMemcpyAsync(out_gid, m->o_gid, n * 4, cudaMemcpyDeviceToHost, m->st));
    if (out_val) CUDA_CHECK(cudaMemcpyAsync(out_val, m->o_val, n * 8, cudaMemcpyDeviceToHost, m->st));
    if (out_src) CUDA_CHECK(cudaMemcpyAsync(out_src, m->o_src, n * 4, cudaMemcpyDeviceToHost, m->st));
  }
  CUDA_CHECK(cudaStreamSynchronize(m->st));
}

void merge_reduce(lk_merge* m, int op, int64_t* n_out, int64_t* out_ts, int32_t* out_gid, double* out_val) {
  LK_CHECK(m->ran, LK_ERR_INVALID, "lk_merge_reduce before lk_merge_run");
  LK_CHECK(op == 0 || op == 1 || op == 2 || op == 3, LK_ERR_INVALID, "op must be 0/1 (add), 2 (min) or 3 (max)");
  CUDA_CHECK(cudaSetDevice(global_options().device));
  *n_out = 0;
  if (m->total == 0) return;
  const uint32_t n = (uint32_t)m->total;
  size_t t1 = 0, t2 = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, (int)n, 0, 64, m->st));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, t2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, m->st));
  const size_t tmp_bytes = std::max(t1, t2);
  // scratch: key_a | key_b (u64) | idx_a | idx_b | head | pos (u32) | out_ts (i64) | out_val (f64) | out_gid (i32) | cub temp
  uint8_t* scratch = nullptr;
  const size_t bytes = (size_t)n * (8 + 8 + 4 * 4 + 8 + 8 + 4) + 512 + tmp_bytes;
  CUDA_CHECK(cudaMallocAsync(&scratch, bytes, m->st));
  unsigned long long* key_a = (unsigned long long*)scratch;
  unsigned long long* key_b = key_a + n;
  long long* r_ts = (long long*)(key_b + n);
  double* r_val = (double*)(r_ts + n);
  uint32_t* idx_a = (uint32_t*)(r_val + n);
  uint32_t* idx_b = idx_a + n;
  uint32_t* head = idx_b + n;
  uint32_t* pos = head + n;
  int* r_gid = (int*)(pos + n);
  void* tmp = (void*)(((uintptr_t)(r_gid + n) + 255) & ~(uintptr_t)255);
  const int grid = (int)((n + 255) / 256);
  // merged order -> stable by gid -> stable by ts: (ts, gid, arrival) order
  mr_iota_kernel<<<grid, 256, 0, m->st>>>(idx_a, n);
  mr_gid_keys_kernel<<<grid, 256, 0, m->st>>>(m->o_gid, n, key_a);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, t1, key_a, key_b, idx_a, idx_b, (int)n, 0, 33, m->st));
  mr_ts_keys_kernel<<<grid, 256, 0, m->st>>>(m->o_ts, idx_b, n, m->reverse ? 1 : 0, key_a);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, t1, key_a, key_b, idx_b, idx_a, (int)n, 0, 64, m->st));
  mr_heads_kernel<<<grid, 256, 0, m->st>>>(key_b, m->o_gid, idx_a, n, head);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, t2, head, pos, (int)n, m->st));
  mr_fold_kernel<<<grid, 256, 0, m->st>>>(key_b, m->o_gid, m->o_val, idx_a, head, pos, n, op, m->reverse ? 1 : 0, r_ts, r_gid, r_val);
  CUDA_CHECK(cudaGetLastError());
  uint32_t last[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(&last[0], pos + (n - 1), 4, cudaMemcpyDeviceToHost, m->st));
  CUDA_CHECK(cudaMemcpyAsync(&last[1], head + (n - 1), 4, cudaMemcpyDeviceToHost, m->st));
  CUDA_CHECK(cudaStreamSynchronize(m->st));
  const size_t ng = (size_t)last[0] + last[1];
  *n_out = (int64_t)ng;
  if (out_ts) CUDA_CHECK(cudaMemcpyAsync(out_ts, r_ts, ng * 8, cudaMemcpyDeviceToHost, m->st));
  if (out_gid) CUDA_CHECK(cudaMemcpyAsync(out_gid, r_gid, ng * 4, cudaMemcpyDeviceToHost, m->st));
  if (out_val) CUDA_CHECK(cudaMemcpyAsync(out_val, r_val, ng * 8, cudaMemcpyDeviceToHost, m->st));
  CUDA_CHECK(cudaStreamSynchronize(m->st));
  CUDA_CHECK(cudaFreeAsync(scratch, m->st));
}

void merge_destroy(lk_merge* m) {
  if (!m) return;
  if (m->st) {
    cudaStreamSynchronize(m->st);
    void* ptrs[] = {m->d_ts, m->d_gid, m->d_val, m->d_offs, m->d_splits, m->d_coarse, m->d_ckey, m->o_ts, m->o_gid, m->o_val, m->o_src,
                    m->r_ctl, m->r_flags, m->r_blocks, m->r_hset, m->r_dlist, m->r_dsorted, m->r_key, m->r_start, m->r_stream, m->r_mat, m->r_parts,
                    m->m_dst, m->m_src, m->m_stream, m->m_key};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, m->st);
    cudaStreamSynchronize(m->st);
    cudaStreamDestroy(m->st);
  }
  for (auto& e : m->ev) if (e) cudaEventDestroy(e);
  if (m->ev_mid) cudaEventDestroy(m->ev_mid);
  if (m->graph) cudaGraphExecDestroy(m->graph);
  delete m;
}

}  // namespace lk

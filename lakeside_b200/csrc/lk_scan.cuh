// Fused decode + filter + step-bucket + group-by aggregate kernel for sm_100a.
//
// Work unit: one WARP owns one tile (<= 512 consecutive rows of one row group; a tile never crosses a page of any
// touched column, so every (tile, column) pair is one contiguous piece of one page).  Warps are fully independent --
// no block-level barrier anywhere -- and pull tiles from a global ticket counter.  Per tile:
//   A  definition-level words: lane l builds the 32-row def word l of a nullable column straight from the bit-packed
//      def runs (funnel shift), a 16-lane shuffle scan of the popcounts gives the row -> value-index mapping
//   B  WHERE: each lane owns 16 consecutive rows, walks the dictionary-index runs of every filter column once
//      (sequential run cursor, no per-row search), maps codes through the per-chunk class table and tests one bit
//      of the pass bitmap; survivors are compacted into shared memory with a warp prefix sum over popc
//   C  one lane per survivor: timestamp -> bucket, group-by codes -> group id, values -> (group x bucket) table:
//      dense planes (global atomics, optional warp pre-reduction of equal cells) or an open-addressing hash table of
//      32/64-byte entries (key + accumulators in one sector pair).
// This replaces DuckDB's execution of the SQL built by BaseExpr.getChartSql
// (core/src/main/scala/com/cardinal/utils/ast/BaseExpr.scala:376-403), reached from Commons.scala:240.
#pragma once
#include <cuda_runtime.h>

#include "lk_device.h"

namespace lk {

constexpr int SCAN_WARPS = 4;
constexpr int SCAN_BLOCK = SCAN_WARPS * 32;
constexpr int SCAN_ROWS_PER_LANE = LK_TILE_ROWS_MAX / 32;  // 16
constexpr int SCAN_WORDS = LK_TILE_ROWS_MAX / 32;          // 16
static_assert(SCAN_ROWS_PER_LANE == 16 && SCAN_WORDS == 16, "the kernel is written for 512-row tiles");

struct WarpSmem {
  ColCursor cur[LK_MAX_PCOLS];
  ChunkInfo ci[LK_MAX_PCOLS];
  uint32_t bits[LK_MAX_PCOLS][SCAN_WORDS];
  uint16_t pref[LK_MAX_PCOLS][SCAN_WORDS];
  uint16_t surv[LK_TILE_ROWS_MAX];
  uint32_t claims[LK_TILE_ROWS_MAX];
};

__device__ __forceinline__ bool col_pos(const WarpSmem& s, int p, uint32_t r, uint32_t& vidx) {
  const ColCursor& c = s.cur[p];
  if (c.flags & CUR_ALL_VALID) { vidx = c.vidx0 + r; return true; }
  if (c.flags & CUR_ALL_NULL) return false;
  uint32_t w = s.bits[p][r >> 5];
  uint32_t b = r & 31;
  vidx = c.vidx0 + s.pref[p][r >> 5] + __popc(w & ((1u << b) - 1));
  return (w >> b) & 1;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// one contribution into an accumulator word (all-zero = empty for every op, see lk_device.h)
__device__ __forceinline__ void acc_update(unsigned long long* word, int op, unsigned long long bits, unsigned long long cnt) {
  switch (op) {
    case AGG_SUM: atomicAdd(reinterpret_cast<double*>(word), __longlong_as_double((long long)bits)); break;
    case AGG_COUNT: atomicAdd(word, cnt); break;
    default: atomicMax(word, bits); break;  // min (complemented key) and max (key)
  }
}

// Sequential cursor over the dictionary-index runs of one column inside a tile.
struct RunCursor {
  const Run* run;      // current run
  const Run* last;     // last run of the tile
  uint32_t start, next_start, kind_value;
  const uint8_t* base;  // packed bytes of a bit-packed run
  __device__ __forceinline__ void load(const uint8_t* chunk) {
    Run r = *run;
    start = r.start;
    kind_value = r.kind_value;
    base = chunk + (r.kind_value & 0x7fffffffu);
    next_start = run < last ? run[1].start : 0xffffffffu;
  }
  __device__ __forceinline__ void seek(const uint8_t* chunk, const Run* runs, const ColCursor& c, uint32_t vidx) {
    const Run* r0 = runs + c.vrun_lo;
    last = r0 + c.vrun_n - 1;
    run = r0 + lk_find_run(r0, c.vrun_n, vidx);
    load(chunk);
  }
  __device__ __forceinline__ uint32_t code(const uint8_t* chunk, uint32_t vidx, uint32_t width, uint32_t mask) {
    while (vidx >= next_start) { run++; load(chunk); }
    if (kind_value >> 31) return kind_value & 0x7fffffffu;
    const uint32_t bitpos = (vidx - start) * width;
    const uint8_t* p = base + (bitpos >> 3);
    const uint64_t a = reinterpret_cast<uint64_t>(p);
    const unsigned sh = (unsigned)(a & 7) * 8 + (bitpos & 7);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~7ull);
    unsigned long long lo = __ldg(q);
    uint32_t x;
    if (sh + width <= 64) x = (uint32_t)(lo >> sh);
    else x = (uint32_t)((lo >> sh) | (__ldg(q + 1) << (64 - sh)));
    return x & mask;
  }
};

__global__ void __launch_bounds__(SCAN_BLOCK) scan_kernel(const __grid_constant__ ScanParams P) {
  __shared__ WarpSmem smem[SCAN_WARPS];
  const int lane = threadIdx.x & 31;
  WarpSmem& s = smem[threadIdx.x >> 5];
  const uint8_t* __restrict__ arena = P.arena;
  const Run* __restrict__ runs = P.runs;
  const unsigned lt_mask = (1u << lane) - 1;
  uint32_t my_phase_min = 0xffffffffu, my_phase_max = 0, my_status = 0;
  unsigned long long my_surv = 0;

  while (true) {
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(P.counters + 4, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= P.ntiles) break;
    const TileDesc td = P.tiles[tile];
    const uint32_t nrows = td.nrows;
    const uint32_t row0 = td.row0;
    __syncwarp();  // every lane is done with the previous tile's shared state
    if (lane < (int)P.npcols) {
      s.cur[lane] = P.cursors[td.cursor0 + lane];
      s.ci[lane] = P.chunks[(size_t)td.rg * P.npcols + lane];
    }
    __syncwarp();

    // ---- phase A: definition-level words + exclusive prefix of their popcounts (two columns per pass) ----
    const uint32_t nwords = (nrows + 31) >> 5;
    {
      uint32_t need = 0;
      for (uint32_t p = 0; p < P.npcols; p++)
        if (!(s.cur[p].flags & (CUR_ALL_VALID | CUR_ALL_NULL))) need |= 1u << p;
      while (need) {
        const int pa = __ffs(need) - 1;
        need &= need - 1;
        int pb = -1;
        if (need) { pb = __ffs(need) - 1; need &= need - 1; }
        const int p = lane < 16 ? pa : pb;
        const uint32_t w = lane & 15;
        uint32_t word = 0;
        if (p >= 0 && w < nwords) word = lk_def_word(arena, runs, s.cur[p], s.ci[p], row0 + 32 * w, min(32u, nrows - 32 * w));
        uint32_t c = __popc(word), incl = c;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, incl, d, 16);
          if ((lane & 15) >= d) incl += o;
        }
        if (p >= 0 && w < nwords) { s.bits[p][w] = word; s.pref[p][w] = (uint16_t)(incl - c); }
      }
    }
    __syncwarp();

    // ---- phase B: WHERE on dictionary codes; lane owns rows [16*lane, 16*lane + 16) ----
    const uint32_t lrow0 = (uint32_t)lane * SCAN_ROWS_PER_LANE;
    const uint32_t lrows = lrow0 >= nrows ? 0u : min((uint32_t)SCAN_ROWS_PER_LANE, nrows - lrow0);
    const uint32_t rowmask = (1u << lrows) - 1;
    uint32_t passmask = 0;
    if (lrows) {
      uint32_t idx[SCAN_ROWS_PER_LANE];
#pragma unroll
      for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) idx[j] = 0;
      for (int f = 0; f < P.n_filter; f++) {
        const int p = P.filter[f].pcol;
        const uint32_t stride = P.filter[f].stride;
        const uint32_t null_cls = P.filter[f].null_cls;
        const ColCursor& c = s.cur[p];
        uint32_t defbits, vidx;
        if (c.flags & CUR_ALL_VALID) { defbits = rowmask; vidx = c.vidx0 + lrow0; }
        else if (c.flags & CUR_ALL_NULL) { defbits = 0; vidx = 0; }
        else {
          const uint32_t w = s.bits[p][lane >> 1];
          const uint32_t sh = 16 * (lane & 1);
          defbits = (w >> sh) & rowmask;
          vidx = c.vidx0 + s.pref[p][lane >> 1] + __popc(w & ((1u << sh) - 1));
        }
        if (P.filter[f].numeric) {
          const ChunkInfo& ci = s.ci[p];
#pragma unroll 4
          for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
            uint32_t cls = null_cls;
            if ((defbits >> j) & 1) {
              uint32_t bad = 0;
              uint64_t bits = lk_value_bits(arena, runs, c, ci, vidx++, &bad);
              if (bad) my_status |= ST_BAD_CODE;
              cls = lk_numeric_class(P.filter[f], lk_bits_to_f64(bits, ci.phys_type));
            }
            idx[j] += cls * stride;
          }
        } else {
          const uint32_t dict_n = s.ci[p].dict_n;
          const uint8_t* __restrict__ lut = P.lut_cls + s.ci[p].lut_cls;
          const uint32_t width = c.width;
          const uint32_t mask = (1u << width) - 1;  // width <= 31 (checked by the host index)
          const uint8_t* chunk = arena + s.ci[p].base_off;
          RunCursor rc;
          if (defbits) rc.seek(chunk, runs, c, vidx);
#pragma unroll 4
          for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
            uint32_t cls = null_cls;
            if ((defbits >> j) & 1) {
              const uint32_t code = rc.code(chunk, vidx++, width, mask);
              if (code < dict_n) cls = __ldg(lut + code);
              else my_status |= ST_BAD_CODE;
            }
            idx[j] += cls * stride;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
        const uint32_t i = idx[j];
        passmask |= ((__ldg(P.pass_bits + (i >> 5)) >> (i & 31)) & 1u) << j;
      }
      passmask &= rowmask;
      if (passmask && P.notnull_pcol >= 0) {
        const ColCursor& c = s.cur[P.notnull_pcol];
        if (c.flags & CUR_ALL_NULL) passmask = 0;
        else if (!(c.flags & CUR_ALL_VALID)) passmask &= s.bits[P.notnull_pcol][lane >> 1] >> (16 * (lane & 1));
      }
    }
    // compaction: warp exclusive scan of the per-lane survivor counts
    uint32_t cnt = __popc(passmask), incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const uint32_t nsurv = __shfl_sync(0xffffffffu, incl, 31);
    {
      uint32_t o = incl - cnt;
      uint32_t m = passmask;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        s.surv[o++] = (uint16_t)(lrow0 + j);
      }
    }
    __syncwarp();

    // ---- phase C: one lane per survivor ----
    uint32_t nclaims = 0;
    for (uint32_t i0 = 0; i0 < nsurv; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool active = i < nsurv;
      unsigned long long cell = 0;
      unsigned long long vbits[LK_MAX_AGGS];
      bool vvalid[LK_MAX_AGGS];
#pragma unroll
      for (int a = 0; a < LK_MAX_AGGS; a++) { vbits[a] = 0; vvalid[a] = false; }
      if (active) {
        const uint32_t r = s.surv[i];
        uint32_t vidx;
        active = col_pos(s, P.ts_pcol, r, vidx);  // NULL timestamp: `ts >= S` is not TRUE
        if (active) {
          uint32_t bad = 0;
          const int64_t ts = (int64_t)lk_value_bits(arena, runs, s.cur[P.ts_pcol], s.ci[P.ts_pcol], vidx, &bad);
          active = ts >= P.ts_lo && ts < P.ts_hi;
          if (active) {
            const uint64_t rel = (uint64_t)(ts - P.base);
            const uint64_t bucket = rel / (uint64_t)P.step;
            if (P.is_metrics) {
              const uint32_t ph = (uint32_t)(rel - bucket * (uint64_t)P.step);
              my_phase_min = min(my_phase_min, ph);
              my_phase_max = max(my_phase_max, ph);
            }
            uint64_t gid = 0;
            for (int k = 0; k < P.n_keys; k++) {
              const int p = P.keys[k].pcol;
              uint32_t gcode = P.keys[k].null_code;
              if (col_pos(s, p, r, vidx)) {
                const uint32_t code = lk_dict_code(arena, runs, s.cur[p], s.ci[p], vidx);
                if (code >= s.ci[p].dict_n) bad = 1;
                else gcode = __ldg(P.lut_gcode + s.ci[p].lut_gcode + code);
              }
              gid += (uint64_t)gcode * P.keys[k].stride;
            }
            cell = bucket * P.n_groups + gid;
#pragma unroll
            for (int a = 0; a < LK_MAX_AGGS; a++) {
              if (a < P.n_aggs) {
                const int p = P.aggs[a].pcol;
                vvalid[a] = col_pos(s, p, r, vidx);
                if (vvalid[a]) {
                  const uint64_t raw = lk_value_bits(arena, runs, s.cur[p], s.ci[p], vidx, &bad);
                  const double x = lk_bits_to_f64(raw, s.ci[p].phys_type);
                  const unsigned long long xb = (unsigned long long)__double_as_longlong(x);
                  const int op = P.aggs[a].op;
                  vbits[a] = op == AGG_MIN ? lk_min_encode(xb) : op == AGG_MAX ? lk_max_encode(xb) : xb;
                }
              }
            }
            my_surv++;
          }
          if (bad) my_status |= ST_BAD_CODE;
        }
      }

      if (P.path == 0) {
        // dense planes; optionally pre-reduce lanes that hit the same cell (few groups => long same-cell runs)
        bool todo = active;
        if (P.warp_agg) {
          unsigned remaining = __ballot_sync(0xffffffffu, todo);
          for (int it = 0; it < 4 && remaining; it++) {
            const int leader = __ffs(remaining) - 1;
            const unsigned long long c = __shfl_sync(0xffffffffu, cell, leader);
            const bool mine = todo && cell == c;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (__popc(m) > 1) {
#pragma unroll
              for (int a = 0; a < LK_MAX_AGGS; a++) {
                if (a < P.n_aggs) {
                  const int op = P.aggs[a].op;
                  const bool has = mine && vvalid[a];
                  const unsigned hm = __ballot_sync(0xffffffffu, has);
                  unsigned long long red = 0;
                  if (op == AGG_SUM) {
                    double v = has ? __longlong_as_double((long long)vbits[a]) : 0.0;
#pragma unroll
                    for (int d = 16; d; d >>= 1) v += shfl_xor_f64(v, d);
                    red = (unsigned long long)__double_as_longlong(v);
                  } else if (op != AGG_COUNT) {
                    unsigned long long v = has ? vbits[a] : 0ull;
#pragma unroll
                    for (int d = 16; d; d >>= 1) { unsigned long long o = shfl_xor_u64(v, d); v = o > v ? o : v; }
                    red = v;
                  }
                  if (lane == leader && hm) acc_update(P.acc[a] + c, op, red, (unsigned long long)__popc(hm));
                }
              }
              if (lane == leader) atomicAdd(P.rowcnt + c, (unsigned long long)__popc(m));
              if (mine) todo = false;
            }
            remaining &= ~m;
          }
        }
        if (todo) {
          atomicAdd(P.rowcnt + cell, 1ull);
#pragma unroll
          for (int a = 0; a < LK_MAX_AGGS; a++)
            if (a < P.n_aggs && vvalid[a]) acc_update(P.acc[a] + cell, P.aggs[a].op, vbits[a], 1ull);
        }
      } else {
        // open addressing with linear probing; entry = {key = cell + 1, acc[n_aggs]}
        bool claimed = false;
        uint32_t claimed_slot = 0;
        if (active) {
          const unsigned long long key = cell + 1;
          uint64_t slot = lk_hash64(cell) & P.h_mask;
          unsigned long long* entry = nullptr;
          for (int probe = 0; probe < 4096; probe++) {
            unsigned long long* e = reinterpret_cast<unsigned long long*>(P.h_entries + slot * P.h_stride);
            unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(e);
            if (k == LK_EMPTY_KEY) {
              k = atomicCAS(e, (unsigned long long)LK_EMPTY_KEY, key);
              if (k == LK_EMPTY_KEY) { claimed = true; claimed_slot = (uint32_t)slot; entry = e; break; }
            }
            if (k == key) { entry = e; break; }
            slot = (slot + 1) & P.h_mask;
          }
          if (!entry) my_status |= ST_HASH_FULL;
          else {
#pragma unroll
            for (int a = 0; a < LK_MAX_AGGS; a++)
              if (a < P.n_aggs && vvalid[a]) acc_update(entry + 1 + a, P.aggs[a].op, vbits[a], 1ull);
          }
        }
        const unsigned cm = __ballot_sync(0xffffffffu, claimed);
        if (claimed) s.claims[nclaims + __popc(cm & lt_mask)] = claimed_slot;
        nclaims += __popc(cm);
      }
    }

    if (P.path == 1 && nclaims) {
      // publish the slots this tile claimed: one global atomic per tile
      __syncwarp();
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.counters + 3, nclaims);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < nclaims; i += 32)
        if (base + i < P.h_occ_cap) P.h_occ[base + i] = s.claims[i];
        else my_status |= ST_HASH_FULL;
    }
  }

  // ---- per-warp epilogue: status flags, timestamp phase range, survivor count ----
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    my_phase_min = min(my_phase_min, __shfl_xor_sync(0xffffffffu, my_phase_min, d));
    my_phase_max = max(my_phase_max, __shfl_xor_sync(0xffffffffu, my_phase_max, d));
    my_status |= __shfl_xor_sync(0xffffffffu, my_status, d);
    my_surv += shfl_xor_u64(my_surv, d);
  }
  if (lane == 0) {
    if (my_status) atomicOr(P.counters + 0, my_status);
    if (my_phase_min != 0xffffffffu) { atomicMin(P.counters + 1, my_phase_min); atomicMax(P.counters + 2, my_phase_max); }
    if (my_surv) atomicAdd(P.survivors, my_surv);
  }
}

}  // namespace lk

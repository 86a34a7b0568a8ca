// Fused decode + filter + step-bucket + group-by aggregate kernel for sm_100a.
//
// One CTA works on one tile at a time (<= 2048 consecutive rows of one row group; a tile never crosses a page of any
// touched column, so every (tile, column) pair is one contiguous piece of one page).  Phases per tile:
//   A  definition-level bitmaps: one 32-row word per thread straight from the bit-packed def runs (funnel shift),
//      then a warp scan of the word popcounts -> row -> value-index mapping for nullable columns
//   B  WHERE: every row decodes only the filter columns' dictionary codes, maps them through the per-chunk class
//      tables and tests one bit of the pass bitmap; survivors are compacted with ballot/popc into shared memory
//   C  survivors decode timestamp (-> bucket), group-by codes (-> group id) and values, then update the
//      (group x bucket) table: dense planes (global atomics, optional warp pre-reduction) or an open-addressing
//      hash table of 32/64-byte entries.
// This replaces DuckDB's execution of the SQL built by BaseExpr.getChartSql
// (core/src/main/scala/com/cardinal/utils/ast/BaseExpr.scala:376-403), reached from Commons.scala:240.
#pragma once
#include <cuda_runtime.h>

#include "lk_device.h"

namespace lk {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_WORDS = LK_TILE_ROWS_MAX / 32;

struct ScanSmem {
  TileDesc td;
  ColCursor cur[LK_MAX_PCOLS];
  ChunkInfo ci[LK_MAX_PCOLS];
  uint32_t bits[LK_MAX_PCOLS][SCAN_WORDS];
  uint16_t pref[LK_MAX_PCOLS][SCAN_WORDS];
  uint16_t surv[LK_TILE_ROWS_MAX];
  uint32_t claims[LK_TILE_ROWS_MAX];
  uint32_t nsurv, nclaims, claim_base, next_tile;
  uint32_t phase_min, phase_max, status;
};

__device__ __forceinline__ bool col_pos(const ScanSmem& s, int p, uint32_t r, uint32_t& vidx) {
  const ColCursor& c = s.cur[p];
  if (c.flags & CUR_ALL_VALID) { vidx = c.vidx0 + r; return true; }
  if (c.flags & CUR_ALL_NULL) return false;
  uint32_t w = s.bits[p][r >> 5];
  uint32_t b = r & 31;
  vidx = c.vidx0 + s.pref[p][r >> 5] + __popc(w & ((1u << b) - 1));
  return (w >> b) & 1;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

// accumulate one (already warp-reduced or single-row) contribution into accumulator words
__device__ __forceinline__ void acc_update(unsigned long long* word, int op, unsigned long long bits, unsigned long long cnt) {
  switch (op) {
    case AGG_SUM: atomicAdd(reinterpret_cast<double*>(word), __longlong_as_double((long long)bits)); break;
    case AGG_COUNT: atomicAdd(word, cnt); break;
    default: atomicMax(word, bits); break;  // min (complemented key) and max (key)
  }
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_kernel(const __grid_constant__ ScanParams P) {
  __shared__ ScanSmem s;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wid = tid >> 5;
  const uint8_t* __restrict__ arena = P.arena;
  const Run* __restrict__ runs = P.runs;
  if (tid == 0) { s.phase_min = 0xffffffffu; s.phase_max = 0; s.status = 0; }
  uint32_t my_phase_min = 0xffffffffu, my_phase_max = 0, my_status = 0;
  unsigned long long my_surv = 0;

  for (uint32_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    __syncthreads();  // previous tile fully consumed
    if (tid == 0) { s.td = P.tiles[tile]; s.nsurv = 0; s.nclaims = 0; }
    __syncthreads();
    const uint32_t nrows = s.td.nrows;
    const uint32_t row0 = s.td.row0;
    if (tid < (int)P.npcols) {
      s.cur[tid] = P.cursors[s.td.cursor0 + tid];
      s.ci[tid] = P.chunks[(size_t)s.td.rg * P.npcols + tid];
    }
    __syncthreads();

    // ---- phase A: definition-level bitmaps and word prefix sums ----
    const uint32_t nwords = (nrows + 31) >> 5;
    for (uint32_t item = tid; item < P.npcols * nwords; item += SCAN_BLOCK) {
      uint32_t p = item / nwords, w = item - p * nwords;
      if (s.cur[p].flags & (CUR_ALL_VALID | CUR_ALL_NULL)) continue;
      uint32_t nb = min(32u, nrows - 32 * w);
      s.bits[p][w] = lk_def_word(arena, runs, s.cur[p], row0 + 32 * w, nb);
    }
    __syncthreads();
    for (uint32_t p = wid; p < P.npcols; p += SCAN_BLOCK / 32) {
      if (s.cur[p].flags & (CUR_ALL_VALID | CUR_ALL_NULL)) continue;
      uint32_t running = 0;
      for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
        uint32_t w = w0 + lane;
        uint32_t c = w < nwords ? __popc(s.bits[p][w]) : 0;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (w < nwords) s.pref[p][w] = (uint16_t)(running + incl - c);
        running += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();

    // ---- phase B: WHERE clause on dictionary codes, ballot/popc compaction of the survivors ----
    const uint32_t padded = (nrows + 31) & ~31u;
    for (uint32_t r = tid; r < padded; r += SCAN_BLOCK) {
      bool pass = r < nrows;
      if (pass) {
        uint32_t idx = 0;
        for (int f = 0; f < P.n_filter; f++) {
          const FilterCol& fc = P.filter[f];
          const int p = fc.pcol;
          uint32_t vidx, cls;
          if (!col_pos(s, p, r, vidx)) cls = fc.null_cls;
          else if (fc.numeric) {
            uint32_t bad = 0;
            uint64_t bits = lk_value_bits(arena, runs, s.cur[p], s.ci[p], vidx, &bad);
            if (bad) my_status |= ST_BAD_CODE;
            cls = lk_numeric_class(fc, lk_bits_to_f64(bits, s.ci[p].phys_type));
          } else {
            uint32_t code = lk_dict_code(arena, runs, s.cur[p], vidx);
            if (code >= s.ci[p].dict_n) { my_status |= ST_BAD_CODE; cls = fc.null_cls; }
            else cls = __ldg(P.lut_cls + s.ci[p].lut_cls + code);
          }
          idx += cls * fc.stride;
        }
        pass = (__ldg(P.pass_bits + (idx >> 5)) >> (idx & 31)) & 1;
        if (pass && P.notnull_pcol >= 0) {
          uint32_t v;
          pass = col_pos(s, P.notnull_pcol, r, v);
        }
      }
      unsigned ballot = __ballot_sync(0xffffffffu, pass);
      if (ballot) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s.nsurv, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) s.surv[base + __popc(ballot & ((1u << lane) - 1))] = (uint16_t)r;
      }
    }
    __syncthreads();

    // ---- phase C: survivors -> bucket, group id, values -> aggregate table ----
    const uint32_t nsurv = s.nsurv;
    const uint32_t spad = (nsurv + 31) & ~31u;
    for (uint32_t i = tid; i < spad; i += SCAN_BLOCK) {
      bool active = i < nsurv;
      unsigned long long cell = 0;
      unsigned long long vbits[LK_MAX_AGGS];
      bool vvalid[LK_MAX_AGGS];
      if (active) {
        const uint32_t r = s.surv[i];
        uint32_t vidx;
        active = col_pos(s, P.ts_pcol, r, vidx);  // NULL timestamp: `ts >= S` is not TRUE
        if (active) {
          uint32_t bad = 0;
          int64_t ts = (int64_t)lk_value_bits(arena, runs, s.cur[P.ts_pcol], s.ci[P.ts_pcol], vidx, &bad);
          if (bad) my_status |= ST_BAD_CODE;
          active = ts >= P.ts_lo && ts < P.ts_hi;
          if (active) {
            uint64_t rel = (uint64_t)(ts - P.base);
            uint64_t bucket = rel / (uint64_t)P.step;
            if (P.is_metrics) {
              uint32_t ph = (uint32_t)(rel - bucket * (uint64_t)P.step);
              my_phase_min = min(my_phase_min, ph);
              my_phase_max = max(my_phase_max, ph);
            }
            uint64_t gid = 0;
            for (int k = 0; k < P.n_keys; k++) {
              const int p = P.keys[k].pcol;
              uint32_t gcode = P.keys[k].null_code;
              if (col_pos(s, p, r, vidx)) {
                uint32_t code = lk_dict_code(arena, runs, s.cur[p], vidx);
                if (code >= s.ci[p].dict_n) my_status |= ST_BAD_CODE;
                else gcode = __ldg(P.lut_gcode + s.ci[p].lut_gcode + code);
              }
              gid += (uint64_t)gcode * P.keys[k].stride;
            }
            cell = bucket * P.n_groups + gid;
#pragma unroll
            for (int a = 0; a < LK_MAX_AGGS; a++) {
              if (a >= P.n_aggs) break;
              const int p = P.aggs[a].pcol;
              vvalid[a] = col_pos(s, p, r, vidx);
              vbits[a] = 0;
              if (vvalid[a]) {
                uint32_t bad2 = 0;
                uint64_t raw = lk_value_bits(arena, runs, s.cur[p], s.ci[p], vidx, &bad2);
                if (bad2) my_status |= ST_BAD_CODE;
                double x = lk_bits_to_f64(raw, s.ci[p].phys_type);
                unsigned long long xb = (unsigned long long)__double_as_longlong(x);
                const int op = P.aggs[a].op;
                vbits[a] = op == AGG_MIN ? lk_min_encode(xb) : op == AGG_MAX ? lk_max_encode(xb) : xb;
              }
            }
            my_surv++;
          }
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
                if (a >= P.n_aggs) break;
                const int op = P.aggs[a].op;
                const bool has = mine && vvalid[a];
                const unsigned hm = __ballot_sync(0xffffffffu, has);
                unsigned long long red;
                if (op == AGG_SUM) {
                  double v = has ? __longlong_as_double((long long)vbits[a]) : 0.0;
#pragma unroll
                  for (int d = 16; d; d >>= 1) v += shfl_xor_f64(v, d);
                  red = (unsigned long long)__double_as_longlong(v);
                } else if (op == AGG_COUNT) {
                  red = 0;
                } else {
                  unsigned long long v = has ? vbits[a] : 0ull;
#pragma unroll
                  for (int d = 16; d; d >>= 1) { unsigned long long o = shfl_xor_u64(v, d); v = o > v ? o : v; }
                  red = v;
                }
                if (lane == leader && hm) acc_update(P.acc[a] + c, op, red, (unsigned long long)__popc(hm));
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
          for (int a = 0; a < LK_MAX_AGGS; a++) {
            if (a >= P.n_aggs) break;
            if (vvalid[a]) acc_update(P.acc[a] + cell, P.aggs[a].op, vbits[a], 1ull);
          }
        }
      } else if (active) {
        // open addressing with linear probing; entry = {key = cell + 1, acc[n_aggs]}
        const unsigned long long key = cell + 1;
        uint64_t slot = lk_hash64(cell) & P.h_mask;
        unsigned long long* entry = nullptr;
        for (int probe = 0; probe < 4096; probe++) {
          unsigned long long* e = reinterpret_cast<unsigned long long*>(P.h_entries + slot * P.h_stride);
          unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(e);
          if (k == LK_EMPTY_KEY) {
            k = atomicCAS(e, (unsigned long long)LK_EMPTY_KEY, key);
            if (k == LK_EMPTY_KEY) {
              uint32_t ci = atomicAdd(&s.nclaims, 1u);
              s.claims[ci] = (uint32_t)slot;
              entry = e;
              break;
            }
          }
          if (k == key) { entry = e; break; }
          slot = (slot + 1) & P.h_mask;
        }
        if (!entry) my_status |= ST_HASH_FULL;
        else {
#pragma unroll
          for (int a = 0; a < LK_MAX_AGGS; a++) {
            if (a >= P.n_aggs) break;
            if (vvalid[a]) acc_update(entry + 1 + a, P.aggs[a].op, vbits[a], 1ull);
          }
        }
      }
    }

    if (P.path == 1) {
      // publish the slots this tile claimed: one global atomic per tile
      __syncthreads();
      const uint32_t nc = s.nclaims;
      if (nc) {
        if (tid == 0) s.claim_base = atomicAdd(P.counters + 3, nc);
        __syncthreads();
        const uint32_t base = s.claim_base;
        for (uint32_t i = tid; i < nc; i += SCAN_BLOCK)
          if (base + i < P.h_occ_cap) P.h_occ[base + i] = s.claims[i];
          else my_status |= ST_HASH_FULL;
      }
    }
  }

  // ---- per-CTA epilogue: status flags, timestamp phase range, survivor count ----
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    my_phase_min = min(my_phase_min, __shfl_xor_sync(0xffffffffu, my_phase_min, d));
    my_phase_max = max(my_phase_max, __shfl_xor_sync(0xffffffffu, my_phase_max, d));
    my_status |= __shfl_xor_sync(0xffffffffu, my_status, d);
    my_surv += shfl_xor_u64(my_surv, d);
  }
  if (lane == 0) {
    if (my_phase_min != 0xffffffffu) { atomicMin(&s.phase_min, my_phase_min); atomicMax(&s.phase_max, my_phase_max); }
    if (my_status) atomicOr(&s.status, my_status);
    if (my_surv) atomicAdd(P.survivors, my_surv);
  }
  __syncthreads();
  if (tid == 0) {
    if (s.status) atomicOr(P.counters + 0, s.status);
    if (s.phase_min != 0xffffffffu) { atomicMin(P.counters + 1, s.phase_min); atomicMax(P.counters + 2, s.phase_max); }
  }
}

}  // namespace lk

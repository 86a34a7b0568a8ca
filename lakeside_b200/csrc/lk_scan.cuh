// Fused decode + filter + step-bucket + group-by aggregate kernel for sm_100a.
//
// Work unit: one WARP owns one tile (<= 512 consecutive rows of one row group; a tile never crosses a page of any
// touched column, so every (tile, column) pair is one contiguous piece of one page).  Warps are fully independent --
// no block-level barrier anywhere -- and pull one tile at a time from a global ticket counter (the ticket after the
// next is requested a tile ahead).  Lane l owns rows [16 l, 16 l + 16).  What is per row group (chunk
// descriptors, the per-code pass bits of phase B) is reloaded only when the row group changes.
// The kernel is latency bound with an instruction budget (profiles/README.md): the SM's instruction cache makes ~3 k SASS
// instructions per instantiation the ceiling, hence the compile-time specialisation and the rolled column loops.
//   A  definition levels: 16 bits per lane from the flat bitmap def_expand_kernel (lk_engine.cu) decoded the hybrid
//      RLE/bit-packed levels into; a warp scan of the popcounts gives every lane its first value index
//   B  WHERE: the lane slides a 64-bit window over the bit-packed dictionary indices of its rows (one funnel shift
//      per value, run changes are rare); with one filter column the class table and the pass bitmap are folded into
//      one bit per dictionary code per row group; with several, every row's class index is accumulated in shared
//      memory and looked up in the pass bitmap; survivors are compacted into shared memory by a warp prefix sum
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
static_assert(SCAN_ROWS_PER_LANE == 16, "the kernel is written for 512-row tiles");
constexpr uint32_t SCAN_CODEPASS_MAX = 1024;
#ifndef SCAN_MIN_CTAS
// resident CTAs per SM the register allocation is held to: 9 x 4 warps at 56 registers measured 4 % faster than 8 x 4 at
// 64 and 8 % faster than 10 x 4 at 48 (spills); tuning builds: make EXTRA=-DSCAN_MIN_CTAS=n
#define SCAN_MIN_CTAS 9
#endif
// Measured and rejected on B200 (r1, C2 workload, this kernel at 1.30 ms): prefetch.global.L1/.L2 of a survivor's nine
// gather addresses ahead of the dependent loads (+7 %: the extra address arithmetic costs more issue slots than the
// latency it hides); batched branch-free gathers in registers (3400 SASS instructions: +16 %, instruction fetch);
// cp.async staging of the seek index one tile ahead (long-scoreboard stalls 7.7 -> 2.2 per issue but +50 % instructions).
// After the definition bitmaps (r1i, 0.84 ms): a rolled software pipeline over the key columns (index words of key k + 1
// requested before the group-code lookup of key k: +3 %, spills) and prefetch.L2 of a survivor's PLAIN value sectors at
// the top of phase C (+-0); prefetch.global.L1 instead of .L2 for the per-tile requests (+10 %); the first four key
// columns as one unrolled batch (all index words requested before the first use: +14 %, 3.0 k instructions and spills);
// cudaLimitMaxL2FetchGranularity 32 / 128 (+-0); r1x (0.80 ms): two key columns at a time, branch-free, both columns'
// index words requested before either is used (+5 %); 10 / 8 CTAs per SM (+4 % / +5 %); r2c (0.72 ms): record slots
// reserved right after the timestamp check so that the counter's round trip overlaps the key / value decode (+2 %:
// the reservation's three values live across the whole decode and spill).

struct WarpSmem {
  ColCursor cur[LK_MAX_PCOLS];
  ChunkInfo ci[LK_MAX_PCOLS];
  uint16_t defb[LK_MAX_PCOLS][32];  // definition bits of the 16 rows of every lane (every column: all-valid tiles hold ones)
  uint16_t vpre[LK_MAX_PCOLS][32];  // non-null values of the tile before the lane's first row
  alignas(16) uint32_t vrs[LK_MAX_PCOLS][4];  // start index of the tile's first 4 dictionary-index runs (0xffffffff = none)
  uint32_t vrk[LK_MAX_PCOLS][4];    // RLE run: the repeated index; bit-packed run: 8 * byte offset - start * width, so that
                                    // index v of the run sits vrk + v * width bits after the chunk's first byte
  uint8_t vrle[LK_MAX_PCOLS];       // bit j set <=> run j is an RLE run
  uint32_t codepass[SCAN_CODEPASS_MAX / 32];  // single filter column: bit c set <=> dictionary code c passes the WHERE
  uint16_t surv[LK_TILE_ROWS_MAX];
};

// the generic (several filter columns) predicate also keeps one class index per row
template <bool GENERIC>
struct WarpSmemT : WarpSmem {
  uint32_t fidx[GENERIC ? LK_TILE_ROWS_MAX : 1];
};

// 32 bits starting `bit` bits after byte address p (bit may exceed 7)
__device__ __forceinline__ uint32_t load_bits32(const uint8_t* __restrict__ p, uint32_t bit) {
  const uint64_t a = reinterpret_cast<uint64_t>(p + (bit >> 3));
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~3ull);
  const uint32_t sh = (uint32_t)(a & 3) * 8 + (bit & 7);
  return __funnelshift_r(__ldg(q), __ldg(q + 1), sh);
}

// request a line the tile will read a few hundred instructions from now (tuning builds: make EXTRA=-DSCAN_PREFETCH_L1)
__device__ __forceinline__ void prefetch_line(const void* p) {
#ifdef SCAN_PREFETCH_L1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// (valid, value index) of row r of column p
__device__ __forceinline__ bool col_pos(const WarpSmem& s, int p, uint32_t r, uint32_t& vidx) {
  // branch-free: phase A leaves defb / vpre filled for every column (all-valid tiles: ones and 16 * lane)
  const uint32_t db = s.defb[p][r >> 4];
  const uint32_t j = r & 15;
  vidx = s.cur[p].vidx0 + s.vpre[p][r >> 4] + __popc(db & ((1u << j) - 1));
  return (db >> j) & 1;
}

// index (0..3) of the dictionary-index run holding value `vidx`, for columns with at most 4 runs in the tile
__device__ __forceinline__ uint32_t fast_run(const WarpSmem& s, int p, uint32_t vidx) {
  const uint4 st = *reinterpret_cast<const uint4*>(s.vrs[p]);  // one 16-byte shared load
  return (uint32_t)(vidx >= st.y) + (uint32_t)(vidx >= st.z) + (uint32_t)(vidx >= st.w);
}

// more than 4 runs of the column in the tile (rare): binary search in the global run pool; one out-of-line copy
__device__ __noinline__ uint32_t dict_code_slow(const uint8_t* arena, const Run* runs, const ColCursor* c, const ChunkInfo* ci, uint32_t vidx) {
  return lk_dict_code(arena, runs, *c, *ci, vidx);
}

// dictionary index of value `vidx`: run descriptors from shared memory when the tile has <= 4 runs of this column
__device__ __forceinline__ uint32_t dict_code(const WarpSmem& s, int p, const uint8_t* __restrict__ arena, const Run* __restrict__ runs, uint32_t vidx) {
  const ColCursor& c = s.cur[p];
  if (c.vrun_n > 4) return dict_code_slow(arena, runs, &c, &s.ci[p], vidx);
  const uint32_t ri = fast_run(s, p, vidx);
  const uint32_t kv = s.vrk[p][ri];
  if ((s.vrle[p] >> ri) & 1) return kv;
  const uint32_t w = c.width;
  const uint32_t bit = kv + vidx * w;  // from the chunk's first byte, which is 256-byte aligned
  const uint32_t* q = reinterpret_cast<const uint32_t*>(arena + s.ci[p].base_off) + (bit >> 5);
  return __funnelshift_r(__ldg(q), __ldg(q + 1), bit) & ((1u << w) - 1);
}

__device__ __forceinline__ uint64_t value_bits(const WarpSmem& s, int p, const uint8_t* __restrict__ arena, const Run* __restrict__ runs, uint32_t vidx,
                                               uint32_t& bad) {
  const ColCursor& c = s.cur[p];
  const ChunkInfo& ci = s.ci[p];
  const unsigned esz = (ci.phys_type == 1 || ci.phys_type == 4) ? 4 : 8;
  uint64_t off;
  if (c.flags & CUR_DICT) {
    uint32_t code = dict_code(s, p, arena, runs, vidx);
    if (code >= ci.dict_n) { bad = 1; code = 0; }
    off = ci.dict_off + (uint64_t)code * esz;
  } else {
    off = c.plain_off + (uint64_t)(vidx - c.vidx0) * esz;
  }
  const uint64_t x = lk_load_u64(arena, off);
  return esz == 4 ? (x & 0xffffffffull) : x;
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

// Sequential reader of consecutive dictionary indices of one column: a 64-bit window (two 32-bit words) slides over a
// bit-packed run, one funnel shift per value; run changes are rare (a bit-packed run holds up to 504 values).
struct SeqReader {
  const Run* r0;       // the column's runs in the global pool (more than 4 runs in the tile)
  const uint32_t* fs;  // or: starts / kind words of its <= 4 runs in shared memory (absent runs start at 0xffffffff)
  const uint32_t* fk;
  const uint8_t* chunk;
  const uint32_t* wbase;  // the chunk's first word; `cur` is word widx of it
  uint32_t widx;
  uint32_t ri, n, next_start, cur, nxt, sh, width, mask, frle;
  uint32_t adv;  // bits one index occupies: `width` in a bit-packed run, 0 in an RLE run (whose window holds the value)
  __device__ __forceinline__ void open_run(uint32_t vidx) {
    uint32_t bitpos;  // of index vidx, from the chunk's first byte (4-byte aligned)
    uint32_t rle;
    bool is_rle;
    if (fs) {
      rle = fk[ri];
      is_rle = (frle >> ri) & 1;
      bitpos = rle + vidx * width;
      next_start = ri < 3 ? fs[ri + 1] : 0xffffffffu;
    } else {
      const Run r = r0[ri];
      next_start = ri + 1 < n ? r0[ri + 1].start : 0xffffffffu;
      is_rle = r.kind_value >> 31;
      rle = r.kind_value & 0x7fffffffu;
      bitpos = r.kind_value * 8 + (vidx - r.start) * width;
    }
    // an RLE run is a window that never moves: the repeated index sits in `cur`, every step advances by 0 bits
    sh = 0;
    adv = 0;
    cur = rle;
    nxt = 0;
    if (!is_rle) {
      wbase = reinterpret_cast<const uint32_t*>(chunk);
      widx = bitpos >> 5;
      sh = bitpos & 31;
      adv = width;
      cur = __ldg(wbase + widx);
      nxt = __ldg(wbase + widx + 1);
    }
  }
  __device__ __forceinline__ void seek(const WarpSmem& s, int p, const uint8_t* arena, const Run* runs, uint32_t vidx) {
    const ColCursor& c = s.cur[p];
    r0 = runs + c.vrun_lo;
    n = c.vrun_n;
    const bool fast = n <= 4;
    fs = fast ? s.vrs[p] : nullptr;
    fk = s.vrk[p];
    frle = s.vrle[p];
    chunk = arena + s.ci[p].base_off;
    width = c.width;
    mask = (1u << width) - 1;  // width <= 31 (checked by the host index)
    ri = fast ? fast_run(s, p, vidx) : lk_find_run(r0, n, vidx);
    open_run(vidx);
  }
  // the index under the cursor / step over it (take = false: stay, for a NULL row)
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(cur, nxt, sh) & mask; }
  __device__ __forceinline__ void advance(bool take) {
    sh += take ? adv : 0u;
    if (sh >= 32) { sh -= 32; cur = nxt; nxt = __ldg(wbase + (++widx) + 1); }
  }
  __device__ __forceinline__ uint32_t next(uint32_t vidx) {
    if (vidx >= next_start) { ri++; open_run(vidx); }
    const uint32_t code = peek();
    advance(true);
    return code;
  }
};

// Specialised at compile time on the aggregate-table layout, the single-filter-column fast path and the record-emit
// mode of exact_sums: every instantiation carries only its own code (the generic kernel overflowed the instruction
// cache: ncu showed 4.2 no-instruction stall cycles per issue).
// NA bounds the unrolled aggregate slots (4 or LK_MAX_AGGS): each slot is a full copy of the value decode.
// LK_SCAN_STOP_AFTER (profiling aid: leave each tile after the prologue / phase A / phase B, skip the table update) is
// only compiled into tuning builds (make EXTRA=-DLK_SCAN_PROFILING): the checks cost ~1 % of the instructions
#ifdef LK_SCAN_PROFILING
#define STOP_AFTER P.stop_after
#else
#define STOP_AFTER 0
#endif
template <int PATH, bool SINGLE, bool EMIT, int NA>
__global__ void __launch_bounds__(SCAN_BLOCK, SCAN_MIN_CTAS) scan_kernel(const __grid_constant__ ScanParams P) {
  __shared__ WarpSmemT<!SINGLE> smem[SCAN_WARPS];
  const int lane = threadIdx.x & 31;
  WarpSmemT<!SINGLE>& s = smem[threadIdx.x >> 5];
  const uint8_t* __restrict__ arena = P.arena;
  const Run* __restrict__ runs = P.runs;
  const unsigned lt_mask = (1u << lane) - 1;
  uint32_t my_phase_min = 0xffffffffu, my_phase_max = 0, my_status = 0;
  unsigned long long my_surv = 0;

  // tiles are handed out one at a time by a global ticket counter (B200, C2: 8 tiles per ticket 0.885 ms, 4: 0.834, 2: 0.814,
  // 1: 0.797 -- a tile is ~20 us of a warp's time, so coarser tickets leave a visible tail).  The ticket after the next is
  // requested a tile ahead (lane 0 holds it, nobody waits for it), and the next tile's descriptor and cursors are
  // prefetched while this one is processed.
  // (the atomic's result goes straight into lane 0's `pending`: nothing merges it with another value, so nothing waits)
  uint32_t pending = 0;
  if (lane == 0) pending = atomicAdd(P.counters + 4, 1u);
  uint32_t tile = __shfl_sync(0xffffffffu, pending, 0);
  if (lane == 0) pending = atomicAdd(P.counters + 4, 1u);
  uint32_t cached_rg = 0xffffffffu;
  uint32_t filled_valid = 0, filled_null = 0;  // columns whose defb / vpre rows hold the all-valid / all-NULL pattern

  for (;;) {
    if (tile >= P.ntiles) break;
    const uint32_t tile_next = __shfl_sync(0xffffffffu, pending, 0);  // requested one tile ago
    if (lane == 0) pending = atomicAdd(P.counters + 4, 1u);
    if (tile_next < P.ntiles) {
      const uint8_t* c0 = reinterpret_cast<const uint8_t*>(P.cursors + (size_t)tile_next * P.npcols);
      if ((uint32_t)lane * 128 < P.npcols * (uint32_t)sizeof(ColCursor) + 127) prefetch_line(c0 + lane * 128);
      if (lane == 31) prefetch_line(P.tiles + tile_next);
    }
    const uint32_t tile_cur = tile;
    tile = tile_next;  // (`continue` below moves on to it)
    const TileDesc td = P.tiles[tile_cur];
    const uint32_t nrows = td.nrows;
    const uint32_t row0 = td.row0;
    __syncwarp();  // every lane is done with the previous tile's shared state
    uint32_t need = 0;      // columns with some (not all) NULLs in this tile
    uint32_t allvalid = 0;  // columns without NULLs in this tile
    {
      bool mine = false, mine_valid = false;
      if (lane < (int)P.npcols) {
        const ColCursor c = P.cursors[(size_t)tile_cur * P.npcols + lane];  // == td.cursor0 + lane: no wait for td
        s.cur[lane] = c;
        if (td.rg != cached_rg) s.ci[lane] = P.chunks[(size_t)td.rg * P.npcols + lane];
        mine = !(c.flags & (CUR_ALL_VALID | CUR_ALL_NULL));
        mine_valid = c.flags & CUR_ALL_VALID;
        // the tile's 512 definition bits of this column (phase A reads them lane by lane): on their way while the
        // run descriptors below are fetched
        if (mine) {
          const uint32_t* dw = P.defbm + s.ci[lane].defbm_word0 + (td.row0 >> 5);
          prefetch_line(dw);
          prefetch_line(dw + 16);
        }
        if ((c.flags & CUR_DICT) && c.nvals) {
          Run r0v;
          uint32_t rle_mask = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            Run r;
            r.start = 0xffffffffu;
            r.kind_value = 0;
            if (j < (int)c.vrun_n) r = runs[c.vrun_lo + j];
            const uint32_t rl = r.kind_value >> 31;
            s.vrs[lane][j] = r.start;
            s.vrk[lane][j] = rl ? (r.kind_value & 0x7fffffffu) : r.kind_value * 8u - r.start * (uint32_t)c.width;
            rle_mask |= rl << j;
            if (j == 0) r0v = r;
          }
          s.vrle[lane] = (uint8_t)rle_mask;
          // every line of a dictionary-coded column's indices is needed by phase B / C (0.5-1 byte per row): request the
          // tile's lines now, one column per lane, instead of one dependent DRAM round trip per column later
          if (!(r0v.kind_value >> 31)) {
            const uint8_t* cb = arena + s.ci[lane].base_off + r0v.kind_value + (((c.vidx0 - r0v.start) * (uint32_t)c.width) >> 3);
            const uint32_t nbytes = (c.nvals * (uint32_t)c.width + 7) >> 3;
            for (uint32_t o = 0; o < nbytes + 127; o += 128) prefetch_line(cb + o);
          }
        }
      }
      need = __ballot_sync(0xffffffffu, mine);
      allvalid = __ballot_sync(0xffffffffu, mine_valid);
    }
    const bool new_rg = td.rg != cached_rg;  // chunk descriptors (and the per-code pass bits) are per row group
    cached_rg = td.rg;
    __syncwarp();
    if (STOP_AFTER == 1) continue;

    const uint32_t lrow0 = (uint32_t)lane * SCAN_ROWS_PER_LANE;
    const uint32_t lrows = lrow0 >= nrows ? 0u : min((uint32_t)SCAN_ROWS_PER_LANE, nrows - lrow0);
    const uint32_t rowmask = (1u << lrows) - 1;

    // ---- phase A: definition bits of my 16 rows + value-index prefix of every nullable column ----
    auto def_bits = [&](int p) -> uint32_t {
      return load_bits32(reinterpret_cast<const uint8_t*>(P.defbm + s.ci[p].defbm_word0), row0 + lrow0);
    };
    const uint32_t need_all = need;
    // three columns per warp scan: the lanes' popcounts (<= 16, prefix <= 512) ride in 10-bit fields of one word, and the
    // three columns' loads are in flight together
    while (need) {
      const int p0 = __ffs(need) - 1;
      need &= need - 1;
      const int p1 = need ? __ffs(need) - 1 : -1;
      need &= need - 1;  // (0 stays 0)
      const int p2 = need ? __ffs(need) - 1 : -1;
      need &= need - 1;
      const uint32_t b0 = def_bits(p0) & rowmask;
      const uint32_t b1 = p1 >= 0 ? def_bits(p1) & rowmask : 0u;
      const uint32_t b2 = p2 >= 0 ? def_bits(p2) & rowmask : 0u;
      const uint32_t cnt = __popc(b0) | (__popc(b1) << 10) | (__popc(b2) << 20);
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const uint32_t excl = incl - cnt;
      s.defb[p0][lane] = (uint16_t)b0;
      s.vpre[p0][lane] = (uint16_t)(excl & 1023u);
      if (p1 >= 0) {
        s.defb[p1][lane] = (uint16_t)b1;
        s.vpre[p1][lane] = (uint16_t)((excl >> 10) & 1023u);
      }
      if (p2 >= 0) {
        s.defb[p2][lane] = (uint16_t)b2;
        s.vpre[p2][lane] = (uint16_t)(excl >> 20);
      }
    }
    // the other columns get the same shape (all valid: ones and 16 * lane; all NULL: zeros), so that nothing downstream
    // branches on the cursor flags
    // (their patterns do not depend on the tile -- rows beyond the tile are masked where it matters -- so a column is
    // only rewritten when its kind changes: filled_valid / filled_null remember what the warp's rows hold)
    {
      const uint32_t allnull = ((1u << P.npcols) - 1) & ~need_all & ~allvalid;
      for (uint32_t m = allvalid & ~filled_valid; m; m &= m - 1) {
        const int p = __ffs(m) - 1;
        s.defb[p][lane] = 0xffffu;
        s.vpre[p][lane] = (uint16_t)lrow0;
      }
      for (uint32_t m = allnull & ~filled_null; m; m &= m - 1) {
        const int p = __ffs(m) - 1;
        s.defb[p][lane] = 0;
        s.vpre[p][lane] = 0;
      }
      filled_valid = allvalid;
      filled_null = allnull;
    }
    __syncwarp();
    if (STOP_AFTER == 2) continue;

    // ---- phase B: WHERE on dictionary codes ----
    uint32_t passmask = 0;
    // definition bits and first value index of this lane's rows in column p
    auto lane_def = [&](int p, uint32_t& defbits, uint32_t& vidx) {
      defbits = s.defb[p][lane] & rowmask;
      vidx = s.cur[p].vidx0 + s.vpre[p][lane];
    };
    // SINGLE: one string filter column whose dictionaries all have <= SCAN_CODEPASS_MAX entries (checked by the host)
    if constexpr (SINGLE) {
      // one filter column: fold class table and pass bitmap into one bit per dictionary code (per tile: every row
      // group has its own dictionary); a row passes iff bit[code] -- or the NULL class bit -- is set
      const int p = P.filter[0].pcol;
      const uint32_t dict_n = s.ci[p].dict_n;
      const uint8_t* __restrict__ lut = P.lut_cls + s.ci[p].lut_cls;
      for (uint32_t c0 = 0; new_rg && c0 < dict_n; c0 += 32) {  // the bits only change with the dictionary
        const uint32_t code = c0 + lane;
        bool ok = false;
        if (code < dict_n) {
          const uint32_t cls = __ldg(lut + code);
          ok = (__ldg(P.pass_bits + (cls >> 5)) >> (cls & 31)) & 1;
        }
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s.codepass[c0 >> 5] = b;
      }
      __syncwarp();
      if (lrows) {
        const uint32_t ncls = P.filter[0].null_cls;
        const uint32_t nullpass = (__ldg(P.pass_bits + (ncls >> 5)) >> (ncls & 31)) & 1u;
        uint32_t defbits, vidx;
        lane_def(p, defbits, vidx);
        passmask = nullpass ? (~defbits & rowmask) : 0u;
        if (defbits) {
          SeqReader rd;
          rd.seek(s, p, arena, runs, vidx);
          rd.mask &= SCAN_CODEPASS_MAX - 1;  // (the host admits only dictionaries of <= SCAN_CODEPASS_MAX entries here)
          // one iteration per ROW slot (no find-first-set walk over the valid rows, no divergence between lanes); a NULL
          // row looks at the index under the cursor, ignores it and does not step over it.  The pass bits are shifted in
          // from the top: after 16 slots they sit in bits 16..31.
          uint32_t maxcode = 0, acc = 0;
#pragma unroll 4
          for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
            const uint32_t v = (defbits >> j) & 1u;
            if (v && vidx >= rd.next_start) { rd.ri++; rd.open_run(vidx); }
            const uint32_t code = rd.peek();
            maxcode = max(maxcode, v ? code : 0u);
            acc = __funnelshift_r(acc, (s.codepass[code >> 5] >> (code & 31)) & v, 1);
            vidx += v;
            rd.advance(v);
          }
          passmask |= acc >> 16;
          if (maxcode >= dict_n) my_status |= ST_BAD_CODE;
        }
      }
    } else if (lrows) {
      // generic predicate: class index over all filter columns, accumulated per row in shared memory (row j of this
      // lane at fidx[32 j + lane]: conflict-free) so that every loop stays rolled -- the unrolled form was a 10 k-
      // instruction kernel, far beyond the instruction cache
      uint32_t* const fidx = s.fidx + lane;
#pragma unroll 1
      for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) fidx[32 * j] = 0;
#pragma unroll 1
      for (int f = 0; f < P.n_filter; f++) {
        const int p = P.filter[f].pcol;
        const uint32_t stride = P.filter[f].stride;
        const uint32_t null_cls = P.filter[f].null_cls;
        uint32_t defbits, vidx;
        lane_def(p, defbits, vidx);
        if (P.filter[f].numeric) {
#pragma unroll 1
          for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
            uint32_t cls = null_cls;
            if ((defbits >> j) & 1) {
              uint32_t bad = 0;
              const uint64_t bits = value_bits(s, p, arena, runs, vidx++, bad);
              if (bad) my_status |= ST_BAD_CODE;
              cls = lk_numeric_class(P.filter[f], lk_bits_to_f64(bits, s.ci[p].phys_type));
            }
            fidx[32 * j] += cls * stride;
          }
        } else {
          const uint32_t dict_n = s.ci[p].dict_n;
          const uint8_t* __restrict__ lut = P.lut_cls + s.ci[p].lut_cls;
          SeqReader rd;
          if (defbits) rd.seek(s, p, arena, runs, vidx);
#pragma unroll 1
          for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
            uint32_t cls = null_cls;
            if ((defbits >> j) & 1) {
              const uint32_t code = rd.next(vidx++);
              if (code < dict_n) cls = __ldg(lut + code);
              else my_status |= ST_BAD_CODE;
            }
            fidx[32 * j] += cls * stride;
          }
        }
      }
#pragma unroll 1
      for (int j = 0; j < SCAN_ROWS_PER_LANE; j++) {
        const uint32_t i = fidx[32 * j];
        passmask |= ((__ldg(P.pass_bits + (i >> 5)) >> (i & 31)) & 1u) << j;
      }
    }
    passmask &= rowmask;
    if (P.notnull_pcol >= 0) passmask &= s.defb[P.notnull_pcol][lane];
    // compaction: warp exclusive scan of the per-lane survivor counts
    uint32_t nsurv;
    {
      uint32_t cnt = __popc(passmask), incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      nsurv = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t o = incl - cnt;
      for (uint32_t m = passmask; m; m &= m - 1) s.surv[o++] = (uint16_t)(lrow0 + __ffs(m) - 1);
    }
    __syncwarp();
    if (STOP_AFTER == 3) { my_surv += (lane == 0) ? nsurv : 0; continue; }

    // ---- phase C: one lane per survivor ----
    for (uint32_t i0 = 0; i0 < nsurv; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool active = i < nsurv;
      unsigned long long cell = 0, seq = 0;
      uint64_t slot = 0;
      uint32_t bucket32 = 0;
      unsigned long long probe_key = 1;  // neither empty nor a key: only looked at by lanes that set it
      (void)slot;
      (void)probe_key;
      unsigned long long vbits[NA];
      bool vvalid[NA];
#pragma unroll
      for (int a = 0; a < NA; a++) { vbits[a] = 0; vvalid[a] = false; }
      if (active) {
        const uint32_t r = s.surv[i];
        uint32_t vidx;
        active = col_pos(s, P.ts_pcol, r, vidx);  // NULL timestamp: `ts >= S` is not TRUE
        if (active) {
          uint32_t bad = 0;
          const int64_t ts = (int64_t)value_bits(s, P.ts_pcol, arena, runs, vidx, bad);
          active = ts >= P.ts_lo && ts < P.ts_hi;
          if (active) {
            const uint64_t rel = (uint64_t)(ts - P.base);
            uint64_t bucket;
            uint32_t ph;
            if (P.fits32) {  // (endTs - base) and step fit 32 bits: one 32-bit division instead of a 64-bit one
              const uint32_t b32 = (uint32_t)rel / (uint32_t)P.step;
              ph = (uint32_t)rel - b32 * (uint32_t)P.step;
              bucket = b32;
            } else {
              bucket = rel / (uint64_t)P.step;
              ph = (uint32_t)(rel - bucket * (uint64_t)P.step);
            }
            if (P.is_metrics) {
              my_phase_min = min(my_phase_min, ph);
              my_phase_max = max(my_phase_max, ph);
            }
            uint64_t gid = 0;
#pragma unroll 1
            for (int k = 0; k < P.n_keys; k++) {
              const int p = P.keys[k].pcol;
              uint32_t gcode = P.keys[k].null_code;
              if (col_pos(s, p, r, vidx)) {
                const uint32_t code = dict_code(s, p, arena, runs, vidx);
                if (code >= s.ci[p].dict_n) bad = 1;
                else gcode = __ldg(P.lut_gcode + s.ci[p].lut_gcode + code);
              }
              gid += (uint64_t)gcode * P.keys[k].stride;
            }
            // record path: the key keeps bucket and group id in separate bit fields (finalize bins on them without a division)
            if constexpr (PATH >= 2) cell = (bucket << P.rec_gid_bits) | gid;
            else cell = bucket * P.n_groups + gid;
            bucket32 = (uint32_t)bucket;
            seq = P.seq_offset + s.ci[P.ts_pcol].seq_base + row0 + r;
            if constexpr (PATH == 1 && !EMIT) {
              // first probe of the hash table: in flight while the values are gathered
              slot = lk_hash64(cell) & P.h_mask;
              if (STOP_AFTER < 4) probe_key = *reinterpret_cast<volatile unsigned long long*>(P.h_entries + slot * P.h_stride);
            }
#pragma unroll
            for (int a = 0; a < NA; a++) {
              if (a < P.n_aggs && STOP_AFTER != 5) {
                const int p = P.aggs[a].pcol;
                vvalid[a] = col_pos(s, p, r, vidx);
                if (vvalid[a]) {
                  const uint64_t raw = value_bits(s, p, arena, runs, vidx, bad);
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

      if constexpr (EMIT && PATH < 2) {
        // fixed-order summation pass: the survivors are not aggregated here but written out as (cell, global row
        // sequence, value) records; they are sorted and folded strictly in row order afterwards (lk_exact.cu)
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (am) {
          const int leader = __ffs(am) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(P.counters + 5, (uint32_t)__popc(am));
          base = __shfl_sync(0xffffffffu, base, leader);
          const uint32_t o = base + __popc(am & lt_mask);
          if (active && o < P.rec_cap) {
            P.rec_cell[o] = cell;
            P.rec_seq[o] = seq;
#pragma unroll
            for (int a = 0; a < NA; a++)
              if (a < P.n_aggs && P.aggs[a].op == AGG_SUM) P.rec_val[a][o] = vvalid[a] ? vbits[a] : 0ull;
          }
        }
        continue;
      }
      if (STOP_AFTER >= 4) {  // profiling aid: no table update
        unsigned long long x = cell;
#pragma unroll
        for (int a = 0; a < NA; a++) x ^= vbits[a];
        if (x == 0x123456789abcdefull) my_status |= 4;
        continue;
      }
      if constexpr (PATH == 0) {
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
              for (int a = 0; a < NA; a++) {
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
          for (int a = 0; a < NA; a++)
            if (a < P.n_aggs && vvalid[a]) acc_update(P.acc[a] + cell, P.aggs[a].op, vbits[a], 1ull);
        }
      } else if constexpr (PATH == 1) {
        // open addressing with linear probing; entry = {key = cell + 1, acc[n_aggs]}.  A plain load first: an atomic on a
        // line that is not yet in L2 measured slower than load + CAS (B200, r1).
        bool claimed = false;
        uint32_t claimed_slot = 0;
        if (active) {
          const unsigned long long key = cell + 1;
          unsigned long long* entry = nullptr;
          unsigned long long k = probe_key;
          for (int probe = 0; probe < 4096; probe++) {
            unsigned long long* e = reinterpret_cast<unsigned long long*>(P.h_entries + slot * P.h_stride);
            if (probe) k = *reinterpret_cast<volatile unsigned long long*>(e);
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
            for (int a = 0; a < NA; a++)
              if (a < P.n_aggs && vvalid[a]) acc_update(entry + 1 + a, P.aggs[a].op, vbits[a], 1ull);
          }
        }
        // publish the slots claimed by this batch of survivors: one global atomic per warp iteration
        const unsigned cm = __ballot_sync(0xffffffffu, claimed);
        if (cm) {
          const int leader = __ffs(cm) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(P.counters + 3, (uint32_t)__popc(cm));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (claimed) {
            const uint32_t o = base + __popc(cm & lt_mask);
            if (o < P.h_occ_cap) { P.h_occ[o] = claimed_slot; P.h_bkt[o] = bucket32; }
            else my_status |= ST_HASH_FULL;
          }
        }
      } else if constexpr (PATH == 2) {
        // record path (selective filter, high-cardinality result): nothing is aggregated here.  Every survivor appends
        // its (bucket, group) key to rec_cell[] and its n_aggs accumulator words (same encodings as the tables: all-zero = no value)
        // to one row of rec_vals[]; finalize groups equal keys and writes the rows.  Appends are sequential, coalesced writes:
        // no 8 GB table, no random sector per survivor, nothing to clear afterwards.
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (am) {
          const int leader = __ffs(am) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(P.counters + 5, (uint32_t)__popc(am));
          base = __shfl_sync(0xffffffffu, base, leader);
          const uint32_t o = base + __popc(am & lt_mask);
          if (active) {
            if (o < P.rec_cap) {
              P.rec_cell[o] = cell << P.rec_idx_bits;
              // EMIT (exact_sums): the row starts with the record's global sequence number; finalize sorts by (key, sequence)
              unsigned long long* rec = P.rec_vals + (size_t)o * (P.n_aggs + (EMIT ? 1 : 0));
              if constexpr (EMIT) *rec++ = seq;
#pragma unroll
              for (int a = 0; a < NA; a++)
                if (a < P.n_aggs) rec[a] = !vvalid[a] ? 0ull : P.aggs[a].op == AGG_COUNT ? 1ull : vbits[a];
            } else my_status |= ST_HASH_FULL;
          }
        }
      } else {
        // sharded record path: the record goes straight into this rank's region of the receive pool of the rank that owns its
        // cell (stores over NVLink / NVSwitch, fire and forget: the transfer overlaps the scan).  The lanes of one
        // destination elect a leader that takes their slots from a LOCAL counter; the accumulator row travels as one 32-byte
        // store (a peer write is a link transaction of its own: four 8-byte stores per record quadrupled them).
        const XchgParams& X = P.x;
        uint32_t dest = 0xffffffffu;
        if (active) dest = X.world > 1 ? __umulhi((uint32_t)(lk_hash64(cell) >> 32), X.world) : 0u;  // (independent of the key table's hash)
        const unsigned peers = __match_any_sync(0xffffffffu, dest);
        if (__any_sync(0xffffffffu, active)) {
          const int leader = __ffs(peers) - 1;
          uint32_t pos = 0;
          if (active && lane == leader) pos = atomicAdd(X.count + dest, (uint32_t)__popc(peers));
          pos = __shfl_sync(0xffffffffu, pos, leader);
          if (active) {
            const uint32_t o = pos + (uint32_t)__popc(peers & lt_mask);
            if (o < X.region_cap) {
              X.keys[dest][o] = cell;
              unsigned long long w[NA];
#pragma unroll
              for (int a = 0; a < NA; a++) w[a] = (a >= P.n_aggs || !vvalid[a]) ? 0ull : P.aggs[a].op == AGG_COUNT ? 1ull : vbits[a];
              unsigned long long* rec = X.vals[dest] + (size_t)o * (P.n_aggs + (EMIT ? 1 : 0));
              if constexpr (EMIT) *rec++ = seq;  // exact_sums: sequence number first (finalize sorts by (key, sequence))
              if (!EMIT && NA == 4 && P.n_aggs == 4) {
                asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "l"(w[0]), "l"(w[1]), "l"(w[2]), "l"(w[3]) : "memory");
              } else {
#pragma unroll
                for (int a = 0; a < NA; a++)
                  if (a < P.n_aggs) rec[a] = w[a];
              }
            } else my_status |= ST_HASH_FULL;  // this rank's region of the owner's pool is full: reported, the query fails
          }
        }
      }
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

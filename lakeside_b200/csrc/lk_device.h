// Structures shared by the host planner and the sm_100a kernels, plus the per-value decode primitives
// (host/device inline so the CPU unit tests can exercise exactly the arithmetic the kernels run).
//
// HBM layout of a prepared query ("arena"): the raw bytes of every touched Parquet column chunk, exactly as they are
// in the file (dictionary page + data pages, page headers included), each chunk at a 256-byte aligned offset.
// Nothing is re-encoded on the host; PLAIN values are therefore at arbitrary byte alignment and are read with
// lk_load_u64 (two aligned 8-byte loads + funnel shift).  Around it, small index pools:
//   runs[]     one 16-byte entry per run of every RLE/bit-packed hybrid stream (definition levels, dictionary indices)
//   tiles[]    one entry per tile (<= tile_rows consecutive rows of one row group, never crossing a page of any column)
//   cursors[]  per (tile, column): where the column starts inside the tile
//   chunks[]   per (row group, column): dictionary location and the offsets of its code -> class / group-code tables
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LK_HD __host__ __device__ __forceinline__
#else
#define LK_HD inline
#endif

namespace lk {

constexpr int LK_MAX_PCOLS = 16;   // distinct physical columns a query may touch
constexpr int LK_MAX_FILTER = 8;   // distinct filter columns
constexpr int LK_MAX_KEYS = 8;     // name + group-by columns
constexpr int LK_MAX_AGGS = 7;     // aggregates of one fused pass
constexpr int LK_MAX_NUMLEAF = 4;  // numeric comparison leaves per column
constexpr int LK_TILE_ROWS_MAX = 512;  // one warp owns one tile: 16 consecutive rows per lane
constexpr uint64_t LK_EMPTY_KEY = 0ull;  // hash entries store cell + 1; an all-zero arena is clean

enum AggOp : uint8_t { AGG_SUM = 0, AGG_COUNT = 1, AGG_MIN = 2, AGG_MAX = 3 };

struct Run {
  uint32_t start;       // chunk-level index of the run's first element (row for def levels, value index for codes)
  uint32_t kind_value;  // bit 31 set: RLE run, bits 0..30 = the repeated value;
                        // clear: bit-packed run, bits 0..30 = byte offset of its packed bytes from the chunk's first byte
};
static_assert(sizeof(Run) == 8, "Run must be 8 bytes");

enum : uint8_t { CUR_ALL_VALID = 1, CUR_ALL_NULL = 2, CUR_DICT = 4 };

struct ColCursor {
  uint64_t plain_off;  // PLAIN page: arena offset of the tile's first value
  uint32_t vidx0;      // chunk-level value index of the tile's first non-null value
  uint32_t vrun_lo;    // first value run intersecting the tile
  uint32_t drun_lo;    // first definition-level run intersecting the tile
  uint16_t vrun_n, drun_n;
  uint32_t nvals;      // non-null values in the tile
  uint8_t flags;
  uint8_t width;       // bit width of the dictionary indices
  uint16_t pad;
};
static_assert(sizeof(ColCursor) == 32, "ColCursor must be 32 bytes");

struct TileDesc {
  uint32_t row0;     // chunk-level row of the tile's first row
  uint32_t nrows;
  uint32_t rg;       // row-group slot
  uint32_t cursor0;  // index of the tile's first cursor (one per physical column)
};

struct ChunkInfo {
  uint64_t base_off;   // arena offset of the chunk's first byte (bit-packed run offsets are relative to it)
  uint64_t dict_off;   // numeric dictionary: arena offset of its PLAIN values
  uint32_t dict_n;
  uint32_t lut_cls;    // offset into the class-table pool (string filter columns)
  uint32_t lut_gcode;  // offset into the group-code pool (key columns)
  uint32_t phys_type;  // parquet physical type
  uint64_t seq_base;   // global sequence number of the row group's first row (fixed-order sums)
  uint32_t defbm_word0; // first 32-bit word of the chunk's expanded definition bitmap (bit r = row r is non-null)
  uint32_t pad;        // 48 bytes
};

// One column chunk whose definition levels the device expands into a flat bitmap before the scan (def_expand_kernel):
// its definition-level runs in the run pool and where its bitmap starts.  One CTA expands LK_DEF_BLOCK_RUNS consecutive
// runs of one chunk; `cum` = number of CTAs of all earlier entries (a CTA finds its chunk by a binary search over it).
constexpr uint32_t LK_DEF_BLOCK_RUNS = 1024;  // measured on B200 (C2): 512 -> 168 us, 1024 -> 145 us, 2048 -> 205 us
struct DefChunk {
  uint64_t base_off;  // arena offset of the chunk's first byte
  uint32_t run_lo, run_n;
  uint32_t word0;     // first word of the chunk's bitmap
  uint32_t num_rows;
  uint32_t cum;
  uint32_t pad;
};
static_assert(sizeof(DefChunk) == 32, "DefChunk must be 32 bytes");
static_assert(sizeof(ChunkInfo) == 48, "ChunkInfo must be 48 bytes");

struct FilterCol {
  uint8_t pcol;
  uint8_t numeric;   // 0: class = lut_cls[code]; 1: class = bitmask of comparison leaves
  uint8_t n_leaves;  // numeric only
  uint8_t null_cls;  // class of SQL NULL
  uint32_t stride;   // multiplier of this column's class in the pass-table index
  uint8_t ops[LK_MAX_NUMLEAF];  // 0 gt, 1 ge, 2 lt, 3 le
  double consts[LK_MAX_NUMLEAF];
};

struct KeyCol {
  uint64_t stride;     // multiplier of this column's group code in the group id
  uint32_t null_code;  // group code of SQL NULL (= dictionary size)
  uint8_t pcol;
};

struct AggSlot {
  uint8_t op;    // AggOp
  uint8_t pcol;  // value column
};

enum : uint32_t { ST_HASH_FULL = 1, ST_BAD_CODE = 2, ST_XCHG_TIMEOUT = 8, ST_XCHG_PHASE = 16 };

// ---- sharded evaluation: the exchange of the record path (one process / GPU per rank, all on one NVLink / NVSwitch node) ----
// Every rank owns two receive pools (even / odd epochs) in one cudaMalloc'ed block that its peers map (CUDA IPC).  A pool is
// divided into one REGION per source rank.  During the scan a survivor record goes straight into the sender's region of the
// pool of the rank that owns its cell (stores over NVLink, fire and forget); the slot comes from a LOCAL counter per
// destination -- nothing on the data path waits for a peer.  At the end the sender publishes how many records it wrote.
// (A first version handed out 256-record chunks of a shared pool with one remote atomic per chunk: at 4 G records/s per
// destination a chunk lasts 60 ns and every sender sat in the remote atomic's round trip -- 121 ms per step at N = 2.)
// See lk_scan.cuh (append) and lk_engine.cu (publish / wait).
constexpr int LK_MAX_RANKS = 16;
struct CommCtrl {  // control words at the start of a rank's exchange block; peers write into it
  uint32_t flag[LK_MAX_RANKS];      // epoch whose records source s has delivered completely
  uint32_t count[2][LK_MAX_RANKS];  // records source s wrote into its region of pool 0 / 1
  uint32_t status[LK_MAX_RANKS];    // scan status flags of source s in that epoch
  uint32_t phase_min[LK_MAX_RANKS], phase_max[LK_MAX_RANKS];  // its timestamp phase range (metrics)
  uint32_t prefix[LK_MAX_RANKS + 1];  // local: records of sources 0..s-1 (written by the wait kernel for the finalize passes)
};
struct XchgParams {
  uint32_t world, rank;
  uint32_t region_cap;                     // records per (source, pool) region
  unsigned long long* keys[LK_MAX_RANKS];  // this rank's region in this epoch's pool of rank d
  unsigned long long* vals[LK_MAX_RANKS];  // ... and its accumulator rows (n_aggs words each)
  CommCtrl* ctrl[LK_MAX_RANKS];            // control words of rank d
  uint32_t* count;                         // local, per destination: records written so far
};

struct ScanParams {
  const uint8_t* arena;
  const TileDesc* tiles;
  const ColCursor* cursors;
  const Run* runs;
  const ChunkInfo* chunks;
  const uint32_t* defbm;      // expanded definition bitmaps of the nullable chunks (ChunkInfo::defbm_word0)
  const uint8_t* lut_cls;
  const uint32_t* lut_gcode;
  const uint32_t* pass_bits;  // bit i set <=> class combination i satisfies the WHERE clause
  uint32_t ntiles, npcols;
  int ts_pcol;
  int n_filter, n_keys, n_aggs;
  FilterCol filter[LK_MAX_FILTER];
  KeyCol keys[LK_MAX_KEYS];
  AggSlot aggs[LK_MAX_AGGS];
  uint32_t def_mask;       // physical columns that can contain NULLs somewhere (need a def bitmap)
  int64_t ts_lo, ts_hi;    // [startTs, endTs)
  int64_t base, step;      // bucket = (ts - base) / step
  uint32_t nbuckets;
  uint64_t n_groups;
  int notnull_pcol;        // chart-field filter `field$type IS NOT NULL` (BaseExpr.scala:407-426), or -1
  int is_metrics;          // metrics: GROUP BY raw timestamp => (ts - base) % step must be one constant
  int path;                // 0 dense, 1 hash, 2 records (sort-based aggregation)
  int warp_agg;            // pre-reduce equal cells inside a warp before the global atomics
  int fits32;              // (endTs - base) and step are below 2^32: 32-bit bucket arithmetic
  int stop_after;          // profiling aid (LK_SCAN_STOP_AFTER=1..5, builds with -DLK_SCAN_PROFILING only): leave each tile after prologue / phase A / phase B; 0 = full
  // dense path: cell = bucket * n_groups + group
  unsigned long long* rowcnt;
  unsigned long long* acc[LK_MAX_AGGS];
  // hash path: open addressing, entries of h_stride bytes {u64 key; u64 acc[n_aggs]}
  uint8_t* h_entries;
  uint32_t h_stride;
  uint64_t h_mask;
  uint32_t* h_occ;      // slots claimed by this query
  uint32_t* h_bkt;      // and the time bucket of each of them (the ORDER BY pass sorts on it without re-reading the table)
  uint32_t h_occ_cap;
  // fixed-order summation (exact_sums): survivors are emitted as records instead of being aggregated
  int emit_records;
  uint32_t rec_cap;
  unsigned long long* rec_cell;
  unsigned long long* rec_seq;  // global row sequence number: segment order, then row order
  unsigned long long* rec_val[LK_MAX_AGGS];
  // record path (path 2): rec_cell[i] = ((bucket << rec_gid_bits | group) << rec_idx_bits) | i and rec_vals[i * n_aggs + a],
  // appended by the scan; finalize partitions the keys into (bucket, hash slice) bins and folds each bin in shared memory
  unsigned long long* rec_vals;
  uint32_t rec_idx_bits;
  uint32_t rec_gid_bits;
  XchgParams x;         // path 3 (sharded record path): records are appended to the owner rank's pool instead
  unsigned long long seq_offset;  // exact_sums, sharded: global sequence number of this shard's first row (shards hold contiguous
                                  // blocks of the request's segment order)
  uint32_t* counters;   // [0] status flags, [1] phase min, [2] phase max, [3] #claimed slots, [4] tile ticket, [5] #records
  unsigned long long* survivors;  // [0] rows that passed the WHERE clause
};

// ---- order-preserving map double -> uint64 (DuckDB total order: -inf < ... < -0.0 < +0.0 < ... < +inf < NaN) ----
LK_HD uint64_t lk_f64_key(uint64_t bits) {
  // all NaNs collapse into one canonical key just below the "empty" sentinels
  if ((bits & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) return 0xfffffffffffffffeull;
  return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}
LK_HD uint64_t lk_key_f64(uint64_t key) {
  if (key == 0xfffffffffffffffeull) return 0x7ff8000000000000ull;
  return (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
}
// Accumulator words are designed so that ALL-ZERO is the empty state of every aggregate (tables are cleared with a
// memset): sum -> +0.0, count -> 0, max -> atomicMax over lk_f64_key (every key is > 0), min -> atomicMax over the
// COMPLEMENT of the key (so "no value yet" is 0 again).
LK_HD uint64_t lk_min_encode(uint64_t f64_bits) { return ~lk_f64_key(f64_bits); }
LK_HD uint64_t lk_max_encode(uint64_t f64_bits) { return lk_f64_key(f64_bits); }
LK_HD uint64_t lk_min_decode(uint64_t stored) { return lk_key_f64(~stored); }
LK_HD uint64_t lk_max_decode(uint64_t stored) { return lk_key_f64(stored); }

// ---- loads ----
LK_HD uint64_t lk_ld64(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
  return *reinterpret_cast<const uint64_t*>(p);
#endif
}

// 8 bytes at an arbitrary byte offset (little endian).  The arena is padded so the second word is always readable.
LK_HD uint64_t lk_load_u64(const uint8_t* arena, uint64_t off) {
  uint64_t a = off & ~7ull;
  unsigned sh = (unsigned)(off & 7) * 8;
  uint64_t lo = lk_ld64(arena + a);
  if (sh == 0) return lo;
  uint64_t hi = lk_ld64(arena + a + 8);
  return (lo >> sh) | (hi << (64 - sh));
}

// index of the last run whose start <= idx (runs sorted by start, n >= 1)
LK_HD uint32_t lk_find_run(const Run* runs, uint32_t n, uint32_t idx) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (runs[mid].start <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

// definition bits of rows [row, row + nbits) (nbits <= 32) of a column that has a def-level stream
LK_HD uint32_t lk_def_word(const uint8_t* arena, const Run* runs, const ColCursor& c, const ChunkInfo& ci, uint32_t row,
                           uint32_t nbits) {
  const Run* r = runs + c.drun_lo;
  uint32_t ri = lk_find_run(r, c.drun_n, row);
  uint32_t w = 0, filled = 0;
  while (filled < nbits) {
    Run run = r[ri];
    uint32_t next = (ri + 1 < c.drun_n) ? r[ri + 1].start : 0xffffffffu;
    uint32_t avail = next - (row + filled);
    if (avail > nbits - filled) avail = nbits - filled;
    uint32_t m = avail >= 32 ? 0xffffffffu : ((1u << avail) - 1);
    if (run.kind_value >> 31) {
      if (run.kind_value & 1) w |= m << filled;
    } else {
      uint32_t bit = row + filled - run.start;
      uint64_t x = lk_load_u64(arena, ci.base_off + run.kind_value + (bit >> 3)) >> (bit & 7);
      w |= ((uint32_t)x & m) << filled;
    }
    filled += avail;
    ri++;
  }
  return w;
}

// 32 bits starting at byte address p (arbitrary alignment); the arena is padded so the second word is always readable
LK_HD uint32_t lk_load_u32_unaligned(const uint8_t* p) {
  const uint64_t a = reinterpret_cast<uint64_t>(p);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~3ull);
  const unsigned sh = (unsigned)(a & 3) * 8;
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(__ldg(q), __ldg(q + 1), sh);
#else
  return sh ? (q[0] >> sh) | (q[1] << (32 - sh)) : q[0];
#endif
}

// Expansion of ONE definition-level run (rows [r.start, next) of its chunk) into the chunk's flat bitmap (bit r = row r
// is non-null).  Bit-packed levels of width 1 already are the bitmap bits; an RLE run of 1 is a range of ones.  Words
// shared with neighbouring runs go through or_word(word, mask) (atomicOr on the device); the whole words inside a long
// RLE run belong to it alone and go through fill(first, last_exclusive).  Shared by def_expand_kernel (lk_engine.cu) and
// the CPU emulator of the tests.
// `first` = the first 32 payload bits of a bit-packed run when the caller has already loaded them (have_first).
template <class OrWord, class Fill>
LK_HD void lk_def_expand_run(const uint8_t* arena, uint64_t base_off, Run r, uint32_t next, OrWord or_word, Fill fill, bool have_first = false,
                             uint32_t first = 0) {
  if (next <= r.start) return;
  if (r.kind_value >> 31) {
    if (!(r.kind_value & 1)) return;
    const uint32_t b0 = r.start, b1 = next;
    const uint32_t w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
    const uint32_t m0 = 0xffffffffu << (b0 & 31), m1 = 0xffffffffu >> (31 - ((b1 - 1) & 31));
    if (w0 == w1) { or_word(w0, m0 & m1); return; }
    or_word(w0, m0);
    or_word(w1, m1);
    if (w1 - w0 > 1) fill(w0 + 1, w1);
    return;
  }
  const uint8_t* src = arena + base_off + r.kind_value;
  const uint32_t n = next - r.start;
  for (uint32_t o = 0; o < n; o += 32) {
    const uint32_t c = n - o < 32 ? n - o : 32;
    uint32_t x = (o == 0 && have_first) ? first : lk_load_u32_unaligned(src + (o >> 3));
    if (c < 32) x &= (1u << c) - 1;
    const uint32_t pos = r.start + o, sh = pos & 31;
    if (x << sh) or_word(pos >> 5, x << sh);
    if (sh && (x >> (32 - sh))) or_word((pos >> 5) + 1, x >> (32 - sh));
  }
}

// dictionary index of chunk-level value `vidx`
LK_HD uint32_t lk_dict_code(const uint8_t* arena, const Run* runs, const ColCursor& c, const ChunkInfo& ci, uint32_t vidx) {
  const Run* r = runs + c.vrun_lo;
  Run run = r[lk_find_run(r, c.vrun_n, vidx)];
  if (run.kind_value >> 31) return run.kind_value & 0x7fffffffu;
  uint64_t bitpos = (uint64_t)(vidx - run.start) * c.width;
  uint64_t x = lk_load_u64(arena, ci.base_off + run.kind_value + (bitpos >> 3)) >> (bitpos & 7);
  return (uint32_t)x & (c.width >= 32 ? 0xffffffffu : ((1u << c.width) - 1));
}

// raw little-endian bits of value `vidx` (PLAIN page or numeric dictionary); 4-byte types in the low half
LK_HD uint64_t lk_value_bits(const uint8_t* arena, const Run* runs, const ColCursor& c, const ChunkInfo& ci, uint32_t vidx,
                             uint32_t* bad) {
  unsigned esz = (ci.phys_type == 1 || ci.phys_type == 4) ? 4 : 8;
  uint64_t off;
  if (c.flags & CUR_DICT) {
    uint32_t code = lk_dict_code(arena, runs, c, ci, vidx);
    if (code >= ci.dict_n) { *bad = 1; code = 0; }
    off = ci.dict_off + (uint64_t)code * esz;
  } else {
    off = c.plain_off + (uint64_t)(vidx - c.vidx0) * esz;
  }
  uint64_t x = lk_load_u64(arena, off);
  return esz == 4 ? (x & 0xffffffffull) : x;
}

LK_HD double lk_bits_to_f64(uint64_t bits, uint32_t phys_type) {
  union { uint64_t u; double d; } a;
  union { uint32_t u; float f; } b;
  switch (phys_type) {
    case 5: a.u = bits; return a.d;
    case 4: b.u = (uint32_t)bits; return (double)b.f;
    case 2: return (double)(int64_t)bits;
    default: return (double)(int32_t)(uint32_t)bits;
  }
}

LK_HD int64_t lk_bits_to_i64(uint64_t bits, uint32_t phys_type) {
  return phys_type == 2 ? (int64_t)bits : (int64_t)(int32_t)(uint32_t)bits;
}

// class of a numeric filter column: bit l set <=> comparison leaf l is TRUE (DuckDB order: NaN greatest, NaN == NaN)
LK_HD uint32_t lk_numeric_class(const FilterCol& f, double x) {
  uint32_t cls = 0;
  bool xn = x != x;
  for (int l = 0; l < f.n_leaves; l++) {
    double c = f.consts[l];
    bool cn = c != c;
    int r = (xn || cn) ? ((int)xn - (int)cn) : ((x > c) - (x < c));
    bool t = f.ops[l] == 0 ? r > 0 : f.ops[l] == 1 ? r >= 0 : f.ops[l] == 2 ? r < 0 : r <= 0;
    cls |= (uint32_t)t << l;
  }
  return cls;
}

// ---- seek index built on the device (lk_engine.cu: idx_* kernels) ----
// The host walks footers and PAGE headers only; the run headers of every hybrid stream (definition levels, dictionary
// indices) are walked by the device once the column chunks are in HBM: one thread per page, a count pass and a fill pass,
// then one thread per (tile, column) computes the cursor.  Replaces ~0.6 host core-seconds per 105 M rows.
struct IdxPage {
  uint64_t def_off, def_end;  // arena byte range of the definition-level stream (def_end == def_off: REQUIRED column, all rows valid)
  uint64_t val_off, val_end;  // arena byte range of the values: PLAIN values, or hybrid dictionary indices (after the width byte)
  uint32_t first_row, num_rows;
  uint32_t chunk;             // index into IdxChunk[]
  uint8_t dict_coded, bit_width, pad[2];
  // filled by the device
  uint32_t nn;                // non-null values of the page
  uint32_t first_vidx;        // non-null values of the chunk before the page
  uint32_t def_runs, val_runs;
  uint32_t def_run0, val_run0;  // first run of the page in the run pool
};
struct IdxChunk {
  uint64_t base_off;          // arena offset of the chunk's first byte
  uint32_t page0, npages;
  uint32_t present, max_def, esz, num_rows, dict_n, string_typed;
  // filled by the device
  uint32_t def_run0, def_runs, val_run0, val_runs;
  uint32_t nn, mixed;         // non-null values; some tile of the chunk mixes NULLs and values (needs a definition bitmap)
};
enum : uint32_t { IDX_ST_TRUNCATED = 1, IDX_ST_BAD_RUN = 2, IDX_ST_BAD_CODE = 4, IDX_ST_PLAIN_STRING = 8, IDX_ST_BAD_SNAPPY = 16 };

// One compressed page (SNAPPY) of a touched column chunk.  The device inflates [src_off, src_off + src_len) into
// [dst_off, dst_off + dst_len) -- room reserved behind the chunk in the same device block -- and then reads from the inflated
// bytes what the host reads from an uncompressed page's first bytes: the length prefix of a V1 page's definition levels and
// the bit width of a dictionary-coded page, completing the page's IdxPage.
enum : uint32_t { ZP_DECODE = 1 /* not inflated yet (a cached block already is) */, ZP_V1_DEF = 2 /* V1 page of an OPTIONAL column: 4-byte length + definition levels first */,
                  ZP_DICT_CODED = 4 /* values start with the bit-width byte */, ZP_VALUES_ONLY = 8 /* V2: the inflated bytes are the values section only */ };
struct ZPage {
  uint64_t src_off, dst_off;  // arena offsets
  uint32_t src_len, dst_len;
  uint32_t page;              // index into IdxPage[] (0xffffffff: a numeric dictionary page, nothing to complete)
  uint32_t flags;
};

struct HybridRun {
  uint32_t n;        // elements the run contributes (clipped to what the page still needs)
  uint32_t is_rle, value;
  uint64_t payload;  // bit-packed: offset of the packed bytes
  uint64_t next;     // offset of the next run header
};
// One run header of an RLE / bit-packed hybrid stream at `off` (< end); `remaining` elements are still expected.
// Returns 0 or an IDX_ST_* flag.  The host's page walker and the device's index kernels share it.
LK_HD uint32_t lk_hybrid_next(const uint8_t* base, uint64_t off, uint64_t end, uint32_t bit_width, uint32_t remaining, HybridRun& r) {
  uint64_t h = 0;
  uint32_t shift = 0;
  for (;;) {
    if (off >= end) return IDX_ST_TRUNCATED;
    const uint32_t b = base[off++];
    h |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
    if (shift > 63) return IDX_ST_BAD_RUN;
  }
  if (h & 1) {
    const uint64_t groups = h >> 1, n = groups * 8, bytes = groups * (uint64_t)bit_width;
    if (groups == 0) return IDX_ST_BAD_RUN;
    const uint32_t take = n < (uint64_t)remaining ? (uint32_t)n : remaining;
    // a writer may truncate the padding of the final group; require the bytes that carry real values
    if (off + ((uint64_t)take * bit_width + 7) / 8 > end) return IDX_ST_TRUNCATED;
    r.n = take;
    r.is_rle = 0;
    r.value = 0;
    r.payload = off;
    r.next = off + (bytes < end - off ? bytes : end - off);
  } else {
    const uint64_t n = h >> 1;
    const uint32_t vbytes = (bit_width + 7) / 8;
    if (n == 0) return IDX_ST_BAD_RUN;
    if (off + vbytes > end) return IDX_ST_TRUNCATED;
    uint32_t v = 0;
    for (uint32_t i = 0; i < vbytes; i++) v |= (uint32_t)base[off + i] << (8 * i);
    r.n = n < (uint64_t)remaining ? (uint32_t)n : remaining;
    r.is_rle = 1;
    r.value = v;
    r.payload = 0;
    r.next = off + vbytes;
  }
  return 0;
}
// set bits among the first nbits bits at byte address p (arbitrary alignment)
LK_HD uint32_t lk_popcount_bits(const uint8_t* p, uint32_t nbits) {
  uint32_t c = 0, i = 0;
  const uint32_t nbytes = nbits >> 3;
  for (; i < nbytes; i++) {
#if defined(__CUDA_ARCH__)
    c += __popc((uint32_t)p[i]);
#else
    c += (uint32_t)__builtin_popcount(p[i]);
#endif
  }
  if (nbits & 7) {
    const uint32_t x = p[nbytes] & ((1u << (nbits & 7)) - 1);
#if defined(__CUDA_ARCH__)
    c += __popc(x);
#else
    c += (uint32_t)__builtin_popcount(x);
#endif
  }
  return c;
}

// ---- record path finalize ----
LK_HD uint32_t lk_rf_mix(uint64_t gid) {  // 32 well-mixed bits of a group id: the slot inside its bucket's region of the key table
  uint32_t h = (uint32_t)gid ^ (uint32_t)(gid >> 32) * 0x9E3779B1u;
  h *= 0x9E3779B1u; h ^= h >> 16;
  h *= 0x85EBCA6Bu; h ^= h >> 13;
  h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

LK_HD uint64_t lk_hash64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

}  // namespace lk

// Device layer of liblakeside_b200: HBM residency of a prepared query, kernel launches, result compaction.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>

#include "lk_cache.h"
#include "lk_engine.h"
#include "lk_scan.cuh"

#include <cub/device/device_radix_sort.cuh>  // exact_sums (the optional fixed-order pass) only; nothing on the default path

namespace lk {

static void device_exact_sums(Query& q, const ScanParams& base);
static void rec_exact_finalize(Query& q);
static void rec_clear_scratch(Query& q, cudaStream_t st);
void device_resolve(Query& q);

#define CUDA_CHECK(x)                                                                                        \
  do {                                                                                                       \
    cudaError_t err__ = (x);                                                                                 \
    if (err__ != cudaSuccess) ::lk::fail(LK_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------------------
// device / pools
// ------------------------------------------------------------------------------------------------------------
static std::mutex g_mu;
static bool g_inited = false;
static int g_num_sms = 148;
static int64_t g_default_cache_bytes = 0;
static cudaStream_t g_cache_stream = nullptr;

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void device_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_inited) { CUDA_CHECK(cudaSetDevice(global_options().device)); return; }
  LK_CHECK(device_count() > 0, LK_ERR_CUDA, "no CUDA device visible: liblakeside_b200 has no CPU fallback");
  CUDA_CHECK(cudaSetDevice(global_options().device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, global_options().device));
  g_num_sms = prop.multiProcessorCount;
  if (const char* g = getenv("LK_L2_FETCH")) {  // tuning aid: L2 fetch granularity hint (32 / 64 / 128 bytes)
    size_t got = 0;
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "[lk] cudaLimitMaxL2FetchGranularity = %zu\n", got);
    cudaGetLastError();
  }
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, global_options().device));
  uint64_t thresh = ~0ull;  // keep freed blocks cached: repeated queries re-use them without cudaMalloc
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  PoolAlloc::alloc = pinned_alloc;   // index pools of later queries live in pinned memory
  PoolAlloc::release = pinned_free;
  // HBM-resident segment cache (lk_cache.h): a third of the device's memory unless lk_init said otherwise
  int64_t cache_bytes = global_options().segment_cache_bytes;
  g_default_cache_bytes = (int64_t)(prop.totalGlobalMem / 3);
  if (cache_bytes < 0) cache_bytes = g_default_cache_bytes;
  segment_cache().set_capacity((size_t)cache_bytes);
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_cache_stream, cudaStreamNonBlocking));
  g_inited = true;
}

// device blocks of cached columns are released here: nothing uses a block any more when its last reference goes away
// (queries drop theirs after synchronising their stream)
void cache_free_device(void* p) {
  if (!p) return;
  cudaSetDevice(global_options().device);  // (the last reference may go away on any thread)
  if (g_cache_stream) cudaFreeAsync(p, g_cache_stream);
  else cudaFree(p);
  cudaGetLastError();
}

int num_sms() { return g_num_sms; }
int64_t device_default_cache_bytes() { return g_default_cache_bytes; }

// pinned host memory, cached by power-of-two size class
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<size_t, void*>> free_list;
  std::vector<std::pair<void*, size_t>> live;
} g_pinned;

void* pinned_alloc(size_t bytes) {
  size_t cap = 4096;
  while (cap < bytes) cap <<= 1;
  {
    std::lock_guard<std::mutex> lk(g_pinned.mu);
    for (size_t i = 0; i < g_pinned.free_list.size(); i++)
      if (g_pinned.free_list[i].first == cap) {
        void* p = g_pinned.free_list[i].second;
        g_pinned.free_list.erase(g_pinned.free_list.begin() + i);
        g_pinned.live.emplace_back(p, cap);
        return p;
      }
  }
  void* p = nullptr;
  // (worker threads of the planner and of the sparse segment reads come here too: a fresh thread's current device is 0, and an
  // allocation there would create a context on GPU 0 from every rank of a multi-GPU box)
  cudaSetDevice(global_options().device);
  cudaError_t e = cudaHostAlloc(&p, cap, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); fail(e == cudaErrorMemoryAllocation ? LK_ERR_NOMEM : LK_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
  std::lock_guard<std::mutex> lk(g_pinned.mu);
  g_pinned.live.emplace_back(p, cap);
  return p;
}

void pinned_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pinned.mu);
  for (size_t i = 0; i < g_pinned.live.size(); i++)
    if (g_pinned.live[i].first == p) {
      g_pinned.free_list.emplace_back(g_pinned.live[i].second, p);
      g_pinned.live.erase(g_pinned.live.begin() + i);
      return;
    }
}

void pinned_release_all() {
  std::lock_guard<std::mutex> lk(g_pinned.mu);
  for (auto& f : g_pinned.free_list) cudaFreeHost(f.second);
  g_pinned.free_list.clear();
}

// hash arenas stay allocated and CLEAN (all zero) between queries: a query only ever touches the slots it claims and
// the emit kernel zeroes them again, so no O(capacity) clear sits on the query path.
struct HashArena {
  uint8_t* entries = nullptr;
  uint32_t* occ = nullptr;
  uint32_t* occ_bkt = nullptr;  // time bucket of every claimed slot (written at claim time: the emit sort never re-reads the entries)
  uint64_t slots = 0;
  uint32_t stride = 0;
  bool busy = false;
};
static std::vector<HashArena> g_arenas;

static HashArena* arena_acquire(uint64_t slots, uint32_t stride, cudaStream_t st) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& a : g_arenas)
      if (!a.busy && a.slots == slots && a.stride == stride) { a.busy = true; return &a; }
  }
  HashArena a;
  a.slots = slots;
  a.stride = stride;
  cudaError_t e = cudaMalloc(&a.entries, slots * stride);
  if (e == cudaSuccess) e = cudaMalloc(&a.occ, slots * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&a.occ_bkt, slots * sizeof(uint32_t));
  if (e != cudaSuccess) { cudaGetLastError(); if (a.entries) cudaFree(a.entries); if (a.occ) cudaFree(a.occ); fail(LK_ERR_NOMEM, strf("hash arena of %llu slots: %s", (unsigned long long)slots, cudaGetErrorString(e))); }
  CUDA_CHECK(cudaMemsetAsync(a.entries, 0, slots * stride, st));
  a.busy = true;
  std::lock_guard<std::mutex> lk(g_mu);
  g_arenas.reserve(64);
  LK_CHECK(g_arenas.size() < 64, LK_ERR_NOMEM, "too many live hash arenas");
  g_arenas.push_back(a);
  return &g_arenas.back();
}

static void arena_release(HashArena* a) {
  std::lock_guard<std::mutex> lk(g_mu);
  a->busy = false;
}

void device_shutdown() {
  segment_cache().clear();
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& a : g_arenas) { cudaFree(a.entries); cudaFree(a.occ); cudaFree(a.occ_bkt); }
  g_arenas.clear();
  pinned_release_all();
}

// ------------------------------------------------------------------------------------------------------------
// per-query device state
// ------------------------------------------------------------------------------------------------------------
struct Query::Device {
  cudaStream_t st = nullptr;
  cudaEvent_t ev[12] = {};
  uint8_t* arena = nullptr;
  TileDesc* tiles = nullptr;
  ColCursor* cursors = nullptr;
  Run* runs = nullptr;
  ChunkInfo* chunks = nullptr;
  DefChunk* def_chunks = nullptr;
  uint32_t* defbm = nullptr;  // expanded definition bitmaps (rewritten by every execute)
  uint8_t* lut_cls = nullptr;
  uint32_t* lut_gcode = nullptr;
  uint32_t* pass_bits = nullptr;
  uint32_t* counters = nullptr;           // 8 x u32
  unsigned long long* survivors = nullptr;
  unsigned long long* planes = nullptr;   // dense: (1 + n_aggs) planes of n_cells words
  HashArena* harena = nullptr;
  bool resident = false, executed = false, group_tables_stale = false;
  size_t uploads_done = 0;  // entries of Query::uploads already on their way
  // compaction scratch + device result
  uint32_t* block_counts = nullptr;
  size_t block_counts_cap = 0;
  uint8_t* dres = nullptr;
  size_t dres_cap = 0;
  // record path: appended (cell, accumulator words) records + sort scratch
  unsigned long long* rec_cell = nullptr;
  unsigned long long* rec_vals = nullptr;
  size_t rec_cap = 0;
  bool rec_borrowed = false;  // rec_cell / rec_vals point into the communicator's receive pool
  // record path finalize: key table (two slots per record) + per-bucket counters (capacity learned from the first finalize,
  // checked on the device)
  unsigned long long* rf_sorted = nullptr;
  uint32_t* rf_tables = nullptr;
  // records partitioned by bucket before grouping (rec_scatter_kernel): second copy of the record arrays, fin_cap records
  unsigned long long* rec2_cell = nullptr;
  unsigned long long* rec2_vals = nullptr;
  bool rec_scatter = false;   // decided when the scratch is sized (rec_finalize_size)
  uint32_t rec_slices = 1;    // partitions per time bucket on the partitioned path
  uint32_t rec_smem_slots = 0;  // shared-memory key table of rec_group_bucket_kernel: slots per CTA
  struct RecFin* fin = nullptr;
  uint32_t* fin_host = nullptr;  // pinned: [0..7] RecFin, [8..15] the scan's counters, copied back at the end of finalize
  // the finalize scratch is cleared on a side stream WHILE the scan runs (execute forks, finalize joins)
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_clear = nullptr;
  bool pre_cleared = false;
  size_t fin_cap = 0;            // records the finalize scratch and the result buffer hold
  bool fin_pending = false;      // finalize kernels enqueued, row count / status not yet read back
  size_t dres_stride = 0;        // rows per result column in dres
  uint8_t* sparse_out = nullptr;  // partitioned copy of the claimed entries (sparse exchange)
  size_t sparse_cap = 0;
  int64_t n_rows = 0;
  uint32_t phase = 0;
  // sharded dense / hash paths: the phase range over ALL ranks, handed in by the host (lk_query_set_phase) between the
  // exchange and the finalize; valid for the current execute only
  bool phase_global = false;
  uint32_t phase_gmin = 0xffffffffu, phase_gmax = 0;
  uint32_t h_counters[8] = {};
  bool finalized_device = false;
};

Query::Query() = default;
Query::~Query() {
  const auto t_destroy0 = std::chrono::steady_clock::now();
  if (dev) {
    Device& d = *dev;
    cudaSetDevice(global_options().device);  // (a query may be destroyed on another thread than the one that ran it)
    if (d.st2) cudaStreamSynchronize(d.st2);
    if (d.st) cudaStreamSynchronize(d.st);
    auto fr = [&](void* p) { if (p) cudaFreeAsync(p, d.st); };
    fr(d.arena); fr(d.tiles); fr(d.cursors); fr(d.runs); fr(d.chunks); fr(d.def_chunks); fr(d.defbm); fr(d.lut_cls); fr(d.lut_gcode); fr(d.pass_bits);
    fr(d.counters); fr(d.survivors); fr(d.planes); fr(d.block_counts); fr(d.dres); fr(d.sparse_out);
    if (!d.rec_borrowed) { fr(d.rec_cell); fr(d.rec_vals); }
    fr(d.rf_sorted); fr(d.rf_tables); fr(d.fin); fr(d.rec2_cell); fr(d.rec2_vals);
    if (d.fin_host) pinned_free(d.fin_host);
    if (d.harena) {
      // an arena that was written but never emitted is dirty: clear it before handing it back
      if (d.executed && !d.finalized_device) cudaMemsetAsync(d.harena->entries, 0, d.harena->slots * d.harena->stride, d.st);
      if (d.st) cudaStreamSynchronize(d.st);
      arena_release(d.harena);
    }
    for (auto& e : d.ev) if (e) cudaEventDestroy(e);
    if (d.st2) { cudaStreamSynchronize(d.st2); cudaStreamDestroy(d.st2); }
    if (d.ev_fork) cudaEventDestroy(d.ev_fork);
    if (d.ev_clear) cudaEventDestroy(d.ev_clear);
    if (d.st) { cudaStreamSynchronize(d.st); cudaStreamDestroy(d.st); }
  }
  cache_fresh.clear();  // the stream is idle: cached columns this query pinned may go (freed when no one else holds them)
  cache_refs.clear();
  for (auto& s : segs) if (s.owned_pinned) pinned_free(s.owned_pinned);
  if (getenv("LK_PLAN_TRACE"))
    fprintf(stderr, "[lk destroy] device part %23.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_destroy0).count());
}

template <class T>
static void upload_vec(T*& dptr, const std::vector<T>& v, cudaStream_t st) {
  if (dptr) { CUDA_CHECK(cudaFreeAsync(dptr, st)); dptr = nullptr; }
  size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  CUDA_CHECK(cudaMallocAsync(&dptr, bytes, st));
  if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

template <class T>
static void upload_vec(T*& dptr, const RawVec<T>& v, cudaStream_t st) {
  if (dptr) { CUDA_CHECK(cudaFreeAsync(dptr, st)); dptr = nullptr; }
  size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  CUDA_CHECK(cudaMallocAsync(&dptr, bytes, st));
  if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

static void upload_group_tables(Query& q) {
  Query::Device& d = *q.dev;
  upload_vec(d.lut_gcode, q.lut_gcode, d.st);
  upload_vec(d.chunks, q.chunk_infos, d.st);
  // aggregate table for the (possibly new) group space
  if (d.planes) { CUDA_CHECK(cudaFreeAsync(d.planes, d.st)); d.planes = nullptr; }
  if (d.harena) { arena_release(d.harena); d.harena = nullptr; }
  if (q.n_cells > 0) {
    if (q.path == 0) CUDA_CHECK(cudaMallocAsync(&d.planes, (1 + q.aggs.size()) * q.n_cells * sizeof(unsigned long long), d.st));
    else if (q.path == 1) d.harena = arena_acquire(q.hash_slots, q.hash_stride, d.st);
    // path 2 (records) has no table: its record arrays are sized by the row count at the first execute
  }
  d.group_tables_stale = false;
}

// Reads a segment file into pinned memory owned by the query (the whole file: a cache miss on one column usually comes with
// misses on the others, and H2D copies want pinned source bytes).
void segment_load(SegmentInput& s) {
  if (s.data) return;
  int fd = open(s.name.c_str(), O_RDONLY);
  LK_CHECK(fd >= 0, LK_ERR_IO, "IO Error: cannot open " + s.name);
  void* buf = nullptr;
  try {
    buf = pinned_alloc(s.len + 16);
  } catch (...) { close(fd); throw; }
  size_t got = 0;
  while (got < s.len) {
    ssize_t r = read(fd, (char*)buf + got, s.len - got);
    if (r <= 0) break;
    got += (size_t)r;
  }
  close(fd);
  if (got != s.len) { pinned_free(buf); fail(LK_ERR_IO, "IO Error: short read on " + s.name); }
  s.data = (const uint8_t*)buf;
  s.owned_pinned = buf;
}

// A segment file that the cache does not know: only its footer is read now (from the file's tail); the column chunks the query
// touches follow in device_layout, once the plan knows which they are -- a wide segment (hundreds of tag columns) is never read
// whole for the ten columns a chart needs.
void segment_open(SegmentInput& s) {
  int fd = open(s.name.c_str(), O_RDONLY);
  LK_CHECK(fd >= 0, LK_ERR_IO, "IO Error: cannot open " + s.name);
  std::vector<uint8_t> tail;
  size_t want = std::min<size_t>(s.len, 64 << 10);
  try {
    for (int attempt = 0;; attempt++) {
      tail.resize(want);
      size_t got = 0;
      while (got < want) {
        ssize_t r = pread(fd, tail.data() + got, want - got, (off_t)(s.len - want + got));
        if (r <= 0) break;
        got += (size_t)r;
      }
      LK_CHECK(got == want, LK_ERR_IO, "IO Error: short read on " + s.name);
      size_t need = 0;
      if (parse_footer_tail(tail.data(), tail.size(), s.len, &s.meta, &need)) break;
      LK_CHECK(attempt == 0 && need <= s.len, LK_ERR_IO, "parquet: bad footer length");
      want = need;
    }
    char magic[4] = {0, 0, 0, 0};
    LK_CHECK(pread(fd, magic, 4, 0) == 4 && memcmp(magic, "PAR1", 4) == 0, LK_ERR_IO, "not a Parquet file (bad magic)");
  } catch (...) { close(fd); throw; }
  close(fd);
  s.sparse = true;
}

// reads the byte ranges of `slots` (all of one sparse segment) into one pinned block, back to back, and points the slots at them
static void segment_read_chunks(SegmentInput& s, std::vector<ChunkSlot*>& slots) {
  if (slots.empty()) return;
  size_t total = 0;
  for (auto* sl : slots) total += ((size_t)sl->len + 15) & ~(size_t)15;
  int fd = open(s.name.c_str(), O_RDONLY);
  LK_CHECK(fd >= 0, LK_ERR_IO, "IO Error: cannot open " + s.name);
  uint8_t* buf = nullptr;
  try {
    buf = static_cast<uint8_t*>(pinned_alloc(total + 16));
    size_t at = 0;
    for (auto* sl : slots) {
      size_t got = 0;
      while (got < sl->len) {
        ssize_t r = pread(fd, buf + at + got, sl->len - got, (off_t)(sl->file_off + got));
        if (r <= 0) break;
        got += (size_t)r;
      }
      LK_CHECK(got == sl->len, LK_ERR_IO, "IO Error: short read on " + s.name);
      sl->host = buf + at;
      at += ((size_t)sl->len + 15) & ~(size_t)15;
    }
  } catch (...) { close(fd); if (buf) pinned_free(buf); throw; }
  close(fd);
  s.owned_pinned = buf;  // (a sparse segment owns nothing else)
  s.sparse_bytes = total;
}

static SegmentIdentity segment_identity(const SegmentInput& s) {
  SegmentIdentity id;
  id.path = s.name;
  id.size = s.len;
  id.mtime_ns = s.id_mtime_ns;
  id.ino = s.id_ino;
  return id;
}

static void ensure_stream(Query& q) {
  device_init();
  if (!q.dev) q.dev = std::make_unique<Query::Device>();
  Query::Device& d = *q.dev;
  if (!d.st) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    for (auto& e : d.ev) CUDA_CHECK(cudaEventCreate(&e));
  }
}

// plan_query's on_layout hook: decides where every touched column chunk lives in device memory and starts the copies.
// Chunks of identified segment files go through the segment cache -- one device block per (file, column), found there
// (nothing to read, parse or copy) or allocated now and published once the query is resident; everything else (caller
// buffers) is laid out in the query's private arena.  Kernels address all of it as `arena + offset`: a cached chunk's
// offset is its address minus the private arena's, in 64-bit wrap-around arithmetic, so nothing downstream knows the
// difference.
void device_layout(Query& q) {
  ensure_stream(q);
  Query::Device& d = *q.dev;
  SegmentCache& cache = segment_cache();
  const int np = (int)q.pcols.size();
  std::vector<uint8_t> placed(q.slots.size(), 0);
  std::vector<uint64_t> abs_addr(q.slots.size(), 0);
  std::vector<uint8_t> fresh(q.slots.size(), 0);
  q.cache_fresh.clear();
  bool any_file = false;
  for (auto& sg : q.segs) any_file = any_file || sg.has_identity;
  const bool use_cache = any_file && q.device_index && cache.capacity() > 0;
  if (use_cache) {
    std::map<std::pair<int, int>, std::vector<size_t>> groups;  // (segment, touched column) -> its chunk slots, one per row group
    for (size_t k = 0; k < q.slots.size(); k++)
      if (q.segs[q.slots[k].seg].has_identity) groups[{q.slots[k].seg, q.slots[k].pcol}].push_back(k);
    for (auto& g : groups) {
      SegmentInput& seg = q.segs[g.first.first];
      const int leaf = q.slots[g.second[0]].leaf;
      std::shared_ptr<CachedColumn> col = seg.cached ? cache.column(seg.cached, leaf) : nullptr;
      if (col)
        for (size_t k : g.second)
          if (col->chunk_off[q.rgs[q.slots[k].rgi].rg] == ~0ull) { col = nullptr; break; }
      const bool hit = col != nullptr;
      if (!hit) {
        col = std::make_shared<CachedColumn>();
        col->chunk_off.assign(seg.meta.row_groups.size(), ~0ull);
        col->index.resize(seg.meta.row_groups.size());
        uint64_t off = 0;
        for (size_t k : g.second) {
          const ChunkSlot& sl = q.slots[k];
          off = (off + 255) & ~255ull;
          col->chunk_off[q.rgs[sl.rgi].rg] = off;
          off += sl.len;
          if (sl.reserve) off = ((off + 7) & ~7ull) + sl.reserve;
        }
        col->bytes = ((off + 255) & ~255ull) + 256;  // tail padding, as for the private arena
        cudaError_t e = cudaMallocAsync(&col->dev, col->bytes, d.st);
        if (e != cudaSuccess) { cudaGetLastError(); col->dev = nullptr; fail(LK_ERR_NOMEM, strf("segment cache block of %zu bytes: %s", col->bytes, cudaGetErrorString(e))); }
        q.cache_fresh.push_back({g.first.first, g.first.second, leaf, col});
      }
      for (size_t k : g.second) {
        const ChunkSlot& sl = q.slots[k];
        RowGroupPlan& rp = q.rgs[sl.rgi];
        abs_addr[k] = (uint64_t)(uintptr_t)col->dev + col->chunk_off[rp.rg];
        placed[k] = 1;
        fresh[k] = hit ? 0 : 1;
        if (hit) {
          rp.chunks[sl.pcol] = col->index[rp.rg];
          rp.from_cache[sl.pcol] = 1;
        }
      }
      q.cache_refs.push_back(col);
    }
  }
  // host-built index (LK_HOST_INDEX): the planner reads run payloads anywhere in a chunk through the whole-file pointer
  if (!q.device_index)
    for (auto& sg : q.segs)
      if (!sg.data) { segment_load(sg); sg.sparse = false; }
  for (auto& sl : q.slots)
    if (!sl.host && q.segs[sl.seg].data) sl.host = q.segs[sl.seg].data + sl.file_off;
  // sparse files: the chunks that are not resident are read now, all segments in parallel
  {
    std::vector<std::vector<ChunkSlot*>> todo(q.segs.size());
    for (size_t k = 0; k < q.slots.size(); k++) {
      ChunkSlot& sl = q.slots[k];
      const bool resident = placed[k] && !fresh[k];
      if (!sl.host && !resident) todo[sl.seg].push_back(&sl);
    }
    parallel_for((int)q.segs.size(), global_options().host_threads, [&](int i) { segment_read_chunks(q.segs[i], todo[i]); });
  }
  layout_private_arena(q, placed);
  if (!d.arena) {
    CUDA_CHECK(cudaEventRecord(d.ev[0], d.st));
    CUDA_CHECK(cudaMallocAsync(&d.arena, q.arena_bytes, d.st));
  }
  const uint64_t origin = (uint64_t)(uintptr_t)d.arena;
  for (size_t k = 0; k < q.slots.size(); k++) {
    if (!placed[k]) continue;
    const ChunkSlot& sl = q.slots[k];
    q.rgs[sl.rgi].arena_base[sl.pcol] = abs_addr[k] - origin;
    if (fresh[k]) q.uploads.push_back({sl.seg, sl.file_off, sl.len, abs_addr[k] - origin});
  }
  // the copies read from where the chunks are in host memory (whole file or sparse block)
  {
    std::map<std::pair<int, uint64_t>, const uint8_t*> host_of;
    for (auto& sl : q.slots) if (sl.host) host_of[{sl.seg, sl.file_off}] = sl.host;
    for (auto& u : q.uploads)
      if (!u.src) {
        auto it = host_of.find({u.seg, u.file_off});
        LK_CHECK(it != host_of.end(), LK_ERR_IO, "segment '" + q.segs[u.seg].name + "': column chunk bytes were not read");
        u.src = it->second;
      }
  }
  (void)np;
  device_begin_upload(q);
}

// the query is resident: the columns it brought in become visible to later queries
static void cache_publish(Query& q) {
  if (q.cache_fresh.empty()) return;
  SegmentCache& cache = segment_cache();
  for (auto& f : q.cache_fresh) {
    const SegmentInput& seg = q.segs[f.seg];
    for (auto& rp : q.rgs) {
      if (rp.seg != f.seg) continue;
      ChunkIndex ci = rp.chunks[f.pcol];
      std::vector<uint8_t>().swap(ci.synth);  // already in the block, behind the chunk
      f.col->index[rp.rg] = std::move(ci);
    }
    cache.publish(segment_identity(seg), seg.meta, f.leaf, f.col);
  }
  q.cache_fresh.clear();
}

// Starts the H2D copies of the touched column chunks (asynchronous; called from plan_query's on_layout hook so that
// the copies run while the host still walks page and run headers).
bool device_is_resident(const Query& q) { return q.dev && q.dev->resident; }

void device_begin_upload(Query& q) {
  ensure_stream(q);
  Query::Device& d = *q.dev;
  if (!d.arena) {
    CUDA_CHECK(cudaEventRecord(d.ev[0], d.st));
    CUDA_CHECK(cudaMallocAsync(&d.arena, q.arena_bytes, d.st));
  }
  // (called again after the page walk: re-encoded PLAIN string pages join the list then)
  // LK_UPLOAD_STREAMS=n (tuning aid, default 1): the copies go round robin over n side streams that the query's stream then
  // waits for.  Measured on B200 / PCIe 5 x16 (C2, 1000 copies, 3.02 GB from pinned memory): see DESIGN.md §4.
  static const int n_up = std::min(4, std::max(1, getenv("LK_UPLOAD_STREAMS") ? atoi(getenv("LK_UPLOAD_STREAMS")) : 1));
  if (n_up <= 1 || d.uploads_done >= q.uploads.size()) {
    for (; d.uploads_done < q.uploads.size(); d.uploads_done++) {
      const Query::Upload& u = q.uploads[d.uploads_done];
      CUDA_CHECK(cudaMemcpyAsync(d.arena + u.arena_off, u.src ? u.src : q.segs[u.seg].data + u.file_off, u.len, cudaMemcpyHostToDevice, d.st));
    }
    return;
  }
  cudaStream_t up[4] = {};
  cudaEvent_t fork = nullptr, join[4] = {};
  CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(fork, d.st));  // the destination blocks are allocated in front of this point
  for (int k = 0; k < n_up; k++) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&up[k], cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&join[k], cudaEventDisableTiming));
    CUDA_CHECK(cudaStreamWaitEvent(up[k], fork, 0));
  }
  for (size_t k = 0; d.uploads_done < q.uploads.size(); d.uploads_done++, k++) {
    const Query::Upload& u = q.uploads[d.uploads_done];
    CUDA_CHECK(cudaMemcpyAsync(d.arena + u.arena_off, u.src ? u.src : q.segs[u.seg].data + u.file_off, u.len, cudaMemcpyHostToDevice, up[k % n_up]));
  }
  for (int k = 0; k < n_up; k++) {
    CUDA_CHECK(cudaEventRecord(join[k], up[k]));
    CUDA_CHECK(cudaStreamWaitEvent(d.st, join[k], 0));
    cudaEventDestroy(join[k]);   // (deferred by the runtime until the event has completed)
    cudaStreamDestroy(up[k]);    // (likewise: the stream's work drains first)
  }
  cudaEventDestroy(fork);
}

// ------------------------------------------------------------------------------------------------------------
// seek index on the device (Query::device_index): run tables of the hybrid streams + per-(tile, column) cursors
// ------------------------------------------------------------------------------------------------------------
// The host has walked the PAGE headers (lk_parquet.cpp, walk_runs = false); everything that used to cost ~0.6 host
// core-seconds per 105 M rows -- every run header of every definition-level and dictionary-index stream, the non-null
// prefix counts, the cursors of every tile -- is computed here, from the column chunks already in HBM:
//   idx_def_count    one thread per page: walk the definition-level runs, count them and the non-null values
//   idx_chunk_prefix one thread per chunk: first value index of every page
//   idx_val_count    one thread per page: walk the dictionary-index runs (needs the page's value count)
//   idx_layout       one block: lay the chunks' run lists out in the pool (definition runs, then value runs, per chunk)
//   idx_fill         one thread per page: walk both streams again, write the Run entries (+ non-null count before each def run)
//   idx_cursor       one thread per (tile, column): the ColCursor the host planner would have written (lk_plan.cpp)
// A stream is inherently sequential (varint headers), so the parallelism is pages x columns: ~1000 threads for 100 segments;
// the whole build is a few milliseconds on the GPU, overlapped with nothing yet (it needs the bytes in HBM).
struct IdxTotals { uint32_t n_runs, status, pad[2]; };

// ---- SNAPPY pages (format breadth): inflated on the device, a warp per page ----
// The element stream of a page is sequential (a tag says how many literal bytes follow or which earlier output bytes to
// repeat), so the parallelism is pages x columns -- thousands for a glob -- plus the 32 lanes that move each element's
// bytes together.  All lanes parse the tags redundantly (uniform loads), so no lane ever waits for a broadcast.
__global__ void __launch_bounds__(128) snappy_decode_kernel(uint8_t* arena, const ZPage* __restrict__ zp, uint32_t n, IdxTotals* __restrict__ tot) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const ZPage z = zp[w];
  if (!(z.flags & ZP_DECODE)) return;
  const uint8_t* __restrict__ src = arena + z.src_off;
  uint8_t* dst = arena + z.dst_off;
  const uint32_t slen = z.src_len, dlen = z.dst_len;
  uint32_t ip = 0, op = 0;
  bool bad = false;
  {  // preamble: the uncompressed length
    uint64_t ulen = 0;
    int sh = 0;
    while (true) {
      if (ip >= slen || sh > 35) { bad = true; break; }
      const uint8_t c = src[ip++];
      ulen |= (uint64_t)(c & 0x7f) << sh;
      if (!(c & 0x80)) break;
      sh += 7;
    }
    if (ulen != dlen) bad = true;
  }
  while (!bad && ip < slen) {
    const uint32_t tag = src[ip++];
    uint32_t len, off = 0;
    if ((tag & 3) == 0) {
      len = (tag >> 2) + 1;
      if (len > 60) {
        const uint32_t nb = len - 60;
        if (ip + nb > slen) { bad = true; break; }
        len = 0;
        for (uint32_t i = 0; i < nb; i++) len |= (uint32_t)src[ip + i] << (8 * i);
        len += 1;
        ip += nb;
      }
      if (len > slen - ip || len > dlen - op) { bad = true; break; }
      const uint8_t* s = src + ip;
      uint8_t* d = dst + op;
      // long literals (incompressible PLAIN values): 4-byte words once source and destination are aligned alike
      uint32_t i0 = 0;
      if (len >= 256 && (((uintptr_t)s ^ (uintptr_t)d) & 3) == 0) {
        const uint32_t head = (4 - ((uintptr_t)d & 3)) & 3;
        if (lane < head) d[lane] = s[lane];
        const uint32_t words = (len - head) >> 2;
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s + head);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(d + head);
        for (uint32_t i = lane; i < words; i += 32) d4[i] = s4[i];
        i0 = head + (words << 2);
      }
      for (uint32_t i = i0 + lane; i < len; i += 32) d[i] = s[i];
      ip += len;
      op += len;
    } else {
      if ((tag & 3) == 1) {
        if (ip + 1 > slen) { bad = true; break; }
        len = 4 + ((tag >> 2) & 7);
        off = ((tag >> 5) << 8) | src[ip];
        ip += 1;
      } else if ((tag & 3) == 2) {
        if (ip + 2 > slen) { bad = true; break; }
        len = (tag >> 2) + 1;
        off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
        ip += 2;
      } else {
        if (ip + 4 > slen) { bad = true; break; }
        len = (tag >> 2) + 1;
        off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
        ip += 4;
      }
      if (off == 0 || off > op || len > dlen - op) { bad = true; break; }
      // earlier output, possibly overlapping what is being written (a pattern of period `off`): every lane reads only
      // bytes in front of `op`, which the __syncwarp below has made visible
      const uint8_t* from = dst + (op - off);
      if (off >= len) { for (uint32_t i = lane; i < len; i += 32) dst[op + i] = from[i]; }
      else { for (uint32_t i = lane; i < len; i += 32) dst[op + i] = from[i % off]; }
      op += len;
    }
    __syncwarp();
  }
  if (!bad && op != dlen) bad = true;
  if (bad && lane == 0) atomicOr(&tot->status, (uint32_t)IDX_ST_BAD_SNAPPY);
}

// completes the IdxPage of every SNAPPY page from its inflated bytes: what the host reads from an uncompressed page's
// first bytes (lk_parquet.cpp: index_chunk) -- the length prefix of a V1 page's definition levels, the bit-width byte
__global__ void zpage_parse_kernel(const uint8_t* __restrict__ arena, const ZPage* __restrict__ zp, uint32_t n, IdxPage* __restrict__ pages, IdxTotals* __restrict__ tot) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ZPage z = zp[i];
  if (z.page == 0xffffffffu) return;
  IdxPage& pg = pages[z.page];
  uint64_t q = z.dst_off;
  const uint64_t end = z.dst_off + z.dst_len;
  uint32_t status = 0;
  if (z.flags & ZP_V1_DEF) {
    if (q + 4 > end) status |= IDX_ST_TRUNCATED;
    else {
      const uint32_t dl = (uint32_t)arena[q] | ((uint32_t)arena[q + 1] << 8) | ((uint32_t)arena[q + 2] << 16) | ((uint32_t)arena[q + 3] << 24);
      if (q + 4 + (uint64_t)dl > end) status |= IDX_ST_TRUNCATED;
      else { pg.def_off = q + 4; pg.def_end = q + 4 + dl; q = pg.def_end; }
    }
  }
  if ((z.flags & ZP_DICT_CODED) && q < end) {
    pg.bit_width = arena[q];
    if (pg.bit_width > 31) status |= IDX_ST_BAD_RUN;
    q += 1;
  }
  pg.val_off = q;
  pg.val_end = end;
  if (status) atomicOr(&tot->status, status);
}

__global__ void idx_def_count_kernel(const uint8_t* __restrict__ arena, IdxPage* __restrict__ pages, uint32_t npages, IdxTotals* __restrict__ tot) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npages) return;
  IdxPage pg = pages[i];
  uint32_t runs = 0, nn = pg.num_rows, status = 0;
  if (pg.def_end > pg.def_off) {
    nn = 0;
    uint64_t off = pg.def_off;
    uint32_t done = 0;
    while (done < pg.num_rows) {
      HybridRun r;
      status |= lk_hybrid_next(arena, off, pg.def_end, 1, pg.num_rows - done, r);
      if (status) break;
      nn += r.is_rle ? ((r.value & 1) ? r.n : 0u) : lk_popcount_bits(arena + r.payload, r.n);
      runs++;
      done += r.n;
      off = r.next;
    }
  }
  pages[i].def_runs = runs;
  pages[i].nn = nn;
  if (status) atomicOr(&tot->status, status);
}

__global__ void idx_chunk_prefix_kernel(IdxPage* __restrict__ pages, IdxChunk* __restrict__ chunks, uint32_t nchunks) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const IdxChunk ch = chunks[c];
  uint32_t v = 0;
  for (uint32_t k = 0; k < ch.npages; k++) {
    pages[ch.page0 + k].first_vidx = v;
    v += pages[ch.page0 + k].nn;
  }
  chunks[c].nn = v;
}

__global__ void idx_val_count_kernel(const uint8_t* __restrict__ arena, IdxPage* __restrict__ pages, const IdxChunk* __restrict__ chunks,
                                     uint32_t npages, IdxTotals* __restrict__ tot) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npages) return;
  const IdxPage pg = pages[i];
  uint32_t runs = 0, status = 0;
  if (pg.dict_coded) {
    uint64_t off = pg.val_off;
    uint32_t done = 0;
    const uint32_t dict_n = chunks[pg.chunk].dict_n;
    while (done < pg.nn) {
      HybridRun r;
      status |= lk_hybrid_next(arena, off, pg.val_end, pg.bit_width, pg.nn - done, r);
      if (status) break;
      if (r.is_rle && r.value >= dict_n) status |= IDX_ST_BAD_CODE;
      runs++;
      done += r.n;
      off = r.next;
    }
  } else {
    const IdxChunk ch = chunks[pg.chunk];
    if (ch.string_typed && pg.nn > 0) status |= IDX_ST_PLAIN_STRING;
    else if (pg.val_off + (uint64_t)pg.nn * ch.esz > pg.val_end) status |= IDX_ST_TRUNCATED;
  }
  pages[i].val_runs = runs;
  if (status) atomicOr(&tot->status, status);
}

// one block: per chunk the number of def / value runs, then an exclusive scan over the chunks (a chunk's runs are contiguous:
// its definition runs page by page, then its value runs page by page)
__global__ void __launch_bounds__(1024) idx_layout_kernel(IdxPage* __restrict__ pages, IdxChunk* __restrict__ chunks, uint32_t nchunks, IdxTotals* __restrict__ tot) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nchunks; base += 1024) {
    const uint32_t c = base + threadIdx.x;
    uint32_t nd = 0, nv = 0;
    IdxChunk ch;
    if (c < nchunks) {
      ch = chunks[c];
      for (uint32_t k = 0; k < ch.npages; k++) { nd += pages[ch.page0 + k].def_runs; nv += pages[ch.page0 + k].val_runs; }
    }
    const uint32_t x = nd + nv;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const uint32_t t = warp_tot[lane];
      uint32_t ti = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += o; }
      warp_tot[lane] = ti - t;
    }
    __syncthreads();
    uint32_t run = carry_s + warp_tot[wid] + incl - x;
    if (c < nchunks) {
      chunks[c].def_run0 = run;
      chunks[c].def_runs = nd;
      for (uint32_t k = 0; k < ch.npages; k++) { pages[ch.page0 + k].def_run0 = run; run += pages[ch.page0 + k].def_runs; }
      chunks[c].val_run0 = run;
      chunks[c].val_runs = nv;
      for (uint32_t k = 0; k < ch.npages; k++) { pages[ch.page0 + k].val_run0 = run; run += pages[ch.page0 + k].val_runs; }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) tot->n_runs = carry_s;
}

__global__ void idx_fill_kernel(const uint8_t* __restrict__ arena, const IdxPage* __restrict__ pages, const IdxChunk* __restrict__ chunks, uint32_t npages,
                                Run* __restrict__ runs, uint32_t* __restrict__ nn_before) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npages) return;
  const IdxPage pg = pages[i];
  const uint64_t base = chunks[pg.chunk].base_off;
  if (pg.def_end > pg.def_off) {
    uint64_t off = pg.def_off;
    uint32_t done = 0, nn = pg.first_vidx, k = pg.def_run0;
    while (done < pg.num_rows) {
      HybridRun r;
      if (lk_hybrid_next(arena, off, pg.def_end, 1, pg.num_rows - done, r)) break;
      Run run;
      run.start = pg.first_row + done;
      run.kind_value = r.is_rle ? (0x80000000u | (r.value & 1)) : (uint32_t)(r.payload - base);
      runs[k] = run;
      nn_before[k] = nn;
      k++;
      nn += r.is_rle ? ((r.value & 1) ? r.n : 0u) : lk_popcount_bits(arena + r.payload, r.n);
      done += r.n;
      off = r.next;
    }
  }
  if (pg.dict_coded) {
    uint64_t off = pg.val_off;
    uint32_t done = 0, k = pg.val_run0;
    while (done < pg.nn) {
      HybridRun r;
      if (lk_hybrid_next(arena, off, pg.val_end, pg.bit_width, pg.nn - done, r)) break;
      Run run;
      run.start = pg.first_vidx + done;
      run.kind_value = r.is_rle ? (0x80000000u | (r.value & 0x7fffffffu)) : (uint32_t)(r.payload - base);
      runs[k++] = run;
      done += r.n;
      off = r.next;
    }
  }
}

// non-null values of the chunk before row r (r may equal num_rows), and the definition run that holds r
__device__ __forceinline__ uint32_t idx_vidx_at(const uint8_t* __restrict__ arena, const Run* __restrict__ dr, const uint32_t* __restrict__ nnb, uint32_t nd,
                                                uint64_t base, uint32_t r, uint32_t& run_index) {
  uint32_t lo = 0, hi = nd;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (dr[mid].start <= r) lo = mid; else hi = mid;
  }
  run_index = lo;
  const Run run = dr[lo];
  const uint32_t k = r - run.start;
  if (run.kind_value >> 31) return nnb[lo] + ((run.kind_value & 1) ? k : 0u);
  return nnb[lo] + lk_popcount_bits(arena + base + run.kind_value, k);
}

__global__ void idx_cursor_kernel(const uint8_t* __restrict__ arena, const TileDesc* __restrict__ tiles, uint32_t ntiles, uint32_t np,
                                  const IdxPage* __restrict__ pages, IdxChunk* __restrict__ chunks, const Run* __restrict__ runs,
                                  const uint32_t* __restrict__ nn_before, ColCursor* __restrict__ cursors) {
  const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= (uint64_t)ntiles * np) return;
  const uint32_t t = (uint32_t)(gi / np), p = (uint32_t)(gi - (uint64_t)t * np);
  const TileDesc td = tiles[t];
  const uint32_t cidx = td.rg * np + p;
  const IdxChunk ch = chunks[cidx];
  ColCursor c;
  memset(&c, 0, sizeof c);
  if (!ch.present) {
    c.flags = CUR_ALL_NULL;
    cursors[gi] = c;
    return;
  }
  const uint32_t r0 = td.row0, r1 = td.row0 + td.nrows;
  uint32_t v0, v1, d0 = 0, d1 = 0;
  const Run* dr = runs + ch.def_run0;
  const uint32_t* nnb = nn_before + ch.def_run0;
  if (ch.max_def == 0 || ch.def_runs == 0) {
    v0 = ch.max_def == 0 ? r0 : 0u;
    v1 = ch.max_def == 0 ? r1 : 0u;
  } else {
    v0 = idx_vidx_at(arena, dr, nnb, ch.def_runs, ch.base_off, r0, d0);
    v1 = idx_vidx_at(arena, dr, nnb, ch.def_runs, ch.base_off, r1, d1);
  }
  c.vidx0 = v0;
  c.nvals = v1 - v0;
  if (c.nvals == r1 - r0) c.flags |= CUR_ALL_VALID;
  else if (c.nvals == 0) c.flags |= CUR_ALL_NULL;
  else {
    if (dr[d1].start >= r1) d1--;  // the run holding row r1 - 1
    c.drun_lo = ch.def_run0 + d0;
    c.drun_n = (uint16_t)(d1 - d0 + 1);
    chunks[cidx].mixed = 1;  // (benign race: every writer stores 1)
  }
  if (c.nvals > 0) {
    uint32_t pi = 0;
    while (pi + 1 < ch.npages && pages[ch.page0 + pi + 1].first_row <= r0) pi++;
    const IdxPage pg = pages[ch.page0 + pi];
    if (pg.dict_coded) {
      c.flags |= CUR_DICT;
      c.width = pg.bit_width;
      const Run* vr = runs + ch.val_run0;
      uint32_t lo = 0, hi = ch.val_runs;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (vr[mid].start <= v0) lo = mid; else hi = mid; }
      uint32_t zlo = lo, zhi = ch.val_runs;
      while (zhi - zlo > 1) { const uint32_t mid = (zlo + zhi) >> 1; if (vr[mid].start <= v1 - 1) zlo = mid; else zhi = mid; }
      c.vrun_lo = ch.val_run0 + lo;
      c.vrun_n = (uint16_t)(zlo - lo + 1);
    } else {
      c.plain_off = pg.val_off + (uint64_t)(v0 - pg.first_vidx) * ch.esz;
    }
  }
  cursors[gi] = c;
}

// Builds runs / cursors / definition-chunk list on the device (after the column chunks are in HBM) and completes what the host
// plan left open.  Two small read-backs: the run total (sizes the pool) and the per-chunk run ranges / mixed flags.
static void device_build_index(Query& q) {
  Query::Device& d = *q.dev;
  const uint32_t npages = (uint32_t)q.idx_pages.size(), nchunks = (uint32_t)q.idx_chunks.size(), np = (uint32_t)q.pcols.size();
  const uint32_t ntiles = (uint32_t)q.tiles.size();
  IdxPage* dpages = nullptr;
  IdxChunk* dchunks = nullptr;
  IdxTotals* dtot = nullptr;
  uint32_t* nn_before = nullptr;
  CUDA_CHECK(cudaMallocAsync(&dpages, std::max<size_t>(1, npages) * sizeof(IdxPage), d.st));
  CUDA_CHECK(cudaMallocAsync(&dchunks, std::max<size_t>(1, nchunks) * sizeof(IdxChunk), d.st));
  CUDA_CHECK(cudaMallocAsync(&dtot, sizeof(IdxTotals), d.st));
  CUDA_CHECK(cudaMemsetAsync(dtot, 0, sizeof(IdxTotals), d.st));
  if (npages) CUDA_CHECK(cudaMemcpyAsync(dpages, q.idx_pages.data(), npages * sizeof(IdxPage), cudaMemcpyHostToDevice, d.st));
  if (nchunks) CUDA_CHECK(cudaMemcpyAsync(dchunks, q.idx_chunks.data(), nchunks * sizeof(IdxChunk), cudaMemcpyHostToDevice, d.st));
  IdxTotals tot{};
  ZPage* dz = nullptr;
  if (npages && !q.zpages.empty()) {  // SNAPPY pages first: inflate (what no cached block holds yet), then complete their IdxPage
    const uint32_t nz = (uint32_t)q.zpages.size();
    CUDA_CHECK(cudaMallocAsync(&dz, nz * sizeof(ZPage), d.st));
    CUDA_CHECK(cudaMemcpyAsync(dz, q.zpages.data(), nz * sizeof(ZPage), cudaMemcpyHostToDevice, d.st));
    snappy_decode_kernel<<<(int)((nz + 3) / 4), 128, 0, d.st>>>(d.arena, dz, nz, dtot);
    zpage_parse_kernel<<<(int)((nz + 127) / 128), 128, 0, d.st>>>(d.arena, dz, nz, dpages, dtot);
    CUDA_CHECK(cudaGetLastError());
  }
  if (npages) {
    const int pg_grid = (int)((npages + 31) / 32), ch_grid = (int)((nchunks + 63) / 64);
    idx_def_count_kernel<<<pg_grid, 32, 0, d.st>>>(d.arena, dpages, npages, dtot);
    idx_chunk_prefix_kernel<<<ch_grid, 64, 0, d.st>>>(dpages, dchunks, nchunks);
    idx_val_count_kernel<<<pg_grid, 32, 0, d.st>>>(d.arena, dpages, dchunks, npages, dtot);
    idx_layout_kernel<<<1, 1024, 0, d.st>>>(dpages, dchunks, nchunks, dtot);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(&tot, dtot, sizeof tot, cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    LK_CHECK(!(tot.status & IDX_ST_BAD_SNAPPY), LK_ERR_IO, "parquet: malformed SNAPPY page");
    LK_CHECK(!(tot.status & IDX_ST_PLAIN_STRING), LK_ERR_UNSUPPORTED, "a string column has PLAIN (non-dictionary) pages");
    LK_CHECK(!(tot.status & IDX_ST_BAD_CODE), LK_ERR_IO, "parquet: dictionary index out of range");
    LK_CHECK(!(tot.status & IDX_ST_BAD_RUN), LK_ERR_IO, "parquet: malformed run header in a hybrid stream");
    LK_CHECK(!(tot.status & IDX_ST_TRUNCATED), LK_ERR_IO, "parquet: hybrid stream / PLAIN page ends before all values are covered");
  }
  q.n_runs = tot.n_runs;
  if (d.runs) { CUDA_CHECK(cudaFreeAsync(d.runs, d.st)); d.runs = nullptr; }
  if (d.cursors) { CUDA_CHECK(cudaFreeAsync(d.cursors, d.st)); d.cursors = nullptr; }
  CUDA_CHECK(cudaMallocAsync(&d.runs, std::max<size_t>(1, tot.n_runs) * sizeof(Run), d.st));
  CUDA_CHECK(cudaMallocAsync(&nn_before, std::max<size_t>(1, tot.n_runs) * sizeof(uint32_t), d.st));
  CUDA_CHECK(cudaMallocAsync(&d.cursors, std::max<size_t>(1, (size_t)ntiles * np) * sizeof(ColCursor), d.st));
  if (npages) {
    idx_fill_kernel<<<(int)((npages + 31) / 32), 32, 0, d.st>>>(d.arena, dpages, dchunks, npages, d.runs, nn_before);
    if (ntiles) {
      const uint64_t n = (uint64_t)ntiles * np;
      idx_cursor_kernel<<<(int)((n + 127) / 128), 128, 0, d.st>>>(d.arena, d.tiles, ntiles, np, dpages, dchunks, d.runs, nn_before, d.cursors);
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(q.idx_chunks.data(), dchunks, nchunks * sizeof(IdxChunk), cudaMemcpyDeviceToHost, d.st));
  }
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  // definition bitmaps for the chunks that have a tile mixing NULLs and values (same rule as the host planner)
  std::vector<DefChunk> def_tmp(nchunks, DefChunk{});
  uint32_t def_mask = 0;
  for (uint32_t k = 0; k < nchunks; k++) {
    const IdxChunk& ic = q.idx_chunks[k];
    if (!ic.present || !ic.mixed) continue;
    DefChunk& dc = def_tmp[k];
    dc.base_off = ic.base_off;
    dc.run_lo = ic.def_run0;
    dc.run_n = ic.def_runs;
    dc.num_rows = ic.num_rows;
    def_mask |= 1u << (k % np);
  }
  layout_def_chunks(q, def_tmp);
  q.params.def_mask = def_mask;
  refresh_info_json(q);
  if (dz) CUDA_CHECK(cudaFreeAsync(dz, d.st));
  CUDA_CHECK(cudaFreeAsync(dpages, d.st));
  CUDA_CHECK(cudaFreeAsync(dchunks, d.st));
  CUDA_CHECK(cudaFreeAsync(dtot, d.st));
  CUDA_CHECK(cudaFreeAsync(nn_before, d.st));
}

void device_upload(Query& q) {
  device_begin_upload(q);
  Query::Device& d = *q.dev;
  upload_vec(d.tiles, q.tiles, d.st);
  if (q.device_index) device_build_index(q);  // runs, cursors and the definition-chunk list come from the device
  else {
    upload_vec(d.cursors, q.cursors, d.st);
    upload_vec(d.runs, q.runs, d.st);
  }
  upload_vec(d.def_chunks, q.def_chunks, d.st);
  if (d.defbm) { CUDA_CHECK(cudaFreeAsync(d.defbm, d.st)); d.defbm = nullptr; }
  CUDA_CHECK(cudaMallocAsync(&d.defbm, std::max<size_t>(q.defbm_words * 4, 16), d.st));
  upload_vec(d.lut_cls, q.lut_cls, d.st);
  upload_vec(d.pass_bits, q.pass_bits, d.st);
  CUDA_CHECK(cudaMallocAsync(&d.counters, 8 * sizeof(uint32_t), d.st));
  CUDA_CHECK(cudaMallocAsync(&d.survivors, sizeof(unsigned long long), d.st));
  upload_group_tables(q);
  CUDA_CHECK(cudaEventRecord(d.ev[1], d.st));
  // the borrowed host buffers may be released by the caller once prepare returns
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  d.resident = true;
  cache_publish(q);
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
  q.t_ms[0] = ms;
}

void device_mark_group_tables_stale(Query& q) {
  if (q.dev) q.dev->group_tables_stale = true;
}

// The single-filter fast path needs every dictionary of the one (string) filter column to fit the per-warp
// code->pass bitmap; one oversized dictionary anywhere in the glob sends the whole query down the generic path.
static bool single_filter_ok(const Query& q) {
  const ScanParams& P = q.params;
  if (P.n_filter != 1 || P.filter[0].numeric) return false;
  const int np = (int)q.pcols.size(), p = P.filter[0].pcol;
  for (size_t i = 0; i < q.rgs.size(); i++)
    if (q.chunk_infos[i * np + p].dict_n > SCAN_CODEPASS_MAX) return false;
  return true;
}

template <int PATH, bool SINGLE, bool EMIT, int NA>
static void launch_scan_variant(const ScanParams& P, cudaStream_t st) {
  // persistent grid: every warp pulls 512-row tiles from the ticket counter until none are left
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, scan_kernel<PATH, SINGLE, EMIT, NA>, SCAN_BLOCK, 0));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  const int grid = (int)std::min<uint32_t>((P.ntiles + SCAN_WARPS - 1) / SCAN_WARPS, (uint32_t)(num_sms() * ctas_per_sm));
  scan_kernel<PATH, SINGLE, EMIT, NA><<<grid, SCAN_BLOCK, 0, st>>>(P);
  CUDA_CHECK(cudaGetLastError());
}

template <int PATH, bool EMIT>
static void launch_scan_table(const ScanParams& P, bool single, cudaStream_t st) {
  const bool few = P.n_aggs <= 4;  // the common case gets the kernel with 4 unrolled aggregate slots
  if (single) {
    if (few) launch_scan_variant<PATH, true, EMIT, 4>(P, st);
    else launch_scan_variant<PATH, true, EMIT, LK_MAX_AGGS>(P, st);
  } else {
    if (few) launch_scan_variant<PATH, false, EMIT, 4>(P, st);
    else launch_scan_variant<PATH, false, EMIT, LK_MAX_AGGS>(P, st);
  }
}


// ------------------------------------------------------------------------------------------------------------
// definition levels: hybrid RLE / bit-packed stream -> one bit per row
// ------------------------------------------------------------------------------------------------------------
// A nullable column's definition levels alternate between short bit-packed groups (the 8 rows around a NULL) and RLE
// runs of "valid": ~20 runs per 512-row tile at 5 % NULLs.  Walking them inside the scan costs every tile a run search and
// 2-3 dependent loads per column; instead the runs are expanded into a flat bitmap first (bit-packed, width 1: the payload
// bits are the bitmap bits; RLE of 1: a range of ones) and the scan reads 16 bits per lane.
// One CTA takes LK_DEF_BLOCK_RUNS consecutive runs of one chunk, every thread DX_PER_THREAD consecutive ones: a thread's
// runs cover one contiguous row range, so it assembles the bitmap words sequentially in a register (RLE and bit-packed
// runs share one code path: the pattern is the payload bits, all ones or all zeros) and hands every finished word to the
// CTA's window in shared memory; the window is stored once, and only its first and last word -- shared with the
// neighbouring CTAs -- go through a global atomicOr.  Runs reaching beyond the window (long stretches without NULLs)
// are written to global memory directly: edge words by atomicOr, whole words by the CTA together.
constexpr int DX_BLOCK = 256;
constexpr uint32_t DX_WIN = 1024;  // window words: 32768 rows (1024 runs at 5 % NULLs cover ~12 k)
constexpr int DX_PER_THREAD = LK_DEF_BLOCK_RUNS / DX_BLOCK;
constexpr uint32_t DX_QUEUE = 128;
struct DxOr {
  uint32_t* w;
  __device__ __forceinline__ void operator()(uint32_t word, uint32_t mask) const { atomicOr(w + word, mask); }
};
struct DxFillQueue {  // the whole words inside a long run: short stretches by the thread, long ones queued for the CTA
  uint32_t* w;
  uint2* q;
  uint32_t* n;
  __device__ __forceinline__ void operator()(uint32_t a, uint32_t b) const {
    if (b - a > 64) {
      const uint32_t slot = atomicAdd(n, 1u);
      if (slot < DX_QUEUE) { q[slot] = make_uint2(a, b); return; }
    }
    for (uint32_t i = a; i < b; i++) w[i] = 0xffffffffu;
  }
};
// a run that does not fit the shared window (rare): one out-of-line copy
__device__ __noinline__ void def_expand_direct(const uint8_t* arena, uint64_t base_off, Run r, uint32_t next, uint32_t* w, uint2* q, uint32_t* nq) {
  lk_def_expand_run(arena, base_off, r, next, DxOr{w}, DxFillQueue{w, q, nq});
}
__global__ void __launch_bounds__(DX_BLOCK) def_expand_kernel(const uint8_t* __restrict__ arena, const Run* __restrict__ runs, const DefChunk* __restrict__ dcs,
                                                              uint32_t dc0, uint32_t* __restrict__ bm) {
  __shared__ uint32_t win[DX_WIN];
  __shared__ uint2 fillq[DX_QUEUE];
  __shared__ uint32_t sh_row_lo, sh_row_hi, nfill, ndirect;
  const DefChunk dc = dcs[dc0 + blockIdx.y];
  const uint32_t k0 = blockIdx.x * LK_DEF_BLOCK_RUNS;
  if (k0 >= dc.run_n) return;  // the grid is as wide as the chunk with the most runs
  const uint32_t k1 = min(dc.run_n, k0 + LK_DEF_BLOCK_RUNS);
  const Run* __restrict__ r0 = runs + dc.run_lo;
  uint32_t* __restrict__ w = bm + dc.word0;
  // my DX_PER_THREAD consecutive runs, the start of the one after them, and the first payload word of the bit-packed ones
  const uint32_t kt = k0 + threadIdx.x * DX_PER_THREAD;
  uint32_t start[DX_PER_THREAD + 1], kv[DX_PER_THREAD], first[DX_PER_THREAD];
#pragma unroll
  for (int j = 0; j < DX_PER_THREAD; j++) {
    Run r;
    r.start = dc.num_rows;
    r.kind_value = 0x80000000u;  // beyond the chunk: an empty RLE run of 0
    if (kt + j < dc.run_n) r = r0[kt + j];
    start[j] = r.start;
    kv[j] = r.kind_value;
  }
  start[DX_PER_THREAD] = kt + DX_PER_THREAD < dc.run_n ? r0[kt + DX_PER_THREAD].start : dc.num_rows;
  for (uint32_t i = threadIdx.x; i < DX_WIN; i += DX_BLOCK) win[i] = 0;
  if (threadIdx.x == 0) { sh_row_lo = start[0]; nfill = 0; ndirect = 0; }
#pragma unroll
  for (int j = 0; j < DX_PER_THREAD; j++)
    if (kt + j + 1 == k1) sh_row_hi = start[j + 1];  // the thread holding the CTA's last run
#pragma unroll
  for (int j = 0; j < DX_PER_THREAD; j++) {
    first[j] = 0;
    if (!(kv[j] >> 31) && start[j + 1] > start[j]) first[j] = lk_load_u32_unaligned(arena + dc.base_off + kv[j]);
  }
  __syncthreads();
  const uint32_t row_lo = sh_row_lo, row_hi = sh_row_hi;
  const uint32_t wbase = row_lo >> 5;
  // sequential assembly: `acc` holds my bits of word (cur >> 5) below bit (cur & 31)
  uint32_t acc = 0, cur = start[0];
  auto emit = [&]() {  // hand the word under the cursor to the window
    if (acc) atomicOr(win + ((cur >> 5) - wbase), acc);
    acc = 0;
  };
#pragma unroll
  for (int j = 0; j < DX_PER_THREAD; j++) {
    if (kt + j >= k1 || start[j + 1] <= start[j]) continue;
    const uint32_t n = start[j + 1] - start[j];
    if (((start[j + 1] - 1) >> 5) - wbase >= DX_WIN) {  // does not fit the window
      if (((cur >> 5) - wbase) < DX_WIN) emit(); else acc = 0;
      ndirect = 1;
      Run r;
      r.start = start[j];
      r.kind_value = kv[j];
      def_expand_direct(arena, dc.base_off, r, start[j + 1], w, fillq, &nfill);
      cur = start[j + 1];
      continue;
    }
    const uint32_t sh = cur & 31;
    if (kv[j] >> 31) {
      // RLE: n equal bits in constant time -- the rest of the word under the cursor, whole words (no one else writes
      // them: plain stores), and the low bits of the last word stay in the accumulator
      const uint32_t ones = (kv[j] & 1) ? 0xffffffffu : 0u;
      if (sh + n < 32) acc |= (ones & ((1u << n) - 1)) << sh;
      else {
        acc |= ones << sh;
        emit();
        const uint32_t wi = (cur >> 5) - wbase, full = ((sh + n) >> 5) - 1;
        if (ones) for (uint32_t i = 1; i <= full; i++) win[wi + i] = 0xffffffffu;
        const uint32_t rem = (sh + n) & 31;
        acc = rem ? (ones & ((1u << rem) - 1)) : 0u;
      }
      cur += n;
      continue;
    }
    for (uint32_t o = 0; o < n; o += 32) {  // bit-packed: one trip unless the run is longer than 32 rows
      uint32_t pat = o == 0 ? first[j] : lk_load_u32_unaligned(arena + dc.base_off + kv[j] + (o >> 3));
      const uint32_t c = min(n - o, 32u);
      if (c < 32) pat &= (1u << c) - 1;
      const uint32_t s2 = cur & 31;
      acc |= pat << s2;
      if (s2 + c >= 32) {  // the word is complete
        emit();
        acc = s2 ? pat >> (32 - s2) : 0u;
      }
      cur += c;
    }
  }
  if ((cur & 31) && ((cur >> 5) - wbase) < DX_WIN) emit();
  __syncthreads();
  if (row_hi <= row_lo) return;
  const uint32_t nq = min(nfill, DX_QUEUE);
  const bool all_atomic = ndirect != 0;  // a direct run may own words of this range
  const uint32_t nw = min(DX_WIN, ((row_hi - 1) >> 5) - wbase + 1);
  for (uint32_t i = threadIdx.x; i < nw; i += DX_BLOCK) {
    const uint32_t v = win[i];
    if (i == 0 || i + 1 == nw || all_atomic) { if (v) atomicOr(w + wbase + i, v); }
    else w[wbase + i] = v;
  }
  for (uint32_t e = 0; e < nq; e++) {
    const uint2 f = fillq[e];
    for (uint32_t i = f.x + threadIdx.x; i < f.y; i += DX_BLOCK) w[i] = 0xffffffffu;
  }
}

static void launch_def_expand(const Query& q, const ScanParams& P) {
  const Query::Device& d = *q.dev;
  if (q.def_chunks.empty()) return;
  CUDA_CHECK(cudaMemsetAsync(d.defbm, 0, q.defbm_words * 4, d.st));
  uint32_t max_runs = 0;
  for (auto& dc : q.def_chunks) max_runs = std::max(max_runs, dc.run_n);
  const uint32_t gx = (max_runs + LK_DEF_BLOCK_RUNS - 1) / LK_DEF_BLOCK_RUNS;
  for (size_t c0 = 0; c0 < q.def_chunks.size(); c0 += 65535) {  // grid.y = chunk
    const uint32_t gy = (uint32_t)std::min<size_t>(65535, q.def_chunks.size() - c0);
    def_expand_kernel<<<dim3(gx, gy), DX_BLOCK, 0, d.st>>>(P.arena, P.runs, d.def_chunks, (uint32_t)c0, d.defbm);
  }
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// communicator of the sharded record path (replaces the HTTP/SSE fan-in of SegmentSequencer.scala:137-158 +
// the re-aggregation of TimeGroupedSketchAggregator.scala:157-177 for partial results of the same query on several GPUs)
// ------------------------------------------------------------------------------------------------------------
// block of rank r: CommCtrl (4 KB) | keys pool 0 | keys pool 1 | vals pool 0 | vals pool 1; a pool = world regions of
// `region_cap` records, region s written by source s only
struct Comm {
  int rank = 0, world = 1, max_aggs = 4, device = 0;
  size_t region_cap = 0;
  uint8_t* block = nullptr;
  size_t block_bytes = 0;
  uint8_t* peer[LK_MAX_RANKS] = {};
  bool peer_ipc[LK_MAX_RANKS] = {};
  uint32_t* send_count = nullptr;  // local: records written per destination in the current epoch
  uint32_t epoch = 0;
  bool connected = false;
  std::string blob;
  size_t cap() const { return region_cap * (size_t)world; }
  CommCtrl* ctrl(int r) const { return reinterpret_cast<CommCtrl*>(peer[r]); }
  unsigned long long* keys(int r, uint32_t pool) const { return reinterpret_cast<unsigned long long*>(peer[r] + 4096) + (size_t)pool * cap(); }
  unsigned long long* vals(int r, uint32_t pool) const {
    return reinterpret_cast<unsigned long long*>(peer[r] + 4096) + 2 * cap() + (size_t)pool * cap() * max_aggs;
  }
};
struct CommBlob {
  uint32_t magic;
  int32_t rank, world, device, max_aggs;
  int64_t pid;
  uint64_t raw_ptr, region_cap;
  cudaIpcMemHandle_t handle;
};
constexpr uint32_t COMM_MAGIC = 0x6c6b636du;

int comm_world(const Comm* c) { return c ? c->world : 1; }

Comm* comm_create(int rank, int world, int64_t pool_records, int max_aggs) {
  device_init();
  LK_CHECK(world >= 1 && world <= LK_MAX_RANKS && rank >= 0 && rank < world, LK_ERR_INVALID, "lk_comm_create: bad rank / world");
  LK_CHECK(max_aggs >= 1 && max_aggs <= LK_MAX_AGGS, LK_ERR_INVALID, "lk_comm_create: bad max_aggs");
  LK_CHECK(pool_records > 0 && pool_records < (1ll << 31), LK_ERR_INVALID, "lk_comm_create: pool_records out of range");
  auto c = std::make_unique<Comm>();
  c->rank = rank;
  c->world = world;
  c->max_aggs = max_aggs;
  c->device = global_options().device;
  static_assert(sizeof(CommCtrl) <= 4096, "control block");
  c->region_cap = (((size_t)pool_records + world - 1) / world + 63) & ~(size_t)63;  // every source may send this many records to one owner
  c->block_bytes = 4096 + 2 * c->cap() * 8 * (size_t)(1 + max_aggs);
  cudaError_t e = cudaMalloc(&c->block, c->block_bytes);  // (not from the stream-ordered pool: IPC handles need a plain allocation)
  if (e != cudaSuccess) { cudaGetLastError(); fail(LK_ERR_NOMEM, strf("lk_comm_create: %zu bytes of receive pools: %s", c->block_bytes, cudaGetErrorString(e))); }
  CUDA_CHECK(cudaMemset(c->block, 0, 4096));
  CUDA_CHECK(cudaMalloc(&c->send_count, LK_MAX_RANKS * sizeof(uint32_t)));
  CommBlob b;
  memset(&b, 0, sizeof b);
  b.magic = COMM_MAGIC;
  b.rank = rank;
  b.world = world;
  b.device = c->device;
  b.max_aggs = max_aggs;
  b.pid = (int64_t)getpid();
  b.raw_ptr = (uint64_t)(uintptr_t)c->block;
  b.region_cap = c->region_cap;
  CUDA_CHECK(cudaIpcGetMemHandle(&b.handle, c->block));
  c->blob.assign(reinterpret_cast<const char*>(&b), sizeof b);
  c->peer[rank] = c->block;
  CUDA_CHECK(cudaDeviceSynchronize());
  return c.release();
}

void comm_handle(Comm* c, const void** blob, size_t* len) {
  *blob = c->blob.data();
  *len = c->blob.size();
}

void comm_connect(Comm* c, const void* blobs, size_t len_each) {
  LK_CHECK(len_each == sizeof(CommBlob), LK_ERR_INVALID, "lk_comm_connect: handle blobs of the wrong size");
  CUDA_CHECK(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; r++) {
    CommBlob b;
    memcpy(&b, static_cast<const uint8_t*>(blobs) + (size_t)r * len_each, sizeof b);
    LK_CHECK(b.magic == COMM_MAGIC && b.rank == r && b.world == c->world, LK_ERR_INVALID, strf("lk_comm_connect: blob %d is not rank %d of %d", r, r, c->world));
    LK_CHECK(b.region_cap == c->region_cap && b.max_aggs == c->max_aggs, LK_ERR_INVALID, "lk_comm_connect: ranks created their pools with different sizes");
    if (r == c->rank) continue;
    if (b.pid == (int64_t)getpid()) {
      // same process (one thread per GPU, or the single-GPU tests): the allocation is addressable as it is
      if (b.device != c->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) fail(LK_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      c->peer[r] = reinterpret_cast<uint8_t*>((uintptr_t)b.raw_ptr);
    } else {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); fail(LK_ERR_CUDA, strf("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e))); }
      c->peer[r] = static_cast<uint8_t*>(p);
      c->peer_ipc[r] = true;
    }
  }
  c->connected = true;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; r++)
    if (c->peer_ipc[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->send_count) cudaFree(c->send_count);
  if (c->block) cudaFree(c->block);
  cudaGetLastError();
  delete c;
}

// end of the scan: tell every owner how many records this rank wrote into its region of the owner's pool (the status flags
// and the timestamp phase of the local scan ride along), then raise this rank's flag there.  The stores of the scan kernel
// are complete when this kernel starts (stream order); the fences order count / status before the flag.
__global__ void comm_publish_kernel(const __grid_constant__ XchgParams X, uint32_t epoch, const uint32_t* __restrict__ counters) {
  const uint32_t d = threadIdx.x;
  if (d >= X.world) return;
  CommCtrl* c = X.ctrl[d];
  const uint32_t n = X.count[d];
  c->count[epoch & 1][X.rank] = min(n, X.region_cap);
  c->status[X.rank] = counters[0] | (n > X.region_cap ? (uint32_t)ST_HASH_FULL : 0u);
  c->phase_min[X.rank] = counters[1];
  c->phase_max[X.rank] = counters[2];
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t*>(&c->flag[X.rank]) = epoch;
}

// before finalize: wait until every source has delivered this epoch (device-side, no host in the loop), then fold the sources'
// status flags and timestamp phases into the local counters and lay the sources' record counts out as one list
__global__ void comm_wait_kernel(CommCtrl* mine, uint32_t world, uint32_t epoch, uint32_t* counters) {
  const uint32_t lane = threadIdx.x;
  uint32_t status = 0, pmin = 0xffffffffu, pmax = 0, cnt = 0;
  if (lane < world) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(*reinterpret_cast<volatile uint32_t*>(&mine->flag[lane]) - epoch) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 20000000000ull) { status |= ST_XCHG_TIMEOUT; break; }  // 20 s: a peer died or never ran this query
      __nanosleep(100);
    }
    __threadfence_system();
    status |= *reinterpret_cast<volatile uint32_t*>(&mine->status[lane]);
    pmin = *reinterpret_cast<volatile uint32_t*>(&mine->phase_min[lane]);
    pmax = *reinterpret_cast<volatile uint32_t*>(&mine->phase_max[lane]);
    cnt = *reinterpret_cast<volatile uint32_t*>(&mine->count[epoch & 1][lane]);
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= (uint32_t)dd) incl += o; }
  if (lane <= LK_MAX_RANKS) mine->prefix[lane] = incl - cnt;  // (lanes >= world hold the total)
#pragma unroll
  for (int dd = 16; dd; dd >>= 1) {
    status |= __shfl_xor_sync(0xffffffffu, status, dd);
    pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, dd));
    pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, dd));
  }
  if (lane == 31) {
    counters[0] = status;
    counters[1] = pmin;
    counters[2] = pmax;
    counters[5] = incl;  // records of all sources
  }
}

static void comm_fill_params(const Comm& c, XchgParams& X, uint32_t pool, size_t n_aggs) {
  memset(&X, 0, sizeof X);
  X.world = (uint32_t)c.world;
  X.rank = (uint32_t)c.rank;
  X.region_cap = (uint32_t)c.region_cap;
  for (int r = 0; r < c.world; r++) {
    X.keys[r] = c.keys(r, pool) + (size_t)c.rank * c.region_cap;
    X.vals[r] = c.vals(r, pool) + (size_t)c.rank * c.region_cap * n_aggs;  // rows of n_aggs words: row of record p = vals + p * n_aggs
    X.ctrl[r] = c.ctrl(r);
  }
  X.count = c.send_count;
}

static void launch_scan(const Query& q, const ScanParams& P, bool emit) {
  cudaStream_t st = q.dev->st;
  const bool single = single_filter_ok(q);
  if (emit && P.path < 2) launch_scan_table<0, true>(P, single, st);  // record pass of exact_sums over a table: cells are written out
  else if (P.path == 0) launch_scan_table<0, false>(P, single, st);
  else if (P.path == 1) launch_scan_table<1, false>(P, single, st);
  else if (P.path == 2 && !emit) launch_scan_table<2, false>(P, single, st);
  else if (P.path == 2) launch_scan_table<2, true>(P, single, st);   // exact_sums on the record path: rows carry their sequence number
  else if (!emit) launch_scan_table<3, false>(P, single, st);        // record path, sharded: appends go to the owner ranks' pools
  else launch_scan_table<3, true>(P, single, st);
}

void device_execute(Query& q) {
  LK_CHECK(q.dev && q.dev->resident, LK_ERR_INVALID, "lk_query_execute before lk_query_prepare");
  Query::Device& d = *q.dev;
  CUDA_CHECK(cudaSetDevice(global_options().device));
  if (d.group_tables_stale) upload_group_tables(q);
  if (d.executed && !d.finalized_device && d.harena)  // re-execute without emit: the arena still holds the last run
    CUDA_CHECK(cudaMemsetAsync(d.harena->entries, 0, d.harena->slots * d.harena->stride, d.st));
  ScanParams P = q.params;
  P.arena = d.arena;
  P.tiles = d.tiles;
  P.cursors = d.cursors;
  P.runs = d.runs;
  P.chunks = d.chunks;
  P.defbm = d.defbm;
  P.lut_cls = d.lut_cls;
  P.lut_gcode = d.lut_gcode;
  P.pass_bits = d.pass_bits;
  P.counters = d.counters;
  P.survivors = d.survivors;
  CUDA_CHECK(cudaEventRecord(d.ev[8], d.st));
  launch_def_expand(q, P);  // part of every execute: the definition levels are decoded on the device, inside the timed step
  CUDA_CHECK(cudaEventRecord(d.ev[9], d.st));
  static const bool no_preclear = getenv("LK_NO_PRECLEAR") != nullptr;  // tuning aid: clear the finalize scratch inline instead
  if (q.path == 2 && !q.exact_sums && !no_preclear && d.fin_cap > 0 && d.rf_sorted && d.rf_tables && d.fin) {
    // record path, scratch already sized by an earlier finalize: its clears (56 MB of key table for C2) do not wait for
    // the scan -- fork here, join at the start of the finalize
    if (!d.st2) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&d.st2, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_clear, cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventRecord(d.ev_fork, d.st));  // the previous finalize (reader of the scratch) is in front of this point
    CUDA_CHECK(cudaStreamWaitEvent(d.st2, d.ev_fork, 0));
    rec_clear_scratch(q, d.st2);
    CUDA_CHECK(cudaEventRecord(d.ev_clear, d.st2));
    d.pre_cleared = true;
  }
  static const uint32_t init_counters[8] = {0, 0xffffffffu, 0, 0, 0, 0, 0, 0};
  CUDA_CHECK(cudaMemcpyAsync(d.counters, init_counters, sizeof init_counters, cudaMemcpyHostToDevice, d.st));
  CUDA_CHECK(cudaMemsetAsync(d.survivors, 0, sizeof(unsigned long long), d.st));
  d.finalized_device = false;
  d.phase_global = false;
  d.fin_pending = false;  // a finalize nobody waited for is superseded by this execute
  d.executed = true;
  const bool sharded = q.comm != nullptr;
  if (sharded) {
    // every rank runs the same sequence of queries on a communicator: one epoch per execute, pools alternate
    Comm& c = *q.comm;
    LK_CHECK(c.connected || c.world == 1, LK_ERR_INVALID, "lk_query_execute: the attached lk_comm is not connected");
    LK_CHECK(q.path == 2, LK_ERR_UNSUPPORTED, "an attached lk_comm exchanges the record path only (dense planes: lk_query_partial_dense + reduce)");
    LK_CHECK((int)q.aggs.size() + (q.exact_sums ? 1 : 0) <= c.max_aggs, LK_ERR_INVALID,
             "the lk_comm was created for fewer aggregates than this query has (exact_sums needs one more word per record)");
    LK_CHECK(!d.rec_cell || d.rec_borrowed, LK_ERR_INVALID, "lk_query_set_comm after an unsharded execute");
    c.epoch++;
    const uint32_t pool = c.epoch & 1;
    comm_fill_params(c, P.x, pool, q.aggs.size() + (q.exact_sums ? 1 : 0));
    P.seq_offset = q.seq_offset;
    P.path = 3;
    d.rec_cell = c.keys(c.rank, pool);
    d.rec_vals = c.vals(c.rank, pool);
    d.rec_cap = c.cap();
    d.rec_borrowed = true;
    CUDA_CHECK(cudaMemsetAsync(c.send_count, 0, LK_MAX_RANKS * sizeof(uint32_t), d.st));
  }
  if (q.n_cells > 0 && P.ntiles > 0) {
    if (q.path == 0) {
      CUDA_CHECK(cudaMemsetAsync(d.planes, 0, (1 + q.aggs.size()) * q.n_cells * sizeof(unsigned long long), d.st));
      P.rowcnt = d.planes;
      for (size_t a = 0; a < q.aggs.size(); a++) P.acc[a] = d.planes + (1 + a) * q.n_cells;
    } else if (q.path == 1) {
      P.h_entries = d.harena->entries;
      P.h_occ = d.harena->occ;
      P.h_bkt = d.harena->occ_bkt;
      P.h_occ_cap = (uint32_t)std::min<uint64_t>(d.harena->slots, 0xffffffffu);
    } else if (!q.comm) {
      // one record per survivor at most: sized by the row count, so the scan can never overflow it
      const size_t cap = (size_t)std::max<int64_t>(q.total_rows, 1);
      if (d.rec_cap < cap) {
        if (d.rec_cell) CUDA_CHECK(cudaFreeAsync(d.rec_cell, d.st));
        if (d.rec_vals) CUDA_CHECK(cudaFreeAsync(d.rec_vals, d.st));
        d.rec_cell = d.rec_vals = nullptr;
        CUDA_CHECK(cudaMallocAsync(&d.rec_cell, cap * 8, d.st));
        CUDA_CHECK(cudaMallocAsync(&d.rec_vals, cap * 8 * (q.aggs.size() + (q.exact_sums ? 1 : 0)), d.st));
        d.rec_cap = cap;
      }
      P.rec_cell = d.rec_cell;
      P.rec_vals = d.rec_vals;
      P.rec_cap = (uint32_t)std::min<size_t>(d.rec_cap, 0xffffffffu);
    }
    CUDA_CHECK(cudaEventRecord(d.ev[2], d.st));
    launch_scan(q, P, q.exact_sums && q.path == 2);
    CUDA_CHECK(cudaEventRecord(d.ev[3], d.st));
    if (q.exact_sums && q.path != 2) device_exact_sums(q, P);
  } else {
    CUDA_CHECK(cudaEventRecord(d.ev[2], d.st));
    CUDA_CHECK(cudaEventRecord(d.ev[3], d.st));
  }
  if (sharded) {  // (also for a shard without tiles: its peers wait for its flag)
    comm_publish_kernel<<<1, 32, 0, d.st>>>(P.x, q.comm->epoch, d.counters);
    CUDA_CHECK(cudaGetLastError());
  }
}

void device_sync(Query& q) {
  LK_CHECK(q.dev && q.dev->st, LK_ERR_INVALID, "query has no device state");
  CUDA_CHECK(cudaStreamSynchronize(q.dev->st));
  device_resolve(q);  // a pending record-path finalize reports its row count / errors here
}

void* device_stream(Query& q) { return q.dev ? (void*)q.dev->st : nullptr; }

// timestamp phase range of this rank's last scan (metrics: (ts - startTs) mod step over its surviving rows; min = 0xffffffff
// when it kept none) and the global range the host reduced over the ranks
void device_phase(Query& q, uint32_t* pmin, uint32_t* pmax) {
  LK_CHECK(q.dev && q.dev->executed, LK_ERR_INVALID, "lk_query_phase before lk_query_execute");
  Query::Device& d = *q.dev;
  uint32_t h[3] = {0, 0xffffffffu, 0};
  CUDA_CHECK(cudaMemcpyAsync(h, d.counters, sizeof h, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  *pmin = q.is_metrics ? h[1] : 0xffffffffu;
  *pmax = q.is_metrics ? h[2] : 0u;
}
void device_set_phase(Query& q, uint32_t pmin, uint32_t pmax) {
  LK_CHECK(q.dev && q.dev->executed && !q.dev->finalized_device, LK_ERR_INVALID, "lk_query_set_phase belongs between lk_query_execute and the finalize");
  q.dev->phase_global = true;
  q.dev->phase_gmin = pmin;
  q.dev->phase_gmax = pmax;
}

void device_partial_dense(Query& q, int64_t* n_cells, int* n_planes, void** ptrs, int* ops) {
  LK_CHECK(q.dev && q.dev->executed, LK_ERR_INVALID, "lk_query_partial_dense before lk_query_execute");
  LK_CHECK(q.path == 0, LK_ERR_INVALID, "query uses the hash path; use lk_query_partial_sparse");
  *n_cells = (int64_t)q.n_cells;
  *n_planes = 1 + (int)q.aggs.size();
  ptrs[0] = q.dev->planes;
  ops[0] = AGG_COUNT;
  for (size_t a = 0; a < q.aggs.size(); a++) {
    ptrs[1 + a] = q.dev->planes ? q.dev->planes + (1 + a) * q.n_cells : nullptr;
    ops[1 + a] = q.aggs[a].op == AGG_MIN ? AGG_MAX : q.aggs[a].op;  // min is stored complemented: combine with max
  }
}

// ------------------------------------------------------------------------------------------------------------
// compaction of the aggregate table into result rows
// ------------------------------------------------------------------------------------------------------------
struct EmitParams {
  int n_aggs, n_keys;
  uint8_t ops[LK_MAX_AGGS];
  double divisor[LK_MAX_AGGS];
  uint64_t key_stride[LK_MAX_KEYS];
  uint32_t key_null[LK_MAX_KEYS];
  uint64_t magic_stride[LK_MAX_KEYS], magic_radix[LK_MAX_KEYS];  // floor(2^64 / d) + 1 (0: d == 1): n / d = umul64hi(n, magic) for n, d < 2^32
  int64_t base, step;
  uint32_t phase;
  const uint32_t* phase_ptr;  // record path: the scan's phase counter, read on the device (nothing waits for the host); else null
  uint64_t n_groups;
  int cells32, groups32;
  int64_t* ts;
  double* val[LK_MAX_AGGS];
  uint8_t* nul[LK_MAX_AGGS];
  int32_t* code[LK_MAX_KEYS];
};

// `acc(a)` returns the accumulator word of aggregate a; the loop over aggregates is unrolled so that a caller holding
// the words in registers (hash path) indexes them statically
template <class AccFn>
__device__ __forceinline__ void emit_row_bg(const EmitParams& E, uint64_t out, uint64_t bucket, uint64_t gid, AccFn acc) {
  uint32_t phase = E.phase;
  if (E.phase_ptr) { phase = *E.phase_ptr; if (phase == 0xffffffffu) phase = 0; }
  E.ts[out] = E.base + (int64_t)bucket * E.step + (int64_t)phase;
#pragma unroll
  for (int a = 0; a < LK_MAX_AGGS; a++) {
    if (a >= E.n_aggs) break;
    const unsigned long long w = acc(a);
    double v;
    uint8_t isnull = 0;
    switch (E.ops[a]) {
      case AGG_SUM: v = __longlong_as_double((long long)w); break;
      case AGG_COUNT: v = (double)w; break;
      case AGG_MIN: if (w == 0) { v = 0.0; isnull = 1; } else v = __longlong_as_double((long long)lk_min_decode(w)); break;
      default: if (w == 0) { v = 0.0; isnull = 1; } else v = __longlong_as_double((long long)lk_max_decode(w)); break;
    }
    if (E.divisor[a] != 1.0 && !isnull) v = v / E.divisor[a];
    E.val[a][out] = v;
    E.nul[a][out] = isnull;
  }
  for (int k = 0; k < E.n_keys; k++) {
    uint32_t g;
    if (E.groups32) {  // division by the invariant stride / radix as a 64 x 64 -> high 64 multiply (a 32-bit division costs ~25 instructions)
      const uint64_t q = E.magic_stride[k] ? __umul64hi(gid, E.magic_stride[k]) : gid;
      const uint64_t qq = E.magic_radix[k] ? __umul64hi(q, E.magic_radix[k]) : q;  // (radix 1: q / 1)
      g = (uint32_t)(q - qq * ((uint64_t)E.key_null[k] + 1));
    } else g = (uint32_t)((gid / E.key_stride[k]) % ((uint64_t)E.key_null[k] + 1));
    E.code[k][out] = g == E.key_null[k] ? -1 : (int32_t)g;
  }
}

template <class AccFn>
__device__ __forceinline__ void emit_row(const EmitParams& E, uint64_t out, uint64_t cell, AccFn acc) {
  uint64_t bucket, gid;
  if (E.cells32) {  // the whole (group x bucket) space fits 32 bits: 32-bit divisions (a 64-bit one costs ~100 instructions)
    const uint32_t b32 = (uint32_t)cell / (uint32_t)E.n_groups;
    gid = (uint32_t)cell - b32 * (uint32_t)E.n_groups;
    bucket = b32;
  } else {
    bucket = cell / E.n_groups;
    gid = cell - bucket * E.n_groups;
  }
  emit_row_bg(E, out, bucket, gid, acc);
}

constexpr int CMP_BLOCK = 256;
constexpr int CMP_PER_THREAD = 8;
constexpr int CMP_CHUNK = CMP_BLOCK * CMP_PER_THREAD;

__global__ void __launch_bounds__(CMP_BLOCK) dense_count_kernel(const unsigned long long* __restrict__ rowcnt, uint64_t n_cells, uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t warp_sums[CMP_BLOCK / 32];
  const uint64_t first = (uint64_t)blockIdx.x * CMP_CHUNK + (uint64_t)threadIdx.x * CMP_PER_THREAD;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < CMP_PER_THREAD; k++)
    if (first + k < n_cells) c += rowcnt[first + k] != 0;
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < CMP_BLOCK / 32; w++) t += warp_sums[w];
    block_counts[blockIdx.x] = t;
  }
}

// single-block exclusive scan of n u32 counters (in place); the grand total goes to *total
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t x = i < n ? v[i] : 0;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const uint32_t t = warp_tot[lane];
      uint32_t ti = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += o; }
      warp_tot[lane] = ti - t;
    }
    __syncthreads();
    const uint32_t excl = carry + warp_tot[wid] + incl - x;
    if (i < n) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(CMP_BLOCK) dense_emit_kernel(const unsigned long long* __restrict__ planes, uint64_t n_cells,
                                                               const uint32_t* __restrict__ block_offsets, const __grid_constant__ EmitParams E) {
  __shared__ uint32_t warp_sums[CMP_BLOCK / 32];
  const uint64_t first = (uint64_t)blockIdx.x * CMP_CHUNK + (uint64_t)threadIdx.x * CMP_PER_THREAD;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t present = 0, c = 0;
#pragma unroll
  for (int k = 0; k < CMP_PER_THREAD; k++)
    if (first + k < n_cells && planes[first + k] != 0) { present |= 1u << k; c++; }
  uint32_t incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < wid; w++) woff += warp_sums[w];
  uint64_t out = (uint64_t)block_offsets[blockIdx.x] + woff + incl - c;
  for (int k = 0; k < CMP_PER_THREAD; k++)
    if (present & (1u << k)) {
      const uint64_t cell = first + k;
      const unsigned long long* acc0 = planes + n_cells + cell;
      emit_row(E, out++, cell, [&](int a) { return acc0[(size_t)a * n_cells]; });
    }
}

// ---- hash path: counting sort of the claimed slots by bucket (ORDER BY timestamp), then a coalesced emit ----
constexpr int HS_BLOCK = 256;
constexpr int HS_CHUNK = 8192;   // claimed slots per block
constexpr int HS_MAXB = 4096;    // buckets histogrammed in shared memory; beyond that global atomics are uncontended enough

// Warp-aggregated counter increment: the lanes that target the same counter elect a leader that adds their count once
// (with 2-8 partitions a plain shared-memory atomic per element serialises 32-way and took 3.4 ms for 6 M entries).
// Must be called by all 32 lanes; lanes with valid == false get no slot.  Returns the lane's position.
__device__ __forceinline__ uint32_t warp_agg_inc(uint32_t* ctrs, uint32_t idx, bool valid) {
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? idx : 0xffffffffu);
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (valid && lane == leader) base = atomicAdd(&ctrs[idx], (uint32_t)__popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(peers & ((1u << lane) - 1));
}

// pass 1: per-bucket counts (block-private histogram, one global atomic per bucket per block).  The bucket of every claimed
// slot was recorded when the slot was claimed (scan kernel / sparse merge): no random read of the table here.
__global__ void __launch_bounds__(HS_BLOCK) hash_hist_kernel(const uint32_t* __restrict__ bucket_of, uint32_t n, uint32_t nbuckets,
                                                             uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[HS_MAXB];
  const bool priv = nbuckets <= HS_MAXB;
  if (priv) {
    for (uint32_t b = threadIdx.x; b < nbuckets; b += HS_BLOCK) h[b] = 0;
    __syncthreads();
  }
  const uint32_t lo = blockIdx.x * HS_CHUNK, hi = min(n, lo + HS_CHUNK);
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {  // uniform trip count: warp_agg_inc needs the whole warp
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    warp_agg_inc(priv ? h : hist, valid ? bucket_of[i] : 0u, valid);
  }
  if (priv) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbuckets; b += HS_BLOCK)
      if (h[b]) atomicAdd(&hist[b], h[b]);
  }
}

// pass 2: each block reserves one range per bucket (cursor = exclusive scan of hist) and scatters its slots into it
__global__ void __launch_bounds__(HS_BLOCK) hash_scatter_kernel(const uint32_t* __restrict__ occ, const uint32_t* __restrict__ bucket_of, uint32_t n,
                                                                uint32_t nbuckets, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  __shared__ uint32_t cnt[HS_MAXB];
  __shared__ uint32_t base[HS_MAXB];
  const bool priv = nbuckets <= HS_MAXB;
  const uint32_t lo = blockIdx.x * HS_CHUNK, hi = min(n, lo + HS_CHUNK);
  if (!priv) {
    for (uint32_t i = lo + threadIdx.x; i < hi; i += HS_BLOCK) sorted[atomicAdd(&cursor[bucket_of[i]], 1u)] = occ[i];
    return;
  }
  for (uint32_t b = threadIdx.x; b < nbuckets; b += HS_BLOCK) cnt[b] = 0;
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    warp_agg_inc(cnt, valid ? bucket_of[i] : 0u, valid);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nbuckets; b += HS_BLOCK) {
    const uint32_t c = cnt[b];
    if (c) base[b] = atomicAdd(&cursor[b], c);
    cnt[b] = 0;
  }
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    const uint32_t b = valid ? bucket_of[i] : 0u;
    const uint32_t pos = warp_agg_inc(cnt, b, valid);
    if (valid) sorted[base[b] + pos] = occ[i];
  }
}

// pass 3: row i <- slot sorted[i] (rows of one bucket are contiguous; their order inside a bucket is arbitrary, like the
// reference's: BaseExpr.scala:394, 403 order by the time column only).  Each emitted entry is zeroed: the arena stays clean.
__global__ void __launch_bounds__(HS_BLOCK) hash_emit_kernel(const uint32_t* __restrict__ sorted, uint32_t n, uint8_t* __restrict__ entries,
                                                             uint32_t stride, const __grid_constant__ EmitParams E) {
  const uint32_t i = blockIdx.x * HS_BLOCK + threadIdx.x;
  if (i >= n) return;
  ulonglong2* z = reinterpret_cast<ulonglong2*>(entries + (uint64_t)sorted[i] * stride);
  unsigned long long w[8];
  {
    const ulonglong2 a = z[0], b = z[1];  // one 32-byte sector
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
    // only the 16-byte pieces that hold the key and the n_aggs accumulators are read and cleared: the rest of a
    // 64-byte entry is never written
    w[4] = w[5] = w[6] = w[7] = 0;
    if (E.n_aggs > 3) { const ulonglong2 c = z[2]; w[4] = c.x; w[5] = c.y; }
    if (E.n_aggs > 5) { const ulonglong2 d = z[3]; w[6] = d.x; w[7] = d.y; }
  }
  const uint32_t used16 = ((uint32_t)E.n_aggs + 2) / 2;  // ceil((1 + n_aggs) * 8 / 16)
  for (uint32_t k = 0; k < max(used16, 2u); k++) z[k] = make_ulonglong2(0ull, 0ull);
  emit_row(E, i, w[0] - 1, [&](int a) { return w[1 + a]; });
}

// ---- record path: the scan appended one (key, accumulator words) record per survivor, key = (bucket, group id, record
// index).  Finalize groups equal (bucket, group) keys through a global table of 8-byte entries and writes the rows:
//   rec_bhist    records per bucket (warp-aggregated counters: neighbouring records share their bucket)
//   rec_regions  prefix over the buckets: bucket b owns table slots [2 start_b, 2 start_b + 2 n_b) -- load factor 0.5 in
//                every bucket whatever the skew, and a region never overflows into its neighbour
//   rec_group    every record inserts its key into its bucket's region (CAS, linear probing inside the region).  The first
//                record of a cell becomes its OWNER; a later one folds its accumulator words into the owner's row of
//                rec_vals[] (add / add / max on the table encodings) and marks itself consumed.  Records arrive clustered
//                by time, so the regions being probed at any moment are a few hundred KB that stay in L2.
//   rec_rowscan  prefix over the owners per bucket = first row of every bucket
//   rec_emit     owners copy their row out: bucket-major = timestamp order (ORDER BY timestamp: BaseExpr.scala:394, 403;
//                the order inside a bucket is unspecified there too), positions from one warp-aggregated cursor per bucket
// No library call and no host round trip: every size the kernels need is computed on the device.
constexpr int RF_BLOCK = 256;
#ifndef REC_EMIT_CTAS
#define REC_EMIT_CTAS 5
#endif
struct RecFin {  // device-resident bookkeeping of one finalize
  uint32_t nrec, nrows, status, pad;
};
enum : uint32_t { RF_ST_CAP = 1, RF_ST_TABLE = 2 };
constexpr unsigned long long RF_CONSUMED = ~0ull;  // key of a record folded into its owner (real keys use at most 63 bits)
struct RecGeom {
  uint32_t idx_bits, gid_bits, nbuckets;  // nbuckets: partitions of the list = time buckets x slices
  uint32_t slices;   // partitions per time bucket (by a hash of the group id): > 1 only on the partitioned (scatter) path, so that a
                     // partition's key table fits shared memory however many records a time bucket holds (C4: 139 k)
  uint32_t rec_cap;  // capacity of the record arrays
  uint32_t fin_cap;  // records the key table (2 slots each) and the result columns hold
  uint32_t world, region_cap;  // sharded: the list is the concatenation of one region per source rank (world <= 1: one plain list)
  const uint32_t* prefix;      // ... prefix[s] = records of sources 0..s-1 (CommCtrl::prefix)
  uint32_t fp_shift; // key table entry = fingerprint << fp_shift | (record index + 1); 32 = no room for a fingerprint
  uint32_t cstride;  // words between the per-bucket counters that are bumped atomically: with few buckets every counter
                     // gets its own 128-byte line (atomics on one line serialise at ~7 ns each on B200: 360 counters packed
                     // into 12 lines kept ONE L2 slice 97 % busy and cost 120 us per pass over 6.2 M records)
};

// partition of a record: its time bucket, or (bucket, hash slice of its group id)
__device__ __forceinline__ uint32_t rec_part(unsigned long long key, const RecGeom& G) {
  const unsigned long long cellx = key >> G.idx_bits;
  const uint32_t bucket = (uint32_t)(cellx >> G.gid_bits);
  if (G.slices <= 1) return bucket;
  const uint32_t h = lk_rf_mix(cellx & ((1ull << G.gid_bits) - 1)) * 0x85EBCA6Bu + 0x6A09E667u;  // (other bits than the table slot uses)
  return bucket * G.slices + __umulhi(h ^ (h >> 15), G.slices);
}

// position of record i of the list in the record arrays
__device__ __forceinline__ uint32_t rec_phys(uint32_t i, const RecGeom& G) {
  if (G.world <= 1) return i;
  // source region of record i: count the region starts at or below i -- independent loads (one cache line, all hits) instead
  // of a walk whose every step waits for the previous one (three kernels do this once per record)
  uint32_t s = 0;
#pragma unroll
  for (int k = 1; k < LK_MAX_RANKS; k++)
    if (k < (int)G.world) s += i >= __ldg(G.prefix + k) ? 1u : 0u;
  return s * G.region_cap + (i - __ldg(G.prefix + s));
}

// Per-bucket counters are bumped once per warp and distinct bucket: the lanes of a warp that hold the same bucket elect a
// leader (__match_any_sync) who adds their count.  (A 32-record stretch of the list holds ~2.5 buckets: the scan appends up to
// 32 survivors of one tile at a time, and neighbouring appends come from tiles anywhere in the time range.)
// Returns the lane's rank among its peers; *base receives what the leader's atomicAdd returned (when want_base).
template <bool WANT_BASE>
__device__ __forceinline__ uint32_t warp_bucket_bump(uint32_t* ctr, uint32_t cs, uint32_t bucket, bool valid, uint32_t* base) {
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? bucket : 0xffffffffu);
  const int leader = __ffs(peers) - 1;
  uint32_t b = 0;
  if (valid && lane == leader) {
    if (WANT_BASE) b = atomicAdd(&ctr[(size_t)bucket * cs], (uint32_t)__popc(peers));
    else atomicAdd(&ctr[(size_t)bucket * cs], (uint32_t)__popc(peers));
  }
  if (WANT_BASE) *base = __shfl_sync(0xffffffffu, b, leader);
  return __popc(peers & ((1u << lane) - 1));
}

__global__ void __launch_bounds__(RF_BLOCK) rec_bhist_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counters,
                                                             const __grid_constant__ RecGeom G, uint32_t* __restrict__ bkt_recs) {
  const uint32_t nrec = min(counters[5], G.rec_cap);
  if (nrec > G.fin_cap) return;  // rec_regions_kernel raises RF_ST_CAP; the host grows the scratch and finalizes again
  const uint32_t n32 = (nrec + 31u) & ~31u;  // whole warps
  for (uint32_t i = blockIdx.x * RF_BLOCK + threadIdx.x; i < n32; i += gridDim.x * RF_BLOCK) {
    const unsigned long long key = i < nrec ? keys[rec_phys(i, G)] : RF_CONSUMED;
    const bool valid = key != RF_CONSUMED;
    warp_bucket_bump<false>(bkt_recs, G.cstride, valid ? rec_part(key, G) : 0u, valid, nullptr);
  }
}

// one block: out[b] = sum of in[0..b) (`in` strided), out[n] = total; optionally zeroes the strided `zero[0..n)` on the way
constexpr int BS_BLOCK = 1024;
constexpr int BS_PER = 8;
__device__ __forceinline__ uint32_t block_exclusive_scan_u32(const uint32_t* __restrict__ in, uint32_t in_stride, uint32_t n, uint32_t* __restrict__ out,
                                                             uint32_t* __restrict__ zero, uint32_t zero_stride) {
  __shared__ uint32_t warp_tot[BS_BLOCK / 32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += BS_BLOCK * BS_PER) {
    const uint32_t first = base + threadIdx.x * BS_PER;
    uint32_t c[BS_PER], local = 0;
#pragma unroll
    for (int k = 0; k < BS_PER; k++) { c[k] = first + k < n ? in[(size_t)(first + k) * in_stride] : 0u; local += c[k]; }
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const uint32_t t = warp_tot[lane];
      uint32_t ti = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += o; }
      warp_tot[lane] = ti - t;
    }
    __syncthreads();
    uint32_t run = carry_s + warp_tot[wid] + incl - local;
#pragma unroll
    for (int k = 0; k < BS_PER; k++) {
      if (first + k < n) { out[first + k] = run; if (zero) zero[(size_t)(first + k) * zero_stride] = 0; }
      run += c[k];
    }
    __syncthreads();
    if (threadIdx.x == BS_BLOCK - 1) carry_s = run;
    __syncthreads();
  }
  const uint32_t total = carry_s;
  if (threadIdx.x == 0) out[n] = total;
  return total;
}

__global__ void __launch_bounds__(BS_BLOCK) rec_regions_kernel(const uint32_t* __restrict__ counters, const __grid_constant__ RecGeom G,
                                                               const uint32_t* __restrict__ bkt_recs, uint32_t* __restrict__ rec_start,
                                                               uint32_t* __restrict__ bkt_rows, RecFin* __restrict__ fin) {
  const uint32_t nrec = min(counters[5], G.rec_cap);
  if (nrec > G.fin_cap) {
    if (threadIdx.x == 0) { fin->nrec = nrec; fin->status = RF_ST_CAP; fin->nrows = 0; }
    return;
  }
  block_exclusive_scan_u32(bkt_recs, G.cstride, G.nbuckets, rec_start, bkt_rows, G.cstride);
  if (threadIdx.x == 0) { fin->nrec = nrec; fin->nrows = 0; }
}

// The key table holds 32-bit entries: (fingerprint of the group id) << fp_shift | (record index + 1) of the cell's owner.  At
// two slots per record it is 8 bytes per record, so table and key list together stay inside the 126 MB L2 for lists of ~7 M
// records.  A probe that meets an occupied slot with another fingerprint moves on without touching memory; an equal
// fingerprint is confirmed against the owner's key in keys[] (an owner's key is never rewritten).
// One record per thread and iteration, small code, full occupancy: the batched variants of this pass (8 records per lane in
// flight, 120 registers, 5-14 k SASS instructions) measured 330-400 us per 6.2 M records on B200, stalled on instruction fetch
// and on the longest probe chain of every 8 x 32 batch.
// SORTED: the list is partitioned by bucket (rec_scatter_kernel ran).  Every warp then walks one contiguous stretch of the list
// -- one or two buckets -- and counts its owners in registers, adding them once at the end: with the records of a bucket side
// by side, one counter bump per warp and 32 records would put hundreds of atomics on ONE address back to back (measured:
// 208 us for this kernel against 153 us on the unsorted list; atomics on one line serialise at ~7 ns each).
template <bool SORTED>
__global__ void __launch_bounds__(RF_BLOCK, 8) rec_group_kernel(unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals,
                                                             const uint32_t* __restrict__ rec_start, uint32_t* __restrict__ table,
                                                             uint32_t* __restrict__ bkt_rows, RecFin* __restrict__ fin, const __grid_constant__ RecGeom G,
                                                             const __grid_constant__ EmitParams E) {
  if (fin->status & RF_ST_CAP) return;
  const uint32_t nrec = fin->nrec;
  const uint32_t n32 = (nrec + 31u) & ~31u;
  const unsigned long long gid_mask = (1ull << G.gid_bits) - 1;
  const uint32_t fp_shift = G.fp_shift, idx_field = fp_shift < 32 ? (1u << fp_shift) - 1 : 0xffffffffu;  // fp_shift == 32: no fingerprint bits
  uint32_t my_status = 0;
  // Two records ahead of the one being inserted the key is requested, one record ahead its table slot is pulled into L2
  // (prefetch.global.L2): the CAS of a record then meets a resident sector instead of paying the DRAM round trip of a random
  // 32-byte access itself (ncu: 76 % of this kernel's stall samples sat behind that one CAS).
  uint32_t stride = gridDim.x * RF_BLOCK, i0 = blockIdx.x * RF_BLOCK + threadIdx.x, i_end = n32;
  if (SORTED) {
    const uint32_t warps = gridDim.x * (RF_BLOCK / 32), gw = (blockIdx.x * RF_BLOCK + threadIdx.x) >> 5;
    const uint32_t per = ((n32 >> 5) + warps - 1) / warps;  // 32-record groups per warp
    stride = 32;
    i0 = min(gw * per * 32u, n32) + (threadIdx.x & 31);
    i_end = min((gw + 1) * per * 32u, n32);
  }
  uint32_t cur_b = 0xffffffffu, cur_n = 0;  // SORTED: owners of bucket cur_b seen by this lane, not yet added to its counter
  auto load_key = [&](uint32_t i, uint32_t& pi) -> unsigned long long {
    pi = i < nrec ? rec_phys(i, G) : 0u;
    return i < nrec ? keys[pi] : RF_CONSUMED;
  };
  auto slot_of = [&](unsigned long long key, uint32_t& bucket, uint32_t& width, uint32_t*& region, uint32_t& h) -> uint32_t {
    const unsigned long long cellx = key >> G.idx_bits;
    bucket = (uint32_t)(cellx >> G.gid_bits);  // (this kernel only sees lists partitioned by time bucket alone: slices == 1)
    const uint32_t s0 = __ldg(rec_start + bucket), s1 = __ldg(rec_start + bucket + 1);
    width = 2u * (s1 - s0);  // >= 2: this record is one of the bucket's
    region = table + 2ull * s0;
    h = lk_rf_mix(cellx & gid_mask);
    return __umulhi(h, width);
  };
  uint32_t pi_a = 0, pi_b = 0, pi_c = 0;
  unsigned long long key_a = i0 < i_end ? load_key(i0, pi_a) : RF_CONSUMED;       // record being inserted
  unsigned long long key_b = i0 + stride < i_end ? load_key(i0 + stride, pi_b) : RF_CONSUMED;  // next
  for (uint32_t i = i0; i < i_end; i += stride) {
    unsigned long long key_c = RF_CONSUMED;                                       // the one after
    if (i + 2 * stride < i_end) key_c = load_key(i + 2 * stride, pi_c);
    if (key_b != RF_CONSUMED) {
      uint32_t b2, w2, h2;
      uint32_t* r2;
      const uint32_t sl = slot_of(key_b, b2, w2, r2, h2);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + sl));
    }
    const uint32_t pi = pi_a;
    const unsigned long long key = key_a;
    const bool valid = key != RF_CONSUMED;
    bool owner = false;
    uint32_t bucket = 0;
    if (valid) {
      const unsigned long long cellx = key >> G.idx_bits;
      uint32_t width, h;
      uint32_t* region;
      uint32_t slot = slot_of(key, bucket, width, region, h), probe = 0;
      const uint32_t entry = (fp_shift < 32 ? (h * 0x2545F491u) >> fp_shift << fp_shift : 0u) | (pi + 1);
      for (; probe < width; probe++) {
        const uint32_t prev = atomicCAS(region + slot, 0u, entry);
        if (prev == 0u) { owner = true; break; }
        if (((prev ^ entry) & ~idx_field) == 0u) {  // same fingerprint: look at the owner's key
          const uint32_t oi = (prev & idx_field) - 1;
          if ((keys[oi] >> G.idx_bits) == cellx) {
            // a later record of the cell: fold into the owner's row, leave the list
            const unsigned long long* mine = vals + (size_t)pi * E.n_aggs;
            unsigned long long* own = vals + (size_t)oi * E.n_aggs;
            for (int a = 0; a < E.n_aggs; a++) {
              const unsigned long long x = mine[a];
              if (E.ops[a] == AGG_SUM) atomicAdd(reinterpret_cast<double*>(own + a), __longlong_as_double((long long)x));
              else if (E.ops[a] == AGG_COUNT) atomicAdd(own + a, x);
              else if (x) atomicMax(own + a, x);  // min (complemented key) and max (key)
            }
            keys[pi] = RF_CONSUMED;
            break;
          }
        }
        slot = slot + 1 == width ? 0u : slot + 1;  // another cell's slot: linear probing inside the region
      }
      if (probe == width) my_status |= RF_ST_TABLE;  // cannot happen: the region has two slots per record of its bucket
    }
    if (SORTED) {
      if (owner) {
        if (bucket != cur_b) {  // (rare: a bucket boundary inside this warp's stretch)
          if (cur_n) atomicAdd(&bkt_rows[(size_t)cur_b * G.cstride], cur_n);
          cur_b = bucket;
          cur_n = 0;
        }
        cur_n++;
      }
    } else warp_bucket_bump<false>(bkt_rows, G.cstride, bucket, owner, nullptr);
    key_a = key_b; pi_a = pi_b;
    key_b = key_c; pi_b = pi_c;
  }
  if (SORTED) {  // one add per warp and bucket for the whole stretch
    const unsigned peers = __match_any_sync(0xffffffffu, cur_n ? cur_b : 0xffffffffu);
    const uint32_t sum = __reduce_add_sync(peers, cur_n);
    if (cur_n && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&bkt_rows[(size_t)cur_b * G.cstride], sum);
  }
  if (my_status) atomicOr(&fin->status, my_status);
}

// Grouping of a bucket-partitioned list: ONE CTA PER BUCKET with the bucket's key table in SHARED memory (two 32-bit slots per
// record: 17 k records of a C2 bucket = 138 KB), so that no insert leaves the SM -- on the partitioned list the global table is
// the wrong tool: the ~300 k inserts in flight all land in the regions of a few neighbouring buckets, a dozen atomics per
// 128-byte line, and atomics on one line serialise (measured: 208 us against 153 us on the unpartitioned list).  Entry =
// 16-bit fingerprint << 16 | (index in the bucket + 1).  The owners are counted per CTA and stored, not added.  A bucket too
// large for shared memory (more than 3/4 of RG_SLOTS records) is grouped in its region of the global table by the same CTA.
constexpr uint32_t RG_SLOTS = 53248;  // 208 KB
constexpr int RG_BLOCK = 1024;
__global__ void __launch_bounds__(RG_BLOCK) rec_group_bucket_kernel(unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals,
                                                                    const uint32_t* __restrict__ rec_start, uint32_t* __restrict__ table,
                                                                    uint32_t* __restrict__ bkt_rows, RecFin* __restrict__ fin,
                                                                    const __grid_constant__ RecGeom G, const __grid_constant__ EmitParams E,
                                                                    uint32_t smem_slots) {
  extern __shared__ uint32_t stab[];
  __shared__ uint32_t s_owners[RG_BLOCK / 32];
  if (fin->status & RF_ST_CAP) return;
  const uint32_t b = blockIdx.x;
  if (b >= G.nbuckets) return;
  const uint32_t s0 = rec_start[b], s1 = rec_start[b + 1], n = s1 - s0;
  if (n == 0) return;  // (its counter was zeroed by rec_regions)
  const bool in_smem = n <= smem_slots / 4 * 3 && n < 0xffffu;
  const uint32_t width = in_smem ? min(2u * n, smem_slots) : 2u * n;
  uint32_t* const tab = in_smem ? stab : table + 2ull * s0;  // (the global region is already zero)
  if (in_smem) {
    for (uint32_t k = threadIdx.x; k < width; k += RG_BLOCK) stab[k] = 0;
    __syncthreads();
  }
  const unsigned long long gid_mask = (1ull << G.gid_bits) - 1;
  const uint32_t idx_field = in_smem ? 0xffffu : (G.fp_shift < 32 ? (1u << G.fp_shift) - 1 : 0xffffffffu);
  const uint32_t fp_shift = in_smem ? 16u : G.fp_shift;
  uint32_t owners = 0, my_status = 0;
  unsigned long long key_n = threadIdx.x < n ? keys[s0 + threadIdx.x] : 0ull;
  for (uint32_t k = threadIdx.x; k < n; k += RG_BLOCK) {
    const unsigned long long key = key_n;
    if (k + RG_BLOCK < n) key_n = keys[s0 + k + RG_BLOCK];  // the next record's key is on its way while this one is inserted
    const uint32_t pi = s0 + k;
    const unsigned long long cellx = key >> G.idx_bits;
    const uint32_t h = lk_rf_mix(cellx & gid_mask);
    const uint32_t entry = (fp_shift < 32 ? (h * 0x2545F491u) >> fp_shift << fp_shift : 0u) | ((in_smem ? k : pi) + 1);
    uint32_t slot = __umulhi(h, width), probe = 0;
    for (; probe < width; probe++) {
      const uint32_t prev = atomicCAS(tab + slot, 0u, entry);
      if (prev == 0u) { owners++; break; }
      if (((prev ^ entry) & ~idx_field) == 0u) {  // same fingerprint: look at the owner's key
        const uint32_t oi = (in_smem ? s0 : 0u) + (prev & idx_field) - 1;
        if ((keys[oi] >> G.idx_bits) == cellx) {
          const unsigned long long* mine = vals + (size_t)pi * E.n_aggs;
          unsigned long long* own = vals + (size_t)oi * E.n_aggs;
          for (int a = 0; a < E.n_aggs; a++) {
            const unsigned long long x = mine[a];
            if (E.ops[a] == AGG_SUM) atomicAdd(reinterpret_cast<double*>(own + a), __longlong_as_double((long long)x));
            else if (E.ops[a] == AGG_COUNT) atomicAdd(own + a, x);
            else if (x) atomicMax(own + a, x);
          }
          keys[pi] = RF_CONSUMED;
          break;
        }
      }
      slot = slot + 1 == width ? 0u : slot + 1;
    }
    if (probe == width) my_status |= RF_ST_TABLE;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) owners += __shfl_xor_sync(0xffffffffu, owners, d);
  if ((threadIdx.x & 31) == 0) s_owners[threadIdx.x >> 5] = owners;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < RG_BLOCK / 32; w++) t += s_owners[w];
    bkt_rows[(size_t)b * G.cstride] = t;
  }
  if (my_status) atomicOr(&fin->status, my_status);
}

// Partition of the record list by bucket (counting sort: rec_bhist counted, rec_regions laid the buckets out) into a second
// copy of the record arrays.  The list arrives in the order the scans appended -- 32 survivors of one tile at a time on one
// GPU, a few records per tile and destination when sharded -- from tiles anywhere in the time range, so a warp's 32 records
// belong to several buckets (~2.5 on one GPU, ~8 at 8 ranks): after this pass they belong to ONE, which is what the grouping
// and the emit want (one counter bump per warp, table probes inside one ~140 KB region, rows stored in whole lines).  One
// 8-byte key + one 32-byte accumulator row moved per record instead of 13 scattered column stores later.
__global__ void __launch_bounds__(RF_BLOCK) rec_scatter_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ vals,
                                                               const uint32_t* __restrict__ rec_start, uint32_t* __restrict__ cursor,
                                                               const RecFin* __restrict__ fin, const __grid_constant__ RecGeom G, int n_aggs,
                                                               unsigned long long* __restrict__ keys2, unsigned long long* __restrict__ vals2) {
  if (fin->status & RF_ST_CAP) return;
  const uint32_t nrec = fin->nrec;
  const uint32_t n32 = (nrec + 31u) & ~31u;
  const uint32_t stride = gridDim.x * RF_BLOCK;
  const bool four = n_aggs == 4;
  auto load_rec = [&](uint32_t i, uint32_t& pi, ulonglong2& lo, ulonglong2& hi) -> unsigned long long {
    pi = i < nrec ? rec_phys(i, G) : 0u;
    if (i >= nrec) return RF_CONSUMED;
    if (four) {
      const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(vals + (size_t)pi * 4);
      lo = rec[0];
      hi = rec[1];
    }
    return keys[pi];
  };
  uint32_t pi_n = 0;
  ulonglong2 lo_n = make_ulonglong2(0, 0), hi_n = make_ulonglong2(0, 0);
  unsigned long long key_n = load_rec(blockIdx.x * RF_BLOCK + threadIdx.x, pi_n, lo_n, hi_n);
  for (uint32_t i = blockIdx.x * RF_BLOCK + threadIdx.x; i < n32; i += stride) {
    const unsigned long long key = key_n;
    const uint32_t pi = pi_n;
    const ulonglong2 lo = lo_n, hi = hi_n;
    if (i + stride < n32) key_n = load_rec(i + stride, pi_n, lo_n, hi_n);
    const bool valid = i < nrec;
    const uint32_t bucket = valid ? rec_part(key, G) : 0u;
    uint32_t base = 0;
    const uint32_t rank = warp_bucket_bump<true>(cursor, G.cstride, bucket, valid, &base);
    if (valid) {
      const uint32_t out = __ldg(rec_start + bucket) + base + rank;  // < nrec <= fin_cap
      keys2[out] = key;
      if (four) {
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(vals2 + (size_t)out * 4);
        dst[0] = lo;
        dst[1] = hi;
      } else {
        for (int a = 0; a < n_aggs; a++) vals2[(size_t)out * n_aggs + a] = vals[(size_t)pi * n_aggs + a];
      }
    }
  }
}

__global__ void __launch_bounds__(BS_BLOCK) rec_rowscan_kernel(const uint32_t* __restrict__ bkt_rows, uint32_t nbuckets, uint32_t cstride,
                                                               uint32_t* __restrict__ row_start, uint32_t* __restrict__ row_cursor, RecFin* __restrict__ fin) {
  if (fin->status & RF_ST_CAP) return;
  const uint32_t total = block_exclusive_scan_u32(bkt_rows, cstride, nbuckets, row_start, row_cursor, cstride);
  if (threadIdx.x == 0) fin->nrows = total;
}

// rows: every owner copies its record out at (first row of its bucket) + (cursor of the bucket, bumped once per warp and bucket)
__global__ void __launch_bounds__(RF_BLOCK, REC_EMIT_CTAS) rec_emit_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ vals,
                                                            const uint32_t* __restrict__ row_start, uint32_t* __restrict__ row_cursor,
                                                            const RecFin* __restrict__ fin, const __grid_constant__ RecGeom G,
                                                            const __grid_constant__ EmitParams E) {
  if (fin->status & RF_ST_CAP) return;
  const uint32_t nrec = fin->nrec;
  const uint32_t n32 = (nrec + 31u) & ~31u;
  const unsigned long long gid_mask = (1ull << G.gid_bits) - 1;
  // the next record's key and accumulator row (address known without the key) are requested before this record's cursor
  // round trip and row stores: two records in flight per thread
  const uint32_t stride = gridDim.x * RF_BLOCK;
  const bool four = E.n_aggs == 4;  // one 32-byte sector per record
  auto load_rec = [&](uint32_t i, uint32_t& pi, ulonglong2& lo, ulonglong2& hi) -> unsigned long long {
    pi = i < nrec ? rec_phys(i, G) : 0u;
    if (i >= nrec) return RF_CONSUMED;
    if (four) {
      const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(vals + (size_t)pi * 4);
      lo = rec[0];
      hi = rec[1];
    }
    return keys[pi];
  };
  uint32_t pi_n = 0;
  ulonglong2 lo_n = make_ulonglong2(0, 0), hi_n = make_ulonglong2(0, 0);
  unsigned long long key_n = load_rec(blockIdx.x * RF_BLOCK + threadIdx.x, pi_n, lo_n, hi_n);
  for (uint32_t i = blockIdx.x * RF_BLOCK + threadIdx.x; i < n32; i += stride) {
    const unsigned long long key = key_n;
    const uint32_t pi = pi_n;
    const ulonglong2 lo = lo_n, hi = hi_n;
    if (i + stride < n32) key_n = load_rec(i + stride, pi_n, lo_n, hi_n);
    const bool owner = key != RF_CONSUMED;
    const unsigned long long cellx = key >> G.idx_bits;
    const uint32_t bucket = owner ? rec_part(key, G) : 0u;  // the partition: cursor and first row are per partition
    unsigned long long w[LK_MAX_AGGS];
    if (owner) {
      if (four) {
        w[0] = lo.x; w[1] = lo.y; w[2] = hi.x; w[3] = hi.y;
        w[4] = w[5] = w[6] = 0;
      } else {
        const unsigned long long* rec = vals + (size_t)pi * E.n_aggs;
#pragma unroll
        for (int a = 0; a < LK_MAX_AGGS; a++) w[a] = a < E.n_aggs ? rec[a] : 0ull;
      }
    }
    uint32_t base = 0;
    const uint32_t rank = warp_bucket_bump<true>(row_cursor, G.cstride, bucket, owner, &base);
    if (owner) {
      const uint32_t out = __ldg(row_start + bucket) + base + rank;
      if (out < G.fin_cap) emit_row_bg(E, out, (uint32_t)(cellx >> G.gid_bits), cellx & gid_mask, [&](int a) { return w[a]; });
    }
  }
}

// device result layout for `stride` rows per column: ts[stride] | val[a][stride] | code[k][stride] | nul[a][stride]
static size_t result_bytes(const Query& q, int64_t n) {
  return (size_t)n * (8 + 8 * q.aggs.size() + 4 * q.key_pcols.size() + q.aggs.size()) + 64;
}

static void fill_emit_params(const Query& q, EmitParams& E, uint8_t* basep, int64_t n) {
  memset(&E, 0, sizeof E);
  E.n_aggs = (int)q.aggs.size();
  E.n_keys = (int)q.key_pcols.size();
  for (int a = 0; a < E.n_aggs; a++) { E.ops[a] = q.aggs[a].op; E.divisor[a] = q.aggs[a].divisor; }
  auto magic = [](uint64_t dv) -> uint64_t { return dv <= 1 ? 0ull : (uint64_t)((((unsigned __int128)1) << 64) / dv) + 1; };
  for (int k = 0; k < E.n_keys; k++) {
    E.key_stride[k] = q.params.keys[k].stride;
    E.key_null[k] = q.params.keys[k].null_code;
    E.magic_stride[k] = magic(E.key_stride[k]);
    E.magic_radix[k] = magic((uint64_t)E.key_null[k] + 1);  // radix >= 2 unless the dictionary is empty (radix 1: q % 1 = 0 = the NULL code)
  }
  E.base = q.base;
  E.step = q.step;
  E.phase = q.dev->phase;
  E.phase_ptr = nullptr;
  E.n_groups = q.n_groups;
  E.cells32 = q.n_cells < (1ull << 32);
  E.groups32 = q.n_groups < (1ull << 32);
  uint8_t* p = basep;
  E.ts = (int64_t*)p; p += 8 * n;
  for (int a = 0; a < E.n_aggs; a++) { E.val[a] = (double*)p; p += 8 * n; }
  for (int k = 0; k < E.n_keys; k++) { E.code[k] = (int32_t*)p; p += 4 * n; }
  for (int a = 0; a < E.n_aggs; a++) { E.nul[a] = p; p += n; }
}

static void ensure_dres(Query::Device& d, size_t bytes) {
  if (d.dres_cap >= bytes) return;
  if (d.dres) CUDA_CHECK(cudaFreeAsync(d.dres, d.st));
  d.dres = nullptr;
  CUDA_CHECK(cudaMallocAsync(&d.dres, bytes, d.st));
  d.dres_cap = bytes;
}

static void ensure_block_counts(Query::Device& d, size_t n) {
  if (d.block_counts_cap >= n) return;
  if (d.block_counts) CUDA_CHECK(cudaFreeAsync(d.block_counts, d.st));
  d.block_counts = nullptr;
  CUDA_CHECK(cudaMallocAsync(&d.block_counts, n * sizeof(uint32_t), d.st));
  d.block_counts_cap = n;
}

static void clear_arena_if_any(Query::Device& d) {
  if (d.harena) CUDA_CHECK(cudaMemsetAsync(d.harena->entries, 0, d.harena->slots * d.harena->stride, d.st));
}

// status flags / timestamp phase of the scan, common to all paths (host copies of the counters in h)
static void check_scan_status(Query& q, const uint32_t* h) {
  Query::Device& d = *q.dev;
  LK_CHECK(!(h[0] & ST_XCHG_TIMEOUT), LK_ERR_CUDA, "sharded exchange: a peer rank did not deliver its records within 20 s");
  if (h[0] & ST_HASH_FULL) {
    clear_arena_if_any(d);
    d.finalized_device = true;
    fail(LK_ERR_NOMEM, q.path == 2 ? std::string("record buffer overflowed: this shard received more records than its buffer holds")
                                   : strf("aggregate hash table of %llu slots overflowed; raise max_hash_slots in lk_init", (unsigned long long)q.hash_slots));
  }
  LK_CHECK(!(h[0] & ST_BAD_CODE), LK_ERR_IO, "corrupt segment: dictionary index out of range");
  uint32_t ph[3] = {0, h[1], h[2]};
  if (d.phase_global) { ph[1] = d.phase_gmin; ph[2] = d.phase_gmax; }  // what all ranks saw, not just this one
  h = ph;
  if (q.is_metrics && h[1] != 0xffffffffu) {
    // GROUP BY "_cardinalhq.timestamp": metric segments are pre-rolled to the step grid (QueryEngineV2.scala:746-752);
    // every timestamp must sit at one offset from startTs modulo step, otherwise buckets would merge distinct rows.
    if (h[1] != h[2]) {
      clear_arena_if_any(d);
      d.finalized_device = true;
      fail(LK_ERR_UNSUPPORTED, "metric timestamps are not on one step-aligned grid; cannot bucket GROUP BY timestamp densely");
    }
    d.phase = h[1];
  } else d.phase = 0;
}

// ---- record path: sizes the scratch (from the record count of the first finalize; later finalizes re-use it and the
// kernels verify that it still fits) and enqueues the kernels; nothing here waits for the device once sized ----
static size_t rec_parts(const Query& q) { return (size_t)q.nbuckets * (q.dev ? q.dev->rec_slices : 1u); }
static size_t rec_counter_stride(const Query& q) { return rec_parts(q) <= 8192 ? 32 : 1; }
// Partition the records by bucket before grouping them?  Measured on B200, C2: one GPU 0.395 ms without / 0.48 ms with (the
// extra pass costs more than the locality buys while a warp's 32 records already share ~2.5 buckets); 4 ranks 0.574 / 0.525;
// the list of a rank is then a mix of short appends from all sources and the gap widens with the rank count (2 ranks, after the
// partitioned path got its shared-memory grouping: 0.49 / 0.45).  So: whenever the query is sharded.  LK_REC_SCATTER=0|1
// overrides (tests run both ways).
// ... and on any number of ranks when the list is long (>= 16 M records: a key table of 128 MB and more lives in DRAM, every
// insert of the unpartitioned grouping is a DRAM round trip: C4, 49.9 M records, 4.25 ms) -- then with several partitions per time
// bucket, so that each still fits the shared-memory table of rec_group_bucket_kernel.
static bool rec_scatter_on(const Query& q, uint32_t nrec) {
  if (const char* e = getenv("LK_REC_SCATTER")) return atoi(e) != 0;
  return (q.comm != nullptr && q.comm->world >= 2) || nrec >= (16u << 20);
}

// the three clears a record finalize starts from: per-bucket counters, the key table, the bookkeeping block
static void rec_clear_scratch(Query& q, cudaStream_t st) {
  Query::Device& d = *q.dev;
  const size_t nb = rec_parts(q) + 1, cs = rec_counter_stride(q);
  CUDA_CHECK(cudaMemsetAsync(d.rf_tables, 0, nb * cs * 4, st));
  CUDA_CHECK(cudaMemsetAsync(d.rf_tables + 3 * nb * cs + 2 * nb, 0, nb * cs * 4, st));  // scatter cursors
  CUDA_CHECK(cudaMemsetAsync(d.rf_sorted, 0, 2 * d.fin_cap * 4, st));
  CUDA_CHECK(cudaMemsetAsync(d.fin, 0, sizeof(RecFin), st));
}

static void rec_finalize_size(Query& q, uint32_t nrec) {
  Query::Device& d = *q.dev;
  if (d.st2) CUDA_CHECK(cudaStreamSynchronize(d.st2));  // a clear of the old scratch may still be running there
  d.pre_cleared = false;
  d.rec_scatter = rec_scatter_on(q, nrec);
  d.rec_slices = 1;
  if (d.rec_scatter && q.nbuckets > 0) {
    // records per partition to aim at (its key table: two slots per record, 1.5 x headroom for uneven partitions; at most 8192
    // partitions, so that every counter keeps its own 128-byte line).  Measured on B200: C2 (6.2 M records, 17 k per time
    // bucket) 0.405 ms with one partition per bucket, 0.385-0.39 with two (8.6 k each: two CTAs per SM), 0.426 with six;
    // C4 (49.9 M records, 139 k per bucket) 2.98 ms with 6 partitions per bucket, 3.20 with 9, 3.37 with 17 -- the partition
    // pass pays for every extra partition with less coalesced stores.
    const uint64_t target = getenv("LK_REC_PART_TARGET") ? (uint64_t)std::max(256, atoi(getenv("LK_REC_PART_TARGET")))
                                                          : (nrec >= (16u << 20) ? 24576 : 12288);
    const uint64_t per_bucket = ((uint64_t)nrec + q.nbuckets - 1) / q.nbuckets;
    uint64_t sl = (per_bucket + target - 1) / target;
    if (const char* e = getenv("LK_REC_SLICES")) sl = (uint64_t)std::max(1, atoi(e));  // tests
    sl = std::min<uint64_t>(std::max<uint64_t>(sl, 1), std::max<uint64_t>(1, 8192 / q.nbuckets));
    d.rec_slices = (uint32_t)sl;
    const uint64_t per_part = (per_bucket + sl - 1) / sl;
    d.rec_smem_slots = (uint32_t)std::min<uint64_t>(RG_SLOTS, std::max<uint64_t>(4096, (3 * per_part + 1023) & ~1023ull));
  }
  const size_t cap = std::min<size_t>(std::max<size_t>(d.rec_cap, 1), (size_t)nrec + nrec / 8 + 4096);
  if (d.rf_sorted) CUDA_CHECK(cudaFreeAsync(d.rf_sorted, d.st));
  if (d.rf_tables) CUDA_CHECK(cudaFreeAsync(d.rf_tables, d.st));
  d.rf_sorted = nullptr;
  d.rf_tables = nullptr;
  CUDA_CHECK(cudaMallocAsync(&d.rf_sorted, 2 * cap * 4 + 64, d.st));  // the key table: two 32-bit slots per record
  // u32 words: bkt_recs[nb * cs] | bkt_rows[nb * cs] | row_cursor[nb * cs] | rec_start[nb] | row_start[nb]
  const size_t nb = rec_parts(q) + 1, cs = rec_counter_stride(q);
  CUDA_CHECK(cudaMallocAsync(&d.rf_tables, (4 * nb * cs + 2 * nb) * 4 + 64, d.st));  // (+ the scatter cursors, last)
  if (d.rec2_cell) CUDA_CHECK(cudaFreeAsync(d.rec2_cell, d.st));
  if (d.rec2_vals) CUDA_CHECK(cudaFreeAsync(d.rec2_vals, d.st));
  d.rec2_cell = d.rec2_vals = nullptr;
  if (d.rec_scatter) {
    CUDA_CHECK(cudaMallocAsync(&d.rec2_cell, cap * 8 + 64, d.st));
    CUDA_CHECK(cudaMallocAsync(&d.rec2_vals, cap * 8 * q.aggs.size() + 64, d.st));
  }
  if (!d.fin) CUDA_CHECK(cudaMallocAsync(&d.fin, sizeof(RecFin), d.st));
  if (!d.fin_host) d.fin_host = static_cast<uint32_t*>(pinned_alloc(64));
  d.fin_cap = cap;
  ensure_dres(d, result_bytes(q, (int64_t)cap));
}

static void rec_finalize_launch(Query& q) {
  Query::Device& d = *q.dev;
  RecGeom G;
  G.idx_bits = q.params.rec_idx_bits;
  G.gid_bits = q.params.rec_gid_bits;
  G.nbuckets = (uint32_t)rec_parts(q);
  G.slices = d.rec_slices;
  G.rec_cap = (uint32_t)std::min<size_t>(d.rec_cap, 0xffffffffu);
  G.fin_cap = (uint32_t)std::min<size_t>(d.fin_cap, 0xffffffffu);
  G.world = q.comm ? (uint32_t)q.comm->world : 1u;
  G.region_cap = q.comm ? (uint32_t)q.comm->region_cap : 0u;
  G.prefix = q.comm ? q.comm->ctrl(q.comm->rank)->prefix : nullptr;
  G.fp_shift = 1;
  const bool scatter = d.rec_scatter && d.rec2_cell != nullptr;
  const uint64_t max_index = (G.world > 1 && !scatter) ? (uint64_t)d.rec_cap : (uint64_t)d.fin_cap;
  while (G.fp_shift < 32 && (max_index + 1) >> G.fp_shift) G.fp_shift++;  // bits of (largest record position + 1)
  const size_t nb = rec_parts(q) + 1, cs = rec_counter_stride(q);
  G.cstride = (uint32_t)cs;
  uint32_t* bkt_recs = d.rf_tables;
  uint32_t* bkt_rows = bkt_recs + nb * cs;
  uint32_t* row_cursor = bkt_rows + nb * cs;
  uint32_t* rec_start = row_cursor + nb * cs;
  uint32_t* row_start = rec_start + nb;
  if (d.pre_cleared) {  // cleared on the side stream while the scan ran (device_execute)
    CUDA_CHECK(cudaStreamWaitEvent(d.st, d.ev_clear, 0));
    d.pre_cleared = false;
  } else rec_clear_scratch(q, d.st);
  EmitParams E;
  fill_emit_params(q, E, d.dres, (int64_t)d.fin_cap);
  E.phase_ptr = d.counters + 1;  // the scan's timestamp phase (metrics), read on the device: nothing waits for the host
  d.dres_stride = d.fin_cap;
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((d.fin_cap + RF_BLOCK * 16 - 1) / (RF_BLOCK * 16), (size_t)num_sms() * 8));
  rec_bhist_kernel<<<grid, RF_BLOCK, 0, d.st>>>(d.rec_cell, d.counters, G, bkt_recs);
  rec_regions_kernel<<<1, BS_BLOCK, 0, d.st>>>(d.counters, G, bkt_recs, rec_start, bkt_rows, d.fin);
  unsigned long long* gk = d.rec_cell;
  unsigned long long* gv = d.rec_vals;
  RecGeom G2 = G;
  if (scatter) {
    uint32_t* scat_cursor = rec_start + 2 * nb;
    rec_scatter_kernel<<<grid, RF_BLOCK, 0, d.st>>>(d.rec_cell, d.rec_vals, rec_start, scat_cursor, d.fin, G, (int)q.aggs.size(), d.rec2_cell, d.rec2_vals);
    gk = d.rec2_cell;
    gv = d.rec2_vals;
    G2.world = 1;  // one plain list from here on
    G2.rec_cap = G.fin_cap;
  }
  if (scatter) {
    static bool attr = false;
    if (!attr) { CUDA_CHECK(cudaFuncSetAttribute(rec_group_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RG_SLOTS * 4))); attr = true; }
    const uint32_t slots = d.rec_smem_slots ? d.rec_smem_slots : RG_SLOTS;
    rec_group_bucket_kernel<<<std::max<uint32_t>(G.nbuckets, 1), RG_BLOCK, (size_t)slots * 4, d.st>>>(gk, gv, rec_start, reinterpret_cast<uint32_t*>(d.rf_sorted), bkt_rows, d.fin, G2,
                                                                                                    E, slots);
  } else rec_group_kernel<false><<<grid, RF_BLOCK, 0, d.st>>>(gk, gv, rec_start, reinterpret_cast<uint32_t*>(d.rf_sorted), bkt_rows, d.fin, G2, E);
  rec_rowscan_kernel<<<1, BS_BLOCK, 0, d.st>>>(bkt_rows, G.nbuckets, G.cstride, row_start, row_cursor, d.fin);
  rec_emit_kernel<<<grid, RF_BLOCK, 0, d.st>>>(gk, gv, row_start, row_cursor, d.fin, G2, E);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(d.fin_host, d.fin, sizeof(RecFin), cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaMemcpyAsync(d.fin_host + 8, d.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.st));
  d.fin_pending = true;
}

// Waits for a pending record-path finalize and reads back what it left for the host: the row count, the scan's status
// flags and timestamp phase.  A finalize whose scratch turned out too small (more records than the capacity learned from
// an earlier run) is repeated once with the right size: the appended records are still in place.
void device_resolve(Query& q) {
  if (!q.dev || !q.dev->fin_pending) return;
  Query::Device& d = *q.dev;
  CUDA_CHECK(cudaSetDevice(global_options().device));
  for (int attempt = 0;; attempt++) {
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    d.fin_pending = false;
    const RecFin* f = reinterpret_cast<const RecFin*>(d.fin_host);
    check_scan_status(q, d.fin_host + 8);
    if (f->status & RF_ST_CAP) {
      LK_CHECK(attempt == 0, LK_ERR_NOMEM, "record finalize: scratch still too small after regrowing");
      rec_finalize_size(q, f->nrec);
      rec_finalize_launch(q);
      CUDA_CHECK(cudaEventRecord(d.ev[5], d.st));
      continue;
    }
    LK_CHECK(!(f->status & RF_ST_TABLE), LK_ERR_CUDA, "record finalize: key table region overflowed (internal error)");
    d.n_rows = f->nrows;
    return;
  }
}

void device_finalize_device(Query& q) {
  LK_CHECK(q.dev && q.dev->executed, LK_ERR_INVALID, "lk_query_finalize before lk_query_execute");
  Query::Device& d = *q.dev;
  CUDA_CHECK(cudaSetDevice(global_options().device));
  if (d.finalized_device) return;
  CUDA_CHECK(cudaEventRecord(d.ev[4], d.st));
  d.n_rows = 0;
  d.fin_pending = false;
  if ((q.n_cells == 0 && !q.comm) || (q.path != 2 && q.tiles.empty()) || (q.path == 2 && d.rec_cap == 0)) {
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    d.finalized_device = true;
    CUDA_CHECK(cudaEventRecord(d.ev[5], d.st));
    return;
  }
  if (q.path == 2) {
    if (q.comm) {
      const Comm& c = *q.comm;
      CUDA_CHECK(cudaEventRecord(d.ev[10], d.st));
      comm_wait_kernel<<<1, 32, 0, d.st>>>(c.ctrl(c.rank), (uint32_t)c.world, c.epoch, d.counters);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaEventRecord(d.ev[11], d.st));
    }
    if (q.exact_sums) {
      rec_exact_finalize(q);
      d.finalized_device = true;
      CUDA_CHECK(cudaEventRecord(d.ev[5], d.st));
      return;
    }
    if (d.fin_cap == 0) {
      // first finalize of this query: one read-back of the record count sizes the scratch and the result buffer
      CUDA_CHECK(cudaMemcpyAsync(d.h_counters, d.counters, sizeof d.h_counters, cudaMemcpyDeviceToHost, d.st));
      CUDA_CHECK(cudaStreamSynchronize(d.st));
      check_scan_status(q, d.h_counters);
      rec_finalize_size(q, d.h_counters[5]);
    }
    rec_finalize_launch(q);
    d.finalized_device = true;
    CUDA_CHECK(cudaEventRecord(d.ev[5], d.st));
    return;
  }
  CUDA_CHECK(cudaMemcpyAsync(d.h_counters, d.counters, sizeof d.h_counters, cudaMemcpyDeviceToHost, d.st));
  uint32_t nblocks = 0;
  if (q.path == 0) {
    nblocks = (uint32_t)((q.n_cells + CMP_CHUNK - 1) / CMP_CHUNK);
    ensure_block_counts(d, (size_t)nblocks + 1);
    dense_count_kernel<<<nblocks, CMP_BLOCK, 0, d.st>>>(d.planes, q.n_cells, d.block_counts);
    exclusive_scan_kernel<<<1, 1024, 0, d.st>>>(d.block_counts, nblocks, d.block_counts + nblocks);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(&d.h_counters[7], d.block_counts + nblocks, 4, cudaMemcpyDeviceToHost, d.st));
  }
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  check_scan_status(q, d.h_counters);
  const int64_t n = q.path == 1 ? (int64_t)d.h_counters[3] : (int64_t)d.h_counters[7];
  d.n_rows = n;
  d.dres_stride = (size_t)n;
  if (n > 0) {
    ensure_dres(d, result_bytes(q, n));
    EmitParams E;
    fill_emit_params(q, E, d.dres, n);
    if (q.path == 0) {
      dense_emit_kernel<<<nblocks, CMP_BLOCK, 0, d.st>>>(d.planes, q.n_cells, d.block_counts, E);
    } else {
      // scratch: hist/cursor[nbuckets + 1] | sorted[n]
      ensure_block_counts(d, (size_t)q.nbuckets + 1 + (size_t)n);
      uint32_t* hist = d.block_counts;
      const uint32_t* bucket_of = d.harena->occ_bkt;
      uint32_t* sorted = hist + q.nbuckets + 1;
      CUDA_CHECK(cudaMemsetAsync(hist, 0, ((size_t)q.nbuckets + 1) * 4, d.st));
      const int hgrid = (int)((n + HS_CHUNK - 1) / HS_CHUNK);
      hash_hist_kernel<<<hgrid, HS_BLOCK, 0, d.st>>>(bucket_of, (uint32_t)n, q.nbuckets, hist);
      exclusive_scan_kernel<<<1, 1024, 0, d.st>>>(hist, q.nbuckets, hist + q.nbuckets);
      hash_scatter_kernel<<<hgrid, HS_BLOCK, 0, d.st>>>(d.harena->occ, bucket_of, (uint32_t)n, q.nbuckets, hist, sorted);
      hash_emit_kernel<<<(int)((n + HS_BLOCK - 1) / HS_BLOCK), HS_BLOCK, 0, d.st>>>(sorted, (uint32_t)n, d.harena->entries, d.harena->stride, E);
    }
    CUDA_CHECK(cudaGetLastError());
  }
  d.finalized_device = true;
  CUDA_CHECK(cudaEventRecord(d.ev[5], d.st));
}

// ---- BaseExpr.eval on the reduced rows (BaseExpr.scala:665-695, 47-95; ASTUtils.scala:190-219) ----
// value = getFromSketch(map sketch, aggregation) -- `avg` = sum / count -- then the chart/metric-type transform:
// x * (step / 1000) or x / (step / 1000) with the INTEGER division of step done first, plain IEEE double arithmetic
// (a zero divisor gives +-Inf / NaN exactly as on the JVM).
__global__ void __launch_bounds__(256) eval_transform_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, int avg,
                                                             int transform, double secs, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double v = a ? a[i] : __longlong_as_double(0x7ff8000000000000ll);  // aggregation missing from the sketch: NaN
  if (avg) v = v / (b ? b[i] : __longlong_as_double(0x7ff8000000000000ll));
  if (transform == 1) v = v * secs;
  else if (transform == 2) v = v / secs;
  out[i] = v;
}

int64_t device_eval(Query& q, const std::string& aggregation, const std::string& chart_type, const std::string& metric_type, double* out, int64_t cap) {
  LK_CHECK(q.dev && q.dev->finalized_device, LK_ERR_INVALID, "lk_query_eval needs a finalized query");
  device_resolve(q);
  Query::Device& d = *q.dev;
  const int64_t n = d.n_rows;
  LK_CHECK(out && cap >= n, LK_ERR_INVALID, "lk_query_eval: output buffer too small");
  if (n == 0) return 0;
  CUDA_CHECK(cudaSetDevice(global_options().device));
  EmitParams E;
  fill_emit_params(q, E, d.dres, (int64_t)d.dres_stride);
  // map-sketch key -> aggregate slot.  Metrics are pre-rolled: the slot over `rollup_<key>` is the key's value
  // (count = sum(rollup_count)); otherwise the first slot whose aggregation has that name.
  auto column = [&](const char* name) -> const double* {
    if (q.is_metrics)
      for (size_t a = 0; a < q.aggs.size(); a++)
        if (q.aggs[a].value_column == std::string("rollup_") + name) return E.val[a];
    for (size_t a = 0; a < q.aggs.size(); a++)
      if (q.aggs[a].aggregation == name) return E.val[a];
    return nullptr;
  };
  const bool avg = aggregation == "avg";
  const double* a = avg ? column("sum") : column(aggregation.c_str());
  const double* b = avg ? column("count") : nullptr;
  // getTransformerFunc (ASTUtils.scala:190-219)
  int transform = 0;
  if (q.is_metrics) {
    if (chart_type == "count" && metric_type == "rate") transform = 1;
    else if (chart_type == "rate" && metric_type == "count") transform = 2;
  } else if (chart_type == "rate") transform = 2;
  const double secs = (double)(q.step / 1000);
  double* dev_out = nullptr;
  CUDA_CHECK(cudaMallocAsync(&dev_out, (size_t)n * 8, d.st));
  eval_transform_kernel<<<(int)((n + 255) / 256), 256, 0, d.st>>>(a, b, n, avg ? 1 : 0, transform, secs, dev_out);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(out, dev_out, (size_t)n * 8, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaFreeAsync(dev_out, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  return n;
}

// ------------------------------------------------------------------------------------------------------------
// Formula.eval (core/.../utils/ast/Formula.scala:32-69) over the reduced rows of two finalized queries, still in HBM
// ------------------------------------------------------------------------------------------------------------
// Per SketchGroup (= timestamp) the reference evaluates both sides into maps groupKey -> EvalResult -- BaseExpr.eval keys a
// side by ITS sorted group-by values joined with ":" (tags dropped by toDataPoint -- NULL, "", "null" -- read as "";
// ASTUtils.scala:87-89), a later row of the same key replaces an earlier one -- and combines equal keys: add / sub / mul /
// div, `add` fills a missing side with 0, a zero divisor gives no result.  Here: every row gets the integer key
// (timestamp - base) * G + mixed radix of its group-by values in dictionaries COMMON to both sides; both sides meet in one
// hash table {key, last row of side 1, last row of side 2} (atomicMax: "the later row wins"); the winners are compacted in
// row order (= timestamp order) into (timestamp, value, side whose tags the result carries, row in that side's result).
struct FormulaSlot { unsigned long long key; uint32_t row1, row2; };
struct FormulaKeyParams {
  int n_pos;                           // group-by positions (sorted group-by names)
  const int32_t* code[LK_MAX_KEYS];    // result column of the side's key column at this position (null: the group-by does not exist -> "")
  const uint32_t* remap[LK_MAX_KEYS];  // side's dictionary code -> common id (one more entry at the end: SQL NULL)
  uint32_t null_slot[LK_MAX_KEYS];     // index of that last entry
  uint64_t stride[LK_MAX_KEYS];
  uint64_t groups;                     // product of the common dictionary sizes
  int64_t base;
};

__device__ __forceinline__ unsigned long long formula_key(const FormulaKeyParams& K, const int64_t* __restrict__ ts, int64_t i) {
  unsigned long long g = 0;
  for (int j = 0; j < K.n_pos; j++) {
    uint32_t id = 0;
    if (K.code[j]) { const int32_t c = K.code[j][i]; id = K.remap[j][c < 0 ? K.null_slot[j] : (uint32_t)c]; }
    g += (unsigned long long)id * K.stride[j];
  }
  return (unsigned long long)(ts[i] - K.base) * K.groups + g;
}

__device__ __forceinline__ uint32_t formula_find(FormulaSlot* __restrict__ tab, uint32_t mask, unsigned long long key, bool insert) {
  unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  uint32_t s = (uint32_t)(h >> 32) & mask;
  const unsigned long long want = key + 1;  // 0 = empty
  while (true) {
    unsigned long long cur = *(volatile unsigned long long*)&tab[s].key;
    if (cur == 0 && insert) cur = atomicCAS(&tab[s].key, 0ull, want), cur = cur == 0 ? want : cur;
    if (cur == want) return s;
    if (cur == 0) return 0xffffffffu;
    s = (s + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) formula_insert_kernel(const __grid_constant__ FormulaKeyParams K, const int64_t* __restrict__ ts, int64_t n, int side,
                                                             FormulaSlot* __restrict__ tab, uint32_t mask) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = formula_find(tab, mask, formula_key(K, ts, i), true);
  atomicMax(side == 1 ? &tab[s].row1 : &tab[s].row2, (uint32_t)i + 1);
}

// pass 0 counts the results of every 256-row block, pass 1 (after the prefix over the blocks) writes them in row order
struct FormulaEmit {
  int op;            // 0 add 1 sub 2 mul 3 div
  int side;          // rows of which side this launch walks (1 or 2)
  int other_const;   // the other side is a constant
  double cval;
  const double* v1;  // per-row values of side 1 / side 2 (null for a constant side)
  const double* v2;
  int const_is_e1;   // constant side is e1 (the walked side is then e2)
  int64_t* out_ts; double* out_val; int32_t* out_side; int64_t* out_row;
};

__device__ __forceinline__ bool formula_row(const FormulaKeyParams& K, const FormulaEmit& F, const int64_t* __restrict__ ts, int64_t i,
                                            const FormulaSlot* __restrict__ tab, uint32_t mask, double& value) {
  const uint32_t s = formula_find(const_cast<FormulaSlot*>(tab), mask, formula_key(K, ts, i), false);
  const FormulaSlot sl = tab[s];
  if (F.side == 1) {
    if (sl.row1 != (uint32_t)i + 1) return false;  // a later row of the same key replaced this one
    double a = F.v1[i], b;
    if (F.other_const) b = F.cval;
    else if (sl.row2) b = F.v2[sl.row2 - 1];
    else if (F.op == 0) b = 0.0;  // add: the missing side counts as 0 (Formula.scala:45-47)
    else return false;
    if (F.const_is_e1) { const double t = a; a = b; b = t; }
    if (F.op == 3 && b == 0.0) return false;  // divide by zero = missing data (Formula.scala:60-64)
    value = F.op == 0 ? a + b : F.op == 1 ? a - b : F.op == 2 ? a * b : a / b;
    return true;
  }
  // side 2: only what `add` contributes when side 1 has no row for the key
  if (sl.row2 != (uint32_t)i + 1 || sl.row1 != 0 || F.op != 0) return false;
  value = 0.0 + F.v2[i];
  return true;
}

__global__ void __launch_bounds__(256) formula_emit_kernel(const __grid_constant__ FormulaKeyParams K, const __grid_constant__ FormulaEmit F,
                                                           const int64_t* __restrict__ ts, int64_t n, const FormulaSlot* __restrict__ tab, uint32_t mask,
                                                           uint32_t* __restrict__ block_counts, int pass, uint32_t out_base) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  double value = 0;
  const bool keep = i < n && formula_row(K, F, ts, i, tab, mask, value);
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  __shared__ uint32_t wcnt[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) wcnt[wid] = __popc(bal);
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (int w = 0; w < 8; w++) { if (w < wid) before += wcnt[w]; total += wcnt[w]; }
  if (pass == 0) { if (threadIdx.x == 0) block_counts[blockIdx.x] = total; return; }
  if (!keep) return;
  const uint32_t o = out_base + block_counts[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1));
  F.out_ts[o] = ts[i];
  F.out_val[o] = value;
  // whose tags: e1's -- for a constant e1 those of the sketch input it was keyed by (= this e2 row), or none at all without
  // group-bys (ASTUtils.scala:51-55: tags = Map())
  F.out_side[o] = F.const_is_e1 ? (K.n_pos ? 2 : 0) : F.side;
  F.out_row[o] = i;
}

namespace {
struct FormulaSideSpec {
  bool constant = false;
  double cval = 0;
  std::string aggregation = "sum", chart_type = "line", metric_type = "gauge";
  std::vector<std::string> group_bys;
  bool has_group_bys = false;
};
FormulaSideSpec parse_formula_side(const Json* j, const char* which) {
  FormulaSideSpec s;
  LK_CHECK(j && j->is_obj(), LK_ERR_INVALID, std::string("formula spec needs an object '") + which + "'");
  if (const Json* c = j->get("constant")) {
    s.constant = true;
    s.cval = c->kind == Json::Num ? c->num : c->kind == Json::Str ? atof(c->str.c_str()) : 0.0;
    if (c->kind == Json::Num && c->is_int) s.cval = (double)c->i64;
    return s;
  }
  if (const Json* a = j->get("aggregation"); a && a->text()) s.aggregation = a->str;
  if (const Json* a = j->get("chartType"); a && a->text()) s.chart_type = a->str;
  if (const Json* a = j->get("metricType"); a && a->text()) s.metric_type = a->str;
  if (const Json* g = j->get("groupBys"); g && g->is_arr()) {
    s.has_group_bys = true;
    for (auto& x : g->arr) if (x.text()) s.group_bys.push_back(x.str);
  }
  return s;
}
std::string canon_tag(const std::string& v) { return (v.empty() || v == "null") ? std::string() : v; }  // Commons.scala:433: dropped tags read as ""
}  // namespace

// the per-row value of a BaseExpr side: getFromSketch + getTransformerFunc, as device_eval
static double* formula_side_values(Query& q, const FormulaSideSpec& sp, cudaStream_t st) {
  Query::Device& d = *q.dev;
  const int64_t n = d.n_rows;
  EmitParams E;
  fill_emit_params(q, E, d.dres, (int64_t)d.dres_stride);
  auto column = [&](const char* name) -> const double* {
    if (q.is_metrics)
      for (size_t a = 0; a < q.aggs.size(); a++)
        if (q.aggs[a].value_column == std::string("rollup_") + name) return E.val[a];
    for (size_t a = 0; a < q.aggs.size(); a++)
      if (q.aggs[a].aggregation == name) return E.val[a];
    return nullptr;
  };
  const bool avg = sp.aggregation == "avg";
  const double* a = avg ? column("sum") : column(sp.aggregation.c_str());
  const double* b = avg ? column("count") : nullptr;
  int transform = 0;
  if (q.is_metrics) {
    if (sp.chart_type == "count" && sp.metric_type == "rate") transform = 1;
    else if (sp.chart_type == "rate" && sp.metric_type == "count") transform = 2;
  } else if (sp.chart_type == "rate") transform = 2;
  double* out = nullptr;
  CUDA_CHECK(cudaMallocAsync(&out, std::max<size_t>((size_t)n * 8, 16), st));
  if (n) eval_transform_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(a, b, n, avg ? 1 : 0, transform, (double)(q.step / 1000), out);
  return out;
}

int64_t device_formula(Query* q1, Query* q2, const std::string& spec_json, int64_t cap, int64_t* out_ts, double* out_val, int32_t* out_side, int64_t* out_row) {
  Json spec = parse_json(spec_json);
  LK_CHECK(spec.is_obj(), LK_ERR_INVALID, "formula spec must be a JSON object");
  const Json* opj = spec.get("op");
  LK_CHECK(opj && opj->text(), LK_ERR_INVALID, "formula spec needs op");
  const int op = opj->str == "add" ? 0 : opj->str == "sub" ? 1 : opj->str == "mul" ? 2 : opj->str == "div" ? 3 : -1;
  LK_CHECK(op >= 0, LK_ERR_INVALID, "formula op must be add | sub | mul | div");
  FormulaSideSpec sp[2] = {parse_formula_side(spec.get("e1"), "e1"), parse_formula_side(spec.get("e2"), "e2")};
  Query* qs[2] = {q1, q2};
  LK_CHECK(!(sp[0].constant && sp[1].constant), LK_ERR_UNSUPPORTED, "formula of two constants has no rows to evaluate on");
  for (int s = 0; s < 2; s++) {
    if (sp[s].constant) continue;
    LK_CHECK(qs[s] && qs[s]->dev && qs[s]->dev->finalized_device, LK_ERR_INVALID, "lk_formula_eval needs finalized queries for its BaseExpr sides");
    LK_CHECK(!qs[s]->tag_query, LK_ERR_INVALID, "lk_formula_eval over a tag query");
    device_resolve(*qs[s]);
    if (!sp[s].has_group_bys) sp[s].group_bys = qs[s]->req.expr.chart.group_bys;
    std::sort(sp[s].group_bys.begin(), sp[s].group_bys.end());
    sp[s].group_bys.erase(std::unique(sp[s].group_bys.begin(), sp[s].group_bys.end()), sp[s].group_bys.end());
  }
  const int walk = sp[0].constant ? 1 : 0;           // the side whose rows drive the output (e1 unless it is the constant)
  const int other = 1 - walk;
  const bool other_const = sp[other].constant;
  Query& qa = *qs[walk];
  CUDA_CHECK(cudaSetDevice(global_options().device));
  cudaStream_t st = qa.dev->st;
  if (!other_const) {
    LK_CHECK(sp[0].group_bys.size() == sp[1].group_bys.size(), LK_ERR_UNSUPPORTED,
             "formula sides with different numbers of group-bys (their group keys can only meet by accident)");
    LK_CHECK(qs[other]->base - qa.base < (1ll << 40) && qa.base - qs[other]->base < (1ll << 40), LK_ERR_UNSUPPORTED, "formula sides over distant time ranges");
    CUDA_CHECK(cudaStreamSynchronize(qs[other]->dev->st));  // its result columns are read on the other query's stream
  }
  const int m = (int)sp[walk].group_bys.size();
  LK_CHECK(m <= LK_MAX_KEYS, LK_ERR_UNSUPPORTED, "too many group-bys in a formula");
  // ---- common dictionaries per group-by position ----
  auto key_col = [](const Query& q, const std::string& name) -> int {
    for (size_t k = 1; k < q.key_names.size(); k++) if (q.key_names[k] == name) return (int)k;  // (slot 0 is "name")
    return -1;
  };
  std::vector<std::vector<std::string>> common(m);
  for (int j = 0; j < m; j++) {
    std::vector<std::string>& c = common[j];
    c.push_back(std::string());
    for (int s = 0; s < 2; s++) {
      if (sp[s].constant) continue;
      const int k = key_col(*qs[s], sp[s].group_bys[j]);
      if (k >= 0) for (auto& v : qs[s]->key_dicts[k]) c.push_back(canon_tag(v));
    }
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    if (m >= 2) for (auto& v : c) LK_CHECK(v.find(':') == std::string::npos, LK_ERR_UNSUPPORTED, "group-by values containing ':' make the reference's joined group keys ambiguous");
  }
  FormulaKeyParams K[2];
  std::vector<uint32_t*> dev_tmp;
  unsigned __int128 groups = 1;
  std::vector<uint64_t> strides(m);
  for (int j = m - 1; j >= 0; j--) { strides[j] = (uint64_t)groups; groups *= common[j].size(); LK_CHECK(groups < ((unsigned __int128)1 << 40), LK_ERR_UNSUPPORTED, "formula group space too large"); }
  int64_t base = qa.base;
  if (!other_const) base = std::min(base, qs[other]->base);
  EmitParams E[2];
  for (int s = 0; s < 2; s++) {
    if (sp[s].constant) continue;
    Query& q = *qs[s];
    fill_emit_params(q, E[s], q.dev->dres, (int64_t)q.dev->dres_stride);
    LK_CHECK((unsigned __int128)((uint64_t)(q.ts_hi - base) + (uint64_t)q.step + 1) * groups < ((unsigned __int128)1 << 62), LK_ERR_UNSUPPORTED, "formula key space too large");
    FormulaKeyParams& P = K[s];
    memset(&P, 0, sizeof P);
    P.n_pos = m;
    P.groups = (uint64_t)groups;
    P.base = base;
    for (int j = 0; j < m; j++) {
      P.stride[j] = strides[j];
      const int k = key_col(q, sp[s].group_bys[j]);
      if (k < 0) continue;  // the group-by does not exist in this side's files: "" for every row
      const auto& dict = q.key_dicts[k];
      std::vector<uint32_t> remap(dict.size() + 1, 0);
      for (size_t c = 0; c <= dict.size(); c++) {
        const std::string v = c < dict.size() ? canon_tag(dict[c]) : std::string();
        remap[c] = (uint32_t)(std::lower_bound(common[j].begin(), common[j].end(), v) - common[j].begin());
      }
      uint32_t* dr = nullptr;
      CUDA_CHECK(cudaMallocAsync(&dr, remap.size() * 4, st));
      CUDA_CHECK(cudaMemcpyAsync(dr, remap.data(), remap.size() * 4, cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaStreamSynchronize(st));  // `remap` is a local
      dev_tmp.push_back(dr);
      P.code[j] = E[s].code[k];
      P.remap[j] = dr;
      P.null_slot[j] = (uint32_t)dict.size();
    }
  }
  // ---- values, table, insert ----
  const int64_t n_walk = qa.dev->n_rows, n_other = other_const ? 0 : qs[other]->dev->n_rows;
  LK_CHECK(n_walk < 0x7fffffff && n_other < 0x7fffffff, LK_ERR_UNSUPPORTED, "formula over more than 2^31 rows");
  double* v[2] = {nullptr, nullptr};
  for (int s = 0; s < 2; s++) if (!sp[s].constant) v[s] = formula_side_values(*qs[s], sp[s], st);
  uint64_t slots = 1024;
  while (slots < 2 * (uint64_t)(n_walk + n_other)) slots <<= 1;
  FormulaSlot* tab = nullptr;
  CUDA_CHECK(cudaMallocAsync(&tab, slots * sizeof(FormulaSlot), st));
  CUDA_CHECK(cudaMemsetAsync(tab, 0, slots * sizeof(FormulaSlot), st));
  const uint32_t mask = (uint32_t)(slots - 1);
  // the walked side is "side 1" of the table, the other one "side 2"
  if (n_walk) formula_insert_kernel<<<(int)((n_walk + 255) / 256), 256, 0, st>>>(K[walk], E[walk].ts, n_walk, 1, tab, mask);
  if (n_other) formula_insert_kernel<<<(int)((n_other + 255) / 256), 256, 0, st>>>(K[other], E[other].ts, n_other, 2, tab, mask);
  // ---- ordered compaction: rows of the walked side, then (add) the rows only the other side has ----
  const size_t cap_rows = (size_t)(n_walk + n_other) + 1;
  int64_t* d_ts = nullptr; double* d_val = nullptr; int32_t* d_side = nullptr; int64_t* d_row = nullptr;
  uint32_t* d_bc = nullptr;
  uint32_t* d_tot = nullptr;
  CUDA_CHECK(cudaMallocAsync(&d_ts, cap_rows * 8, st));
  CUDA_CHECK(cudaMallocAsync(&d_val, cap_rows * 8, st));
  CUDA_CHECK(cudaMallocAsync(&d_side, cap_rows * 4, st));
  CUDA_CHECK(cudaMallocAsync(&d_row, cap_rows * 8, st));
  const uint32_t nb1 = (uint32_t)((n_walk + 255) / 256), nb2 = (uint32_t)((n_other + 255) / 256);
  CUDA_CHECK(cudaMallocAsync(&d_bc, ((size_t)nb1 + nb2 + 2) * 4, st));
  CUDA_CHECK(cudaMallocAsync(&d_tot, 8, st));
  FormulaEmit F;
  memset(&F, 0, sizeof F);
  F.op = op; F.cval = sp[other].constant ? sp[other].cval : 0.0; F.other_const = other_const ? 1 : 0; F.const_is_e1 = walk == 1 ? 1 : 0;
  F.v1 = v[walk]; F.v2 = v[other];
  F.out_ts = d_ts; F.out_val = d_val; F.out_side = d_side; F.out_row = d_row;
  uint32_t counts[2] = {0, 0};
  if (nb1) {
    F.side = 1;
    formula_emit_kernel<<<nb1, 256, 0, st>>>(K[walk], F, E[walk].ts, n_walk, tab, mask, d_bc, 0, 0);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(d_bc, nb1, d_tot);
    formula_emit_kernel<<<nb1, 256, 0, st>>>(K[walk], F, E[walk].ts, n_walk, tab, mask, d_bc, 1, 0);
    CUDA_CHECK(cudaMemcpyAsync(&counts[0], d_tot, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (nb2 && op == 0) {
    F.side = 2;
    // rows only e2 has: their tags are e2's.  (walk == other is impossible here: a constant side has no rows.)
    F.v2 = v[other];
    formula_emit_kernel<<<nb2, 256, 0, st>>>(K[other], F, E[other].ts, n_other, tab, mask, d_bc + nb1 + 1, 0, 0);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(d_bc + nb1 + 1, nb2, d_tot + 1);
    formula_emit_kernel<<<nb2, 256, 0, st>>>(K[other], F, E[other].ts, n_other, tab, mask, d_bc + nb1 + 1, 1, counts[0]);
    CUDA_CHECK(cudaMemcpyAsync(&counts[1], d_tot + 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  CUDA_CHECK(cudaGetLastError());
  const int64_t n_out = (int64_t)counts[0] + counts[1];
  LK_CHECK(cap >= n_out, LK_ERR_INVALID, strf("lk_formula_eval: %lld rows do not fit the output buffers (%lld)", (long long)n_out, (long long)cap));
  // both lists are in timestamp order; the result interleaves them by timestamp (e1's rows first on ties)
  std::vector<int64_t> h_ts((size_t)n_out), h_row((size_t)n_out);
  std::vector<double> h_val((size_t)n_out);
  std::vector<int32_t> h_side((size_t)n_out);
  if (n_out) {
    CUDA_CHECK(cudaMemcpyAsync(h_ts.data(), d_ts, (size_t)n_out * 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_val.data(), d_val, (size_t)n_out * 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_side.data(), d_side, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_row.data(), d_row, (size_t)n_out * 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  size_t a = 0, b = counts[0], o = 0;
  const size_t a_end = counts[0], b_end = (size_t)n_out;
  while (a < a_end || b < b_end) {
    const bool take_a = b >= b_end || (a < a_end && h_ts[a] <= h_ts[b]);
    const size_t i = take_a ? a++ : b++;
    out_ts[o] = h_ts[i]; out_val[o] = h_val[i]; out_side[o] = h_side[i]; out_row[o] = h_row[i];
    o++;
  }
  for (void* p : {(void*)tab, (void*)d_ts, (void*)d_val, (void*)d_side, (void*)d_row, (void*)d_bc, (void*)d_tot, (void*)v[0], (void*)v[1]}) if (p) cudaFreeAsync(p, st);
  for (auto* p : dev_tmp) cudaFreeAsync(p, st);
  CUDA_CHECK(cudaStreamSynchronize(st));
  return n_out;
}

HostResult* device_fetch(Query& q) {
  Query::Device& d = *q.dev;
  LK_CHECK(d.finalized_device, LK_ERR_INVALID, "fetch before finalize");
  device_resolve(q);
  auto r = std::make_unique<HostResult>();
  const int64_t n = d.n_rows;
  r->n = n;
  r->n_values = (int)q.aggs.size();
  r->n_tags = (int)q.key_pcols.size();
  r->tag_query = q.tag_query;
  if (q.tag_query) {  // SELECT "tag" as "tag", COUNT(*) AS count
    r->col_names.push_back(q.key_names[0]);
    r->col_names.push_back("count");
  } else {
    r->col_names.push_back(q.ts_col_name);
    for (size_t a = 0; a < q.aggs.size(); a++)
      r->col_names.push_back(q.is_metrics ? (q.aggs.size() == 1 ? std::string("value") : "value_" + std::to_string(a))
                                           : q.aggs[a].aggregation + "(\"" + q.aggs[a].value_column + "\")");
    for (auto& k : q.key_names) r->col_names.push_back(k);
  }
  r->dicts = q.key_dicts;
  r->dict_ptrs.resize(r->dicts.size());
  for (size_t k = 0; k < r->dicts.size(); k++)
    for (auto& s : r->dicts[k]) r->dict_ptrs[k].push_back(s.c_str());
  CUDA_CHECK(cudaEventRecord(d.ev[6], d.st));
  if (n > 0) {
    size_t bytes = result_bytes(q, n);
    r->pinned = (uint8_t*)pinned_alloc(bytes);
    uint8_t* p = r->pinned;
    r->ts = (int64_t*)p; p += 8 * n;
    for (size_t a = 0; a < q.aggs.size(); a++) { r->values.push_back((double*)p); p += 8 * n; }
    for (size_t k = 0; k < q.key_pcols.size(); k++) { r->codes.push_back((int32_t*)p); p += 4 * n; }
    for (size_t a = 0; a < q.aggs.size(); a++) { r->nulls.push_back(p); p += n; }
    if (d.dres_stride == (size_t)n) {
      CUDA_CHECK(cudaMemcpyAsync(r->pinned, d.dres, bytes - 64, cudaMemcpyDeviceToHost, d.st));
    } else {
      // the device columns are `dres_stride` rows apart (record path: capacity, not the row count): one copy per column
      EmitParams E;
      fill_emit_params(q, E, d.dres, (int64_t)d.dres_stride);
      CUDA_CHECK(cudaMemcpyAsync(r->ts, E.ts, 8 * n, cudaMemcpyDeviceToHost, d.st));
      for (size_t a = 0; a < q.aggs.size(); a++) {
        CUDA_CHECK(cudaMemcpyAsync(r->values[a], E.val[a], 8 * n, cudaMemcpyDeviceToHost, d.st));
        CUDA_CHECK(cudaMemcpyAsync(r->nulls[a], E.nul[a], n, cudaMemcpyDeviceToHost, d.st));
      }
      for (size_t k = 0; k < q.key_pcols.size(); k++) CUDA_CHECK(cudaMemcpyAsync(r->codes[k], E.code[k], 4 * n, cudaMemcpyDeviceToHost, d.st));
    }
  } else {
    for (size_t a = 0; a < q.aggs.size(); a++) { r->values.push_back(nullptr); r->nulls.push_back(nullptr); }
    for (size_t k = 0; k < q.key_pcols.size(); k++) r->codes.push_back(nullptr);
  }
  CUDA_CHECK(cudaEventRecord(d.ev[7], d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, d.ev[2], d.ev[3])); q.t_ms[1] = ms;
  CUDA_CHECK(cudaEventElapsedTime(&ms, d.ev[8], d.ev[9])); q.t_ms[5] = ms;
  CUDA_CHECK(cudaEventElapsedTime(&ms, d.ev[4], d.ev[5])); q.t_ms[2] = ms;
  CUDA_CHECK(cudaEventElapsedTime(&ms, d.ev[6], d.ev[7])); q.t_ms[3] = ms;
  return r.release();
}

void device_timings(Query& q) {
  Query::Device& d = *q.dev;
  float ms = 0;
  if (d.executed && cudaEventElapsedTime(&ms, d.ev[2], d.ev[3]) == cudaSuccess) q.t_ms[1] = ms;
  if (d.executed && cudaEventElapsedTime(&ms, d.ev[8], d.ev[9]) == cudaSuccess) q.t_ms[5] = ms;
  if (d.finalized_device && cudaEventElapsedTime(&ms, d.ev[4], d.ev[5]) == cudaSuccess) q.t_ms[2] = ms;
  if (d.finalized_device && q.comm && q.path == 2 && cudaEventElapsedTime(&ms, d.ev[10], d.ev[11]) == cudaSuccess) q.t_ms[6] = ms;
  cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// sparse exchange (hash path, sharded evaluation): every rank owns the cells whose hash falls into its partition
// ------------------------------------------------------------------------------------------------------------
constexpr int SP_MAXPARTS = 64;
struct AggSlot4 { uint8_t op[LK_MAX_AGGS + 1]; };

__device__ __forceinline__ uint32_t cell_partition(unsigned long long cell, uint32_t nparts) {
  return (uint32_t)((lk_hash64(cell ^ 0x9e3779b97f4a7c15ull) >> 33) % nparts);
}

__global__ void __launch_bounds__(HS_BLOCK) sparse_hist_kernel(const uint32_t* __restrict__ occ, uint32_t n, const uint8_t* __restrict__ entries, uint32_t stride,
                                                               uint32_t nparts, uint32_t* __restrict__ part_of, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[SP_MAXPARTS];
  if (threadIdx.x < SP_MAXPARTS) h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * HS_CHUNK, hi = min(n, lo + HS_CHUNK);
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {  // uniform trip count: warp_agg_inc needs the whole warp
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    uint32_t p = 0;
    if (valid) {
      const unsigned long long key = *reinterpret_cast<const unsigned long long*>(entries + (uint64_t)occ[i] * stride);
      p = cell_partition(key - 1, nparts);
      part_of[i] = p;
    }
    warp_agg_inc(h, p, valid);
  }
  __syncthreads();
  if (threadIdx.x < nparts && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// copies every claimed entry into its partition's range of `out` and returns the table slot to the clean state
__global__ void __launch_bounds__(HS_BLOCK) sparse_scatter_kernel(const uint32_t* __restrict__ occ, const uint32_t* __restrict__ part_of, uint32_t n,
                                                                  uint8_t* __restrict__ entries, uint32_t stride, uint32_t nparts,
                                                                  uint32_t* __restrict__ cursor, uint8_t* __restrict__ out) {
  __shared__ uint32_t cnt[SP_MAXPARTS];
  __shared__ uint32_t base[SP_MAXPARTS];
  if (threadIdx.x < SP_MAXPARTS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * HS_CHUNK, hi = min(n, lo + HS_CHUNK);
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    warp_agg_inc(cnt, valid ? part_of[i] : 0u, valid);
  }
  __syncthreads();
  if (threadIdx.x < nparts) {
    const uint32_t c = cnt[threadIdx.x];
    if (c) base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], c);
    cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += HS_BLOCK) {
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    const uint32_t p = valid ? part_of[i] : 0u;
    const uint32_t pos = base[p] + warp_agg_inc(cnt, p, valid);
    if (!valid) continue;
    ulonglong2* src = reinterpret_cast<ulonglong2*>(entries + (uint64_t)occ[i] * stride);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + (uint64_t)pos * stride);
    // all loads, then all stores of the copy, then the clearing stores: the 16-byte pieces of one output entry reach
    // L2 together (interleaved with the loads they arrived a DRAM latency apart, half-written sectors were evicted and
    // the kernel took 3.5 ms for 6 M entries)
    const uint32_t nw = stride / 16;  // 2 or 4
    ulonglong2 v[4];
#pragma unroll
    for (uint32_t w = 0; w < 4; w++) v[w] = w < nw ? src[w] : make_ulonglong2(0ull, 0ull);
#pragma unroll
    for (uint32_t w = 0; w < 4; w++) if (w < nw) dst[w] = v[w];
#pragma unroll
    for (uint32_t w = 0; w < 4; w++) if (w < nw) src[w] = make_ulonglong2(0ull, 0ull);
  }
}

// merges n foreign entries {key, acc[n_aggs]} into the table: sum/count add, min/max (both stored as "max" keys) max
__global__ void __launch_bounds__(HS_BLOCK) sparse_merge_kernel(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ entries, uint32_t stride,
                                                                uint64_t mask, uint32_t* __restrict__ occ, uint32_t* __restrict__ occ_bkt, uint64_t n_groups, uint32_t occ_cap,
                                                                uint32_t* __restrict__ counters,
                                                                int n_aggs, const __grid_constant__ AggSlot4 ops) {
  const uint32_t i = blockIdx.x * HS_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool claimed = false;
  uint32_t claimed_slot = 0;
  uint32_t status = 0;
  unsigned long long key = 1;
  if (i < n) {
    const unsigned long long* e_in = reinterpret_cast<const unsigned long long*>(in + (uint64_t)i * stride);
    key = e_in[0];
    uint64_t slot = lk_hash64(key - 1) & mask;
    unsigned long long* entry = nullptr;
    for (int probe = 0; probe < 4096; probe++) {
      unsigned long long* e = reinterpret_cast<unsigned long long*>(entries + slot * stride);
      unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(e);
      if (k == LK_EMPTY_KEY) {
        k = atomicCAS(e, (unsigned long long)LK_EMPTY_KEY, key);
        if (k == LK_EMPTY_KEY) { claimed = true; claimed_slot = (uint32_t)slot; entry = e; break; }
      }
      if (k == key) { entry = e; break; }
      slot = (slot + 1) & mask;
    }
    if (!entry) status |= ST_HASH_FULL;
    else
      for (int a = 0; a < n_aggs; a++) {
        const unsigned long long w = e_in[1 + a];
        if (ops.op[a] == AGG_SUM) atomicAdd(reinterpret_cast<double*>(entry + 1 + a), __longlong_as_double((long long)w));
        else if (ops.op[a] == AGG_COUNT) atomicAdd(entry + 1 + a, w);
        else if (w) atomicMax(entry + 1 + a, w);
      }
  }
  const unsigned cm = __ballot_sync(0xffffffffu, claimed);
  if (cm) {
    uint32_t base = 0;
    if (lane == __ffs(cm) - 1) base = atomicAdd(counters + 3, (uint32_t)__popc(cm));
    base = __shfl_sync(0xffffffffu, base, __ffs(cm) - 1);
    if (claimed) {
      const uint32_t idx = base + __popc(cm & ((1u << lane) - 1));
      if (idx < occ_cap) { occ[idx] = claimed_slot; occ_bkt[idx] = (uint32_t)((key - 1) / n_groups); }
      else status |= ST_HASH_FULL;
    }
  }
  if (status) atomicOr(counters + 0, status);
}

void device_partial_sparse(Query& q, int nparts, void** entries_out, int64_t* counts, int* stride_out) {
  LK_CHECK(q.dev && q.dev->executed, LK_ERR_INVALID, "lk_query_partial_sparse before lk_query_execute");
  LK_CHECK(q.path != 0, LK_ERR_INVALID, "query uses the dense path; use lk_query_partial_dense");
  LK_CHECK(q.path != 2, LK_ERR_INVALID, "the record path exchanges its records during the scan: attach an lk_comm (lk_query_set_comm)");
  LK_CHECK(nparts >= 1 && nparts <= SP_MAXPARTS, LK_ERR_INVALID, "nparts must be in [1, 64]");
  Query::Device& d = *q.dev;
  LK_CHECK(!d.finalized_device, LK_ERR_INVALID, "lk_query_partial_sparse after finalize");
  CUDA_CHECK(cudaSetDevice(global_options().device));
  CUDA_CHECK(cudaMemcpyAsync(d.h_counters, d.counters, sizeof d.h_counters, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  if (d.h_counters[0] & ST_HASH_FULL) {
    if (d.harena) CUDA_CHECK(cudaMemsetAsync(d.harena->entries, 0, d.harena->slots * d.harena->stride, d.st));
    d.finalized_device = true;
    fail(LK_ERR_NOMEM, "aggregate hash table overflowed; raise max_hash_slots in lk_init");
  }
  LK_CHECK(!(d.h_counters[0] & ST_BAD_CODE), LK_ERR_IO, "corrupt segment: dictionary index out of range");
  const uint32_t n = q.n_cells ? d.h_counters[3] : 0;
  const uint32_t stride = q.hash_stride;
  for (int p = 0; p < nparts; p++) counts[p] = 0;
  *stride_out = (int)stride;
  *entries_out = nullptr;
  if (n == 0) return;
  // scratch: hist[64] | cursor[64] | part_of[n]
  ensure_block_counts(d, 128 + (size_t)n);
  uint32_t* hist = d.block_counts;
  uint32_t* cursor = hist + 64;
  uint32_t* part_of = hist + 128;
  if (d.sparse_cap < (size_t)n * stride) {
    if (d.sparse_out) CUDA_CHECK(cudaFreeAsync(d.sparse_out, d.st));
    d.sparse_out = nullptr;
    CUDA_CHECK(cudaMallocAsync(&d.sparse_out, (size_t)n * stride, d.st));
    d.sparse_cap = (size_t)n * stride;
  }
  CUDA_CHECK(cudaMemsetAsync(hist, 0, 128 * 4, d.st));
  const int grid = (int)((n + HS_CHUNK - 1) / HS_CHUNK);
  sparse_hist_kernel<<<grid, HS_BLOCK, 0, d.st>>>(d.harena->occ, n, d.harena->entries, stride, (uint32_t)nparts, part_of, hist);
  uint32_t h[64];
  CUDA_CHECK(cudaMemcpyAsync(h, hist, 64 * 4, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  uint32_t starts[64] = {0};
  uint32_t run = 0;
  for (int p = 0; p < nparts; p++) { counts[p] = h[p]; starts[p] = run; run += h[p]; }
  CUDA_CHECK(cudaMemcpyAsync(cursor, starts, 64 * 4, cudaMemcpyHostToDevice, d.st));
  sparse_scatter_kernel<<<grid, HS_BLOCK, 0, d.st>>>(d.harena->occ, part_of, n, d.harena->entries, stride, (uint32_t)nparts, cursor, d.sparse_out);
  // the table is empty and clean again: foreign + own entries come back through lk_query_merge_sparse
  CUDA_CHECK(cudaMemsetAsync(d.counters + 3, 0, 4, d.st));
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  *entries_out = d.sparse_out;
}

void device_merge_sparse(Query& q, const void* dev_entries, int64_t n) {
  LK_CHECK(q.dev && q.dev->executed && !q.dev->finalized_device, LK_ERR_INVALID, "lk_query_merge_sparse needs an executed, not yet finalized query");
  LK_CHECK(q.path != 0, LK_ERR_INVALID, "query uses the dense path");
  LK_CHECK(n >= 0 && n < (int64_t)0xffffffffu, LK_ERR_INVALID, "bad entry count");
  if (n == 0) return;
  Query::Device& d = *q.dev;
  CUDA_CHECK(cudaSetDevice(global_options().device));
  LK_CHECK(q.path != 2, LK_ERR_INVALID, "the record path exchanges its records during the scan: attach an lk_comm (lk_query_set_comm)");
  AggSlot4 ops;
  memset(&ops, 0, sizeof ops);
  for (size_t a = 0; a < q.aggs.size(); a++) ops.op[a] = q.aggs[a].op;
  sparse_merge_kernel<<<(int)((n + HS_BLOCK - 1) / HS_BLOCK), HS_BLOCK, 0, d.st>>>(
      (const uint8_t*)dev_entries, (uint32_t)n, d.harena->entries, d.harena->stride, q.params.h_mask, d.harena->occ,
      d.harena->occ_bkt, q.n_groups, (uint32_t)std::min<uint64_t>(d.harena->slots, 0xffffffffu), d.counters, (int)q.aggs.size(), ops);
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// exact_sums: fixed-order (row order) double sums, bit-identical to a sequential evaluator
// ------------------------------------------------------------------------------------------------------------
// The atomics of the scan add in an arbitrary order (sums agree to ~1e-16 relative, not bitwise).  With exact_sums the
// scan is run a second time in record mode, the (cell, row sequence, value) records are sorted by (cell, sequence)
// -- two stable LSD radix sorts (CUB, a library call on this OPTIONAL slow path only) -- and one thread per cell folds
// its values strictly left to right, then overwrites the accumulator word in the table.
__global__ void iota_kernel(uint32_t* v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void gather_u64_kernel(const unsigned long long* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, unsigned long long* __restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
struct ExactPatch {
  int path, n_aggs;
  uint8_t ops[LK_MAX_AGGS + 1];
  const unsigned long long* rec_val[LK_MAX_AGGS];
  unsigned long long* acc[LK_MAX_AGGS];  // dense planes
  uint8_t* h_entries;
  uint32_t h_stride;
  uint64_t h_mask;
};
__global__ void exact_fold_kernel(const unsigned long long* __restrict__ sorted_cell, const uint32_t* __restrict__ order, uint32_t n,
                                  const __grid_constant__ ExactPatch X) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long cell = sorted_cell[i];
  if (i > 0 && sorted_cell[i - 1] == cell) return;  // not the head of its cell
  unsigned long long* entry = nullptr;
  if (X.path == 1) {
    uint64_t slot = lk_hash64(cell) & X.h_mask;
    for (int probe = 0; probe < 4096; probe++) {
      unsigned long long* e = reinterpret_cast<unsigned long long*>(X.h_entries + slot * X.h_stride);
      if (*e == cell + 1) { entry = e; break; }
      slot = (slot + 1) & X.h_mask;
    }
    if (!entry) return;
  }
  for (int a = 0; a < X.n_aggs; a++) {
    if (X.ops[a] != AGG_SUM) continue;
    double sum = 0.0;
    for (uint32_t j = i; j < n && sorted_cell[j] == cell; j++) sum += __longlong_as_double((long long)X.rec_val[a][order[j]]);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sum);
    if (X.path == 0) X.acc[a][cell] = bits;
    else entry[1 + a] = bits;
  }
}

static void sort_order_by_two_keys(cudaStream_t st, const unsigned long long* primary, const unsigned long long* secondary, uint32_t n,
                                   unsigned long long* key_a, unsigned long long* key_b, uint32_t* idx_a, uint32_t* idx_b, void* tmp, size_t tmp_bytes,
                                   const unsigned long long** sorted_primary, const uint32_t** order) {
  const int grid = (int)((n + 255) / 256);
  iota_kernel<<<grid, 256, 0, st>>>(idx_a, n);
  // pass 1: by the secondary key
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, secondary, key_a, idx_a, idx_b, (int)n, 0, 64, st));
  // pass 2: stable by the primary key
  gather_u64_kernel<<<grid, 256, 0, st>>>(primary, idx_b, n, key_a);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_a, key_b, idx_b, idx_a, (int)n, 0, 64, st));
  *sorted_primary = key_b;
  *order = idx_a;
}

static void device_exact_sums(Query& q, const ScanParams& base) {
  Query::Device& d = *q.dev;
  bool any_sum = false;
  for (auto& a : q.aggs) any_sum |= a.op == AGG_SUM;
  if (!any_sum || q.n_cells == 0 || q.tiles.empty()) return;
  const int64_t nsurv = device_survivors(q);
  if (nsurv <= 0) return;
  LK_CHECK(nsurv < (int64_t)0x7fffffff, LK_ERR_UNSUPPORTED, "exact_sums: too many surviving rows for one pass");
  const uint32_t n = (uint32_t)nsurv;
  const size_t na = q.aggs.size();
  // scratch: rec_cell | rec_seq | rec_val[na] | key_a | key_b (u64 each) | idx_a | idx_b (u32 each) | counters2 | cub temp
  size_t tmp_bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, (int)n, 0, 64, d.st));
  const size_t words = (size_t)n * (4 + na);
  uint8_t* scratch = nullptr;
  const size_t bytes = words * 8 + (size_t)n * 8 + 64 + tmp_bytes + 256;
  CUDA_CHECK(cudaMallocAsync(&scratch, bytes, d.st));
  unsigned long long* rec_cell = (unsigned long long*)scratch;
  unsigned long long* rec_seq = rec_cell + n;
  unsigned long long* rec_val0 = rec_seq + n;
  unsigned long long* key_a = rec_val0 + (size_t)n * na;
  unsigned long long* key_b = key_a + n;
  uint32_t* idx_a = (uint32_t*)(key_b + n);
  uint32_t* idx_b = idx_a + n;
  uint32_t* counters2 = idx_b + n;
  void* tmp = (void*)(((uintptr_t)(counters2 + 16) + 255) & ~(uintptr_t)255);
  ScanParams P = base;
  P.emit_records = 1;
  P.rec_cap = n;
  P.rec_cell = rec_cell;
  P.rec_seq = rec_seq;
  for (size_t a = 0; a < na; a++) P.rec_val[a] = rec_val0 + a * n;
  P.counters = counters2;
  P.survivors = (unsigned long long*)(counters2 + 8);
  static const uint32_t init_counters[16] = {0, 0xffffffffu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  CUDA_CHECK(cudaMemcpyAsync(counters2, init_counters, sizeof init_counters, cudaMemcpyHostToDevice, d.st));
  launch_scan(q, P, true);
  const unsigned long long* sorted_cell = nullptr;
  const uint32_t* order = nullptr;
  sort_order_by_two_keys(d.st, rec_cell, rec_seq, n, key_a, key_b, idx_a, idx_b, tmp, tmp_bytes, &sorted_cell, &order);
  ExactPatch X;
  memset(&X, 0, sizeof X);
  X.path = q.path;
  X.n_aggs = (int)na;
  for (size_t a = 0; a < na; a++) {
    X.ops[a] = q.aggs[a].op;
    X.rec_val[a] = P.rec_val[a];
    X.acc[a] = base.acc[a];
  }
  X.h_entries = base.h_entries;
  X.h_stride = base.h_stride;
  X.h_mask = base.h_mask;
  exact_fold_kernel<<<(int)((n + 255) / 256), 256, 0, d.st>>>(sorted_cell, order, n, X);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaFreeAsync(scratch, d.st));
}

// ---- exact_sums on the record path (single GPU and sharded): every record row starts with its global sequence number
// (segment order, then row order; sharded: + the shard's offset).  The records of this rank -- its own list, or the regions
// its sources filled -- are ordered by (key, sequence) with two stable radix sorts (CUB, this optional pass only), and one
// thread per cell folds its records strictly in that order and writes the row: the sums are bit-identical to a sequential
// evaluator's and do not depend on which rank scanned which segment (TimeGroupedSketchAggregator.scala:76-78 adds partials in
// arrival order; this is the "fixed-order reduction option" of the north star).
__global__ void exact_gather_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ rows, uint32_t n, int row_words,
                                    const __grid_constant__ RecGeom G, unsigned long long* __restrict__ cell, unsigned long long* __restrict__ seq,
                                    uint32_t* __restrict__ phys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t pi = rec_phys(i, G);
  cell[i] = keys[pi];
  seq[i] = rows[(size_t)pi * row_words];
  phys[i] = pi;
}
__global__ void exact_heads_kernel(const unsigned long long* __restrict__ sorted_cell, uint32_t n, uint32_t* __restrict__ block_counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = i < n && (i == 0 || sorted_cell[i] != sorted_cell[i - 1]);
  const int c = __syncthreads_count(head);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(256) exact_emit_kernel(const unsigned long long* __restrict__ sorted_cell, const uint32_t* __restrict__ order,
                                                         const uint32_t* __restrict__ phys, uint32_t n, const unsigned long long* __restrict__ rows, int row_words,
                                                         const uint32_t* __restrict__ block_offsets, const __grid_constant__ RecGeom G,
                                                         const __grid_constant__ EmitParams E) {
  __shared__ uint32_t warp_sums[8];
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned long long cell = i < n ? sorted_cell[i] : 0ull;
  const bool head = i < n && (i == 0 || cell != sorted_cell[i - 1]);
  const unsigned hm = __ballot_sync(0xffffffffu, head);
  if (lane == 0) warp_sums[wid] = (uint32_t)__popc(hm);
  __syncthreads();
  if (!head) return;
  uint32_t out = block_offsets[blockIdx.x] + (uint32_t)__popc(hm & ((1u << lane) - 1));
  for (int w = 0; w < wid; w++) out += warp_sums[w];
  unsigned long long acc[LK_MAX_AGGS];
#pragma unroll
  for (int a = 0; a < LK_MAX_AGGS; a++) acc[a] = 0;
  for (uint32_t j = i; j < n && sorted_cell[j] == cell; j++) {  // strictly in (key, sequence) order
    const unsigned long long* rec = rows + (size_t)phys[order[j]] * row_words + 1;
    for (int a = 0; a < E.n_aggs; a++) {
      const unsigned long long w = rec[a];
      if (E.ops[a] == AGG_SUM) acc[a] = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc[a]) + __longlong_as_double((long long)w));
      else if (E.ops[a] == AGG_COUNT) acc[a] += w;
      else acc[a] = w > acc[a] ? w : acc[a];  // min (complemented key) and max (key): both stored as "max"
    }
  }
  emit_row_bg(E, out, cell >> G.gid_bits, cell & ((1ull << G.gid_bits) - 1), [&](int a) { return acc[a]; });
}

static void rec_exact_finalize(Query& q) {
  Query::Device& d = *q.dev;
  CUDA_CHECK(cudaMemcpyAsync(d.h_counters, d.counters, sizeof d.h_counters, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  check_scan_status(q, d.h_counters);
  const uint32_t n = (uint32_t)std::min<size_t>(d.h_counters[5], d.rec_cap);
  d.n_rows = 0;
  d.dres_stride = 0;
  if (n == 0) return;
  RecGeom G;
  memset(&G, 0, sizeof G);
  G.idx_bits = q.params.rec_idx_bits;
  G.gid_bits = q.params.rec_gid_bits;
  G.nbuckets = q.nbuckets;
  G.world = q.comm ? (uint32_t)q.comm->world : 1u;
  G.region_cap = q.comm ? (uint32_t)q.comm->region_cap : 0u;
  G.prefix = q.comm ? q.comm->ctrl(q.comm->rank)->prefix : nullptr;
  const int row_words = (int)q.aggs.size() + 1;
  size_t tmp_bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, (int)n, 0, 64, d.st));
  // scratch: cell | seq | key_a | key_b (u64 each) | phys | idx_a | idx_b | block counts (u32 each) | cub temp
  const uint32_t nblocks = (n + 255) / 256;
  uint8_t* scratch = nullptr;
  const size_t bytes = (size_t)n * 8 * 4 + (size_t)n * 4 * 3 + ((size_t)nblocks + 2) * 4 + tmp_bytes + 1024;
  CUDA_CHECK(cudaMallocAsync(&scratch, bytes, d.st));
  unsigned long long* cell = (unsigned long long*)scratch;
  unsigned long long* seq = cell + n;
  unsigned long long* key_a = seq + n;
  unsigned long long* key_b = key_a + n;
  uint32_t* phys = (uint32_t*)(key_b + n);
  uint32_t* idx_a = phys + n;
  uint32_t* idx_b = idx_a + n;
  uint32_t* counts = idx_b + n;
  void* tmp = (void*)(((uintptr_t)(counts + nblocks + 2) + 255) & ~(uintptr_t)255);
  exact_gather_kernel<<<nblocks, 256, 0, d.st>>>(d.rec_cell, d.rec_vals, n, row_words, G, cell, seq, phys);
  const unsigned long long* sorted_cell = nullptr;
  const uint32_t* order = nullptr;
  sort_order_by_two_keys(d.st, cell, seq, n, key_a, key_b, idx_a, idx_b, tmp, tmp_bytes, &sorted_cell, &order);
  exact_heads_kernel<<<nblocks, 256, 0, d.st>>>(sorted_cell, n, counts);
  exclusive_scan_kernel<<<1, 1024, 0, d.st>>>(counts, nblocks, counts + nblocks);
  CUDA_CHECK(cudaGetLastError());
  uint32_t n_rows = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_rows, counts + nblocks, 4, cudaMemcpyDeviceToHost, d.st));
  CUDA_CHECK(cudaStreamSynchronize(d.st));
  d.n_rows = n_rows;
  d.dres_stride = n_rows;
  ensure_dres(d, result_bytes(q, n_rows));
  EmitParams E;
  fill_emit_params(q, E, d.dres, n_rows);
  exact_emit_kernel<<<nblocks, 256, 0, d.st>>>(sorted_cell, order, phys, n, d.rec_vals, row_words, counts, G, E);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaFreeAsync(scratch, d.st));
}

int64_t device_survivors(Query& q) {
  unsigned long long v = 0;
  CUDA_CHECK(cudaMemcpyAsync(&v, q.dev->survivors, 8, cudaMemcpyDeviceToHost, q.dev->st));
  CUDA_CHECK(cudaStreamSynchronize(q.dev->st));
  return (int64_t)v;
}

HostResult::~HostResult() { pinned_free(pinned); }

}  // namespace lk

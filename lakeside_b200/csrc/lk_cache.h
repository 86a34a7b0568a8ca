// HBM-resident segment cache (SURVEY.md §8b "Ownership"): the column chunks of sealed segments a query has uploaded stay in
// device memory, keyed by the segment file's identity (path + size + mtime + inode), together with what the host learned
// from the file (footer, page headers, dictionaries), so that the next query over the same segments neither reads the file
// nor crosses PCIe.  The analogue of the worker's Caffeine cache of downloaded segment files
// (query-worker/src/main/scala/com/cardinal/queryworker/WorkerApi.scala:53-64) one level further down the memory hierarchy.
//
// Unit of caching: one column of one file (all its row groups' chunks in one device block).  Entries are shared_ptrs: a
// query pins what it uses, eviction (least recently used segment first) only unlinks, the block is freed when the last
// query that uses it is gone.  All methods are thread-safe.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "lk_parquet.h"

namespace lk {

struct CachedColumn {
  void* dev = nullptr;                // device block (stream-ordered allocation)
  size_t bytes = 0;
  std::vector<uint64_t> chunk_off;    // per row group of the file: offset of its chunk inside the block (~0: none)
  std::vector<ChunkIndex> index;      // per row group: the host's page / dictionary index of the chunk
  ~CachedColumn();
};

struct SegmentIdentity {
  std::string path;
  uint64_t size = 0, mtime_ns = 0, ino = 0;
  bool operator==(const SegmentIdentity& o) const { return path == o.path && size == o.size && mtime_ns == o.mtime_ns && ino == o.ino; }
};

struct CachedSegment {
  SegmentIdentity id;
  FileMeta meta;
  std::map<int, std::shared_ptr<CachedColumn>> cols;  // by leaf index; guarded by the cache's mutex
  size_t bytes = 0;
  uint64_t last_use = 0;
};

struct SegmentCacheStats {
  int64_t capacity_bytes, resident_bytes, segments, column_hits, column_misses, evicted_segments;
};

class SegmentCache {
 public:
  void set_capacity(size_t bytes);
  size_t capacity();
  // the entry of this file, or null (a stale entry -- same path, other size / mtime / inode -- is dropped)
  std::shared_ptr<CachedSegment> lookup(const SegmentIdentity& id);
  std::shared_ptr<CachedColumn> column(const std::shared_ptr<CachedSegment>& seg, int leaf);
  // makes `col` (uploaded, complete) available to later queries; evicts least recently used segments to make room and
  // returns false when the column does not fit the cache at all (it then lives and dies with the query that built it)
  bool publish(const SegmentIdentity& id, const FileMeta& meta, int leaf, const std::shared_ptr<CachedColumn>& col);
  void clear();
  SegmentCacheStats stats();

 private:
  void evict_for(size_t need, const CachedSegment* keep);  // caller holds mu_
  std::mutex mu_;
  std::map<std::string, std::shared_ptr<CachedSegment>> segs_;
  size_t capacity_ = 0, bytes_ = 0;
  uint64_t tick_ = 0;
  int64_t hits_ = 0, misses_ = 0, evicted_ = 0;
};

SegmentCache& segment_cache();
// device blocks of evicted columns are released on this stream (lk_engine.cu)
void cache_free_device(void* p);

}  // namespace lk

// HBM-resident segment cache: bookkeeping (host only; the device blocks are allocated and filled by lk_engine.cu).
#include "lk_cache.h"

namespace lk {

CachedColumn::~CachedColumn() {
  if (dev) cache_free_device(dev);
}

SegmentCache& segment_cache() {
  static SegmentCache c;
  return c;
}

void SegmentCache::set_capacity(size_t bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  capacity_ = bytes;
  if (bytes_ > capacity_) evict_for(0, nullptr);
}

size_t SegmentCache::capacity() {
  std::lock_guard<std::mutex> lk(mu_);
  return capacity_;
}

std::shared_ptr<CachedSegment> SegmentCache::lookup(const SegmentIdentity& id) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!capacity_) return nullptr;
  auto it = segs_.find(id.path);
  if (it == segs_.end()) return nullptr;
  if (!(it->second->id == id)) {  // the file was replaced: what is cached describes other bytes
    bytes_ -= it->second->bytes;
    segs_.erase(it);
    return nullptr;
  }
  it->second->last_use = ++tick_;
  return it->second;
}

std::shared_ptr<CachedColumn> SegmentCache::column(const std::shared_ptr<CachedSegment>& seg, int leaf) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = seg->cols.find(leaf);
  if (it == seg->cols.end()) { misses_++; return nullptr; }
  hits_++;
  return it->second;
}

void SegmentCache::evict_for(size_t need, const CachedSegment* keep) {
  while (bytes_ + need > capacity_) {
    auto victim = segs_.end();
    for (auto it = segs_.begin(); it != segs_.end(); ++it)
      if (it->second.get() != keep && (victim == segs_.end() || it->second->last_use < victim->second->last_use)) victim = it;
    if (victim == segs_.end()) break;
    bytes_ -= victim->second->bytes;
    segs_.erase(victim);  // queries that still use its columns keep them alive through their own references
    evicted_++;
  }
}

bool SegmentCache::publish(const SegmentIdentity& id, const FileMeta& meta, int leaf, const std::shared_ptr<CachedColumn>& col) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!capacity_ || col->bytes > capacity_) return false;
  auto it = segs_.find(id.path);
  if (it != segs_.end() && !(it->second->id == id)) {
    bytes_ -= it->second->bytes;
    segs_.erase(it);
    it = segs_.end();
  }
  if (it != segs_.end() && it->second->cols.count(leaf)) return true;  // another query was faster; ours stays private
  // make room first, never at the expense of the segment being extended
  evict_for(col->bytes, it != segs_.end() ? it->second.get() : nullptr);
  if (bytes_ + col->bytes > capacity_) return false;
  std::shared_ptr<CachedSegment> seg;
  if (it == segs_.end()) {
    seg = std::make_shared<CachedSegment>();
    seg->id = id;
    seg->meta = meta;
    segs_[id.path] = seg;
  } else seg = it->second;
  seg->cols[leaf] = col;
  seg->bytes += col->bytes;
  seg->last_use = ++tick_;
  bytes_ += col->bytes;
  return true;
}

void SegmentCache::clear() {
  std::lock_guard<std::mutex> lk(mu_);
  segs_.clear();
  bytes_ = 0;
}

SegmentCacheStats SegmentCache::stats() {
  std::lock_guard<std::mutex> lk(mu_);
  return SegmentCacheStats{(int64_t)capacity_, (int64_t)bytes_, (int64_t)segs_.size(), hits_, misses_, evicted_};
}

}  // namespace lk

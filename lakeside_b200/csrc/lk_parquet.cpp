#include "lk_parquet.h"

#include <unordered_map>

#include <algorithm>
#include <cstring>

namespace lk {

// ------------------------------------------------------------------------------------------------------------
// Thrift compact protocol reader (just enough for parquet.thrift's FileMetaData and PageHeader)
// ------------------------------------------------------------------------------------------------------------
namespace {

enum TType { T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_BYTE = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6, T_DOUBLE = 7, T_BINARY = 8, T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12 };

struct TReader {
  const uint8_t* p;
  const uint8_t* end;
  TReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  uint8_t byte() {
    LK_CHECK(p < end, LK_ERR_IO, "parquet: truncated thrift data");
    return *p++;
  }
  uint64_t varint() {
    uint64_t r = 0;
    int sh = 0;
    while (true) {
      uint8_t c = byte();
      r |= (uint64_t)(c & 0x7f) << sh;
      if (!(c & 0x80)) return r;
      sh += 7;
      LK_CHECK(sh < 70, LK_ERR_IO, "parquet: bad varint");
    }
  }
  int64_t zigzag() {
    uint64_t v = varint();
    return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
  }
  std::string binary() {
    uint64_t n = varint();
    LK_CHECK(n <= (uint64_t)(end - p), LK_ERR_IO, "parquet: truncated thrift binary");
    std::string s((const char*)p, n);
    p += n;
    return s;
  }
  void skip_binary() {
    uint64_t n = varint();
    LK_CHECK(n <= (uint64_t)(end - p), LK_ERR_IO, "parquet: truncated thrift binary");
    p += n;
  }
  // returns false at STOP; otherwise sets fid/type
  bool field(int& fid, int& type, int& last) {
    uint8_t h = byte();
    if (h == 0) return false;
    type = h & 0x0f;
    int delta = h >> 4;
    fid = delta ? last + delta : (int)zigzag();
    last = fid;
    return true;
  }
  void list_header(int& n, int& et) {
    uint8_t h = byte();
    n = h >> 4;
    et = h & 0x0f;
    if (n == 15) n = (int)varint();
  }
  void skip(int type, int depth = 0) {
    LK_CHECK(depth < 32, LK_ERR_IO, "parquet: thrift nesting too deep");
    switch (type) {
      case T_TRUE: case T_FALSE: break;
      case T_BYTE: byte(); break;
      case T_I16: case T_I32: case T_I64: varint(); break;
      case T_DOUBLE: LK_CHECK(end - p >= 8, LK_ERR_IO, "parquet: truncated"); p += 8; break;
      case T_BINARY: skip_binary(); break;
      case T_LIST: case T_SET: {
        int n, et;
        list_header(n, et);
        for (int i = 0; i < n; i++) {
          if (et == T_TRUE || et == T_FALSE) byte();  // list<bool>: one byte per element
          else skip(et, depth + 1);
        }
        break;
      }
      case T_MAP: {
        uint64_t n = varint();
        if (n) {
          uint8_t kv = byte();
          for (uint64_t i = 0; i < n; i++) { skip(kv >> 4, depth + 1); skip(kv & 0x0f, depth + 1); }
        }
        break;
      }
      case T_STRUCT: {
        int fid, t, last = 0;
        while (field(fid, t, last)) skip(t, depth + 1);
        break;
      }
      default: fail(LK_ERR_IO, "parquet: unknown thrift type");
    }
  }
};

struct SchemaElem {
  int type = -1, repetition = 0, num_children = 0;
  std::string name;
};

SchemaElem read_schema_elem(TReader& r) {
  SchemaElem s;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    switch (fid) {
      case 1: s.type = (int)r.zigzag(); break;
      case 3: s.repetition = (int)r.zigzag(); break;
      case 4: s.name = r.binary(); break;
      case 5: s.num_children = (int)r.zigzag(); break;
      default: r.skip(t);
    }
  }
  return s;
}

ColumnChunkMeta read_column_meta(TReader& r) {
  ColumnChunkMeta m;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    switch (fid) {
      case 1: m.phys_type = (int)r.zigzag(); break;
      case 4: m.codec = (int)r.zigzag(); break;
      case 5: m.num_values = r.zigzag(); break;
      case 6: m.total_uncompressed_size = r.zigzag(); break;
      case 7: m.total_compressed_size = r.zigzag(); break;
      case 9: m.data_page_offset = r.zigzag(); break;
      case 11: m.dictionary_page_offset = r.zigzag(); break;
      case 13: {  // encoding_stats: list<PageEncodingStats {1: page_type, 2: encoding, 3: count}>
        int n, et;
        r.list_header(n, et);
        m.plain_data_pages = 0;
        for (int i = 0; i < n; i++) {
          int f2, t2, l2 = 0, page_type = -1, enc = -1;
          int64_t count = 0;
          while (r.field(f2, t2, l2)) {
            if (f2 == 1) page_type = (int)r.zigzag();
            else if (f2 == 2) enc = (int)r.zigzag();
            else if (f2 == 3) count = r.zigzag();
            else r.skip(t2);
          }
          if ((page_type == 0 || page_type == 3) && enc == ENC_PLAIN) m.plain_data_pages += (int)std::min<int64_t>(count, 1 << 20);
        }
        break;
      }
      default: r.skip(t);
    }
  }
  return m;
}

ColumnChunkMeta read_column_chunk(TReader& r) {
  ColumnChunkMeta m;
  bool have = false;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    if (fid == 3 && t == T_STRUCT) { m = read_column_meta(r); have = true; }
    else r.skip(t);
  }
  LK_CHECK(have, LK_ERR_UNSUPPORTED, "parquet: column chunk without inline meta_data");
  return m;
}

RowGroupMeta read_row_group(TReader& r) {
  RowGroupMeta g;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    if (fid == 1 && t == T_LIST) {
      int n, et;
      r.list_header(n, et);
      for (int i = 0; i < n; i++) g.columns.push_back(read_column_chunk(r));
    } else if (fid == 3) g.num_rows = r.zigzag();
    else r.skip(t);
  }
  return g;
}

}  // namespace

int FileMeta::leaf_index(const std::string& name) const {
  for (size_t i = 0; i < leaves.size(); i++)
    if (leaves[i].flat && leaves[i].name == name) return (int)i;
  return -1;
}

static FileMeta parse_footer_thrift(const uint8_t* begin, const uint8_t* end);

FileMeta parse_footer(const uint8_t* data, size_t len) {
  LK_CHECK(len >= 12 && memcmp(data, "PAR1", 4) == 0 && memcmp(data + len - 4, "PAR1", 4) == 0, LK_ERR_IO,
           "not a Parquet file (bad magic)");
  uint32_t flen;
  memcpy(&flen, data + len - 8, 4);
  LK_CHECK((size_t)flen + 12 <= len, LK_ERR_IO, "parquet: bad footer length");
  return parse_footer_thrift(data + len - 8 - flen, data + len - 8);
}

bool parse_footer_tail(const uint8_t* tail, size_t tail_len, size_t file_len, FileMeta* out, size_t* need) {
  LK_CHECK(file_len >= 12 && tail_len >= 8 && tail_len <= file_len && memcmp(tail + tail_len - 4, "PAR1", 4) == 0, LK_ERR_IO,
           "not a Parquet file (bad magic)");
  uint32_t flen;
  memcpy(&flen, tail + tail_len - 8, 4);
  LK_CHECK((size_t)flen + 12 <= file_len, LK_ERR_IO, "parquet: bad footer length");
  if ((size_t)flen + 8 > tail_len) { *need = (size_t)flen + 8; return false; }
  *out = parse_footer_thrift(tail + tail_len - 8 - flen, tail + tail_len - 8);
  return true;
}

static FileMeta parse_footer_thrift(const uint8_t* begin, const uint8_t* end) {
  TReader r(begin, end);
  FileMeta fm;
  std::vector<SchemaElem> schema;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    if (fid == 2 && t == T_LIST) {
      int n, et;
      r.list_header(n, et);
      for (int i = 0; i < n; i++) schema.push_back(read_schema_elem(r));
    } else if (fid == 3) fm.num_rows = r.zigzag();
    else if (fid == 4 && t == T_LIST) {
      int n, et;
      r.list_header(n, et);
      for (int i = 0; i < n; i++) fm.row_groups.push_back(read_row_group(r));
    } else r.skip(t);
  }
  LK_CHECK(!schema.empty(), LK_ERR_IO, "parquet: empty schema");
  // depth-first walk: leaves in schema order; only direct children of the root are addressable by name
  size_t pos = 1;
  struct Frame { int remaining; int depth; };
  std::vector<Frame> st;
  st.push_back({schema[0].num_children, 0});
  while (pos < schema.size() && !st.empty()) {
    while (!st.empty() && st.back().remaining == 0) st.pop_back();
    if (st.empty()) break;
    st.back().remaining--;
    int depth = (int)st.size();
    const SchemaElem& e = schema[pos++];
    if (e.num_children > 0) { st.push_back({e.num_children, depth}); continue; }
    LeafColumn lc;
    lc.name = e.name;
    lc.phys_type = e.type;
    lc.flat = depth == 1 && e.repetition != 2;
    lc.max_def = e.repetition == 1 ? 1 : 0;
    fm.leaves.push_back(lc);
  }
  for (auto& g : fm.row_groups)
    LK_CHECK(g.columns.size() == fm.leaves.size(), LK_ERR_IO, "parquet: row group column count != schema leaves");
  return fm;
}

// ------------------------------------------------------------------------------------------------------------
// pages and runs
// ------------------------------------------------------------------------------------------------------------
namespace {

struct PageHeader {
  int type = -1;
  int32_t uncompressed = 0, compressed = 0;
  int32_t num_values = 0;
  int encoding = -1;
  int def_encoding = ENC_RLE;
  // v2
  int32_t num_nulls = 0, num_rows = 0, def_len = 0, rep_len = 0;
  bool v2_compressed = true;  // DataPageHeaderV2.is_compressed (values section only; levels are never compressed)
  size_t header_len = 0;
};

PageHeader read_page_header(const uint8_t* p, const uint8_t* end) {
  TReader r(p, end);
  PageHeader h;
  int fid, t, last = 0;
  while (r.field(fid, t, last)) {
    switch (fid) {
      case 1: h.type = (int)r.zigzag(); break;
      case 2: h.uncompressed = (int32_t)r.zigzag(); break;
      case 3: h.compressed = (int32_t)r.zigzag(); break;
      case 5: case 7: {  // DataPageHeader / DictionaryPageHeader
        int f2, t2, l2 = 0;
        while (r.field(f2, t2, l2)) {
          if (f2 == 1) h.num_values = (int32_t)r.zigzag();
          else if (f2 == 2) h.encoding = (int)r.zigzag();
          else if (f2 == 3 && fid == 5) h.def_encoding = (int)r.zigzag();
          else r.skip(t2);
        }
        break;
      }
      case 8: {  // DataPageHeaderV2
        int f2, t2, l2 = 0;
        while (r.field(f2, t2, l2)) {
          switch (f2) {
            case 1: h.num_values = (int32_t)r.zigzag(); break;
            case 2: h.num_nulls = (int32_t)r.zigzag(); break;
            case 3: h.num_rows = (int32_t)r.zigzag(); break;
            case 4: h.encoding = (int)r.zigzag(); break;
            case 5: h.def_len = (int32_t)r.zigzag(); break;
            case 6: h.rep_len = (int32_t)r.zigzag(); break;
            case 7: h.v2_compressed = t2 == T_TRUE; break;
            default: r.skip(t2);
          }
        }
        break;
      }
      default: r.skip(t);
    }
  }
  h.header_len = (size_t)(r.p - p);
  return h;
}

inline uint32_t popcount_bits(const uint8_t* p, uint32_t nbits) {
  uint32_t c = 0;
  uint32_t nbytes = nbits >> 3;
  uint32_t i = 0;
  for (; i + 8 <= nbytes; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    c += (uint32_t)__builtin_popcountll(w);
  }
  for (; i < nbytes; i++) c += (uint32_t)__builtin_popcount(p[i]);
  if (nbits & 7) c += (uint32_t)__builtin_popcount(p[nbytes] & ((1u << (nbits & 7)) - 1));
  return c;
}

// Walks the runs of an RLE/bit-packed hybrid stream holding `count` elements.  cb(start, n, is_rle, value, byte_off)
template <class CB>
void walk_hybrid(const uint8_t* file, uint64_t off, uint64_t end, int bit_width, uint32_t count, CB cb) {
  uint32_t done = 0;
  const int vbytes = (bit_width + 7) / 8;
  while (done < count) {
    LK_CHECK(off < end, LK_ERR_IO, "parquet: hybrid stream ends before all values are covered");
    TReader r(file + off, file + end);
    uint64_t h = r.varint();
    off = (uint64_t)(r.p - file);
    if (h & 1) {
      uint64_t groups = h >> 1;
      uint64_t n = groups * 8;
      uint64_t bytes = groups * (uint64_t)bit_width;
      LK_CHECK(groups > 0, LK_ERR_IO, "parquet: empty bit-packed run");
      // a writer may truncate the padding of the final group; require the bytes that carry real values
      uint32_t take = (uint32_t)std::min<uint64_t>(n, count - done);
      uint64_t need = ((uint64_t)take * bit_width + 7) / 8;
      LK_CHECK(off + need <= end, LK_ERR_IO, "parquet: truncated bit-packed run");
      cb(done, take, false, 0u, off);
      off += std::min<uint64_t>(bytes, end - off);
      done += take;
    } else {
      uint64_t n = h >> 1;
      LK_CHECK(n > 0, LK_ERR_IO, "parquet: empty RLE run");
      LK_CHECK(off + vbytes <= end, LK_ERR_IO, "parquet: truncated RLE run");
      uint32_t v = 0;
      for (int i = 0; i < vbytes; i++) v |= (uint32_t)file[off + i] << (8 * i);
      off += vbytes;
      uint32_t take = (uint32_t)std::min<uint64_t>(n, count - done);
      cb(done, take, true, v, (uint64_t)0);
      done += take;
    }
  }
}

}  // namespace

int ChunkIndex::page_at(uint32_t r) const {
  int lo = 0, hi = (int)pages.size();
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (pages[mid].first_row <= r) lo = mid; else hi = mid;
  }
  return lo;
}

int ChunkIndex::def_run_at(uint32_t r) const {
  int lo = 0, hi = (int)def_runs.size();
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (def_runs[mid].start <= r) lo = mid; else hi = mid;
  }
  return lo;
}

int ChunkIndex::val_run_at(uint32_t v) const {
  int lo = 0, hi = (int)val_runs.size();
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (val_runs[mid].start <= v) lo = mid; else hi = mid;
  }
  return lo;
}

uint32_t ChunkIndex::vidx_at(const uint8_t* file, uint32_t r) const {
  if (max_def == 0) return r;
  if (def_runs.empty()) return 0;
  return vidx_in_run(file, def_run_at(r), r);
}

uint32_t ChunkIndex::vidx_in_run(const uint8_t* file, int i, uint32_t r) const {
  const Run& run = def_runs[i];
  uint32_t nn = def_nn_before[i];
  uint32_t k = r - run.start;
  if (run.kind_value >> 31) return nn + ((run.kind_value & 1) ? k : 0);
  return nn + popcount_bits(file + file_start + run.kind_value, k);
}

bool snappy_uncompress(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  size_t ip = 0;
  uint64_t ulen = 0;
  for (int sh = 0;; sh += 7) {
    if (ip >= n || sh > 35) return false;
    const uint8_t c = src[ip++];
    ulen |= (uint64_t)(c & 0x7f) << sh;
    if (!(c & 0x80)) break;
  }
  if (ulen > (1ull << 31)) return false;
  out.assign((size_t)ulen, 0);
  size_t op = 0;
  while (ip < n) {
    const uint8_t tag = src[ip++];
    size_t len, off = 0;
    if ((tag & 3) == 0) {
      len = (size_t)(tag >> 2) + 1;
      if (len > 60) {
        const size_t nb = len - 60;
        if (ip + nb > n) return false;
        len = 0;
        for (size_t i = 0; i < nb; i++) len |= (size_t)src[ip + i] << (8 * i);
        len += 1;
        ip += nb;
      }
      if (ip + len > n || op + len > ulen) return false;
      memcpy(out.data() + op, src + ip, len);
      ip += len;
      op += len;
      continue;
    }
    if ((tag & 3) == 1) {
      if (ip + 1 > n) return false;
      len = 4 + ((tag >> 2) & 7);
      off = ((size_t)(tag >> 5) << 8) | src[ip];
      ip += 1;
    } else if ((tag & 3) == 2) {
      if (ip + 2 > n) return false;
      len = (size_t)(tag >> 2) + 1;
      off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
      ip += 2;
    } else {
      if (ip + 4 > n) return false;
      len = (size_t)(tag >> 2) + 1;
      off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8) | ((size_t)src[ip + 2] << 16) | ((size_t)src[ip + 3] << 24);
      ip += 4;
    }
    if (off == 0 || off > op || op + len > ulen) return false;
    for (size_t i = 0; i < len; i++) out[op + i] = out[op - off + i];  // byte by byte: the ranges may overlap (run-length patterns)
    op += len;
  }
  return op == ulen;
}

uint64_t synth_reserve(const ColumnChunkMeta& cm) {
  // compressed chunk: room for its inflated pages (the footer's total_uncompressed_size also counts the page headers)
  if (cm.codec == CODEC_SNAPPY) return (uint64_t)std::max<int64_t>(cm.total_uncompressed_size, 0) + 64;
  if (cm.phys_type != PT_BYTE_ARRAY || cm.plain_data_pages == 0 || cm.num_values <= 0) return 0;
  // one bit-packed run per page: at most 4 bytes per value + header and group padding per page (a page holds >= 1 value)
  const uint64_t pages = cm.plain_data_pages > 0 ? (uint64_t)cm.plain_data_pages : (uint64_t)(cm.total_compressed_size >> 10) + 2;
  return 4ull * (uint64_t)cm.num_values + 48ull * pages + 64;
}

void chunk_byte_range(const ColumnChunkMeta& cm, size_t file_len, const std::string& name, uint64_t& start_out, uint64_t& len_out) {
  int64_t start = cm.data_page_offset;
  if (cm.dictionary_page_offset > 0 && cm.dictionary_page_offset < start) start = cm.dictionary_page_offset;
  LK_CHECK(start >= 4 && cm.total_compressed_size >= 0 && (uint64_t)start + (uint64_t)cm.total_compressed_size <= file_len, LK_ERR_IO,
           "parquet: column chunk '" + name + "' out of file bounds");
  LK_CHECK((uint64_t)cm.total_compressed_size < (1ull << 31), LK_ERR_UNSUPPORTED, "column chunk '" + name + "' larger than 2 GiB");
  start_out = (uint64_t)start;
  len_out = (uint64_t)cm.total_compressed_size;
}

ChunkIndex index_chunk(const uint8_t* data, size_t len, const LeafColumn& leaf, const ColumnChunkMeta& cm, int64_t rg_rows,
                       bool want_strings, bool walk_runs) {
  ChunkIndex ci;
  ci.present = true;
  ci.phys_type = cm.phys_type;
  ci.max_def = leaf.max_def;
  ci.total_compressed_size = cm.total_compressed_size;
  LK_CHECK(cm.codec == CODEC_NONE || cm.codec == CODEC_SNAPPY, LK_ERR_UNSUPPORTED,
           "column '" + leaf.name + "': compression codec " + std::to_string(cm.codec) + " is not supported (UNCOMPRESSED and SNAPPY are)");
  const bool snappy = cm.codec == CODEC_SNAPPY;
  LK_CHECK(!snappy || !walk_runs, LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': SNAPPY pages are inflated on the device and need the device-built index");
  LK_CHECK(rg_rows >= 0 && rg_rows < (int64_t)0xfffffff0u, LK_ERR_UNSUPPORTED, "row group too large");
  ci.num_rows = (uint32_t)rg_rows;
  LK_CHECK(cm.num_values == rg_rows, LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': repeated values are not supported");
  chunk_byte_range(cm, len, leaf.name, ci.file_start, ci.file_len);
  const uint64_t cend = ci.file_start + ci.file_len;
  uint64_t p = ci.file_start;
  uint32_t row = 0, vidx = 0;
  std::unordered_map<std::string, uint32_t> synth_lookup;  // string -> dictionary index, once a PLAIN string page shows up
  while (p < cend && row < ci.num_rows) {
    PageHeader h = read_page_header(data + p, data + cend);
    uint64_t payload = p + h.header_len;
    LK_CHECK(h.compressed >= 0 && payload + (uint64_t)h.compressed <= cend, LK_ERR_IO, "parquet: page out of chunk bounds");
    LK_CHECK(snappy || h.compressed == h.uncompressed, LK_ERR_UNSUPPORTED, "parquet: compressed page in an UNCOMPRESSED chunk");
    uint64_t pend = payload + (uint64_t)h.compressed;
    if (snappy && !((h.type == 3 && !h.v2_compressed))) {
      // ---- SNAPPY page: recorded for the device (see ChunkIndex::zpages); only a string dictionary is inflated here ----
      LK_CHECK(h.uncompressed >= 0, LK_ERR_IO, "parquet: negative page size");
      if (ci.zpages.empty() && ci.z_base == 0) ci.z_base = (ci.file_len + 7) & ~7ull;
      auto reserve_z = [&](uint32_t n) {
        const uint64_t virt = ci.file_start + ci.z_base + ci.z_len;
        ci.z_len += n;
        LK_CHECK(ci.z_len + 64 <= synth_reserve(cm), LK_ERR_IO, "parquet: pages inflate to more than the footer's total_uncompressed_size");
        return virt;
      };
      if (h.type == 2) {
        LK_CHECK(h.encoding == ENC_PLAIN || h.encoding == ENC_PLAIN_DICTIONARY, LK_ERR_UNSUPPORTED, "parquet: dictionary page encoding");
        LK_CHECK(!ci.has_dict, LK_ERR_IO, "parquet: two dictionary pages in one chunk");
        ci.has_dict = true;
        ci.dict_n = (uint32_t)h.num_values;
        ci.dict_len = (uint32_t)h.uncompressed;
        if (cm.phys_type == PT_BYTE_ARRAY) {
          if (want_strings) {
            std::vector<uint8_t> buf;
            LK_CHECK(snappy_uncompress(data + payload, (size_t)h.compressed, buf) && buf.size() == (size_t)h.uncompressed, LK_ERR_IO,
                     "parquet: malformed SNAPPY dictionary page");
            ci.dict_strings.reserve(ci.dict_n);
            size_t q = 0;
            for (uint32_t i = 0; i < ci.dict_n; i++) {
              LK_CHECK(q + 4 <= buf.size(), LK_ERR_IO, "parquet: truncated dictionary page");
              uint32_t n;
              memcpy(&n, buf.data() + q, 4);
              q += 4;
              LK_CHECK(q + n <= buf.size(), LK_ERR_IO, "parquet: truncated dictionary page");
              ci.dict_strings.emplace_back((const char*)buf.data() + q, n);
              q += n;
            }
          }
        } else {
          unsigned esz = (cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT) ? 4 : 8;
          LK_CHECK(cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT || cm.phys_type == PT_INT64 || cm.phys_type == PT_DOUBLE,
                   LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': unsupported physical type");
          LK_CHECK((uint64_t)ci.dict_n * esz <= ci.dict_len, LK_ERR_IO, "parquet: truncated numeric dictionary");
          ci.dict_off = reserve_z((uint32_t)h.uncompressed);
          ci.zpages.push_back({payload, (uint32_t)h.compressed, ci.dict_off, (uint32_t)h.uncompressed, -1, ZP_DECODE});
        }
      } else if (h.type == 0 || h.type == 3) {
        LK_CHECK(h.num_values >= 0 && (uint64_t)row + (uint64_t)h.num_values <= ci.num_rows, LK_ERR_IO, "parquet: page rows exceed row group");
        PageInfo pg;
        pg.first_row = row;
        pg.num_rows = (uint32_t)h.num_values;
        pg.first_vidx = vidx;
        pg.deferred = true;
        uint32_t flags = ZP_DECODE;
        uint64_t src = payload;
        uint32_t src_len = (uint32_t)h.compressed, dst_len = (uint32_t)h.uncompressed;
        if (h.type == 3) {  // V2: repetition / definition levels sit uncompressed in front of the compressed values
          LK_CHECK(h.rep_len == 0, LK_ERR_UNSUPPORTED, "parquet: repetition levels");
          LK_CHECK(h.def_len >= 0 && h.def_len <= h.compressed && h.def_len <= h.uncompressed, LK_ERR_IO, "parquet: truncated definition levels");
          if (leaf.max_def > 0) { pg.def_off = payload; pg.def_end = payload + (uint64_t)h.def_len; }
          src += (uint64_t)h.def_len;
          src_len -= (uint32_t)h.def_len;
          dst_len -= (uint32_t)h.def_len;
          flags |= ZP_VALUES_ONLY;
        } else if (leaf.max_def > 0) {
          LK_CHECK(h.def_encoding == ENC_RLE, LK_ERR_UNSUPPORTED, "parquet: BIT_PACKED definition levels");
          flags |= ZP_V1_DEF;
        }
        if (h.encoding == ENC_RLE_DICTIONARY || h.encoding == ENC_PLAIN_DICTIONARY) {
          LK_CHECK(ci.has_dict, LK_ERR_IO, "parquet: dictionary-encoded page without a dictionary page");
          pg.dict_coded = true;
          flags |= ZP_DICT_CODED;
        } else if (h.encoding == ENC_PLAIN) {
          LK_CHECK(cm.phys_type != PT_BYTE_ARRAY, LK_ERR_UNSUPPORTED,
                   "column '" + leaf.name + "': PLAIN (non-dictionary) string pages inside a compressed chunk are not supported");
          LK_CHECK(cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT || cm.phys_type == PT_INT64 || cm.phys_type == PT_DOUBLE, LK_ERR_UNSUPPORTED,
                   "column '" + leaf.name + "': PLAIN pages of this type are not supported");
        } else {
          fail(LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': page encoding " + std::to_string(h.encoding) + " is not supported");
        }
        pg.values_off = reserve_z(dst_len);  // start of the inflated bytes; the device moves it past levels / bit width
        pg.values_len = dst_len;
        ci.zpages.push_back({src, src_len, pg.values_off, dst_len, (int)ci.pages.size(), flags});
        ci.pages.push_back(pg);
        row += pg.num_rows;
      }
      p = pend;
      continue;
    }
    if (h.type == 2) {  // dictionary page
      LK_CHECK(h.encoding == ENC_PLAIN || h.encoding == ENC_PLAIN_DICTIONARY, LK_ERR_UNSUPPORTED, "parquet: dictionary page encoding");
      LK_CHECK(!ci.has_dict, LK_ERR_IO, "parquet: two dictionary pages in one chunk");
      ci.has_dict = true;
      ci.dict_off = payload;
      ci.dict_len = (uint32_t)h.compressed;
      ci.dict_n = (uint32_t)h.num_values;
      if (cm.phys_type == PT_BYTE_ARRAY) {
        if (want_strings) {
          ci.dict_strings.reserve(ci.dict_n);
          uint64_t q = payload;
          for (uint32_t i = 0; i < ci.dict_n; i++) {
            LK_CHECK(q + 4 <= pend, LK_ERR_IO, "parquet: truncated dictionary page");
            uint32_t n;
            memcpy(&n, data + q, 4);
            q += 4;
            LK_CHECK(q + n <= pend, LK_ERR_IO, "parquet: truncated dictionary page");
            ci.dict_strings.emplace_back((const char*)data + q, n);
            q += n;
          }
        }
      } else {
        unsigned esz = (cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT) ? 4 : 8;
        LK_CHECK(cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT || cm.phys_type == PT_INT64 || cm.phys_type == PT_DOUBLE,
                 LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': unsupported physical type");
        LK_CHECK((uint64_t)ci.dict_n * esz <= ci.dict_len, LK_ERR_IO, "parquet: truncated numeric dictionary");
      }
    } else if (h.type == 0 || h.type == 3) {  // data page v1 / v2
      LK_CHECK(h.num_values >= 0 && (uint64_t)row + (uint64_t)h.num_values <= ci.num_rows, LK_ERR_IO, "parquet: page rows exceed row group");
      PageInfo pg;
      pg.first_row = row;
      pg.num_rows = (uint32_t)h.num_values;
      pg.first_vidx = vidx;
      uint64_t q = payload;
      uint32_t nn = pg.num_rows;
      if (h.type == 3) LK_CHECK(h.rep_len == 0, LK_ERR_UNSUPPORTED, "parquet: repetition levels");
      if (leaf.max_def > 0) {
        uint64_t dl_off, dl_end;
        if (h.type == 0) {
          LK_CHECK(h.def_encoding == ENC_RLE, LK_ERR_UNSUPPORTED, "parquet: BIT_PACKED definition levels");
          LK_CHECK(q + 4 <= pend, LK_ERR_IO, "parquet: truncated definition levels");
          uint32_t dl;
          memcpy(&dl, data + q, 4);
          dl_off = q + 4;
          dl_end = dl_off + dl;
          LK_CHECK(dl_end <= pend, LK_ERR_IO, "parquet: truncated definition levels");
          q = dl_end;
        } else {
          dl_off = q;
          dl_end = q + (uint64_t)h.def_len;
          LK_CHECK(dl_end <= pend, LK_ERR_IO, "parquet: truncated definition levels");
          q = dl_end;
        }
        pg.def_off = dl_off;
        pg.def_end = dl_end;
        nn = 0;
        uint32_t base_row = row;
        uint32_t nn_base = vidx;
        if (walk_runs) walk_hybrid(data, dl_off, dl_end, 1, pg.num_rows, [&](uint32_t s, uint32_t n, bool rle, uint32_t v, uint64_t off) {
          Run run;
          run.start = base_row + s;
          run.kind_value = rle ? (0x80000000u | (v & 1)) : (uint32_t)(off - ci.file_start);
          ci.def_runs.push_back(run);
          ci.def_nn_before.push_back(nn_base + nn);
          nn += rle ? ((v & 1) ? n : 0) : popcount_bits(data + off, n);
        });
      }
      pg.nvals = nn;
      if (h.encoding == ENC_PLAIN && cm.phys_type == PT_BYTE_ARRAY) {
        // length-prefixed strings (a string column whose dictionary outgrew its page): extend the chunk's dictionary and
        // re-encode the page as ONE bit-packed run of indices (see ChunkIndex::synth)
        LK_CHECK(want_strings, LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': PLAIN string pages of a column that is not used as a string");
        if (ci.synth.empty()) {
          ci.synth_base = (ci.file_len + 7) & ~7ull;
          for (uint32_t k = 0; k < (uint32_t)ci.dict_strings.size(); k++) synth_lookup.emplace(ci.dict_strings[k], k);
          ci.has_dict = true;
        }
        std::vector<uint32_t> codes;
        uint64_t v = q;
        while (v < pend) {
          LK_CHECK(v + 4 <= pend, LK_ERR_IO, "parquet: truncated PLAIN string page");
          uint32_t n;
          memcpy(&n, data + v, 4);
          v += 4;
          LK_CHECK(v + n <= pend, LK_ERR_IO, "parquet: truncated PLAIN string page");
          std::string str((const char*)data + v, n);
          v += n;
          auto it = synth_lookup.find(str);
          if (it == synth_lookup.end()) {
            it = synth_lookup.emplace(str, (uint32_t)ci.dict_strings.size()).first;
            ci.dict_strings.push_back(std::move(str));
          }
          codes.push_back(it->second);
          if (walk_runs && codes.size() == nn) break;  // (a page may carry padding after its last value)
        }
        LK_CHECK(!walk_runs || codes.size() == nn, LK_ERR_IO, "parquet: PLAIN string page holds fewer values than its definition levels announce");
        LK_CHECK(codes.size() <= pg.num_rows, LK_ERR_IO, "parquet: PLAIN string page holds more values than rows");
        if (!walk_runs) nn = (uint32_t)codes.size();
        pg.nvals = nn;
        ci.dict_n = (uint32_t)ci.dict_strings.size();
        uint32_t width = 1;
        while (width < 31 && (ci.dict_n - 1) >> width) width++;
        pg.dict_coded = true;
        pg.synth = true;
        pg.bit_width = (uint8_t)width;
        // hybrid stream: varint((groups << 1) | 1), then groups * width bytes of LSB-first packed indices
        const uint64_t groups = (codes.size() + 7) / 8;
        const size_t at = ci.synth.size();
        if (groups) {
          uint64_t hdr = (groups << 1) | 1;
          while (hdr >= 0x80) { ci.synth.push_back((uint8_t)(hdr | 0x80)); hdr >>= 7; }
          ci.synth.push_back((uint8_t)hdr);
          const size_t payload = ci.synth.size();
          ci.synth.resize(payload + groups * width, 0);
          uint64_t bit = 0;
          for (uint32_t c : codes) {
            for (uint32_t b = 0; b < width; b++, bit++)
              if ((c >> b) & 1) ci.synth[payload + (bit >> 3)] |= (uint8_t)(1u << (bit & 7));
          }
        }
        pg.values_off = at;  // offset inside ci.synth
        pg.values_len = (uint32_t)(ci.synth.size() - at);
        while (ci.synth.size() & 7) ci.synth.push_back(0);
        if (walk_runs && nn > 0) {
          uint32_t vbase = vidx;
          walk_hybrid(ci.synth.data(), at, at + pg.values_len, pg.bit_width, nn, [&](uint32_t s2, uint32_t n2, bool rle, uint32_t v2, uint64_t off) {
            (void)n2;
            Run run;
            run.start = vbase + s2;
            run.kind_value = rle ? (0x80000000u | (v2 & 0x7fffffffu)) : (uint32_t)(ci.synth_base + off);
            ci.val_runs.push_back(run);
          });
        }
      } else if (h.encoding == ENC_PLAIN) {
        unsigned esz = (cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT) ? 4 : (cm.phys_type == PT_INT64 || cm.phys_type == PT_DOUBLE) ? 8 : 0;
        LK_CHECK(esz != 0, LK_ERR_UNSUPPORTED,
                 "column '" + leaf.name + "': PLAIN pages of this type are not supported");
        LK_CHECK(!walk_runs || q + (uint64_t)nn * esz <= pend, LK_ERR_IO, "parquet: truncated PLAIN page");
        pg.dict_coded = false;
        pg.values_off = q;
        pg.values_len = (uint32_t)(pend - q);
      } else if (h.encoding == ENC_RLE_DICTIONARY || h.encoding == ENC_PLAIN_DICTIONARY) {
        LK_CHECK(ci.has_dict, LK_ERR_IO, "parquet: dictionary-encoded page without a dictionary page");
        pg.dict_coded = true;
        if (nn > 0 || q < pend) {
          LK_CHECK(q + 1 <= pend, LK_ERR_IO, "parquet: truncated dictionary-encoded page");
          pg.bit_width = data[q];
          LK_CHECK(pg.bit_width <= 31, LK_ERR_UNSUPPORTED, "parquet: dictionary index bit width > 31");
          q += 1;
        }
        pg.values_off = q;
        pg.values_len = (uint32_t)(pend - q);
        uint32_t vbase = vidx;
        if (nn > 0 && walk_runs)
          walk_hybrid(data, q, pend, pg.bit_width, nn, [&](uint32_t s, uint32_t n, bool rle, uint32_t v, uint64_t off) {
            (void)n;
            Run run;
            run.start = vbase + s;
            run.kind_value = rle ? (0x80000000u | (v & 0x7fffffffu)) : (uint32_t)(off - ci.file_start);
            if (rle) LK_CHECK(v < ci.dict_n, LK_ERR_IO, "parquet: dictionary index out of range");
            ci.val_runs.push_back(run);
          });
      } else {
        fail(LK_ERR_UNSUPPORTED, "column '" + leaf.name + "': page encoding " + std::to_string(h.encoding) + " is not supported");
      }
      ci.pages.push_back(pg);
      row += pg.num_rows;
      vidx += nn;
    }  // other page types (index pages) are skipped
    p = pend;
  }
  LK_CHECK(row == ci.num_rows, LK_ERR_IO, "parquet: pages do not cover the row group");
  // the kernels address a chunk's bit-packed dictionary indices by a 32-bit BIT offset from the chunk's first byte
  bool any_dict_page = false;
  for (auto& pg : ci.pages) any_dict_page |= pg.dict_coded;
  LK_CHECK(!(any_dict_page || !ci.val_runs.empty()) || ci.file_len + ci.synth.size() + ci.z_len + 16 < (1ull << 29), LK_ERR_UNSUPPORTED,
           "dictionary-coded column chunk of '" + leaf.name + "' is larger than 512 MB");
  return ci;
}

}  // namespace lk

// Host-side Parquet reader: footer (Thrift compact), page headers, dictionary pages and the *seek index* over the
// RLE/bit-packed hybrid streams.  The host never materialises column values: it walks headers (and popcounts the
// definition-level bits) so that the device can start decoding at any row.  Replaces the metadata half of DuckDB's
// parquet scan, reached from core/src/main/scala/com/cardinal/utils/Commons.scala:213-240.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lk_common.h"
#include "lk_device.h"

namespace lk {

enum PhysType : int { PT_BOOLEAN = 0, PT_INT32 = 1, PT_INT64 = 2, PT_INT96 = 3, PT_FLOAT = 4, PT_DOUBLE = 5, PT_BYTE_ARRAY = 6, PT_FLBA = 7 };
enum Codec : int { CODEC_NONE = 0, CODEC_SNAPPY = 1 };
enum Encoding : int { ENC_PLAIN = 0, ENC_PLAIN_DICTIONARY = 2, ENC_RLE = 3, ENC_BIT_PACKED = 4, ENC_RLE_DICTIONARY = 8 };

struct ColumnChunkMeta {
  int phys_type = -1;
  int codec = 0;
  int64_t num_values = 0;
  int64_t total_compressed_size = 0;
  int64_t total_uncompressed_size = 0;
  int64_t data_page_offset = 0;
  int64_t dictionary_page_offset = -1;
  int plain_data_pages = -1;  // data pages encoded PLAIN according to the footer's encoding_stats (-1: the writer left them out)
};

struct RowGroupMeta {
  int64_t num_rows = 0;
  std::vector<ColumnChunkMeta> columns;  // by leaf index
};

struct LeafColumn {
  std::string name;  // flat leaf name (may contain dots: "resource.service.name")
  int phys_type = -1;
  int max_def = 0;   // 0 REQUIRED, 1 OPTIONAL
  bool flat = true;  // false for leaves below a group (not addressable by the reference's SQL either)
};

struct FileMeta {
  int64_t num_rows = 0;
  std::vector<LeafColumn> leaves;
  std::vector<RowGroupMeta> row_groups;
  int leaf_index(const std::string& name) const;
};

// Parses the footer of a Parquet file held in memory.  Throws lk::Error(LK_ERR_IO) if it is not Parquet.
FileMeta parse_footer(const uint8_t* data, size_t len);
// The same from the LAST `tail_len` bytes of a file of `file_len` bytes (a file that is read sparsely: footer first, then the
// touched column chunks).  Returns false when the tail is too short to hold the footer: *need = bytes of tail required.
bool parse_footer_tail(const uint8_t* tail, size_t tail_len, size_t file_len, FileMeta* out, size_t* need);

struct PageInfo {
  uint32_t first_row = 0, num_rows = 0;  // chunk-level row range
  uint32_t first_vidx = 0, nvals = 0;    // chunk-level index of the first non-null value, number of non-null values
  bool dict_coded = false;
  uint8_t bit_width = 0;
  uint64_t values_off = 0;  // file offset of the value bytes (after the bit-width byte for dictionary pages)
  uint32_t values_len = 0;
  uint64_t def_off = 0, def_end = 0;  // file byte range of the definition-level stream (empty for REQUIRED columns)
  bool deferred = false;  // compressed page: def_off / def_end (V1) and bit_width are read by the device from the inflated bytes
  bool synth = false;  // PLAIN string page re-encoded by the host: values_off / values_len address ChunkIndex::synth instead of the file
};

// Index of one column chunk.  Bit-packed run offsets are relative to file_start (the chunk's first byte).
struct ChunkIndex {
  bool present = false;
  int phys_type = -1;
  int max_def = 0;
  uint64_t file_start = 0, file_len = 0;  // byte range of the chunk (dictionary page + data pages)
  int64_t total_compressed_size = 0;
  uint32_t num_rows = 0;
  std::vector<PageInfo> pages;
  std::vector<Run> def_runs;             // start = chunk-level row
  std::vector<uint32_t> def_nn_before;   // non-null values before each def run
  std::vector<Run> val_runs;             // start = chunk-level value index (dictionary-coded pages only)
  bool has_dict = false;
  uint64_t dict_off = 0;  // file offset of the PLAIN dictionary payload
  uint32_t dict_len = 0, dict_n = 0;
  std::vector<std::string> dict_strings;  // BYTE_ARRAY dictionaries, decoded on the host
  // PLAIN (non-dictionary) BYTE_ARRAY data pages -- what a writer falls back to when a chunk's dictionary outgrows its page
  // limit, i.e. high-cardinality tags -- are re-encoded here: their strings extend dict_strings and every such page becomes
  // one bit-packed run of indices in `synth`, a hybrid stream like the file's own.  The bytes are placed right behind the
  // chunk in the arena (offset synth_base from the chunk's first byte), so kernels and index builders see a dictionary page.
  std::vector<uint8_t> synth;
  uint64_t synth_base = 0;
  // SNAPPY chunks: every compressed page is inflated ON THE DEVICE into room reserved behind the chunk (at z_base from the
  // chunk's first byte, pages back to back); page offsets that point into inflated bytes are "virtual file offsets"
  // file_start + z_base + position, so that the rebasing of file offsets to arena offsets works for them unchanged.  The
  // host inflates only BYTE_ARRAY dictionary pages (it needs the strings; they never reach the device).
  struct ZPageInfo { uint64_t src_off; uint32_t src_len; uint64_t dst_virt; uint32_t dst_len; int page; uint32_t flags; };
  std::vector<ZPageInfo> zpages;
  uint64_t z_base = 0, z_len = 0;
  // chunk-level value index of row r (number of non-null values before r); r may equal num_rows
  uint32_t vidx_at(const uint8_t* file, uint32_t r) const;
  // same, when the def run `run_index` holding row r is already known (linear sweeps)
  uint32_t vidx_in_run(const uint8_t* file, int run_index, uint32_t r) const;
  int def_run_at(uint32_t r) const;   // index of the def run containing row r
  int val_run_at(uint32_t v) const;   // index of the value run containing value v
  int page_at(uint32_t r) const;
};

// Arena bytes to reserve behind a BYTE_ARRAY chunk for re-encoded PLAIN pages (0 when the footer says there are none), or
// behind a compressed chunk for its inflated pages (the footer's total_uncompressed_size).
uint64_t synth_reserve(const ColumnChunkMeta& cm);
// SNAPPY block format (length preamble + literal / copy elements); returns false on malformed input.  Host side: dictionary
// pages of string columns only -- data pages are inflated by snappy_decode_kernel (lk_engine.cu).
bool snappy_uncompress(const uint8_t* src, size_t n, std::vector<uint8_t>& out);

// Byte range [start, start + len) of a column chunk inside the file (dictionary page + data pages), from the footer alone.
void chunk_byte_range(const ColumnChunkMeta& cm, size_t file_len, const std::string& name, uint64_t& start, uint64_t& len);

// Walks every page header of the chunk and -- walk_runs -- every run header of its hybrid streams.
// want_strings: decode the BYTE_ARRAY dictionary into dict_strings.
// walk_runs = false (the device builds the run index, lk_engine.cu): pages carry the byte ranges of their streams, nvals /
// first_vidx are left at 0 and def_runs / val_runs stay empty.
ChunkIndex index_chunk(const uint8_t* data, size_t len, const LeafColumn& leaf, const ColumnChunkMeta& cm, int64_t rg_rows,
                       bool want_strings, bool walk_runs = true);

}  // namespace lk

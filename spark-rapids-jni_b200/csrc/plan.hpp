// plan.hpp -- per-schema plan: JCUDF layout (compute_column_information, RC:1332-1371) plus the
// static work schedule the kernels run (columns grouped by width class), mirrored on the device.
#pragma once
#include <stdint.h>

#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "../../include/srj_b200.h"

namespace srj {

// Width classes: element sizes 1, 2, 4, 8, 16 bytes.
constexpr int kNumClasses = 5;
inline int class_of_size(int sz) { return sz == 1 ? 0 : sz == 2 ? 1 : sz == 4 ? 2 : sz == 8 ? 3 : sz == 16 ? 4 : -1; }

// One fixed-width "entry" = one field the fixed-width transpose moves.  A STRING column
// contributes one 4-byte entry for its length word (second uint32 of the pair, RC:2163-2172) in
// from_rows; in to_rows the pair is produced by the string stage instead.
struct Entry {
  int32_t start;   // byte offset in the row
  int32_t column;  // schema column index
};

struct Tiling {
  int32_t tile_rows;    // max rows per tile (multiple of rows_per_item)
  int32_t rows_per_item;  // 8, 16 or 32: lanes of a warp item that map to rows
  int32_t stage_bytes;  // shared-memory bytes per pipeline stage (payload)
  int32_t num_stages;
};

// ---- wide variable-width tables (from_rows_wide.cu) ------------------------------------------------------
// The fixed section of a wide row is cut into byte-range "slabs"; a tile = R rows x one slab, each row's slab
// fetched by its own TMA bulk copy (the variable section of the row is never read by phase 1).
struct WideEntry {
  int32_t start;   // byte offset in the row (STRING: the (offset, len) pair)
  int32_t column;  // schema column index
  int32_t sidx;    // index among the STRING columns, or -1
  int32_t slab;
};
struct WideSlab {
  int32_t begin, end;  // row byte range [begin, end) staged for this slab (multiples of 8)
  int32_t cb[7];       // entry index boundaries of the width classes, processed in the order 16,8,4,2,1,STRING
};
struct WidePlan {
  bool enabled = false;
  int32_t R = 0, G = 0, pitch = 0, nstages = 0, nslabs = 0;
  std::vector<WideEntry> entries;
  std::vector<WideSlab> slabs;
  const WideEntry* d_entries = nullptr;
  const WideSlab* d_slabs    = nullptr;
};

// Per-call pointer tables (column/mask/offset pointers differ on every call) go host -> device through
// a small ring of pinned staging buffers + device buffers owned by the plan: one truly asynchronous
// cudaMemcpyAsync per call, no allocation, re-entrant (a slot is reused only after the event recorded
// behind its last use has completed).
struct TableSlot {
  void* h_pinned = nullptr;
  void* d_buf    = nullptr;
  size_t cap     = 0;
  cudaEvent_t ev = nullptr;
  bool used      = false;
  std::mutex busy;  // held for the duration of one API call's lease
};
struct TableRing {
  static constexpr int kSlots = 8;
  std::mutex mu;
  TableSlot slots[kSlots];
  unsigned next = 0;
};

// Device staging of the host-buffer entry points (srj_convert_*_host): grow-only buffers + streams, pooled per plan.
struct HostArena {
  std::mutex busy;
  cudaStream_t st[3] = {nullptr, nullptr, nullptr};
  void* d_buf[4]     = {nullptr, nullptr, nullptr, nullptr};  // rows | outputs / inputs | misc (offsets, workspace, counters) | chars
  size_t d_cap[4]    = {0, 0, 0, 0};
  void* h_pin        = nullptr;  // pinned host scratch (counters)
  size_t h_cap       = 0;
};
struct HostArenaPool {
  static constexpr int kArenas = 4;
  HostArena a[kArenas];
};

}  // namespace srj

struct srj_plan {
  int32_t device;
  int32_t num_columns;
  int32_t num_string_columns;
  int32_t validity_offset;
  int32_t size_per_row;
  int32_t fixed_row_size;
  std::vector<int32_t> type_ids, scales, col_start, col_size;
  std::vector<int32_t> string_columns;  // schema indices of STRING columns, in order

  // from_rows schedule: entries grouped by width class; STRING columns add a 4-byte length entry
  std::vector<srj::Entry> fr_entries;
  int32_t fr_class_begin[srj::kNumClasses + 1];
  // to_rows schedule: fixed-width columns only
  std::vector<srj::Entry> tr_entries;
  int32_t tr_class_begin[srj::kNumClasses + 1];

  srj::Tiling tiling;
  srj::WidePlan wide;  // from_rows of wide variable-width tables

  // device mirrors (one allocation)
  void* d_blob;
  const srj::Entry* d_fr_entries;
  const srj::Entry* d_tr_entries;
  const int32_t* d_col_start;     // [num_columns]
  const int32_t* d_string_cols;   // [num_string_columns]
  const int32_t* d_string_start;  // [num_string_columns] row byte offset of each pair
  const int32_t* d_tr_chunk_off;  // [tr_entries] staging byte offset per row unit (to_rows2)

  mutable srj::TableRing ring;
  mutable srj::HostArenaPool host_pool;
};

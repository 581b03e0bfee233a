// from_rows.cu -- JCUDF rows -> columns (reference: convert_from_rows, RC:2149-2441).
//
// One persistent, warp-specialised kernel replaces copy_from_rows + copy_validity_from_rows +
// fixup_null_counts (RC:879-969, 987-1094, 2130-2136) and optionally fuses the partition hash:
//
//   producer warp : walks the CTA's contiguous row range tile by tile; each tile is ONE contiguous
//                   byte range of the row buffer, moved global->shared by a single 1-D TMA bulk copy
//                   (cp.async.bulk ... mbarrier::complete_tx) into a multi-stage ring, so ~2 stages
//                   (>=100 KB) per SM are always in flight with no registers or LSU slots spent.
//   consumer warps: lane = row.  For each width class (1/2/4/8/16 B) a warp item reads one field
//                   of 32 consecutive rows from the tile and stores it to the column with one
//                   coalesced store; validity bytes are bit-transposed with __ballot_sync into the
//                   column masks; null counts are popc'd on the way; the row hash of the key
//                   columns is computed in registers from the same tile.
//
// Rows whose tile cannot be staged (row larger than a stage, unaligned buffers) take the SAFE
// path: same code, reading global memory byte-wise.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "hash_device.cuh"
#include "kernels.hpp"
#include "plan.hpp"

namespace srj {

constexpr int kConsumerWarps = 15;
constexpr int kThreads       = (kConsumerWarps + 1) * 32;
constexpr int kMaxStages     = 4;
constexpr int kStageSlack    = 32;  // skew (<=8) + tail

struct StageHdr {
  int64_t r0;        // first table row of the tile
  int32_t rows;      // 0 = end of this CTA's range
  int32_t safe;      // 1 = not staged: read rows from global memory byte-wise
  int32_t skew;      // payload byte offset of global byte `gbase` (fixed-stride tiles)
  int32_t pad;
  int64_t gbase;     // global byte offset (from p.rows) that payload[skew] corresponds to
};

struct FromRowsParams {
  const uint8_t* rows;
  const int32_t* row_offsets;  // NULL => fixed stride
  int64_t rows_bytes;
  int64_t num_rows;
  int64_t rows_per_cta;
  int32_t ncols;
  int32_t validity_offset;
  int32_t row_stride;
  int32_t tile_rows;
  int32_t rpl;  // rows per item: 8, 16 or 32
  int32_t stage_bytes;
  int32_t nstages;
  int32_t nentries;
  int32_t class_begin[kNumClasses + 1];
  const Entry* entries;
  void* const* ent_dst;     // [nentries] column base pointers (STRING: offsets + 1)
  uint32_t* const* masks;   // [ncols]
  unsigned long long* null_counts;  // [ncols] or NULL
  // fused hash
  int32_t hash_kind;
  int32_t hash_nkeys;
  int32_t key_start[16];
  int32_t key_type[16];
  int32_t key_col[16];
  int64_t hash_seed;
  void* hash_out;
};

// ---- element movers ------------------------------------------------------------------------------
template <int W, bool SAFE>
__device__ __forceinline__ void move_elem(const uint8_t* src, uint8_t* dst)
{
  if constexpr (SAFE) {
    uint8_t tmp[W];
#pragma unroll
    for (int i = 0; i < W; ++i) tmp[i] = src[i];
    if constexpr (W == 1) {
      *dst = tmp[0];
    } else if constexpr (W == 2) {
      *reinterpret_cast<uint16_t*>(dst) = static_cast<uint16_t>(tmp[0] | (tmp[1] << 8));
    } else if constexpr (W == 4) {
      *reinterpret_cast<uint32_t*>(dst) = hash::ld_u32_bytes(tmp);
    } else if constexpr (W == 8) {
      *reinterpret_cast<uint64_t*>(dst) = hash::ld_u64_bytes(tmp);
    } else {
      uint4 v;
      v.x = hash::ld_u32_bytes(tmp);
      v.y = hash::ld_u32_bytes(tmp + 4);
      v.z = hash::ld_u32_bytes(tmp + 8);
      v.w = hash::ld_u32_bytes(tmp + 12);
      *reinterpret_cast<uint4*>(dst) = v;
    }
  } else {
    if constexpr (W == 1) {
      *dst = *src;
    } else if constexpr (W == 2) {
      *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(src);
    } else if constexpr (W == 4) {
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
    } else if constexpr (W == 8) {
      *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
    } else {
      // rows are only 8-byte aligned (JCUDF_ROW_ALIGNMENT, RC:63): two 8-byte reads, one 16-byte store
      const uint2 a = *reinterpret_cast<const uint2*>(src);
      const uint2 b = *reinterpret_cast<const uint2*>(src + 8);
      *reinterpret_cast<uint4*>(dst) = make_uint4(a.x, a.y, b.x, b.y);
    }
  }
}

template <bool SAFE>
__device__ __forceinline__ uint64_t load_key(const uint8_t* p, int sz)
{
  if constexpr (SAFE) {
    uint64_t v = 0;
    for (int i = 0; i < sz; ++i) v |= static_cast<uint64_t>(p[i]) << (8 * i);
    return v;
  } else {
    switch (sz) {
      case 1: return *p;
      case 2: return *reinterpret_cast<const uint16_t*>(p);
      case 4: return *reinterpret_cast<const uint32_t*>(p);
      default: {
        const uint2 a = *reinterpret_cast<const uint2*>(p);
        return static_cast<uint64_t>(a.x) | (static_cast<uint64_t>(a.y) << 32);
      }
    }
  }
}

__device__ __forceinline__ int key_size(int32_t t)
{
  switch (t) {
    case SRJ_INT8: case SRJ_UINT8: case SRJ_BOOL8: return 1;
    case SRJ_INT16: case SRJ_UINT16: return 2;
    case SRJ_INT32: case SRJ_UINT32: case SRJ_FLOAT32: case SRJ_TIMESTAMP_DAYS: case SRJ_DURATION_DAYS:
    case SRJ_DECIMAL32: return 4;
    case SRJ_DECIMAL128: return 16;
    default: return 8;
  }
}

// Shared-memory resident copies of the schedule (filled once per CTA).
struct SmemTables {
  const int32_t* ent_start;  // [nentries]
  uint8_t* const* ent_dst;   // [nentries]
  uint32_t* const* masks;    // [ncols]
  int32_t* nulls;            // [ncols]
};

// Process one tile.  `base` + row offset gives the row's first byte:
//   s_off != NULL : base + (uint32)s_off[i]
//   s_off == NULL : base + i * stride
template <bool SAFE>
__device__ __forceinline__ void process_tile(const FromRowsParams& p, const SmemTables& t, const uint8_t* base,
                                             const int32_t* s_off, int64_t stride, int64_t r0, int rows,
                                             int cw /* consumer warp index */)
{
  const int lane = lane_id();
  auto rowptr    = [&](int i) -> const uint8_t* {
    return s_off ? base + static_cast<uint32_t>(s_off[i]) : base + static_cast<int64_t>(i) * stride;
  };

  // ---- fixed-width fields ------------------------------------------------------------------------
  const int rpl     = p.rpl;
  const int cpi     = 32 / rpl;
  const int sub     = lane / rpl;
  const int lr      = lane - sub * rpl;
  const int ngroups = (rows + rpl - 1) / rpl;

  auto run_class = [&](auto wtag, int k) {
    constexpr int W  = decltype(wtag)::value;
    const int nb     = p.class_begin[k];
    const int ne     = p.class_begin[k + 1];
    const int nslots = (ne - nb + cpi - 1) / cpi;
    const int total  = nslots * ngroups;
    for (int item = cw; item < total; item += kConsumerWarps) {
      const int g    = item / nslots;
      const int slot = item - g * nslots;
      const int e    = nb + slot * cpi + sub;
      const int row  = g * rpl + lr;
      if (e < ne && row < rows) {
        const uint8_t* src = rowptr(row) + t.ent_start[e];
        uint8_t* dst       = t.ent_dst[e] + (r0 + row) * W;
        move_elem<W, SAFE>(src, dst);
      }
    }
  };
  run_class(std::integral_constant<int, 16>{}, 4);
  run_class(std::integral_constant<int, 8>{}, 3);
  run_class(std::integral_constant<int, 4>{}, 2);
  run_class(std::integral_constant<int, 2>{}, 1);
  run_class(std::integral_constant<int, 1>{}, 0);

  // ---- validity: bit-transpose row bytes -> column mask words (RC:1062-1071 semantics) ----------
  const int nvb    = (p.ncols + 7) >> 3;
  const int ng32   = (rows + 31) >> 5;
  const int vitems = nvb * ng32;
  for (int item = cw; item < vitems; item += kConsumerWarps) {
    const int g       = item / nvb;
    const int b       = item - g * nvb;
    const int row     = g * 32 + lane;
    const bool active = row < rows;
    uint32_t byte     = 0;
    if (active) byte = rowptr(row)[p.validity_offset + b];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t w = __ballot_sync(0xffffffffu, (byte >> k) & 1u);
      if (lane == k) mine = w;
    }
    const int col = b * 8 + lane;
    if (lane < 8 && col < p.ncols) {
      const int nact       = tmin(32, rows - g * 32);
      const uint32_t amask = nact == 32 ? 0xffffffffu : ((1u << nact) - 1u);
      const int nnull      = __popc(~mine & amask);
      if (nnull && t.nulls) atomicAdd(&t.nulls[col], nnull);
      const int64_t rg   = r0 + g * 32;  // multiple of 8
      uint8_t* mp        = reinterpret_cast<uint8_t*>(t.masks[col]) + (rg >> 3);
      const bool tbl_end = (rg + nact) == p.num_rows;
      int nbytes         = (nact + 7) >> 3;
      // the table's last mask word is written whole so its tail bits are 0 (RC:1081-1090)
      if (tbl_end) nbytes = static_cast<int>(round_up64((rg >> 3) + nbytes, 4) - (rg >> 3));
      if (nbytes == 4 && ((rg & 31) == 0)) {
        *reinterpret_cast<uint32_t*>(mp) = mine;
      } else {
        for (int i = 0; i < nbytes; ++i) mp[i] = static_cast<uint8_t>(static_cast<uint64_t>(mine) >> (8 * i));
      }
    }
  }

  // ---- fused row hash of the key columns ----------------------------------------------------------
  if (p.hash_kind != SRJ_HASH_NONE) {
    for (int g = cw; g < ng32; g += kConsumerWarps) {
      const int row = g * 32 + lane;
      if (row >= rows) continue;
      const uint8_t* rp = rowptr(row);
      uint64_t hx       = static_cast<uint64_t>(p.hash_seed);
      uint32_t hm       = static_cast<uint32_t>(p.hash_seed);
      uint32_t hh       = 0;
      for (int k = 0; k < p.hash_nkeys; ++k) {
        const int c      = p.key_col[k];
        const bool valid = (rp[p.validity_offset + (c >> 3)] >> (c & 7)) & 1u;
        const int32_t ty = p.key_type[k];
        const int sz     = key_size(ty);
        uint64_t v = 0, v2 = 0;
        if (valid) {
          v = load_key<SAFE>(rp + p.key_start[k], sz);
          if (sz == 16) v2 = load_key<SAFE>(rp + p.key_start[k] + 8, 8);
        }
        if (p.hash_kind == SRJ_HASH_XXHASH64) {
          if (valid) hx = hash::xx_fixed(ty, v, v2, hx);
        } else if (p.hash_kind == SRJ_HASH_MURMUR3_32) {
          if (valid) hm = hash::mm_fixed(ty, v, v2, hm);
        } else {
          hh = 31u * hh + (valid ? static_cast<uint32_t>(hash::hive_fixed(ty, v)) : 0u);
        }
      }
      if (p.hash_kind == SRJ_HASH_XXHASH64)
        reinterpret_cast<uint64_t*>(p.hash_out)[r0 + row] = hx;
      else if (p.hash_kind == SRJ_HASH_MURMUR3_32)
        reinterpret_cast<uint32_t*>(p.hash_out)[r0 + row] = hm;
      else
        reinterpret_cast<uint32_t*>(p.hash_out)[r0 + row] = hh;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) from_rows_kernel(const __grid_constant__ FromRowsParams p)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int NS         = p.nstages;
  const int stage_span = p.stage_bytes + kStageSlack;
  uint8_t* payload0    = smem;
  int32_t* soff0       = reinterpret_cast<int32_t*>(smem + static_cast<size_t>(NS) * stage_span);
  const int soff_span  = (p.tile_rows + 4) & ~3;  // ints per stage (tile_rows + 1, padded)
  StageHdr* hdr0       = reinterpret_cast<StageHdr*>(soff0 + static_cast<size_t>(NS) * soff_span);
  uint64_t* full       = reinterpret_cast<uint64_t*>(hdr0 + NS);
  uint64_t* empty      = full + kMaxStages;
  int32_t* s_ent_start = reinterpret_cast<int32_t*>(empty + kMaxStages);
  uint8_t** s_ent_dst  = reinterpret_cast<uint8_t**>(s_ent_start + ((p.nentries + 1) & ~1));
  uint32_t** s_masks   = reinterpret_cast<uint32_t**>(s_ent_dst + p.nentries);
  int32_t* s_nulls     = reinterpret_cast<int32_t*>(s_masks + p.ncols);

  const int tid = threadIdx.x;
  for (int i = tid; i < p.nentries; i += kThreads) {
    s_ent_start[i] = p.entries[i].start;
    s_ent_dst[i]   = static_cast<uint8_t*>(p.ent_dst[i]);
  }
  for (int i = tid; i < p.ncols; i += kThreads) {
    s_masks[i] = p.masks[i];
    s_nulls[i] = 0;
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * p.rows_per_cta;
  const int64_t c1 = tmin(p.num_rows, c0 + p.rows_per_cta);
  const int lane   = lane_id();
  const bool fixed = p.row_offsets == nullptr;
  // the staged fast path needs 8-byte aligned rows
  const bool base_ok = (reinterpret_cast<uintptr_t>(p.rows) & 7) == 0;

  if (warp_id() == 0) {
    // =================================== producer ===================================
    int64_t r = c0;
    int it    = 0;
    for (;; ++it) {
      const int s        = it % NS;
      const uint32_t par = ((it / NS) & 1) ^ 1;
      mbar_wait(&empty[s], par);  // first pass over the ring returns immediately
      uint8_t* pay  = payload0 + static_cast<size_t>(s) * stage_span;
      int32_t* soff = soff0 + static_cast<size_t>(s) * soff_span;
      StageHdr* h   = hdr0 + s;
      if (r >= c1) {
        if (lane == 0) {
          h->rows = 0;
          mbar_arrive(&full[s]);
        }
        break;
      }
      int rows    = static_cast<int>(tmin<int64_t>(p.tile_rows, c1 - r));
      int64_t glo = 0, ghi = 0;  // global byte range [glo, ghi) from p.rows
      bool safe   = !base_ok;
      if (fixed) {
        glo = r * p.row_stride;
        ghi = (r + rows) * static_cast<int64_t>(p.row_stride);
        if (ghi - glo > p.stage_bytes) safe = true;
      } else {
        // load off[r .. r+rows] (coalesced) and pick the largest multiple-of-8 row count that fits
        const int64_t a0   = p.row_offsets[r];
        const int64_t base = a0 - static_cast<int64_t>((reinterpret_cast<uintptr_t>(p.rows) + a0) & 15);
        int fit            = 0;
        bool misaligned    = (a0 & 7) != 0;
        for (int i0 = 0; i0 <= rows; i0 += 32) {
          const int i = i0 + lane;
          int64_t o   = 0;
          if (i <= rows) {
            o       = p.row_offsets[r + i];
            soff[i] = static_cast<int32_t>(o - base);
            if (i < rows && (o & 7)) misaligned = true;
          }
          const bool ok = (i >= 1) && (i <= rows) && (round_up64(o - base, 16) <= p.stage_bytes + 16);
          fit += __popc(__ballot_sync(0xffffffffu, ok));
        }
        misaligned = __any_sync(0xffffffffu, misaligned);
        if (fit < rows) fit &= ~7;
        if (fit == 0 || misaligned || safe) {
          // a row does not fit a stage (or rows are not 8-aligned): SAFE tile of <= 8 rows
          safe = true;
          rows = tmin(rows, 8);
          __syncwarp();
          for (int i = lane; i <= rows; i += 32) soff[i] = p.row_offsets[r + i];  // absolute offsets
        } else {
          rows = fit;
        }
        glo = a0;
        ghi = p.row_offsets[r + rows];
      }
      uint32_t tx     = 0;
      int32_t skew    = 0;
      uintptr_t t_lo  = 0, fl = 0;
      if (!safe) {
        // 16-byte aligned TMA window inside [rows, rows + rows_bytes); <16-byte head/tail remainders
        // (8-byte units) are copied by hand
        const uintptr_t a_lo = reinterpret_cast<uintptr_t>(p.rows) + glo;
        const uintptr_t a_hi = reinterpret_cast<uintptr_t>(p.rows) + ghi;
        const uintptr_t b_lo = reinterpret_cast<uintptr_t>(p.rows);
        const uintptr_t b_hi = b_lo + p.rows_bytes;
        fl                   = a_lo & ~uintptr_t{15};
        skew                 = static_cast<int32_t>(a_lo - fl);  // payload[skew] == global byte glo
        t_lo                 = fl;
        if (t_lo < b_lo) t_lo = fl + 16;                          // cannot read before the buffer
        uintptr_t t_hi = (a_hi + 15) & ~uintptr_t{15};
        if (t_hi > b_hi) t_hi = a_hi & ~uintptr_t{15};            // cannot read past the buffer
        if (t_hi > t_lo) tx = static_cast<uint32_t>(t_hi - t_lo);
        uintptr_t h_end = tmin(tmax(t_lo, a_lo), a_hi);             // head  [a_lo, h_end)
        if (tx == 0) h_end = a_hi;                                // tiny tile: all by hand
        const uintptr_t t_beg = tmax(tmin(t_hi, a_hi), h_end);      // tail  [t_beg, a_hi)
        if (lane == 0) {
          for (uintptr_t a = a_lo; a < h_end; a += 8)
            *reinterpret_cast<uint2*>(pay + (a - fl)) = *reinterpret_cast<const uint2*>(a);
          for (uintptr_t a = t_beg; a < a_hi; a += 8)
            *reinterpret_cast<uint2*>(pay + (a - fl)) = *reinterpret_cast<const uint2*>(a);
        }
      }
      if (lane == 0) {
        h->r0    = r;
        h->rows  = rows;
        h->safe  = safe ? 1 : 0;
        h->skew  = skew;
        h->gbase = glo;
      }
      __syncwarp();
      if (lane == 0) {
        if (tx) {
          mbar_arrive_expect_tx(&full[s], tx);  // release: header / offsets / hand copies are visible
          tma_load_1d(pay + (t_lo - fl), reinterpret_cast<const void*>(t_lo), tx, &full[s]);
        } else {
          mbar_arrive(&full[s]);
        }
      }
      r += rows;
    }
  } else {
    // =================================== consumers ===================================
    const int cw = warp_id() - 1;
    SmemTables t{s_ent_start, s_ent_dst, s_masks, p.null_counts ? s_nulls : nullptr};
    for (int it = 0;; ++it) {
      const int s        = it % NS;
      const uint32_t par = (it / NS) & 1;
      mbar_wait(&full[s], par);
      const StageHdr h = hdr0[s];
      if (h.rows == 0) break;
      const uint8_t* pay  = payload0 + static_cast<size_t>(s) * stage_span;
      const int32_t* soff = soff0 + static_cast<size_t>(s) * soff_span;
      if (!h.safe) {
        if (fixed)
          process_tile<false>(p, t, pay + h.skew, nullptr, p.row_stride, h.r0, h.rows, cw);
        else
          process_tile<false>(p, t, pay, soff, 0, h.r0, h.rows, cw);
      } else {
        if (fixed)
          process_tile<true>(p, t, p.rows + h.gbase, nullptr, p.row_stride, h.r0, h.rows, cw);
        else
          process_tile<true>(p, t, p.rows, soff, 0, h.r0, h.rows, cw);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (p.null_counts) {
    for (int i = tid; i < p.ncols; i += kThreads)
      if (s_nulls[i]) atomicAdd(&p.null_counts[i], static_cast<unsigned long long>(s_nulls[i]));
  }
}

// ---- host launcher -------------------------------------------------------------------------------
size_t from_rows_smem_bytes(const Tiling& tl, int nentries, int ncols)
{
  size_t b = static_cast<size_t>(tl.num_stages) * (tl.stage_bytes + kStageSlack);
  b += static_cast<size_t>(tl.num_stages) * ((tl.tile_rows + 4) & ~3) * 4;
  b += static_cast<size_t>(tl.num_stages) * sizeof(StageHdr);
  b += 2 * kMaxStages * 8;
  b += static_cast<size_t>((nentries + 1) & ~1) * 4;
  b += static_cast<size_t>(nentries) * 8 + static_cast<size_t>(ncols) * 8 + static_cast<size_t>(ncols) * 4;
  return (b + 127) & ~size_t{127};
}

int launch_from_rows(const srj_plan* plan, const uint8_t* rows, const int32_t* row_offsets, int64_t rows_bytes,
                     int64_t num_rows, void* const* d_ent_dst, uint32_t* const* d_masks, int64_t* d_null_counts,
                     const srj_fused_hash* fh, cudaStream_t stream)
{
  if (num_rows == 0) return SRJ_OK;
  FromRowsParams p{};
  p.rows            = rows;
  p.row_offsets     = row_offsets;
  p.rows_bytes      = rows_bytes;
  p.num_rows        = num_rows;
  p.ncols           = plan->num_columns;
  p.validity_offset = plan->validity_offset;
  p.row_stride      = plan->fixed_row_size;
  p.tile_rows       = plan->tiling.tile_rows;
  p.rpl             = plan->tiling.rows_per_item;
  p.stage_bytes     = plan->tiling.stage_bytes;
  p.nstages         = plan->tiling.num_stages;
  p.nentries        = static_cast<int32_t>(plan->fr_entries.size());
  for (int k = 0; k <= kNumClasses; ++k) p.class_begin[k] = plan->fr_class_begin[k];
  p.entries     = plan->d_fr_entries;
  p.ent_dst     = d_ent_dst;
  p.masks       = d_masks;
  p.null_counts = reinterpret_cast<unsigned long long*>(d_null_counts);
  p.hash_kind   = SRJ_HASH_NONE;
  if (fh && fh->kind != SRJ_HASH_NONE) {
    p.hash_kind  = fh->kind;
    p.hash_nkeys = fh->num_keys;
    p.hash_seed  = fh->seed;
    p.hash_out   = fh->out;
    for (int k = 0; k < fh->num_keys; ++k) {
      const int c    = fh->key_columns[k];
      p.key_col[k]   = c;
      p.key_type[k]  = plan->type_ids[c];
      p.key_start[k] = plan->col_start[c];
    }
  }
  int dev = 0, nsm = 0;
  SRJ_CUDA_TRY(cudaGetDevice(&dev));
  SRJ_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  // contiguous row range per CTA, a multiple of the tile height (=> of 32 or 8|16: mask-byte aligned)
  const int64_t T      = p.tile_rows;
  const int64_t ntiles = (num_rows + T - 1) / T;
  int64_t grid         = std::min<int64_t>(nsm, ntiles);
  int64_t tiles_per    = (ntiles + grid - 1) / grid;
  p.rows_per_cta       = tiles_per * T;
  grid                 = (num_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  const size_t smem    = from_rows_smem_bytes(plan->tiling, p.nentries, p.ncols);
  static thread_local int configured_dev = -1;
  static thread_local size_t configured_smem = 0;
  if (configured_dev != dev || configured_smem < smem) {
    SRJ_CUDA_TRY(cudaFuncSetAttribute(from_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    configured_dev  = dev;
    configured_smem = 232448;
  }
  from_rows_kernel<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p);
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

}  // namespace srj

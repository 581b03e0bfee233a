// to_rows_wide.cu -- columns -> JCUDF rows for WIDE variable-width tables (reference: copy_to_rows +
// copy_validity_to_rows + copy_strings_to_rows, RC:574-861).  The mirror of from_rows_wide.cu + strings_wide_kernel:
// the fixed section of the rows and their chars are written by two kernels, each shaped for its own access pattern,
// instead of one kernel that assembles whole ~4 KB rows (to_rows3_kernel: 48 % of the HBM peak on config C3, bound by
// CTA-wide barriers and bank conflicts).
//
//   to_rows_wide_fixed_kernel : a tile = 32 rows x one slab of the fixed section, one shared-memory slot per row.
//       filler warps : lane = row.  A unit = one fixed-width field: a coalesced ld.global of 32 consecutive column
//                      values (eight units in flight per warp: the loads are DRAM round trips), st.shared into the
//                      rows' slots.  Validity: lane = column loads the mask word of the 32 rows, the 32x32 bit
//                      butterfly turns it into 4 validity bytes per row.  STRING columns: the length word of the
//                      (offset, len) pair, from the column offsets.
//       storer warps : a warp owns a row: lane = STRING column scans the row's length words into the pairs' offset
//                      words (size_per_row + the lengths of the preceding STRING columns, RC:842-858), then the warp
//                      writes the row's slab with coalesced 8-byte stores (32 lanes x 8 contiguous bytes) to
//                      out_data + offsets[r] + slab.begin.  Slots are zeroed once per CTA: alignment gaps stay 0.
//       Fillers and storers meet at full / empty mbarriers of a two-stage ring: a stage is stored while the next fills.
//   to_rows_wide_chars_kernel : a tile = 32 rows, owned by a group of up to 8 warps; warp i scatters the STRING
//       columns [i cpw, (i + 1) cpw).  Per column the tile's chars -- one contiguous range of the column -- are pulled
//       into the warp's staging area with 16-byte cp.async (all columns of the warp in flight together, no
//       registers); lane = row then moves its string as aligned 32-bit
//       words, funnel-shifted, into its row's variable-section image (copy_shared_to_staging of strings.cu, the other
//       way round).  After a group barrier the images leave with coalesced 8-byte stores to
//       out_data + offsets[r] + size_per_row; the <= 7 padding bytes of every row are zeroed on the way.
//
// Row sizes / batch cut / LIST offsets are computed before (row_sizes_kernel + scan + batch_cut_kernel +
// batch_offsets_kernel, to_rows.cu).  A tile whose chars do not fit its image buffer raises *fail_flag; the generic
// kernel launched right behind redoes the batch when it sees the flag.
#include <algorithm>
#include <cstdio>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.hpp"
#include "movers.cuh"
#include "plan.hpp"

namespace srj {

// ======================================== fixed sections =============================================================
constexpr int kTwfMaxFill   = 12;  // filler warps (p.nfill <= this)
constexpr int kTwfStore     = 4;   // storer warps: each writes the STRING pairs of a quarter of the columns, then stores 8 rows
constexpr int kTwfMaxStages = 8;   // ring of (tile, slab) stages: as many as shared memory takes (>= 2)
constexpr int kTwfMaxThreads = (kTwfMaxFill + kTwfStore) * 32;
constexpr int kTwfClasses   = 7;   // unit classes of a slab: 16, 8, 4, 2, 1-byte fields, STRING columns, zero pieces

#ifdef SRJ_DEV_KNOBS
__device__ __forceinline__ unsigned long long gtime()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned long long g_twf_trace[12 * 8];
#define TWF_TRACE(tag, itv)                                                                              \
  if (p.trace && blockIdx.x == 3 && lane == 0 && (warp_id() == 0 || warp_id() == F) && (itv) >= 40 && (itv) < 52) \
    g_twf_trace[((itv) - 40) * 8 + (tag)] = gtime();
#else
#define TWF_TRACE(tag, itv)
#endif

struct TwfHdr {
  int64_t r0;
  int32_t rows;  // 0 = end
  int32_t slab;
};

struct TwfParams {
  const void* const* col_data;       // device [ncols]
  const uint32_t* const* masks;      // device [ncols] (entries may be NULL: all valid)
  const int32_t* const* str_offsets; // device [nstr]
  const int32_t* out_offsets;        // batch-relative LIST offsets [row_count + 1]
  uint8_t* out_data;
  int64_t row_start, row_count;
  int32_t ncols, nstr, size_per_row, validity_offset;
  int32_t trace, pf_tiles;
  int32_t pitch, nslabs, nent, nfill, nstages, so_span;   // so_span: bytes of a stage's offsets table
  const WideEntry* entries;          // to_rows units, sorted by (slab, class)
  const WideSlabTr* slabs;           // store ranges [begin, end) of the slabs (a partition of [0, size_per_row))
};

struct __align__(16) TwfDesc {
  const uint8_t* src;  // column data, or the int32 offsets of a STRING column
  int32_t rel;         // field start relative to the slab's begin
  int32_t aux;         // STRING: index among the STRING columns; zero piece: its size
};

template <int W>
__device__ __forceinline__ Reg<W> ldg_elem(const uint8_t* p)
{
  Reg<W> r;
  if constexpr (W == 1) {
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r.v[0]) : "l"(p));
  } else if constexpr (W == 2) {
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(r.v[0]) : "l"(p));
  } else if constexpr (W == 4) {
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.v[0]) : "l"(p));
  } else if constexpr (W == 8) {
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
  }
  return r;
}

template <int W>
__device__ __forceinline__ void sts_elem(uint32_t a, const Reg<W>& r)
{
  if constexpr (W == 1) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(r.v[0]) : "memory");
  } else if constexpr (W == 2) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(r.v[0]) : "memory");
  } else if constexpr (W == 4) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(r.v[0]) : "memory");
  } else {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(r.v[0]), "r"(r.v[1]) : "memory");
  }
}

__device__ __forceinline__ void lds_desc(uint32_t a, const uint8_t*& src, int32_t& rel)
{
  uint32_t x, y, z;
  asm volatile("{ .reg .b32 t; ld.shared.v4.u32 {%0, %1, %2, t}, [%3]; }" : "=r"(x), "=r"(y), "=r"(z) : "r"(a));
  src = reinterpret_cast<const uint8_t*>(static_cast<uint64_t>(x) | (static_cast<uint64_t>(y) << 32));
  rel = static_cast<int32_t>(z);
}

// global -> shared without registers, 4 or 8 bytes per lane (row images are only 8-byte aligned: a 16-byte field
// moves as two halves)
template <int W>
__device__ __forceinline__ void cp_async_elem(uint32_t dst_s, const uint8_t* src)
{
  if constexpr (W == 8) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_s), "l"(src) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s), "l"(src) : "memory");
  }
}
// the mbarrier gets one arrival from this thread once all its cp.async so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A filler warp moves its fields of a slab in ROUNDS: the loads of up to kN16 + kN8 + kN4 fields (this warp's units
// first, first + F, ... of each width class) are issued back to back, then everything is stored into the rows' slots --
// one DRAM round trip per round, and one round per (tile, slab) for schemas up to kN x F fields per width (config C3
// with 12 filler warps).  cp.async would need no registers, but a copy whose 32 lanes land in 32 different rows costs
// ~20 LSU cycles per instruction whatever its size; ld.global + st.shared cost 5-12.
// Loads are never predicated (a slot without a unit repeats the class's last unit, lanes past the tile's rows read
// its first row): the values stay in registers and the loads are not serialised by branches; only the stores are guarded.
constexpr int kN16 = 6, kN8 = 6, kN4 = 6;
constexpr int kFillBatch = 8;

template <int W, int N>
__device__ __forceinline__ void round_load(Reg<W> (&v)[N], uint32_t desc_s, int first, int end, int F, int64_t row)
{
  if (first >= end) return;   // warp-uniform: no unit of this class in this round
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const uint8_t* src;
    int32_t rel;
    lds_desc(desc_s + 16u * tmin(first + j * F, end - 1), src, rel);
    v[j] = ldg_elem<W>(src + row * W);
  }
}

template <int W, int N>
__device__ __forceinline__ void round_store(const Reg<W> (&v)[N], uint32_t desc_s, int first, int end, int F, uint32_t slot_s, bool act)
{
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int u = first + j * F;
    if (u < end && act) {
      const uint32_t a = slot_s + lds_u32(desc_s + 16u * u + 8);
      if constexpr (W == 16) {   // row images are only 8-byte aligned: two halves
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v[j].v[0]), "r"(v[j].v[1]) : "memory");
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a + 8), "r"(v[j].v[2]), "r"(v[j].v[3]) : "memory");
      } else {
        sts_elem<W>(a, v[j]);
      }
    }
  }
}

// STRING columns: the tile's rows + 1 entries of the column's offsets go to the stage's offsets table,
// so[(u - first STRING unit of the slab) * 33 + r] = offsets[row0 + r] -- consecutive lanes, consecutive words
__device__ __forceinline__ void fill_offsets(uint32_t desc_s, int first, int end, int F, int cb5, uint32_t so_s, int64_t row0, int lane, int rows)
{
  for (int u = first; u < end; u += kFillBatch * F) {
    const uint8_t* src[kFillBatch];
    int32_t rel[kFillBatch];
#pragma unroll
    for (int j = 0; j < kFillBatch; ++j) lds_desc(desc_s + 16u * tmin(u + j * F, end - 1), src[j], rel[j]);
#pragma unroll
    for (int j = 0; j < kFillBatch; ++j)
      if (u + j * F < end) {
        const uint32_t d = so_s + static_cast<uint32_t>(((u + j * F - cb5) * 33 + lane) * 4);
        if (lane < rows) cp_async_elem<4>(d, src[j] + (row0 + lane) * 4);
        if (lane == rows - 1) cp_async_elem<4>(d + 4, src[j] + (row0 + lane + 1) * 4);
      }
  }
}

// 1- and 2-byte fields go through registers (cp.async moves >= 4 bytes), two loads in flight
template <int W>
__device__ __forceinline__ void fill_small(uint32_t desc_s, int first, int end, int F, uint32_t slot_s, int64_t row)
{
  for (int u = first; u < end; u += 2 * F) {
    const uint8_t *s0, *s1;
    int32_t r0, r1;
    lds_desc(desc_s + 16u * u, s0, r0);
    lds_desc(desc_s + 16u * tmin(u + F, end - 1), s1, r1);
    const Reg<W> a = ldg_elem<W>(s0 + row * W);
    const Reg<W> b = ldg_elem<W>(s1 + row * W);
    sts_elem<W>(slot_s + static_cast<uint32_t>(r0), a);
    if (u + F < end) sts_elem<W>(slot_s + static_cast<uint32_t>(r1), b);
  }
}

// zero pieces: the alignment gaps between the fields (a slot holds the image at +0 or +8 depending on the row's
// address: what a gap covers was a field of the slot's previous tile)
__device__ __forceinline__ void fill_zero(uint32_t desc_s, int first, int end, int F, uint32_t slot_s)
{
  for (int u = first; u < end; u += F) {
    uint32_t rel, sz;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rel), "=r"(sz) : "r"(desc_s + 16u * u + 8));
    const uint32_t a = slot_s + rel;
    if (sz == 8) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(a), "r"(0u) : "memory");
    else if (sz == 4) sts_u32(a, 0u);
    else if (sz == 2) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(0u) : "memory");
    else sts_u8(a, 0u);
  }
}

__global__ void __launch_bounds__(kTwfMaxThreads, 1) to_rows_wide_fixed_kernel(const __grid_constant__ TwfParams p)
{
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NC     = kTwfClasses;
  const int F          = p.nfill;
  const int NS         = p.nstages;
  const int stage_span = 32 * p.pitch;
  uint8_t* payload0    = smem;
  TwfDesc* s_desc      = reinterpret_cast<TwfDesc*>(smem + static_cast<size_t>(NS) * stage_span);
  int32_t* s_off       = reinterpret_cast<int32_t*>(s_desc + p.nent);                     // [stages][32] row offsets
  TwfHdr* hdr0         = reinterpret_cast<TwfHdr*>(s_off + NS * 32);
  uint64_t* full       = reinterpret_cast<uint64_t*>(hdr0 + NS);
  uint64_t* empty      = full + NS;
  WideSlabTr* s_slabs  = reinterpret_cast<WideSlabTr*>(empty + NS);
  const uint32_t** s_masks = reinterpret_cast<const uint32_t**>((reinterpret_cast<uintptr_t>(s_slabs + p.nslabs) + 7) & ~uintptr_t{7});
  int16_t* s_first     = reinterpret_cast<int16_t*>(s_masks + p.ncols);                   // [nslabs][NC][F]
  int32_t* s_so        = reinterpret_cast<int32_t*>((reinterpret_cast<uintptr_t>(s_first + p.nslabs * NC * F) + 15) & ~uintptr_t{15});
  const uint32_t so_s  = smem_u32(s_so);   // [stages][max STRING columns of a slab][33] column offsets of the tile's rows

  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < p.nslabs; i += nthr) s_slabs[i] = p.slabs[i];
  for (int i = tid; i < p.nent; i += nthr) {
    const WideEntry e = p.entries[i];
    TwfDesc d;
    d.src     = e.column < 0 ? nullptr
                             : (e.sidx >= 0 ? reinterpret_cast<const uint8_t*>(p.str_offsets[e.sidx]) : static_cast<const uint8_t*>(p.col_data[e.column]));
    d.rel     = e.start - p.slabs[e.slab].begin;
    d.aux     = e.sidx;   // zero piece (column < 0): its size
    s_desc[i] = d;
  }
  for (int i = tid; i < p.ncols; i += nthr) s_masks[i] = p.masks[i];
  for (int i = tid; i < p.nslabs * NC * F; i += nthr) {
    // unit u of a slab goes to filler warp (u - first unit of the slab) % F: the first unit of warp w in class k
    const int sb = i / (NC * F), k = (i / F) % NC, w = i % F;
    const int cb = p.slabs[sb].cb[k], cb0 = p.slabs[sb].cb[0];
    s_first[i]   = static_cast<int16_t>(cb + (w + F - (cb - cb0) % F) % F);
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], F * 33);   // per filler lane: its cp.async have landed; per filler warp: its st.shared are done
      mbar_init(&empty[s], kTwfStore);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int lane        = lane_id();
  const int64_t ntiles  = (p.row_count + 31) >> 5;
  const uint32_t desc_s = smem_u32(s_desc);
  const uint32_t odd    = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p.out_data)) & 8u;   // out_data is 8-byte aligned
  if (warp_id() < F) {
    // =================================== fillers ===================================
    const int w  = warp_id();
    const int nq = (p.ncols + 31) >> 5;
    const int q0 = F - 1 - w;   // validity quads from the last warp down: the first warps hold the remainder units
    // what a tile needs before anything can be issued -- the rows' offsets (where the image sits in its slot) and
    // this warp's validity word -- is fetched one tile ahead
    auto prefetch = [&](int64_t T, uint32_t& off, uint32_t& word) {
      const int64_t r0 = T << 5;
      off = word = 0;
      if (T < ntiles) {
        if (r0 + lane < p.row_count) off = static_cast<uint32_t>(p.out_offsets[r0 + lane]);
        if (q0 < nq && q0 * 32 + lane < p.ncols) {
          const uint32_t* m = s_masks[q0 * 32 + lane];
          word              = m ? __ldg(m + ((p.row_start + r0) >> 5)) : 0xffffffffu;   // NULL mask = all valid (RC:757-759)
        }
      }
    };
    // a CTA takes a run of consecutive tiles: what it reads of a column is one contiguous range, pulled into L2 kPfTiles
    // tiles at a time (one prefetch instruction per column: DRAM sees 0.5 - 2 KB bursts instead of 128 - 512 bytes)
    const int64_t tpc = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t T0 = blockIdx.x * tpc, T1 = tmin<int64_t>(ntiles, T0 + tpc);
    uint32_t off_c, word_c, off_n, word_n;
    prefetch(T0, off_c, word_c);
    int it = 0;
    for (int64_t T = T0; T < T1; ++T) {
      prefetch(T + 1 < T1 ? T + 1 : ntiles, off_n, word_n);
      if (p.pf_tiles > 0 && ((T - T0) % p.pf_tiles) == 0) {
        const int64_t rowp  = p.row_start + ((T + (T == T0 ? 0 : p.pf_tiles)) << 5);     // first row of the window
        const int64_t rowe  = p.row_start + tmin<int64_t>(p.row_count, (tmin<int64_t>(T1, T + 2 * p.pf_tiles)) << 5);
        for (int sb = 0; sb < p.nslabs; ++sb) {
          const WideSlabTr sl = s_slabs[sb];
          for (int u = sl.cb[0] + w; u < sl.cb[6]; u += F) {
            const int W = u < sl.cb[1] ? 16 : u < sl.cb[2] ? 8 : u < sl.cb[3] ? 4 : u < sl.cb[4] ? 2 : u < sl.cb[5] ? 1 : 4;
            const uint8_t* src;
            int32_t rel;
            lds_desc(desc_s + 16u * u, src, rel);
            const uint8_t* a = src + rowp * W + 128 * lane;
            if (a < src + rowe * W) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            if (T == T0 && a + 4096 < src + rowe * W) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 4096));
          }
        }
      }
      const int64_t r0      = T << 5;
      const int rows        = static_cast<int>(tmin<int64_t>(32, p.row_count - r0));
      const bool act        = lane < rows;
      const int64_t abs_row = p.row_start + r0 + (act ? lane : 0);
      // the image of a row whose address is 8 mod 16 sits at +8 in its slot: slot and row are congruent mod 16,
      // which is what lets a bulk copy take the row out
      const uint32_t skew = (off_c + odd) & 8u;
      for (int sb = 0; sb < p.nslabs; ++sb, ++it) {
        const int s        = it % NS;
        const uint32_t par = ((it / NS) & 1) ^ 1;
        if (lane == 0) mbar_wait(&empty[s], par);
        __syncwarp();
        TWF_TRACE(0, it);
        const WideSlabTr sl   = s_slabs[sb];
        const uint32_t slot_s = smem_u32(payload0 + static_cast<size_t>(s) * stage_span) + static_cast<uint32_t>(lane * p.pitch) + skew;
        const int16_t* fu     = s_first + (sb * NC) * F + w;
        if (w == F - 1) {
          if (lane == 0) {
            TwfHdr* h = hdr0 + s;
            h->r0     = r0;
            h->rows   = rows;
            h->slab   = sb;
          }
          s_off[s * 32 + lane] = static_cast<int32_t>(off_c);
        }
        fill_offsets(desc_s, fu[5 * F], sl.cb[6], F, sl.cb[5], so_s + static_cast<uint32_t>(s * p.so_span), p.row_start + r0, lane, rows);
        cp_async_arrive(&full[s]);
        TWF_TRACE(1, it);
        // 1- and 2-byte fields, the zero pieces and the validity bytes (st.shared work that waits for no load of this tile)
        auto misc = [&]() {
          if (act) {
            fill_small<2>(desc_s, fu[3 * F], sl.cb[4], F, slot_s, abs_row);
            fill_small<1>(desc_s, fu[4 * F], sl.cb[5], F, slot_s, abs_row);
            fill_zero(desc_s, fu[6 * F], sl.cb[7], F, slot_s);
          }
          if (sb == p.nslabs - 1) {
            // validity: lane = column holds the mask word of the 32 rows, the butterfly hands every row its 4 bytes
            const int nvb  = (p.ncols + 7) >> 3;
            const int vrel = p.validity_offset - sl.begin;
            for (int q = q0; q < nq; q += F) {
              const int col = q * 32 + lane;
              uint32_t word = word_c;
              if (q != q0) {
                word = 0;
                if (col < p.ncols) {
                  const uint32_t* m = s_masks[col];
                  word              = m ? __ldg(m + ((p.row_start + r0) >> 5)) : 0xffffffffu;
                }
              }
              const uint32_t mine = transpose32(word, lane);   // lane = row, bit i = column 32 q + i
              if (act) {
                const int nbv    = tmin(4, nvb - 4 * q);
                const uint32_t a = slot_s + static_cast<uint32_t>(vrel + 4 * q);
                if (nbv == 4 && (vrel & 3) == 0) sts_u32(a, mine);
                else
                  for (int i = 0; i < nbv; ++i) sts_u8(a + i, mine >> (8 * i));
              }
            }
          }
        };
        {
          // fields of 4, 8 and 16 bytes: rounds of loads, then stores
          const int n16 = sl.cb[1] - sl.cb[0], n8 = sl.cb[2] - sl.cb[1], n4 = sl.cb[3] - sl.cb[2];
          const int nrounds = tmax(tmax((n16 + kN16 * F - 1) / (kN16 * F), (n8 + kN8 * F - 1) / (kN8 * F)), (n4 + kN4 * F - 1) / (kN4 * F));
          if (nrounds == 0) misc();
          for (int r = 0; r < nrounds; ++r) {
            const int f16 = fu[0 * F] + r * kN16 * F, f8 = fu[1 * F] + r * kN8 * F, f4 = fu[2 * F] + r * kN4 * F;
            Reg<16> v16[kN16];
            Reg<8> v8[kN8];
            Reg<4> v4[kN4];
            round_load<16, kN16>(v16, desc_s, f16, sl.cb[1], F, abs_row);
            round_load<8, kN8>(v8, desc_s, f8, sl.cb[2], F, abs_row);
            round_load<4, kN4>(v4, desc_s, f4, sl.cb[3], F, abs_row);
            if (r == 0) misc();   // while the loads are in flight
            round_store<16, kN16>(v16, desc_s, f16, sl.cb[1], F, slot_s, act);
            round_store<8, kN8>(v8, desc_s, f8, sl.cb[2], F, slot_s, act);
            round_store<4, kN4>(v4, desc_s, f4, sl.cb[3], F, slot_s, act);
          }
        }
        fence_proxy_async();   // this lane's st.shared -> visible to the storers' bulk copies
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        TWF_TRACE(2, it);
      }
      off_c  = off_n;
      word_c = word_n;
    }
    // end marker
    const int s        = it % NS;
    const uint32_t par = ((it / NS) & 1) ^ 1;
    if (lane == 0) {
      mbar_wait(&empty[s], par);
      if (w == F - 1) hdr0[s].rows = 0;
    }
    __syncwarp();
    cp_async_arrive(&full[s]);
    if (lane == 0) mbar_arrive(&full[s]);
  } else {
    // =================================== storers ===================================
    const int j  = warp_id() - F;
    uint32_t run = 0;   // lane = row: where the next chars of the row go (kept across the slabs of a tile)
    for (int it = 0;; ++it) {
      const int s        = it % NS;
      const uint32_t par = (it / NS) & 1;
      mbar_wait(&full[s], par);
      TWF_TRACE(3, it);
      const TwfHdr h = hdr0[s];
      if (h.rows == 0) break;
      const WideSlabTr sl = s_slabs[h.slab];
      const uint32_t st_s = smem_u32(payload0 + static_cast<size_t>(s) * stage_span);
      if (h.slab == 0) run = static_cast<uint32_t>(p.size_per_row);
      const int ns = sl.cb[6] - sl.cb[5];
      if (ns > 0) {
        // The (offset, len) pairs of the slab's STRING columns, offset = size_per_row + the lengths of the row's
        // preceding STRING columns (RC:842-858).  lane = row; every storer warp takes ceil(ns / 4) consecutive columns
        // and first adds up the lengths of the columns before its own (conflict-free reads of the stage's offsets
        // table: consecutive lanes, consecutive words).
        const int cpw       = (ns + kTwfStore - 1) / kTwfStore;
        const int k0        = tmin(ns, j * cpw), k1 = tmin(ns, k0 + cpw);
        const uint32_t so   = so_s + static_cast<uint32_t>(s * p.so_span + lane * 4);
        const uint32_t slot = st_s + static_cast<uint32_t>(lane * p.pitch) + ((static_cast<uint32_t>(s_off[s * 32 + lane]) + odd) & 8u);
        uint32_t off        = run;
        {
          uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;   // independent chains: the loads pipeline
          int k = 0;
          for (; k + 4 <= k0; k += 4) {
            a0 += lds_u32(so + 132u * k + 4) - lds_u32(so + 132u * k);
            a1 += lds_u32(so + 132u * k + 136) - lds_u32(so + 132u * k + 132);
            a2 += lds_u32(so + 132u * k + 268) - lds_u32(so + 132u * k + 264);
            a3 += lds_u32(so + 132u * k + 400) - lds_u32(so + 132u * k + 396);
          }
          for (; k < k0; ++k) a0 += lds_u32(so + 132u * k + 4) - lds_u32(so + 132u * k);
          off += (a0 + a1) + (a2 + a3);
        }
        if (lane < h.rows)
          for (int k = k0; k < k1; ++k) {
            const uint32_t len = lds_u32(so + 132u * k + 4) - lds_u32(so + 132u * k);
            const uint32_t a   = slot + lds_u32(desc_s + 16u * (sl.cb[5] + k) + 8);
            sts_u32(a, off);
            sts_u32(a + 4, len);
            off += len;
          }
        if (p.nslabs > 1) {   // the row's chars so far, for the next slab
          for (int k = k1; k < ns; ++k) off += lds_u32(so + 132u * k + 4) - lds_u32(so + 132u * k);
          run = off;
        }
        fence_proxy_async();
        named_bar_sync(1, kTwfStore * 32);   // the pairs are in the slots
      }
      // ---- rows out: warp j owns the rows [8 j, 8 j + 8).  A bulk copy takes [16-byte aligned start, 16-byte aligned
      // end) of the row's slab; the <= 8 bytes before and < 16 bytes after go by hand.  Four lanes per row, one per
      // piece, so that the pieces' latencies overlap.
      TWF_TRACE(4, it);
      fence_proxy_async();   // (the fillers' cp.async data, acquired through the mbarrier)
      const int row  = j * 8 + (lane & 7);
      const int part = lane >> 3;
      if (row < h.rows) {
        const uint32_t offr = static_cast<uint32_t>(s_off[s * 32 + row]);
        const uint32_t skew = (offr + odd) & 8u;
        const uint32_t img  = st_s + static_cast<uint32_t>(row * p.pitch) + skew;
        uint8_t* g          = p.out_data + static_cast<int64_t>(offr) + sl.begin;
        const int len       = sl.end - sl.begin;
        const int head      = static_cast<int>(skew);   // 8: the row starts 8 bytes before a 16-byte boundary
        const int core      = (len - head) & ~15;
        const int rest      = head + core;
        if (part == 0 && head) {
          uint32_t a, b;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(img));
          asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(g), "r"(a), "r"(b));
        } else if (part == 1 && core > 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + rest - core), "r"(img + rest - core), "r"(core) : "memory");
        } else if (part == 2 && len - rest >= 8) {
          uint32_t a, b;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(img + rest));
          asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(g + rest), "r"(a), "r"(b));
        } else if (part == 3) {
          for (int d = rest + ((len - rest) & 8); d < len; ++d) {
            uint32_t v;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(img + d));
            asm volatile("st.global.u8 [%0], %1;" ::"l"(g + d), "r"(v));
          }
        }
      }
      tma_store_commit();
      TWF_TRACE(5, it);
      tma_store_wait_read<0>();   // the bulk copies have read their rows: the stage may be refilled
      TWF_TRACE(6, it);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    tma_store_wait_all<0>();
#ifdef SRJ_DEV_KNOBS
    if (p.trace && blockIdx.x == 3 && j == 0 && lane == 0)
      for (int i = 0; i < 12; ++i)
        printf("TWF it%d  empty %6llu issued %6llu arrived %6llu | full %6llu pairs %6llu stores %6llu read %6llu\n", 40 + i,
               g_twf_trace[i * 8 + 0] - g_twf_trace[0], g_twf_trace[i * 8 + 1] - g_twf_trace[0], g_twf_trace[i * 8 + 2] - g_twf_trace[0],
               g_twf_trace[i * 8 + 3] - g_twf_trace[0], g_twf_trace[i * 8 + 4] - g_twf_trace[0], g_twf_trace[i * 8 + 5] - g_twf_trace[0],
               g_twf_trace[i * 8 + 6] - g_twf_trace[0]);
#endif
  }
}

// ======================================== chars ====================================================================
constexpr int kTwcNG      = 3;   // tiles in flight per CTA
constexpr int kTwcMaxWpt  = 8;   // warps per tile
constexpr int kTwcMaxCpw  = 8;   // STRING columns per warp
constexpr int kTwcFront   = 16;  // slack before the first row image / staging area (word reads may start 7 bytes early)
constexpr int kTwcBack    = 64;  // slack after (word reads run up to 44 bytes past a string)
constexpr int kTwcMinArea = 16 + 1024 + kTwcBack;   // a column of 32 strings of <= 32 bytes always fits
constexpr int kTwcMaxThreads = kTwcNG * kTwcMaxWpt * 32;

struct TwcParams {
  const int32_t* const* str_offsets;  // device [nstr]
  const uint8_t* const* str_chars;    // device [nstr]
  const int32_t* out_offsets;
  uint8_t* out_data;
  int64_t row_start, row_count;
  int32_t nstr, size_per_row, wpt, cpw;
  int32_t image_bytes;   // row images of a tile (multiple of 16)
  int32_t area_bytes;    // staging area of a warp (multiple of 16, >= kTwcMinArea)
  int32_t* fail_flag;
};

__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void* src, int src_bytes)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_s), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(kTwcMaxThreads, 1) to_rows_wide_chars_kernel(const __grid_constant__ TwcParams p)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int wpt        = p.wpt;
  const int image_span = kTwcFront + p.image_bytes + kTwcBack;
  uint8_t* image0      = smem;                                                                          // [kTwcNG][image_span]
  int32_t* blk0        = reinterpret_cast<int32_t*>(smem + static_cast<size_t>(kTwcNG) * image_span);   // [kTwcNG][kTwcMaxWpt][32]
  const int32_t** s_so = reinterpret_cast<const int32_t**>(blk0 + kTwcNG * kTwcMaxWpt * 32);
  const uint8_t** s_ch = reinterpret_cast<const uint8_t**>(s_so + p.nstr);
  uint8_t* area0       = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_ch + p.nstr) + 15) & ~uintptr_t{15}) + kTwcFront;

  for (int i = threadIdx.x; i < p.nstr; i += blockDim.x) {
    s_so[i] = p.str_offsets[i];
    s_ch[i] = p.str_chars[i];
  }
  __syncthreads();

  const int lane = lane_id();
  const int cw   = warp_id();
  const int gi   = cw / wpt;
  const int wi   = cw - gi * wpt;
  const int c0   = wi * p.cpw;
  const int ncol = tmax(0, tmin(p.nstr, c0 + p.cpw) - c0);
  const uint32_t area_s = smem_u32(area0 + static_cast<size_t>(cw) * p.area_bytes);
  const int area_cap    = p.area_bytes - kTwcBack;   // bytes of the area chunks may occupy
  const uint32_t img_s  = smem_u32(image0 + static_cast<size_t>(gi) * image_span + kTwcFront);
  int32_t* blk          = blk0 + gi * kTwcMaxWpt * 32;
  const int64_t ntiles  = (p.row_count + 31) >> 5;
  const int phase       = p.size_per_row & 7;   // image byte x of a row <-> global byte offsets[r] + size_per_row - phase + x
  const int bar_id      = 1 + gi;
  const int bar_n       = wpt * 32;

  for (int64_t T = static_cast<int64_t>(blockIdx.x) * kTwcNG + gi; T < ntiles; T += static_cast<int64_t>(gridDim.x) * kTwcNG) {
    const int64_t r0      = T << 5;
    const int rows        = static_cast<int>(tmin<int64_t>(32, p.row_count - r0));
    const int last        = rows - 1;
    const bool act        = lane < rows;
    const int64_t abs_row = p.row_start + r0 + lane;
    // ---- this warp's columns: offsets entries -> lengths ------------------------------------------------------------
    int32_t o0[kTwcMaxCpw], len[kTwcMaxCpw];
#pragma unroll
    for (int j = 0; j < kTwcMaxCpw; ++j) {
      o0[j] = len[j] = 0;
      if (j < ncol && act) {
        const int32_t* so = s_so[c0 + j] + abs_row;
        o0[j]             = static_cast<int32_t>(ldg_elem<4>(reinterpret_cast<const uint8_t*>(so)).v[0]);
        len[j]            = static_cast<int32_t>(ldg_elem<4>(reinterpret_cast<const uint8_t*>(so + 1)).v[0]);
      }
    }
    // ---- geometry: the variable section of row r is [offsets[r] + size_per_row, offsets[r + 1]) --------------------
    int64_t ro0 = 0, ro1 = 0;
    if (act) {
      ro0 = static_cast<uint32_t>(p.out_offsets[r0 + lane]);
      ro1 = static_cast<uint32_t>(p.out_offsets[r0 + lane + 1]);
    }
    const int32_t vlen = act ? static_cast<int32_t>(ro1 - ro0 - p.size_per_row) : 0;   // chars + padding
    const int32_t span = (phase + vlen + 7) & ~7;                                       // image bytes of the row
    int32_t x          = span;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    const int32_t slot = x - span;   // 8-byte aligned
    if (__any_sync(0xffffffffu, act && x > p.image_bytes)) {
      // the tile's rows do not fit the image buffer: the generic kernel redoes the batch (every warp of the group
      // takes this branch, the barriers below stay matched)
      if (wi == 0 && lane == 0) atomicExch(p.fail_flag, 1);
      continue;
    }
    int32_t mysum = 0, mymax = 0;
#pragma unroll
    for (int j = 0; j < kTwcMaxCpw; ++j) {
      if (j < ncol && act) len[j] -= o0[j];
      mysum += len[j];
      mymax = tmax(mymax, len[j]);
    }
    const bool longs = __reduce_max_sync(0xffffffffu, mymax) > 32;   // some string of the warp's columns takes the byte-wise path
    blk[wi * 32 + lane] = mysum;
    named_bar_sync(bar_id, bar_n);   // also: the group's stores of its previous tile are done, the images are free
    int32_t run = phase;
    for (int w2 = 0; w2 < wi; ++w2) run += blk[w2 * 32 + lane];
    const uint32_t row_s = img_s + static_cast<uint32_t>(slot);
    // ---- the tile's chars of every column of the warp -> staging area (cp.async), in as few rounds as the area takes
    int jb = 0;
    while (jb < ncol) {   // warp-uniform
      int pos = 0, je = jb;
      int32_t rsum = 0;   // this lane's bytes in the columns [jb, j) of the round
      int32_t apos[kTwcMaxCpw];
#pragma unroll
      for (int j = 0; j < kTwcMaxCpw; ++j) {
        apos[j] = -1;
        if (j >= jb && j < ncol && j == je) {
          const int32_t base = __shfl_sync(0xffffffffu, o0[j], 0);
          const int32_t T2   = __shfl_sync(0xffffffffu, o0[j] + len[j], last) - base;   // chars of the tile in this column
          const int maxL     = longs ? __reduce_max_sync(0xffffffffu, len[j]) : 0;
          const uint8_t* S   = s_ch[c0 + j] + base;
          const int a        = static_cast<int>(reinterpret_cast<uintptr_t>(S) & 15);
          const int need     = (a + T2 + 15) & ~15;
          if (T2 == 0) {
            je = j + 1;   // nothing to move
          } else if (maxL > 32 || T2 > 1024) {
            if (pos == 0) {   // long strings: byte-wise, lane = row, on their own
              if (act)
                for (int32_t i = 0; i < len[j]; ++i) sts_u8(row_s + static_cast<uint32_t>(run + rsum + i), __ldg(S + (o0[j] - base) + i));
              je = j + 1;
            }
          } else if (pos + need <= area_cap) {
            apos[j] = pos + a;   // area byte of the tile's first char of this column
            // 16-byte chunks c of the aligned range; the first may start before the chars buffer (its first tile
            // only): that one is moved byte-wise; the last is cut at the end of the tile's chars (cp.async src-size)
            const uint8_t* Sal = S - a;
            const int nch      = need >> 4;
            const bool head_in = base >= a;   // the chunk's leading bytes are chars of earlier rows of the column
            for (int c = lane; c < nch; c += 32) {
              const int valid = tmin(16, a + T2 - 16 * c);
              if (c > 0 || head_in || a == 0) cp_async16(area_s + pos + 16 * c, Sal + 16 * c, valid);
            }
            if (!head_in && a != 0) {   // warp-uniform, first tile of a column only
              if (lane >= a && lane < tmin(16, a + T2)) sts_u8(area_s + pos + lane, __ldg(Sal + lane));
            }
            pos += need;
            je = j + 1;
          }
        }
        if (j >= jb && j < je) rsum += len[j];
      }
      cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kTwcMaxCpw; ++j) {
        if (j >= jb && j < je) {
          if (apos[j] >= 0) {
            const int32_t base = __shfl_sync(0xffffffffu, o0[j], 0);
            copy_shared_to_staging(area_s + static_cast<uint32_t>(apos[j] + (o0[j] - base)), row_s + static_cast<uint32_t>(run), act ? len[j] : 0);
          }
          run += len[j];
        }
      }
      __syncwarp();   // the area is rewritten by the next round
      jb = je;
    }
    // the last warp of the group knows where the row's chars end: zero the padding up to the row's 8-byte end
    if (wi == wpt - 1 && act)
      for (int32_t i = run; i < phase + vlen; ++i) sts_u8(row_s + static_cast<uint32_t>(i), 0u);
    named_bar_sync(bar_id, bar_n);
    // ---- store: a warp writes a row's variable section with coalesced 8-byte stores ---------------------------------
    for (int row = wi; row < rows; row += wpt) {
      const int32_t rs   = __shfl_sync(0xffffffffu, slot, row);
      const int64_t g0   = __shfl_sync(0xffffffffu, ro0, row) + p.size_per_row - phase;   // 8-byte aligned
      const int32_t vl   = __shfl_sync(0xffffffffu, vlen, row);
      uint8_t* dst       = p.out_data + g0;                                               // image byte x -> dst + x
      const uint32_t src = img_s + static_cast<uint32_t>(rs);
      const int endb     = phase + vl;                                                    // multiple of 8 (rows are padded to 8)
      // bytes [phase, 8) of the first unit when the section does not start on an 8-byte boundary
      if (phase && lane >= phase && lane < tmin(8, endb)) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(src + lane));
        asm volatile("st.global.u8 [%0], %1;" ::"l"(dst + lane), "r"(v));
      }
      for (int u = (phase ? 1 : 0) + lane; 8 * u + 8 <= endb; u += 32) {
        uint32_t a2, b2;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a2), "=r"(b2) : "r"(src + 8u * u));
        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + 8 * u), "r"(a2), "r"(b2));
      }
    }
    // the next tile's first barrier orders these reads before the images are overwritten
  }
}

// ======================================== host side =================================================================
// bytes of one stage's STRING offsets table: [most STRING columns of a slab][33]
static int twf_so_span(const srj_plan* plan)
{
  int most = 0;
  for (const WideSlabTr& sl : plan->wide.tr_slabs) most = std::max(most, sl.cb[6] - sl.cb[5]);
  return (most * 33 * 4 + 15) & ~15;
}

static size_t twf_smem_bytes(const srj_plan* plan, int nstages)
{
  const WidePlan& wp = plan->wide;
  size_t b = static_cast<size_t>(nstages) * 32 * wp.tr_pitch;
  b += wp.tr_entries.size() * sizeof(TwfDesc) + nstages * 32 * 4 + nstages * sizeof(TwfHdr) + 2 * nstages * 8;
  b += wp.tr_slabs.size() * sizeof(WideSlabTr) + 8 + static_cast<size_t>(plan->num_columns) * 8;
  b += wp.tr_slabs.size() * kTwfClasses * kTwfMaxFill * 2 + 16;
  b += static_cast<size_t>(nstages) * twf_so_span(plan) + 16;
  return (b + 127) & ~size_t{127};
}

// slabs of at most ~cap bytes: cuts at multiples of 16 that split no field
static void plan_wide_to_rows_cap(srj_plan* plan, int cap)
{
  WidePlan& wp = plan->wide;
  wp.tr_entries.clear();
  wp.tr_slabs.clear();
  const int nc = plan->num_columns, nstr = plan->num_string_columns;
  const int spr  = plan->size_per_row;
  const int want = std::max(1, std::min((spr + cap - 1) / cap, 16));
  std::vector<int> cut{0};
  for (int i = 1; i < want; ++i) {
    int x = static_cast<int>(static_cast<int64_t>(i) * spr / want) & ~15;
    for (; x > cut.back(); x -= 16) {
      bool split = x >= plan->validity_offset;   // the validity bytes stay in one slab (the last)
      for (int k = 0; k < nc && !split; ++k) split = plan->col_start[k] < x && plan->col_start[k] + plan->col_size[k] > x;
      if (!split) break;
    }
    if (x > cut.back()) cut.push_back(x);
  }
  cut.push_back(spr);
  const int nslabs = static_cast<int>(cut.size()) - 1;
  std::vector<int> sidx_of(nc, -1);
  for (int s = 0; s < nstr; ++s) sidx_of[plan->string_columns[s]] = s;
  // the bytes no field covers (alignment gaps), as naturally aligned pieces of 8, 4, 2, 1 bytes
  std::vector<uint8_t> covered(static_cast<size_t>(spr), 0);
  for (int c = 0; c < nc; ++c) std::fill(covered.begin() + plan->col_start[c], covered.begin() + plan->col_start[c] + plan->col_size[c], 1);
  std::fill(covered.begin() + plan->validity_offset, covered.end(), 1);
  std::vector<std::pair<int, int>> zeros;   // (start, size)
  for (int x = 0; x < spr;) {
    if (covered[x]) { ++x; continue; }
    int sz = 8;
    while (sz > 1 && ((x & (sz - 1)) != 0 || x + sz > spr || std::count(covered.begin() + x, covered.begin() + x + sz, 0) != sz)) sz >>= 1;
    zeros.emplace_back(x, sz);
    x += sz;
  }
  int maxlen = 0;
  for (int i = 0; i < nslabs; ++i) {
    WideSlabTr sl{};
    sl.begin = cut[i];
    sl.end   = cut[i + 1];
    maxlen   = std::max(maxlen, sl.end - sl.begin);
    for (int k = 0; k < 6; ++k) {
      sl.cb[k] = static_cast<int32_t>(wp.tr_entries.size());
      for (int c = 0; c < nc; ++c) {
        if (plan->col_start[c] < sl.begin || plan->col_start[c] >= sl.end) continue;
        const bool str = sidx_of[c] >= 0;
        const int want_sz = k < 5 ? (16 >> k) : 0;
        if (str ? (k == 5) : (k < 5 && plan->col_size[c] == want_sz)) wp.tr_entries.push_back(WideEntry{plan->col_start[c], c, sidx_of[c], i});
      }
    }
    sl.cb[6] = static_cast<int32_t>(wp.tr_entries.size());
    for (const auto& z : zeros)
      if (z.first >= sl.begin && z.first < sl.end) wp.tr_entries.push_back(WideEntry{z.first, -1, z.second, i});
    sl.cb[7] = static_cast<int32_t>(wp.tr_entries.size());
    wp.tr_slabs.push_back(sl);
  }
  // a slot holds a slab image at +0 or +8; consecutive slots an odd multiple of 16 bytes apart (4 banks)
  wp.tr_pitch = (maxlen + 8 + 15) & ~15;
  if (((wp.tr_pitch >> 4) & 1) == 0) wp.tr_pitch += 16;
  wp.tr_enabled = twf_smem_bytes(plan, 2) <= 232448;
}

// to_rows slabs of a wide plan: a partition of [0, size_per_row) at 16-byte aligned cuts that split no field.
bool plan_wide_to_rows(srj_plan* plan)
{
  WidePlan& wp = plan->wide;
  wp.tr_entries.clear();
  wp.tr_slabs.clear();
  wp.tr_enabled = false;
  if (!wp.enabled) return false;
  if (plan->num_string_columns > kTwcMaxWpt * kTwcMaxCpw) return false;
  // slabs as long as shared memory takes (two stages of 32 slab images + the tables)
  for (int cap = SRJ_KNOB("SRJ_TW_SLABCAP", 3200); cap >= 800 && !wp.tr_enabled; cap -= 800) plan_wide_to_rows_cap(plan, cap);
  return wp.tr_enabled;
}

// Returns SRJ_OK and sets *launched when the two kernels were launched (the caller then launches the generic kernel
// guarded by d_fail_flag); *launched = 0 means the table is not eligible.
int launch_to_rows_wide(const srj_plan* plan, const void* const* d_col_data, const uint32_t* const* d_masks,
                        const int32_t* const* d_str_offsets, const uint8_t* const* d_str_chars, int64_t row_start,
                        int64_t row_count, const int32_t* out_offsets, uint8_t* out_data, int64_t out_bytes,
                        int32_t* d_fail_flag, cudaStream_t stream, const void* const* h_col_data, int* launched)
{
  *launched = 0;
  const WidePlan& wp = plan->wide;
  if (!wp.tr_enabled || row_count == 0 || !d_fail_flag || !h_col_data || SRJ_KNOB("SRJ_TW_OFF", 0)) return SRJ_OK;
  if ((reinterpret_cast<uintptr_t>(out_data) & 7) != 0 || (row_start & 31) != 0) return SRJ_OK;
  for (const Entry& e : plan->tr_entries)
    if (reinterpret_cast<uintptr_t>(h_col_data[e.column]) & static_cast<uintptr_t>(plan->col_size[e.column] - 1)) return SRJ_OK;
  int dev = 0, nsm = 0;
  SRJ_CUDA_TRY(cudaGetDevice(&dev));
  SRJ_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ntiles = (row_count + 31) / 32;
  // ---- chars kernel configuration (decides eligibility: nothing is launched when its images do not fit) ----
  TwcParams c{};
  c.str_offsets  = d_str_offsets;
  c.str_chars    = d_str_chars;
  c.out_offsets  = out_offsets;
  c.out_data     = out_data;
  c.row_start    = row_start;
  c.row_count    = row_count;
  c.nstr         = plan->num_string_columns;
  c.size_per_row = plan->size_per_row;
  c.wpt          = std::min(kTwcMaxWpt, (c.nstr + 3) / 4);
  c.cpw          = (c.nstr + c.wpt - 1) / c.wpt;
  c.fail_flag    = d_fail_flag;
  // shared memory: per tile the row images (the variable sections of 32 average rows + 30 %), per warp a staging
  // area for the tile's chars of its columns (+ 25 %; at least one column of 32 x 32 bytes)
  const int nwarps      = kTwcNG * c.wpt;
  const size_t cfixed   = static_cast<size_t>(kTwcNG) * (kTwcFront + kTwcBack) + static_cast<size_t>(kTwcNG) * kTwcMaxWpt * 32 * 4 +
                          static_cast<size_t>(c.nstr) * 16 + 16 + kTwcFront + 256;
  const int64_t avg_var = std::max<int64_t>(0, out_bytes / row_count - plan->size_per_row) + 16;
  int64_t area          = (avg_var * c.cpw / c.nstr * 40 + 32 * c.cpw + kTwcBack + 15) & ~int64_t{15};
  area                  = std::max<int64_t>(kTwcMinArea, std::min<int64_t>(area, SRJ_KNOB("SRJ_TW_AREA", 6144)));
  const int64_t cap     = ((232448 - static_cast<int64_t>(cfixed) - nwarps * area) / kTwcNG) & ~int64_t{15};
  int64_t img           = (avg_var * 32 * 13 / 10 + 1023) & ~int64_t{1023};
  img                   = std::min(img, cap);
  if (img < 2048) return SRJ_OK;   // no room for the images: the other kernels take the table
  c.image_bytes         = static_cast<int32_t>(img);
  c.area_bytes          = static_cast<int32_t>(area);
  const size_t csmem    = (cfixed + static_cast<size_t>(kTwcNG) * img + static_cast<size_t>(nwarps) * area + 127) & ~size_t{127};

  SRJ_CUDA_TRY(cudaMemsetAsync(d_fail_flag, 0, sizeof(int32_t), stream));
  {
    TwfParams p{};
    p.col_data        = d_col_data;
    p.masks           = d_masks;
    p.str_offsets     = d_str_offsets;
    p.out_offsets     = out_offsets;
    p.out_data        = out_data;
    p.row_start       = row_start;
    p.row_count       = row_count;
    p.ncols           = plan->num_columns;
    p.nstr            = plan->num_string_columns;
    p.size_per_row    = plan->size_per_row;
    p.validity_offset = plan->validity_offset;
    p.pitch           = wp.tr_pitch;
    p.nslabs          = static_cast<int32_t>(wp.tr_slabs.size());
    p.nent            = static_cast<int32_t>(wp.tr_entries.size());
    p.so_span         = twf_so_span(plan);
    p.trace           = SRJ_KNOB("SRJ_TW_TRACE", 0);
    p.pf_tiles        = SRJ_KNOB("SRJ_TW_PF", 4);
    p.nfill           = std::max(1, std::min(kTwfMaxFill, SRJ_KNOB("SRJ_TW_FILL", kTwfMaxFill)));
    p.entries         = wp.d_tr_entries;
    p.slabs           = wp.d_tr_slabs;
    p.nstages = 2;
    while (p.nstages < kTwfMaxStages && twf_smem_bytes(plan, p.nstages + 1) <= 232448) ++p.nstages;
    p.nstages         = std::max(2, std::min(p.nstages, SRJ_KNOB("SRJ_TW_STAGES", kTwfMaxStages)));
    const size_t smem = twf_smem_bytes(plan, p.nstages);
    SRJ_CUDA_TRY(cudaFuncSetAttribute(to_rows_wide_fixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    to_rows_wide_fixed_kernel<<<static_cast<unsigned>(std::min<int64_t>(nsm, ntiles)), (p.nfill + kTwfStore) * 32, smem, stream>>>(p);
  }
  {
    const int64_t grid = std::min<int64_t>(nsm, (ntiles + kTwcNG - 1) / kTwcNG);
    SRJ_CUDA_TRY(cudaFuncSetAttribute(to_rows_wide_chars_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    to_rows_wide_chars_kernel<<<static_cast<unsigned>(grid), kTwcNG * c.wpt * 32, csmem, stream>>>(c);
  }
  SRJ_CUDA_TRY(cudaGetLastError());
  *launched = 1;
  return SRJ_OK;
}

}  // namespace srj

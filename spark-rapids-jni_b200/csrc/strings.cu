// strings.cu -- STRING-column stages of convert_from_rows:
//   (1) lengths -> offsets: batched (all STRING columns in one launch) three-step scan replacing the
//       per-column thrust::exclusive_scan + .element() sync loop of the reference (RC:2375-2395);
//   (2) chars gather: replaces copy_strings_from_rows (RC:1110-1150).
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.hpp"
#include "plan.hpp"

namespace srj {

constexpr int kScanThreads = 256;
constexpr int kScanIter    = kScanThreads * 4;  // elements per block iteration
constexpr int kScanChunk   = kScanIter * 4;     // elements per CTA

__device__ __forceinline__ int64_t block_sum(int64_t v, int64_t* s_warp)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane_id() == 0) s_warp[warp_id()] = v;
  __syncthreads();
  int64_t t = 0;
  if (threadIdx.x < kScanThreads / 32) t = s_warp[threadIdx.x];
  if (warp_id() == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane_id() == 0) s_warp[0] = t;
  }
  __syncthreads();
  t = s_warp[0];
  __syncthreads();
  return t;
}

// Element i of the scan input for a column: 0 for i == 0, else the uint32 length stored at offs[i].
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(int32_t* const* offsets, int64_t n1 /* n + 1 */,
                                                                      int nchunks, int64_t* partials)
{
  __shared__ int64_t s_warp[kScanThreads / 32];
  const int c         = blockIdx.y;
  const int k         = blockIdx.x;
  const int32_t* offs = offsets[c];
  const int64_t beg   = static_cast<int64_t>(k) * kScanChunk;
  const int64_t end   = tmin(n1, beg + kScanChunk);
  int64_t acc         = 0;
  for (int64_t i = beg + threadIdx.x; i < end; i += kScanThreads)
    if (i > 0) acc += static_cast<uint32_t>(offs[i]);
  acc = block_sum(acc, s_warp);
  if (threadIdx.x == 0) partials[static_cast<int64_t>(c) * nchunks + k] = acc;
}

// exclusive scan of each column's chunk sums, in place; writes the column total
__global__ void __launch_bounds__(kScanThreads) scan_chunks_kernel(int64_t* partials, int nchunks,
                                                                    const int32_t* string_cols, int64_t* char_totals,
                                                                    int32_t* error)
{
  __shared__ int64_t s_warp[kScanThreads / 32];
  __shared__ int64_t s_carry;
  const int c  = blockIdx.x;
  int64_t* p   = partials + static_cast<int64_t>(c) * nchunks;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += kScanThreads) {
    const int i     = base + threadIdx.x;
    const int64_t v = i < nchunks ? p[i] : 0;
    // block inclusive scan
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane_id() >= o) x += y;
    }
    if (lane_id() == 31) s_warp[warp_id()] = x;
    __syncthreads();
    int64_t wpre = 0;
    for (int w = 0; w < warp_id(); ++w) wpre += s_warp[w];
    const int64_t carry = s_carry;
    if (i < nchunks) p[i] = carry + wpre + x - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) s_carry = carry + wpre + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t total = s_carry;
    if (char_totals) char_totals[string_cols[c]] = total;
    if (total > INT32_MAX) atomicExch(error, SRJ_EOVERFLOW);  // cudf strings offsets are int32
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(int32_t* const* offsets, int64_t n1, int nchunks,
                                                                   const int64_t* partials)
{
  __shared__ int64_t s_warp[kScanThreads / 32];
  __shared__ int64_t s_carry;
  const int c   = blockIdx.y;
  const int k   = blockIdx.x;
  int32_t* offs = offsets[c];
  if (threadIdx.x == 0) s_carry = partials[static_cast<int64_t>(c) * nchunks + k];
  __syncthreads();
  const int64_t beg = static_cast<int64_t>(k) * kScanChunk;
  const int64_t end = tmin(n1, beg + kScanChunk);
  for (int64_t base = beg; base < end; base += kScanIter) {
    const int64_t i0 = base + threadIdx.x * 4;
    int64_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i0 + j;
      v[j]            = (i < end && i > 0) ? static_cast<int64_t>(static_cast<uint32_t>(offs[i])) : 0;
    }
    const int64_t tsum = v[0] + v[1] + v[2] + v[3];
    int64_t x          = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane_id() >= o) x += y;
    }
    if (lane_id() == 31) s_warp[warp_id()] = x;
    __syncthreads();
    int64_t wpre = 0;
    for (int w = 0; w < warp_id(); ++w) wpre += s_warp[w];
    const int64_t carry = s_carry;
    int64_t run         = carry + wpre + x - tsum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += v[j];
      const int64_t i = i0 + j;
      if (i < end) offs[i] = static_cast<int32_t>(run);
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) s_carry = run;
    __syncthreads();
  }
}

int64_t string_scan_partials_bytes(int nstr, int64_t num_rows)
{
  const int64_t nchunks = (num_rows + 1 + kScanChunk - 1) / kScanChunk;
  return static_cast<int64_t>(nstr) * nchunks * 8;
}

int launch_string_offsets_scan(int32_t* const* d_offsets, const int32_t* d_string_cols, int nstr, int64_t num_rows,
                               int64_t* d_char_totals, int32_t* d_error, void* d_partials, cudaStream_t stream)
{
  if (nstr == 0) return SRJ_OK;
  const int64_t n1  = num_rows + 1;
  const int nchunks = static_cast<int>((n1 + kScanChunk - 1) / kScanChunk);
  auto* partials    = static_cast<int64_t*>(d_partials);
  dim3 grid(nchunks, nstr);
  scan_partials_kernel<<<grid, kScanThreads, 0, stream>>>(d_offsets, n1, nchunks, partials);
  scan_chunks_kernel<<<nstr, kScanThreads, 0, stream>>>(partials, nchunks, d_string_cols, d_char_totals, d_error);
  scan_apply_kernel<<<grid, kScanThreads, 0, stream>>>(d_offsets, n1, nchunks, partials);
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

// --------------------------------------------------------------------------------------------------
// chars gather, GENERIC path (follows the stored pair offsets, RC:1143): a warp owns (32-row tile, STRING column); lane = row reads the (offset,len)
// pair from the row, then the warp copies the tile's chars for that column -- a contiguous
// destination range -- with lane = destination byte, locating the source row by a shuffle search.
// Destination stores are fully coalesced; source reads stay inside 2-3 sectors per instruction.
// --------------------------------------------------------------------------------------------------

// ---- shared-memory helpers of the chars gathers (32-bit shared-space addresses) ----------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// A warp's staging line holds the T chars of one (32-row tile, column) at the byte positions [a, a + T), a = the
// destination's offset inside its 16-byte granule.  Flush: whole 16-byte chunks with one ld.shared.v4 /
// st.global.v4 per lane; the bytes of the two partial chunks at the ends (their neighbours belong to other tiles /
// warps) one per lane: lanes 0-15 the head chunk, lanes 16-31 the tail chunk.
__device__ __forceinline__ void flush_staging_line(uint32_t stg_s, uint8_t* D, int a, int T, int lane)
{
  uint8_t* Dal      = D - a;  // 16-byte aligned
  const int aT      = a + T;
  const int c_first = (a + 15) >> 4;  // first whole chunk
  const int c_end   = aT >> 4;        // one past the last whole chunk
  for (int c = c_first + lane; c < c_end; c += 32) {
    uint32_t v0, v1, v2, v3;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(stg_s + 16 * c));
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(Dal + 16 * c), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
  }
  // head chunk = chunk 0 when a > 0; tail chunk = chunk c_end when aT is not a multiple of 16.  When both are the
  // same chunk (c_end == 0) the head lanes cover all of [a, aT).
  const int hl = lane & 15;
  int bpos;
  bool ok;
  if (lane < 16) {
    bpos = hl;                                   // head chunk bytes [a, min(16, aT))
    ok   = hl >= a && hl < tmin(16, aT) && a > 0;
    if (a == 0 && c_end == 0) ok = hl < aT;      // a single partial chunk starting at an aligned byte
  } else {
    bpos = 16 * c_end + hl;                      // tail chunk bytes [16 * c_end, aT)
    ok   = c_end > 0 && bpos < aT;
  }
  if (ok) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(stg_s + bpos));
    asm volatile("st.global.u8 [%0], %1;" ::"l"(Dal + bpos), "r"(v));
  }
}

// lane's string = L bytes at GLOBAL address S -> staging byte ds (shared address).  Aligned 32-bit source words
// (only words that overlap the string are read), funnel shift to the staging alignment, st.shared.u32 for whole
// words; the <= 3 edge bytes at each end come straight from the source.  maxL (<= 32) is the warp's longest string.
__device__ __forceinline__ void copy_global_to_staging(uint64_t S, uint32_t ds, int L, int maxL)
{
  const int dsh      = static_cast<int>(ds & 3u);
  const int ssh      = static_cast<int>(S & 3u);
  const int dlt      = ssh - dsh;
  const int pre      = ssh + (dlt < 0 ? 4 : 0);
  const uint64_t sp  = S - pre;
  const int sh       = (dlt & 3) * 8;
  const int end      = dsh + L;
  const int kfull1   = end >> 2;
  const uint32_t w0s = ds - dsh;
  const int lim      = L > 0 ? L + pre : 0;
  const int Kmax     = (maxL + 6) >> 2;
  const int nh       = dsh ? tmin(L, 4 - dsh) : 0;
  const int nt       = (kfull1 > 0 || !dsh) ? (end & 3) : 0;
  const uint8_t* Sb  = reinterpret_cast<const uint8_t*>(S);
  uint32_t hb[3], tb[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    hb[t] = tb[t] = 0;
    if (t < nh) hb[t] = __ldg(Sb + t);
    if (t < nt) tb[t] = __ldg(Sb + (L - nt) + t);
  }
  uint32_t w[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    w[k]            = 0;
    const bool need = k == 0 ? (lim > 0 && pre < 4) : (4 * k < lim);
    if (need) w[k] = __ldg(reinterpret_cast<const uint32_t*>(sp + 4 * k));
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t < nh) sts_u8(ds + t, hb[t]);
    if (t < nt) sts_u8(ds + (L - nt) + t, tb[t]);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (k < Kmax) {
      const uint32_t y = __funnelshift_r(w[k], w[k + 1], sh);
      const bool full  = k == 0 ? (dsh == 0 && kfull1 > 0) : (k < kfull1);
      if (full) sts_u32(w0s + 4 * k, y);
    }
  }
}

constexpr int kStrWarps = 8;
constexpr int kStrLine  = 16 + 1024 + 32;  // per-warp staging line

__global__ void __launch_bounds__(kStrWarps * 32) strings_from_rows_kernel(
  const uint8_t* __restrict__ rows, const int32_t* __restrict__ row_offsets, int64_t row_stride, int64_t num_rows,
  int nstr, const int32_t* __restrict__ string_start, const int32_t* const* __restrict__ offsets,
  uint8_t* const* __restrict__ chars, int64_t ntiles, const int64_t* __restrict__ status)
{
  // runs only when phase 1 flagged non-canonical rows (or no status word was passed)
  if (status && !(*status & 1)) return;
  __shared__ __align__(16) uint8_t s_line[kStrWarps * kStrLine];
  const int lane       = lane_id();
  const uint32_t stg_s = smem_u32(s_line + warp_id() * kStrLine);
  // one task = (32-row tile, STRING column); tasks are dealt to the warps of the whole grid, so every warp is busy
  // whatever the number of STRING columns (tables with 1-3 strings are the common case)
  const int64_t ntasks = ntiles * nstr;
  const int64_t wstep  = static_cast<int64_t>(gridDim.x) * kStrWarps;
  for (int64_t task = static_cast<int64_t>(blockIdx.x) * kStrWarps + warp_id(); task < ntasks; task += wstep) {
    const int64_t tile = task / nstr;
    const int s        = static_cast<int>(task - tile * nstr);
    const int64_t r    = tile * 32 + lane;
    const bool active  = r < num_rows;
    const int64_t rsta = active ? (row_offsets ? static_cast<int64_t>(row_offsets[r]) : r * row_stride) : 0;
    {
      uint32_t so = 0, len = 0;
      int32_t d0 = 0;
      if (active) {
        const uint8_t* pp = rows + rsta + string_start[s];
        if ((reinterpret_cast<uintptr_t>(pp) & 3) == 0) {
          so  = *reinterpret_cast<const uint32_t*>(pp);
          len = *reinterpret_cast<const uint32_t*>(pp + 4);
        } else {
          so  = pp[0] | (pp[1] << 8) | (pp[2] << 16) | (static_cast<uint32_t>(pp[3]) << 24);
          len = pp[4] | (pp[5] << 8) | (pp[6] << 16) | (static_cast<uint32_t>(pp[7]) << 24);
        }
        d0                = offsets[s][r];
      }
      const int32_t dbase  = __shfl_sync(0xffffffffu, d0, 0);
      const uint32_t pe    = static_cast<uint32_t>(d0 - dbase);  // exclusive prefix inside the tile
      // total = pe + len of the last active lane
      const int last       = static_cast<int>(tmin<int64_t>(31, num_rows - 1 - tile * 32));
      const uint32_t total = __shfl_sync(0xffffffffu, pe + len, last);
      const int64_t srcoff = rsta + so;
      uint8_t* dst         = chars[s] + dbase;
      const int maxL       = __reduce_max_sync(0xffffffffu, active ? static_cast<int>(tmin<uint32_t>(len, 1u << 20)) : 0);
      if (total == 0) continue;
      if (maxL <= 32 && total <= 1024) {
        // short strings (the common case): lane = row copies its string into the warp's staging line, laid out
        // like the destination, and the line leaves with 16-byte stores
        const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
        copy_global_to_staging(reinterpret_cast<uint64_t>(rows) + static_cast<uint64_t>(srcoff), stg_s + static_cast<uint32_t>(a) + pe,
                               active ? static_cast<int>(len) : 0, maxL);
        __syncwarp();
        flush_staging_line(stg_s, dst, a, static_cast<int>(total), lane);
        __syncwarp();
        continue;
      }
      const uint32_t bound = (total + 31u) & ~31u;  // all lanes take part in the shuffles
      for (uint32_t p = lane; p < bound; p += 32) {
        int j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int cand   = j + step;
          const uint32_t v = __shfl_sync(0xffffffffu, pe, cand & 31);
          if (cand <= last && v <= p) j = cand;
        }
        const uint32_t pj  = __shfl_sync(0xffffffffu, pe, j);
        const int64_t sj   = __shfl_sync(0xffffffffu, srcoff, j);
        if (p < total) dst[p] = rows[sj + (p - pj)];
      }
    }
  }
}


// --------------------------------------------------------------------------------------------------
// chars gather, FAST path (canonical rows: a row's chars follow its fixed section in column order,
// which is what convert_to_rows writes and what phase 1 verified).
//
//   producer warp : per tile of <= 32 rows, one TMA bulk load per row of JUST the row's variable
//                   section (lane = row) into a shared-memory ring, plus the [r0, r0+rows] slice of
//                   every STRING column's offsets via 16-byte cp.async (completion on the same
//                   mbarrier) -- so the consumers never wait on a global load;
//   consumer warps: each owns a block of STRING columns.  lane = row: the string's position in the
//                   row is the running sum of the lengths of the preceding STRING columns; it is
//                   read as aligned words, funnel-shifted to the destination's byte alignment and
//                   written word-wise into a per-warp staging line laid out like the destination;
//                   the line (the tile's chars of that column: one contiguous range of the chars
//                   buffer) is flushed with aligned 16-byte st.global.
// --------------------------------------------------------------------------------------------------
constexpr int kS2Consumers  = 19;  // issue/latency-bound gather: more independent warps (80 registers each)
constexpr int kS2Threads    = (kS2Consumers + 1) * 32;
constexpr int kS2Rows       = 32;
constexpr int kS2Stages     = 3;
constexpr int kS2StageBytes = 44 * 1024;
constexpr int kS2Front      = 16;   // slack before the payload (word reads may start 4 bytes early)
constexpr int kS2Back       = 48;   // slack after it (word reads may run past a string)
constexpr int kS2Slice      = 36;   // ints per offsets slice: rows + 1 <= 33, padded to 16-byte chunks
constexpr int kS2StageLine  = 16 + 1024 + 32;  // per-warp staging line
constexpr int kS2MinCols    = 32;   // fewer STRING columns: the generic task-parallel kernel
constexpr int kS2MaxCols    = 160;  // offsets slices must fit shared memory

struct S2Hdr {
  int64_t r0;
  int32_t rows;
  int32_t safe;
};

struct S2Params {
  const uint8_t* rows;
  const int32_t* row_offsets;
  int64_t rows_bytes;
  int64_t num_rows;
  int64_t super_rows;
  int32_t nstr;
  int32_t size_per_row;
  const int32_t* const* offsets;
  uint8_t* const* chars;
  const int64_t* status;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (does not bump the pending count)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


__global__ void __launch_bounds__(kS2Threads, 1) strings2_kernel(const __grid_constant__ S2Params p)
{
  if (p.status && (*p.status & 1)) return;  // non-canonical rows: the generic kernel does the work
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NS        = kS2Stages;
  constexpr int stage_span = kS2Front + kS2StageBytes + kS2Back;
  uint8_t* payload0  = smem;
  int32_t* rowsm0    = reinterpret_cast<int32_t*>(smem + static_cast<size_t>(NS) * stage_span);  // [NS][36]
  int32_t* slice0    = rowsm0 + NS * kS2Slice;                                                    // [NS][nstr][36]
  const int slice_sp = p.nstr * kS2Slice;
  S2Hdr* hdr0        = reinterpret_cast<S2Hdr*>(slice0 + static_cast<size_t>(NS) * slice_sp);
  uint64_t* full     = reinterpret_cast<uint64_t*>(hdr0 + NS);
  uint64_t* empty    = full + NS;
  const int32_t** s_offs = reinterpret_cast<const int32_t**>(empty + NS);
  uint8_t** s_chars      = reinterpret_cast<uint8_t**>(const_cast<int32_t**>(s_offs) + p.nstr);
  uint8_t* stg0          = reinterpret_cast<uint8_t*>(s_chars + p.nstr);
  stg0                   = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stg0) + 15) & ~uintptr_t{15});

  const int tid  = threadIdx.x;
  const int lane = lane_id();
  for (int i = tid; i < p.nstr; i += kS2Threads) {
    s_offs[i]  = p.offsets[i];
    s_chars[i] = p.chars[i];
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);  // lane 0's arrive.expect_tx; every byte of the stage arrives by TMA
      mbar_init(&empty[s], kS2Consumers);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const uintptr_t b_lo = reinterpret_cast<uintptr_t>(p.rows);
  const uintptr_t b_hi = b_lo + p.rows_bytes;

  if (warp_id() == 0) {
    // =================================== producer ===================================
    int64_t sup = blockIdx.x;
    int64_t r   = sup * p.super_rows;
    int64_t c1  = tmin(p.num_rows, r + p.super_rows);
    // per-lane geometry of the NEXT tile (lane = row)
    int g_rows = 0;
    bool g_end = false, g_safe = false;
    uintptr_t a_lo = 0, a_hi = 0, t_lo = 0, t_hi = 0, fl = 0;
    int32_t slot = 0;
    auto next_geometry = [&]() {
      if (r >= c1) {
        sup += gridDim.x;
        r  = sup * p.super_rows;
        c1 = tmin(p.num_rows, r + p.super_rows);
      }
      g_end = r >= p.num_rows;
      if (g_end) return;
      int rows = static_cast<int>(tmin<int64_t>(kS2Rows, c1 - r));
      int64_t o0 = 0, o1 = 0;
      if (lane < rows) {
        o0 = p.row_offsets[r + lane];
        o1 = p.row_offsets[r + lane + 1];
      }
      int64_t gs = o0 + p.size_per_row, ge = o1;
      if (ge < gs) ge = gs;
      a_lo = b_lo + gs;
      a_hi = b_lo + ge;
      fl   = a_lo & ~uintptr_t{15};
      t_lo = fl < b_lo ? fl + 16 : fl;
      t_hi = (a_hi + 15) & ~uintptr_t{15};
      if (t_hi > b_hi) t_hi = a_hi & ~uintptr_t{15};
      if (t_hi < t_lo) t_hi = t_lo;
      if (lane >= rows || a_hi == a_lo) { t_hi = t_lo; }
      int32_t span = (lane < rows) ? static_cast<int32_t>(tmin<uintptr_t>(((a_hi + 15) & ~uintptr_t{15}) - fl, 1u << 30)) : 0;
      // exclusive scan of the window spans -> slot of each row; rows that fit the stage
      int32_t x = span;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      slot = x - span;
      const bool ok = lane < rows && x <= kS2StageBytes;
      int fit       = __popc(__ballot_sync(0xffffffffu, ok));
      // ballot counts lanes that fit; spans are non-negative so the set is a prefix
      if (fit < rows) fit &= ~7;
      g_safe = false;
      if (fit == 0) {
        g_safe = true;
        rows   = tmin(rows, 8);
      } else {
        rows = fit;
      }
      g_rows = rows;
    };
    next_geometry();

    for (int it = 0;; ++it) {
      const int s        = it % NS;
      const uint32_t par = ((it / NS) & 1) ^ 1;
      if (lane == 0) mbar_wait(&empty[s], par);
      __syncwarp();
      uint8_t* pay   = payload0 + static_cast<size_t>(s) * stage_span + kS2Front;
      int32_t* rowsm = rowsm0 + s * kS2Slice;
      int32_t* slice = slice0 + static_cast<size_t>(s) * slice_sp;
      S2Hdr* h       = hdr0 + s;
      if (g_end) {
        if (lane == 0) {
          h->rows = 0;
          mbar_arrive(&full[s]);
        }
        break;
      }
      const int rows  = g_rows;
      const bool safe = g_safe;
      uint32_t tx     = 0;
      if (!safe && lane < rows) {
        tx           = static_cast<uint32_t>(t_hi - t_lo);
        rowsm[lane]  = slot + static_cast<int32_t>(a_lo - fl);
        // bytes of [a_lo, a_hi) outside the TMA window [t_lo, t_hi): copy by hand
        for (uintptr_t a = a_lo; a < tmin(tmax(t_lo, a_lo), a_hi); ++a) pay[slot + (a - fl)] = *reinterpret_cast<const uint8_t*>(a);
        for (uintptr_t a = tmax(tmin(t_hi, a_hi), tmin(tmax(t_lo, a_lo), a_hi)); a < a_hi; ++a)
          pay[slot + (a - fl)] = *reinterpret_cast<const uint8_t*>(a);
      }
      // offsets slices [r, r + rows] of every STRING column: one small TMA bulk copy per column (r is a
      // multiple of 8, so &offsets[c][r] is 16-byte aligned whenever the buffer is).  Entries past the end
      // of the array (last tile) and unaligned buffers are copied by hand.
      const int need     = rows + 1;                                   // entries wanted
      const int64_t have = p.num_rows + 1 - r;                         // entries that exist from r on
      const int nbulk    = static_cast<int>(tmin<int64_t>((need + 3) & ~3, have & ~int64_t{3}));  // whole 16-byte chunks
      uint32_t stx       = 0;
      for (int sc = lane; sc < p.nstr; sc += 32) {
        const int32_t* src = s_offs[sc] + r;
        int32_t* dst       = slice + sc * kS2Slice;
        int done           = 0;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && nbulk > 0) {
          stx += static_cast<uint32_t>(nbulk) * 4u;
          done = nbulk;
        }
        for (int e = done; e < need && e < have; ++e) dst[e] = src[e];
      }
      uint32_t total = tx + stx;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) {
        h->r0   = r;
        h->rows = rows;
        h->safe = safe ? 1 : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&full[s], total);  // release: header / rowsm / hand copies visible
      __syncwarp();
      if (tx) tma_load_1d(pay + slot + (t_lo - fl), reinterpret_cast<const void*>(t_lo), tx, &full[s]);
      for (int sc = lane; sc < p.nstr; sc += 32) {
        const int32_t* src = s_offs[sc] + r;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && nbulk > 0)
          tma_load_1d(slice + sc * kS2Slice, src, static_cast<uint32_t>(nbulk) * 4u, &full[s]);
      }
      r += rows;
      next_geometry();
    }
  } else {
    // =================================== consumers ===================================
    const int cw       = warp_id() - 1;
    const int s0       = (cw * p.nstr) / kS2Consumers;
    const int s1       = ((cw + 1) * p.nstr) / kS2Consumers;
    uint8_t* stg       = stg0 + static_cast<size_t>(cw) * kS2StageLine;
    const uint32_t stg_s = smem_u32(stg);
    for (int it = 0;; ++it) {
      const int s        = it % NS;
      const uint32_t par = (it / NS) & 1;
      mbar_wait_backoff(&full[s], par, 256);
      const S2Hdr h = hdr0[s];
      if (h.rows == 0) break;
      const uint8_t* pay   = payload0 + static_cast<size_t>(s) * stage_span + kS2Front;
      const int32_t* rowsm = rowsm0 + s * kS2Slice;
      const int32_t* slice = slice0 + static_cast<size_t>(s) * slice_sp;
      const int rows       = h.rows;
      const bool active    = lane < rows;
      const uint32_t pay_s = smem_u32(pay);
      const int32_t rowsm_lane = (!h.safe && active) ? rowsm[lane] : 0;
      // this lane's row: start of its variable section
      const uint8_t* var0;
      if (!h.safe) {
        var0 = pay + (active ? rowsm[lane] : 0);
      } else {
        var0 = p.rows + (active ? static_cast<int64_t>(p.row_offsets[h.r0 + lane]) + p.size_per_row : 0);
      }
      // running position inside the variable section: lengths of the STRING columns before s0
      int32_t run = 0;
      if (active)
        for (int sc = 0; sc < s0; ++sc) run += slice[sc * kS2Slice + lane + 1] - slice[sc * kS2Slice + lane];
      for (int sc = s0; sc < s1; ++sc) {
        const int32_t* sl  = slice + sc * kS2Slice;
        const int32_t base = sl[0];
        const int32_t T    = sl[rows] - base;
        int32_t o0 = 0, L = 0;
        if (active) {
          o0 = sl[lane];
          L  = tmax(sl[lane + 1] - o0, 0);
        }
        const int32_t pe         = o0 - base;
        const int32_t run_before = run;
        const uint8_t* src       = var0 + run;
        run += L;
        if (T <= 0) continue;
        uint8_t* D     = s_chars[sc] + base;
        const int maxL = __reduce_max_sync(0xffffffffu, L);
        if (!h.safe && maxL <= 32 && T <= 1024) {
          // ---- fast: word-granular copy into the staging line, then aligned 16-byte flush ----------
          // everything below addresses shared memory through 32-bit shared-space addresses (explicit
          // ld.shared / st.shared: the generic pointers above would compile to LD/ST + 64-bit math)
          const int a         = static_cast<int>(reinterpret_cast<uintptr_t>(D) & 15);
          const int d         = a + pe;                 // staging byte of this lane's string
          const int dsh       = d & 3;
          const uint32_t srcs = pay_s + static_cast<uint32_t>(rowsm_lane + run_before);
          const int ssh       = static_cast<int>(srcs & 3u);
          const int dlt       = ssh - dsh;
          const int pre       = ssh + (dlt < 0 ? 4 : 0);  // string byte 0 is byte `pre` of the source word stream
          const uint32_t sp   = srcs - pre;
          const int sh        = (dlt & 3) * 8;
          const int end       = dsh + L;                 // one past the last staging byte, relative to word w0
          const int kfull1    = end >> 2;                // full words: [dsh ? 1 : 0, kfull1)
          const uint32_t ds   = stg_s + static_cast<uint32_t>(d);
          const uint32_t w0s  = ds - dsh;
          const int lim       = L > 0 ? L + pre : 0;     // source word k overlaps the string iff 4k < lim
          const int Kmax      = (maxL + 6) >> 2;         // warp-uniform bound on kfull1 (<= 9)
          // edge bytes straight from the source: the first nh bytes when the string starts inside a staging word,
          // the last nt bytes when it ends inside one
          const int nh = dsh ? tmin(L, 4 - dsh) : 0;
          const int nt = (kfull1 > 0 || !dsh) ? (end & 3) : 0;
          uint32_t hb[3], tb[3];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            hb[t] = tb[t] = 0;
            if (t < nh) hb[t] = lds_u8(srcs + t);
            if (t < nt) tb[t] = lds_u8(srcs + (L - nt) + t);
          }
          // all source words first (independent loads), then shift + store
          uint32_t w[10];
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            w[k] = 0;
            const bool need = k == 0 ? (lim > 0 && pre < 4) : (4 * k < lim);
            if (need) w[k] = lds_u32(sp + 4 * k);
          }
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            if (t < nh) sts_u8(ds + t, hb[t]);
            if (t < nt) sts_u8(ds + (L - nt) + t, tb[t]);
          }
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            if (k < Kmax) {
              const uint32_t y = __funnelshift_r(w[k], w[k + 1], sh);
              const bool full  = k == 0 ? (dsh == 0 && kfull1 > 0) : (k < kfull1);
              if (full) sts_u32(w0s + 4 * k, y);
            }
          }
          __syncwarp();
          flush_staging_line(stg_s, D, a, T, lane);
          __syncwarp();
        } else {
          // ---- slow: long strings / SAFE tiles: lane = destination byte, source row by shuffle search ----
          const int last       = rows - 1;
          const uint32_t upe   = static_cast<uint32_t>(pe);
          const uint32_t bound = (static_cast<uint32_t>(T) + 31u) & ~31u;
          const uint64_t sbase = reinterpret_cast<uint64_t>(src);
          for (uint32_t q = lane; q < bound; q += 32) {
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const int cand   = j + step;
              const uint32_t v = __shfl_sync(0xffffffffu, upe, cand & 31);
              if (cand <= last && v <= q) j = cand;
            }
            const uint32_t pj = __shfl_sync(0xffffffffu, upe, j);
            const uint64_t sj = __shfl_sync(0xffffffffu, sbase, j);
            if (q < static_cast<uint32_t>(T)) D[q] = *reinterpret_cast<const uint8_t*>(sj + (q - pj));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

static size_t strings2_smem_bytes(int nstr)
{
  size_t b = static_cast<size_t>(kS2Stages) * (kS2Front + kS2StageBytes + kS2Back);
  b += static_cast<size_t>(kS2Stages) * kS2Slice * 4;
  b += static_cast<size_t>(kS2Stages) * nstr * kS2Slice * 4;
  b += static_cast<size_t>(kS2Stages) * sizeof(S2Hdr) + 2 * kS2Stages * 8;
  b += static_cast<size_t>(nstr) * 16 + 16;
  b += static_cast<size_t>(kS2Consumers) * kS2StageLine;
  return (b + 127) & ~size_t{127};
}

int launch_strings_from_rows(const srj_plan* plan, const uint8_t* rows, const int32_t* row_offsets, int64_t rows_bytes,
                             int64_t num_rows, const int32_t* const* d_offsets, uint8_t* const* d_chars,
                             const int64_t* d_status, cudaStream_t stream)
{
  const int nstr = plan->num_string_columns;
  if (nstr == 0 || num_rows == 0) return SRJ_OK;
  int dev = 0, nsm = 0;
  SRJ_CUDA_TRY(cudaGetDevice(&dev));
  SRJ_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  // strings2 splits the STRING columns of a 32-row tile over its 19 consumer warps: it needs many columns to fill
  // them.  Tables with few STRING columns take the task-parallel generic kernel (which follows the stored offsets,
  // so it serves canonical and non-canonical rows alike).
  const char* e_min  = getenv("SRJ_S2_MINCOLS");  // tuning knob (development)
  const int min_cols = e_min ? atoi(e_min) : kS2MinCols;
  const bool fast    = d_status != nullptr && nstr <= kS2MaxCols && nstr >= min_cols;
  if (fast) {
    S2Params p{};
    p.rows         = rows;
    p.row_offsets  = row_offsets;
    p.rows_bytes   = rows_bytes;
    p.num_rows     = num_rows;
    p.super_rows   = kS2Rows * 8;
    p.nstr         = nstr;
    p.size_per_row = plan->size_per_row;
    p.offsets      = d_offsets;
    p.chars        = d_chars;
    p.status       = d_status;
    const int64_t ns   = (num_rows + p.super_rows - 1) / p.super_rows;
    const int64_t grid = std::min<int64_t>(nsm, ns);
    const size_t smem  = strings2_smem_bytes(nstr);
    SRJ_CUDA_TRY(cudaFuncSetAttribute(strings2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    strings2_kernel<<<static_cast<unsigned>(grid), kS2Threads, smem, stream>>>(p);
  }
  // generic kernel: does the work only when the status word flags non-canonical rows (or is absent)
  const int64_t ntiles = (num_rows + 31) / 32;
  const int64_t grid   = std::min<int64_t>(ntiles, static_cast<int64_t>(nsm) * 8);
  strings_from_rows_kernel<<<static_cast<unsigned>(grid), kStrWarps * 32, 0, stream>>>(
    rows, row_offsets, plan->fixed_row_size, num_rows, nstr, plan->d_string_start, d_offsets, d_chars, ntiles,
    fast ? d_status : nullptr);
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

}  // namespace srj

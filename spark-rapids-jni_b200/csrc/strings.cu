// strings.cu -- STRING-column stages of convert_from_rows:
//   (1) lengths -> offsets: batched (all STRING columns in one launch) three-step scan replacing the
//       per-column thrust::exclusive_scan + .element() sync loop of the reference (RC:2375-2395);
//   (2) chars gather: replaces copy_strings_from_rows (RC:1110-1150).
#include <algorithm>
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
// chars gather (v1): a warp owns (32-row tile, STRING column); lane = row reads the (offset,len)
// pair from the row, then the warp copies the tile's chars for that column -- a contiguous
// destination range -- with lane = destination byte, locating the source row by a shuffle search.
// Destination stores are fully coalesced; source reads stay inside 2-3 sectors per instruction.
// --------------------------------------------------------------------------------------------------
constexpr int kStrWarps = 8;

__global__ void __launch_bounds__(kStrWarps * 32) strings_from_rows_kernel(
  const uint8_t* __restrict__ rows, const int32_t* __restrict__ row_offsets, int64_t row_stride, int64_t num_rows,
  int nstr, const int32_t* __restrict__ string_start, const int32_t* const* __restrict__ offsets,
  uint8_t* const* __restrict__ chars, int64_t ntiles)
{
  const int lane = lane_id();
  const int w    = warp_id();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r    = tile * 32 + lane;
    const bool active  = r < num_rows;
    const int64_t rsta = active ? (row_offsets ? static_cast<int64_t>(row_offsets[r]) : r * row_stride) : 0;
    for (int s = w; s < nstr; s += kStrWarps) {
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

int launch_strings_from_rows(const srj_plan* plan, const uint8_t* rows, const int32_t* row_offsets, int64_t num_rows,
                             const int32_t* const* d_offsets, uint8_t* const* d_chars, cudaStream_t stream)
{
  const int nstr = plan->num_string_columns;
  if (nstr == 0 || num_rows == 0) return SRJ_OK;
  const int64_t ntiles = (num_rows + 31) / 32;
  int dev = 0, nsm = 0;
  SRJ_CUDA_TRY(cudaGetDevice(&dev));
  SRJ_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int64_t grid = std::min<int64_t>(ntiles, static_cast<int64_t>(nsm) * 8);
  strings_from_rows_kernel<<<static_cast<unsigned>(grid), kStrWarps * 32, 0, stream>>>(
    rows, row_offsets, plan->fixed_row_size, num_rows, nstr, plan->d_string_start, d_offsets, d_chars, ntiles);
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

}  // namespace srj

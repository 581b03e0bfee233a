// hash.cu -- standalone Spark row hashes over columns (reference: hash/xxhash64.cu:550-579,
// hash/murmur_hash.cu:191-221, hash/hive_hash.cu:474-500; the reference runs
// thrust::tabulate with a per-row type_dispatcher over a table_device_view).
//
// Here the column descriptors travel as a __grid_constant__ kernel parameter (no
// table_device_view allocation / H2D copy); tables wider than kHashColsPerLaunch are hashed in
// column chunks chained through the output column (the accumulator IS the next chunk's seed, which
// is exactly the Spark chaining rule), so there is no hidden allocation at any width.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "hash_device.cuh"
#include "kernels.hpp"

namespace srj {

constexpr int kHashColsPerLaunch = 48;
constexpr int kHashThreads       = 256;

struct HashCol {
  const uint8_t* data;
  const uint32_t* mask;
  const int32_t* offsets;
  int16_t type;
  int16_t kind;  // 0 = generic dispatch, 1 = plain 4-byte value, 2 = plain 8-byte value (hashed as raw bits)
  int32_t size;  // bytes per element (0 for STRING)
};

struct HashParams {
  HashCol cols[kHashColsPerLaunch];
  int32_t ncols;
  int32_t kind;
  int32_t first;  // 1: accumulator starts from the seed; 0: from out[r] (a previous column chunk)
  int32_t pad;
  int64_t seed;
  int64_t n;
  void* out;
};

__device__ __forceinline__ bool row_valid(const uint32_t* m, int64_t r)
{
  return !m || ((__ldg(m + (r >> 5)) >> (r & 31)) & 1u);
}

__device__ __forceinline__ void load_fixed(const HashCol& c, int64_t r, uint64_t& v, uint64_t& v2)
{
  v2 = 0;
  switch (c.size) {
    case 1: v = __ldg(c.data + r); break;
    case 2: v = __ldg(reinterpret_cast<const uint16_t*>(c.data) + r); break;
    case 4: v = __ldg(reinterpret_cast<const uint32_t*>(c.data) + r); break;
    case 8: v = __ldg(reinterpret_cast<const unsigned long long*>(c.data) + r); break;
    default: {
      const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2*>(c.data) + r);
      v                  = q.x;
      v2                 = q.y;
    }
  }
}

// Each thread hashes kRowsPerThread rows (strided by the block size, so every load stays coalesced): the
// per-column dispatch is paid once per kRowsPerThread rows, the dependent multiply chains of the rows
// overlap (ILP), and the NEXT column's values and mask words are fetched (unconditionally: fixed-width
// reads never depend on validity) while the current column is hashed -- the memory-level parallelism a
// one-row-per-thread loop (thrust::tabulate in the reference) lacks.
constexpr int kRowsPerThread = 4;

struct Fetched {
  uint64_t v[kRowsPerThread];
  uint64_t v2[kRowsPerThread];
  bool ok[kRowsPerThread];
};

__device__ __forceinline__ void fetch_column(const HashCol& col, const int64_t (&r)[kRowsPerThread],
                                             const bool (&live)[kRowsPerThread], Fetched& f)
{
  if (col.type == SRJ_STRING) {  // strings are read where they are hashed
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) f.ok[j] = live[j] && row_valid(col.mask, r[j]);
    return;
  }
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j) {
    f.v[j] = f.v2[j] = 0;
    if (live[j]) {
      if (col.kind == 1) f.v[j] = __ldg(reinterpret_cast<const uint32_t*>(col.data) + r[j]);
      else if (col.kind == 2) f.v[j] = __ldg(reinterpret_cast<const unsigned long long*>(col.data) + r[j]);
      else load_fixed(col, r[j], f.v[j], f.v2[j]);
    }
    f.ok[j] = live[j] && row_valid(col.mask, r[j]);
  }
}

// Each thread makes one pass (load 4 rows of every key column, hash, store), so the kernel lives on occupancy:
// 4 resident CTAs (64 registers) for the multiply-heavy xxhash64 / murmur3, 5 (48 registers) for hive.  Measured on
// 100 M rows x (INT32, INT64): xxhash64 1.26 -> 0.99 ms against the unbounded 80-register build.
template <int KIND>
__global__ void __launch_bounds__(kHashThreads, KIND == SRJ_HASH_HIVE ? 5 : 4) row_hash_kernel(const __grid_constant__ HashParams p)
{
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * (kHashThreads * kRowsPerThread) + threadIdx.x;
  int64_t r[kRowsPerThread];
  bool live[kRowsPerThread];
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j) {
    r[j]    = r0 + static_cast<int64_t>(j) * kHashThreads;
    live[j] = r[j] < p.n;
  }
  if (!live[0]) return;
  using acc_t = typename std::conditional<KIND == SRJ_HASH_XXHASH64, uint64_t, uint32_t>::type;
  acc_t h[kRowsPerThread];
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j) {
    if (p.first || !live[j]) h[j] = KIND == SRJ_HASH_HIVE ? acc_t{0} : static_cast<acc_t>(p.seed);
    else h[j] = reinterpret_cast<const acc_t*>(p.out)[r[j]];
  }
  Fetched cur, nxt;
  fetch_column(p.cols[0], r, live, cur);
  for (int c = 0; c < p.ncols; ++c) {
    const HashCol col = p.cols[c];
    if (c + 1 < p.ncols) fetch_column(p.cols[c + 1], r, live, nxt);  // in flight while this column is hashed
    if (col.type == SRJ_STRING) {
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        if constexpr (KIND == SRJ_HASH_HIVE) {
          uint32_t x = 0;  // null -> 0 (hive_hash.cu:201-203)
          if (cur.ok[j]) {
            const int32_t o0 = __ldg(col.offsets + r[j]), o1 = __ldg(col.offsets + r[j] + 1);
            x                = static_cast<uint32_t>(hash::hive_bytes(col.data + o0, o1 - o0));
          }
          if (live[j]) h[j] = 31u * h[j] + x;
        } else {
          if (!cur.ok[j]) continue;  // null keeps the accumulator (xxhash64.cu:352-353, murmur_hash.cu:111-117)
          const int32_t o0 = __ldg(col.offsets + r[j]), o1 = __ldg(col.offsets + r[j] + 1);
          if constexpr (KIND == SRJ_HASH_XXHASH64) h[j] = hash::xx_bytes(col.data + o0, o1 - o0, h[j]);
          else h[j] = hash::mm_bytes(col.data + o0, o1 - o0, h[j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        if constexpr (KIND == SRJ_HASH_XXHASH64) {
          if (cur.ok[j]) {
            if (col.kind == 1) h[j] = hash::xx_u32(static_cast<uint32_t>(cur.v[j]), h[j]);
            else if (col.kind == 2) h[j] = hash::xx_u64(cur.v[j], h[j]);
            else h[j] = hash::xx_fixed(col.type, cur.v[j], cur.v2[j], h[j]);
          }
        } else if constexpr (KIND == SRJ_HASH_MURMUR3_32) {
          if (cur.ok[j]) {
            if (col.kind == 1) h[j] = hash::mm_u32(static_cast<uint32_t>(cur.v[j]), h[j]);
            else if (col.kind == 2) h[j] = hash::mm_u64(cur.v[j], h[j]);
            else h[j] = hash::mm_fixed(col.type, cur.v[j], cur.v2[j], h[j]);
          }
        } else {
          uint32_t x = 0;
          if (cur.ok[j]) {
            if (col.kind == 1) x = static_cast<uint32_t>(cur.v[j]);                      // INT32 / DATE (hive_hash.cu:97-133)
            else if (col.kind == 2) x = static_cast<uint32_t>((cur.v[j] >> 32) ^ cur.v[j]);  // INT64 (hive_hash.cu:44-47)
            else x = static_cast<uint32_t>(hash::hive_fixed(col.type, cur.v[j]));
          }
          if (live[j]) h[j] = 31u * h[j] + x;  // hive_hash.cu:179-191
        }
      }
    }
    cur = nxt;
  }
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j)
    if (live[j]) reinterpret_cast<acc_t*>(p.out)[r[j]] = h[j];
}


// Every key column is a plain 4- or 8-byte value hashed as stored (32/64-bit integers, dates, timestamps,
// durations, DECIMAL64): no type dispatch, no 16-byte lane, no string branch -- fewer registers and ~25 % fewer
// instructions per row than the general kernel above.  Same structure: 4 rows per thread, next column in flight.
template <int KIND>
__global__ void __launch_bounds__(kHashThreads, KIND == SRJ_HASH_HIVE ? 5 : 4) row_hash_plain_kernel(const __grid_constant__ HashParams p)
{
  // grid-stride over row blocks: a resident CTA keeps going (no relaunch gap between blocks)
  const int64_t nblk = (p.n + kHashThreads * kRowsPerThread - 1) / (kHashThreads * kRowsPerThread);
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
  const int64_t r0 = blk * (kHashThreads * kRowsPerThread) + threadIdx.x;
  int64_t r[kRowsPerThread];
  bool live[kRowsPerThread];
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j) {
    r[j]    = r0 + static_cast<int64_t>(j) * kHashThreads;
    live[j] = r[j] < p.n;
  }
  if (!live[0]) continue;
  using acc_t = typename std::conditional<KIND == SRJ_HASH_XXHASH64, uint64_t, uint32_t>::type;
  acc_t h[kRowsPerThread];
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j) {
    if (p.first || !live[j]) h[j] = KIND == SRJ_HASH_HIVE ? acc_t{0} : static_cast<acc_t>(p.seed);
    else h[j] = reinterpret_cast<const acc_t*>(p.out)[r[j]];
  }
  uint64_t cv[kRowsPerThread], nv[kRowsPerThread];
  bool cok[kRowsPerThread], nok[kRowsPerThread];
  auto fetch = [&](const HashCol& col, uint64_t (&v)[kRowsPerThread], bool (&ok)[kRowsPerThread]) {
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      v[j] = 0;
      if (live[j]) {
        if (col.kind == 1) v[j] = __ldg(reinterpret_cast<const uint32_t*>(col.data) + r[j]);
        else v[j] = __ldg(reinterpret_cast<const unsigned long long*>(col.data) + r[j]);
      }
      ok[j] = live[j] && row_valid(col.mask, r[j]);
    }
  };
  fetch(p.cols[0], cv, cok);
  for (int c = 0; c < p.ncols; ++c) {
    const bool four = p.cols[c].kind == 1;
    if (c + 1 < p.ncols) fetch(p.cols[c + 1], nv, nok);  // in flight while this column is hashed
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      if constexpr (KIND == SRJ_HASH_XXHASH64) {
        const uint64_t t = four ? hash::xx_u32(static_cast<uint32_t>(cv[j]), h[j]) : hash::xx_u64(cv[j], h[j]);
        h[j]             = cok[j] ? t : h[j];  // a null keeps the accumulator (xxhash64.cu:352-353)
      } else if constexpr (KIND == SRJ_HASH_MURMUR3_32) {
        const uint32_t t = four ? hash::mm_u32(static_cast<uint32_t>(cv[j]), h[j]) : hash::mm_u64(cv[j], h[j]);
        h[j]             = cok[j] ? t : h[j];  // murmur_hash.cu:111-117
      } else {
        const uint32_t x = four ? static_cast<uint32_t>(cv[j]) : static_cast<uint32_t>((cv[j] >> 32) ^ cv[j]);
        h[j]             = 31u * h[j] + (cok[j] ? x : 0u);  // hive_hash.cu:179-203 (null -> 0)
      }
    }
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      cv[j]  = nv[j];
      cok[j] = nok[j];
    }
  }
#pragma unroll
  for (int j = 0; j < kRowsPerThread; ++j)
    if (live[j]) reinterpret_cast<acc_t*>(p.out)[r[j]] = h[j];
  }
}

static int elem_size(int32_t t)
{
  switch (t) {
    case SRJ_INT8: case SRJ_UINT8: case SRJ_BOOL8: return 1;
    case SRJ_INT16: case SRJ_UINT16: return 2;
    case SRJ_INT32: case SRJ_UINT32: case SRJ_FLOAT32: case SRJ_TIMESTAMP_DAYS: case SRJ_DURATION_DAYS:
    case SRJ_DECIMAL32: return 4;
    case SRJ_INT64: case SRJ_UINT64: case SRJ_FLOAT64: case SRJ_TIMESTAMP_SECONDS: case SRJ_TIMESTAMP_MILLISECONDS:
    case SRJ_TIMESTAMP_MICROSECONDS: case SRJ_TIMESTAMP_NANOSECONDS: case SRJ_DURATION_SECONDS:
    case SRJ_DURATION_MILLISECONDS: case SRJ_DURATION_MICROSECONDS: case SRJ_DURATION_NANOSECONDS:
    case SRJ_DECIMAL64: return 8;
    case SRJ_DECIMAL128: return 16;
    default: return 0;
  }
}

int launch_hash(int kind, const srj_column* cols, int32_t num_columns, int64_t num_rows, int64_t seed, void* out,
                cudaStream_t stream)
{
  if (num_columns == 0 || num_rows == 0) return SRJ_OK;  // xxhash64.cu:564
  for (int c = 0; c < num_columns; ++c) {
    const int32_t t = cols[c].type_id;
    const bool ok   = (t == SRJ_STRING) || elem_size(t) > 0;
    if (!ok) { set_error("hash: column %d has unsupported type id %d (nested/dictionary types are not on this path)", c, t); return SRJ_EUNSUPPORTED; }
    if (kind == SRJ_HASH_HIVE && !hash::hive_supported(t)) { set_error("hive_hash: column %d has unsupported type id %d (hive_hash.cu:63-66)", c, t); return SRJ_EUNSUPPORTED; }
    if (cols[c].size != num_rows) { set_error("hash: column %d has %lld rows, expected %lld", c, (long long)cols[c].size, (long long)num_rows); return SRJ_EINVAL; }
  }
  const int64_t per_block = static_cast<int64_t>(kHashThreads) * kRowsPerThread;
  const unsigned grid     = static_cast<unsigned>((num_rows + per_block - 1) / per_block);
  for (int c0 = 0; c0 < num_columns; c0 += kHashColsPerLaunch) {
    HashParams p{};
    p.ncols = std::min(kHashColsPerLaunch, num_columns - c0);
    p.kind  = kind;
    p.first = c0 == 0;
    p.seed  = seed;
    p.n     = num_rows;
    p.out   = out;
    for (int i = 0; i < p.ncols; ++i) {
      const srj_column& c = cols[c0 + i];
      const int32_t t     = c.type_id;
      int kind2           = 0;
      if (kind == SRJ_HASH_HIVE) {
        if (t == SRJ_INT32 || t == SRJ_TIMESTAMP_DAYS) kind2 = 1;
        else if (t == SRJ_INT64) kind2 = 2;
      } else {
        if (t == SRJ_INT32 || t == SRJ_UINT32 || t == SRJ_TIMESTAMP_DAYS || t == SRJ_DURATION_DAYS) kind2 = 1;
        else if (elem_size(t) == 8 && t != SRJ_FLOAT64) kind2 = 2;  // ints, timestamps, durations, DECIMAL64: raw 8 bytes
      }
      p.cols[i] = HashCol{static_cast<const uint8_t*>(c.data), c.null_mask, c.offsets, static_cast<int16_t>(t),
                          static_cast<int16_t>(kind2), elem_size(t)};
    }
    bool plain = SRJ_KNOB("SRJ_HASH_GENERAL", 0) == 0;
    for (int i = 0; i < p.ncols; ++i) plain = plain && p.cols[i].kind != 0;
    if (plain) {
      // persistent launch: as many CTAs as stay resident (the kernel strides over the row blocks)
      int dev = 0, nsm = 0;
      SRJ_CUDA_TRY(cudaGetDevice(&dev));
      SRJ_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      // grid = resident CTAs per SM (occupancy) x rounds.  Measured on 100 M rows x (INT32, INT64): the multiply-heavy
      // kernels like 4 rounds (xxhash64 0.87 -> 0.83 ms, murmur3 0.78 -> 0.74 ms against one CTA per row block),
      // hive -- pure streaming -- exactly one (0.67 -> 0.63 ms); a grid that is not a multiple of the resident count
      // leaves a straggler CTA per SM (xxhash64 1.17 ms).
      static int occ_cache[4] = {0, 0, 0, 0};  // per hash kind; the same for every sm_100a device (benign race)
      int occ = occ_cache[kind & 3];
      if (occ == 0) {
        if (kind == SRJ_HASH_XXHASH64)
          SRJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, row_hash_plain_kernel<SRJ_HASH_XXHASH64>, kHashThreads, 0));
        else if (kind == SRJ_HASH_MURMUR3_32)
          SRJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, row_hash_plain_kernel<SRJ_HASH_MURMUR3_32>, kHashThreads, 0));
        else
          SRJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, row_hash_plain_kernel<SRJ_HASH_HIVE>, kHashThreads, 0));
        occ_cache[kind & 3] = occ = std::max(1, occ);
      }
      const int e_w        = SRJ_KNOB("SRJ_HASH_WAVES", -1);  // tuning knob (development builds): CTAs per SM, 0 = one CTA per row block
      const int waves      = e_w >= 0 ? e_w : std::max(1, occ) * (kind == SRJ_HASH_HIVE ? 1 : 4);
      const unsigned pgrid = waves > 0 ? std::min<unsigned>(grid, static_cast<unsigned>(nsm * waves)) : grid;
      if (kind == SRJ_HASH_XXHASH64)
        row_hash_plain_kernel<SRJ_HASH_XXHASH64><<<pgrid, kHashThreads, 0, stream>>>(p);
      else if (kind == SRJ_HASH_MURMUR3_32)
        row_hash_plain_kernel<SRJ_HASH_MURMUR3_32><<<pgrid, kHashThreads, 0, stream>>>(p);
      else
        row_hash_plain_kernel<SRJ_HASH_HIVE><<<pgrid, kHashThreads, 0, stream>>>(p);
    } else if (kind == SRJ_HASH_XXHASH64)
      row_hash_kernel<SRJ_HASH_XXHASH64><<<grid, kHashThreads, 0, stream>>>(p);
    else if (kind == SRJ_HASH_MURMUR3_32)
      row_hash_kernel<SRJ_HASH_MURMUR3_32><<<grid, kHashThreads, 0, stream>>>(p);
    else
      row_hash_kernel<SRJ_HASH_HIVE><<<grid, kHashThreads, 0, stream>>>(p);
  }
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

}  // namespace srj

// hash.cu -- standalone Spark row hashes over columns (reference: hash/xxhash64.cu:550-579,
// hash/murmur_hash.cu:191-221, hash/hive_hash.cu:474-500; the reference runs
// thrust::tabulate with a per-row type_dispatcher over a table_device_view).
//
// Here the column descriptors travel as a __grid_constant__ kernel parameter (no
// table_device_view allocation / H2D copy); tables wider than kHashColsPerLaunch are hashed in
// column chunks chained through the output column (the accumulator IS the next chunk's seed, which
// is exactly the Spark chaining rule), so there is no hidden allocation at any width.
#include <algorithm>
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
  int32_t type;
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

template <int KIND>
__global__ void __launch_bounds__(kHashThreads) row_hash_kernel(const __grid_constant__ HashParams p)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kHashThreads + threadIdx.x;
  if (r >= p.n) return;
  if constexpr (KIND == SRJ_HASH_XXHASH64) {
    uint64_t h = p.first ? static_cast<uint64_t>(p.seed) : reinterpret_cast<const uint64_t*>(p.out)[r];
    for (int c = 0; c < p.ncols; ++c) {
      const HashCol& col = p.cols[c];
      if (!row_valid(col.mask, r)) continue;  // null keeps the accumulator (xxhash64.cu:352-353)
      if (col.type == SRJ_STRING) {
        const int32_t o0 = __ldg(col.offsets + r), o1 = __ldg(col.offsets + r + 1);
        h = hash::xx_bytes(col.data + o0, o1 - o0, h);
      } else {
        uint64_t v, v2;
        load_fixed(col, r, v, v2);
        h = hash::xx_fixed(col.type, v, v2, h);
      }
    }
    reinterpret_cast<uint64_t*>(p.out)[r] = h;
  } else if constexpr (KIND == SRJ_HASH_MURMUR3_32) {
    uint32_t h = p.first ? static_cast<uint32_t>(p.seed) : reinterpret_cast<const uint32_t*>(p.out)[r];
    for (int c = 0; c < p.ncols; ++c) {
      const HashCol& col = p.cols[c];
      if (!row_valid(col.mask, r)) continue;  // murmur_hash.cu:111-117
      if (col.type == SRJ_STRING) {
        const int32_t o0 = __ldg(col.offsets + r), o1 = __ldg(col.offsets + r + 1);
        h = hash::mm_bytes(col.data + o0, o1 - o0, h);
      } else {
        uint64_t v, v2;
        load_fixed(col, r, v, v2);
        h = hash::mm_fixed(col.type, v, v2, h);
      }
    }
    reinterpret_cast<uint32_t*>(p.out)[r] = h;
  } else {
    uint32_t h = p.first ? 0u : reinterpret_cast<const uint32_t*>(p.out)[r];
    for (int c = 0; c < p.ncols; ++c) {
      const HashCol& col = p.cols[c];
      uint32_t x         = 0;  // null -> 0 (hive_hash.cu:201-203)
      if (row_valid(col.mask, r)) {
        if (col.type == SRJ_STRING) {
          const int32_t o0 = __ldg(col.offsets + r), o1 = __ldg(col.offsets + r + 1);
          x = static_cast<uint32_t>(hash::hive_bytes(col.data + o0, o1 - o0));
        } else {
          uint64_t v, v2;
          load_fixed(col, r, v, v2);
          x = static_cast<uint32_t>(hash::hive_fixed(col.type, v));
        }
      }
      h = 31u * h + x;  // hive_hash.cu:179-191
    }
    reinterpret_cast<uint32_t*>(p.out)[r] = h;
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
  const unsigned grid = static_cast<unsigned>((num_rows + kHashThreads - 1) / kHashThreads);
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
      p.cols[i]           = HashCol{static_cast<const uint8_t*>(c.data), c.null_mask, c.offsets, c.type_id, elem_size(c.type_id)};
    }
    if (kind == SRJ_HASH_XXHASH64)
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

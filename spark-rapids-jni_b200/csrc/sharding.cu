// sharding.cu -- multi-GPU configuration (no reference counterpart): after the NCCL all-gather of the
// per-rank column chunks, the STRING offsets of rank r's rows are still local to r's shard.  This kernel turns them
// into offsets of the gathered column: + chars of the column held by the ranks before r.
#include <algorithm>
#include "common.cuh"

namespace srj {

struct RebaseParams {
  uint8_t* gathered;        // [world][slab_bytes]: every rank's packed slab, as all-gathered
  int64_t slab_bytes;
  const int64_t* offs_at;   // device [nstr]: byte offset of each STRING column's int32 offsets[rows + 1] inside a slab
  const int32_t* scol;      // device [nstr]: schema column of each STRING column
  const int64_t* totals;    // device [world][ncols + 1]: per-rank char totals (phase 1 output, all-gathered)
  int64_t rows;             // rows per shard
  int32_t ncols, nstr, world;
};

__global__ void __launch_bounds__(256) shard_rebase_offsets_kernel(const __grid_constant__ RebaseParams p)
{
  const int s = blockIdx.y, r = blockIdx.z;
  if (r == 0) return;  // rank 0's offsets are global already
  int64_t delta = 0;
  for (int q = 0; q < r; ++q) delta += p.totals[static_cast<int64_t>(q) * (p.ncols + 1) + p.scol[s]];
  int32_t* offs = reinterpret_cast<int32_t*>(p.gathered + static_cast<int64_t>(r) * p.slab_bytes + p.offs_at[s]);
  const int32_t d = static_cast<int32_t>(delta);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i <= p.rows; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    offs[i] += d;
}

}  // namespace srj

extern "C" int srj_shard_rebase_offsets(void* gathered, int64_t slab_bytes, const int64_t* d_offs_at, const int32_t* d_scol,
                                        const int64_t* d_totals, int64_t rows_per_shard, int32_t num_columns,
                                        int32_t num_string_columns, int32_t world, void* stream)
{
  using namespace srj;
  if (!gathered || !d_offs_at || !d_scol || !d_totals || world < 1 || num_string_columns < 0) { set_error("shard_rebase_offsets: bad argument"); return SRJ_EINVAL; }
  if (world == 1 || num_string_columns == 0 || rows_per_shard == 0) return SRJ_OK;
  RebaseParams p{static_cast<uint8_t*>(gathered), slab_bytes, d_offs_at, d_scol, d_totals, rows_per_shard, num_columns, num_string_columns, world};
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>(64, (rows_per_shard + 256) / 256));
  shard_rebase_offsets_kernel<<<dim3(gx, num_string_columns, world), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  SRJ_CUDA_TRY(cudaGetLastError());
  return SRJ_OK;
}

"""Dev helper: time srj_convert_to_rows (fixed-width C2 schema) on the device with CUDA events."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spark-rapids-jni_b200")]
import torch

import bench
import srj_b200 as S
from srj_b200 import _native as N

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
types = wl["types"]
plan = S.Plan.get([S.DType(t) for t in types])
rs = plan.layout.fixed_row_size
cols = bench.synth_columns_gpu(torch, S, types, n, wl["null_frac"], 1)
lib = N.lib()
carr = (N.SrjColumn * len(cols))()
for i, c in enumerate(cols):
    carr[i] = c._c()
batches = (N.SrjRowBatch * 64)()
nb = C.c_int32(0)
st = int(torch.cuda.current_stream().cuda_stream)
N.check(lib.srj_to_rows_plan_batches(plan.handle, carr, n, None, batches, 64, C.byref(nb), st))
rows = torch.empty(n * rs, dtype=torch.uint8, device="cuda")
offs = torch.empty(n + nb.value, dtype=torch.int32, device="cuda")
op, dp = (C.c_void_p * nb.value)(), (C.c_void_p * nb.value)()
for b in range(nb.value):
    op[b] = offs.data_ptr() + 4 * (batches[b].row_start + b)
    dp[b] = rows.data_ptr() + batches[b].row_start * rs


def go():
    N.check(lib.srj_convert_to_rows(plan.handle, carr, n, None, batches, nb.value, op, dp, st))


for _ in range(3):
    go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    go()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
bpr = bench.algorithmic_bytes_per_row(types, rs) + 4      # + the int32 row offset written per row
print(f"to_rows {sys.argv[1] if len(sys.argv) > 1 else 'c2'} rows={n} batches={nb.value} ms={ms:.3f} "
      f"GB/s={bpr * n / ms / 1e6:.0f} frac={bpr * n / ms / 1e6 / 6576.1:.3f}")

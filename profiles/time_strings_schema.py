"""Dev helper: time convertToRows / convertFromRows (public API) on a narrow table with STRING columns.

    python profiles/time_strings_schema.py <rows> <schema: e.g. i4,s,i8,d16,s,b1,s,i2>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spark-rapids-jni_b200")]
import torch

import bench
import srj_b200 as S

CODE = {"i1": bench.INT8, "i2": bench.INT16, "i4": bench.INT32, "i8": bench.INT64, "b1": bench.BOOL8, "d16": bench.DEC128,
        "s": bench.STRING}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
schema = (sys.argv[2] if len(sys.argv) > 2 else "i4,s,i8,d16,s,b1,s,i2").split(",")
types = [CODE[c] for c in schema]
g = torch.Generator(device="cuda").manual_seed(5)
fixed = iter(bench.synth_columns_gpu(torch, S, [t for t in types if t != bench.STRING], n, 0.1, 3))
cols = [bench.synth_strings_gpu(torch, S, n, 0.1, g) if t == bench.STRING else next(fixed) for t in types]
tbl = S.Table(cols)
dts = [S.DType(t) for t in types]


def timeit(fn, k=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, out


ms_to, rows = timeit(lambda: S.RowConversion.convertToRows(tbl))
row_bytes = sum(r.child.data.numel() for r in rows)
col_bytes = sum(c.data.numel() + (c.offsets.numel() * 4 if c.offsets is not None else 0) + n // 8 for c in cols)
byts = row_bytes + col_bytes + 4 * n
print(f"to_rows   {','.join(schema)} rows={n} batches={len(rows)} avg_row={row_bytes / n:.0f}B ms={ms_to:.3f} "
      f"GB/s={byts / ms_to / 1e6:.0f} frac={byts / ms_to / 1e6 / 6576.1:.3f} (public API: incl. allocation + batch-plan sync)")
if len(rows) == 1:
    ms_fr, _ = timeit(lambda: S.RowConversion.convertFromRows(rows[0], dts))
    print(f"from_rows {','.join(schema)} rows={n} ms={ms_fr:.3f} GB/s={byts / ms_fr / 1e6:.0f} frac={byts / ms_fr / 1e6 / 6576.1:.3f}")

"""Dev helper: time the standalone columnar hash kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spark-rapids-jni_b200")]
import torch

import bench
import srj_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
g = torch.Generator(device="cuda").manual_seed(1)


def timeit(fn, k=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


INT32, INT64 = 3, 4
cols = bench.synth_columns_gpu(torch, S, [INT32, INT64], n, 0.04, 3)
for name, fn, outb in (("xxhash64", lambda: S.Hash.xxhash64(42, cols), 8), ("murmur3", lambda: S.Hash.murmurHash32(42, cols), 4),
                       ("hive", lambda: S.Hash.hiveHash(cols), 4)):
    ms = timeit(fn)
    byts = n * (12 + 0.25 + outb)
    print(f"{name:9s} keys(int32,int64) rows={n} ms={ms:.3f} GB/s={byts / ms / 1e6:.0f} frac={byts / ms / 1e6 / 6576.1:.3f} rows/s={n / ms * 1e3:.3g}")
ns = int(os.environ.get("SRJ_TH_STR_ROWS", n // 5))
sc = bench.synth_strings_gpu(torch, S, ns, 0.2, g)
chars = sc.data.numel()
for name, fn, outb in (("xxhash64", lambda: S.Hash.xxhash64(42, [sc]), 8), ("murmur3", lambda: S.Hash.murmurHash32(42, [sc]), 4),
                       ("hive", lambda: S.Hash.hiveHash([sc]), 4)):
    ms = timeit(fn)
    byts = chars + ns * (4 + 0.125 + outb)
    print(f"{name:9s} strings rows={ns} chars={chars} ms={ms:.3f} GB/s={byts / ms / 1e6:.0f} frac={byts / ms / 1e6 / 6576.1:.3f}")

"""Timings of the other BASELINE configs on one GPU (development probe): config 3 (1M-cell self-kNN, k=15),
config-5 slice (500k targets x 2M references, cosine, k=30) and the CSR projection feeding it."""
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from nabo_b200 import core, synth

def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

g = 50
ref = torch.from_numpy(synth.pc_mixture(1_000_000, g, 1)).cuda()
ms = timeit(lambda: core.knn(ref, ref, 15, "euclidean", drop_first=True, mode="fast"))
print("config 3: 1M-cell reference self-kNN, k=15, Euclidean: %.1f ms  (%.2e pairs/s, %.2f M cells/s)" % (ms, 1e12 / ms * 1e3, 1e6 / ms * 1e3 / 1e6))
del ref
ref = torch.from_numpy(synth.pc_mixture(2_000_000, g, 1)).cuda()
tgt = torch.from_numpy(synth.pc_mixture(500_000, g, 101)).cuda()
ms = timeit(lambda: core.knn(tgt, ref, 30, "cosine", mode="fast"))
print("config 5 slice: 500k targets x 2M references, cosine, k=30: %.1f ms  (%.2e pairs/s, %.2f M cells/s)" % (ms, 1e12 / ms * 1e3, 5e5 / ms * 1e3 / 1e6))
st = core.knn(tgt, ref, 30, "cosine", mode="fast", return_stats=True)[2]
print("   stats:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
ms = timeit(lambda: core.knn(tgt[:200000], ref[:1250000], 30, "mod_canberra", 0.25, mode="fast"), n=1)
print("config 4 shard shape, modified Canberra: 200k targets x 1.25M references: %.1f ms (%.2e pairs/s)" % (ms, 2.5e11 / ms * 1e3))

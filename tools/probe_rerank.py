"""Re-rank stage time at config 2 and at the config-4 shard shape (development probe)."""
import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
g, k = 50, 30
for n, m in ((100000, 100000), (227328, 1250000)):
    q = synth.pc_mixture_device(n, g, 101, "cuda")
    r = synth.pc_mixture_device(m, g, 1, "cuda")
    for _ in range(2): core.knn(q, r, k, "euclidean")
    acc = {}
    for _ in range(5):
        st = core.knn(q, r, k, "euclidean", return_stats=True)[2]
        for kk in ("main_kernel_ms", "rerank_ms", "prep_ms", "fallback_ms", "rows_exact_fallback"):
            acc[kk] = acc.get(kk, 0.0) + st[kk] / 5
    print(n, m, {kk: round(v, 3) for kk, v in acc.items()})

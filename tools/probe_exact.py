"""Quick device timing of the exact engine (development probe, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n, m, g, k = 8192, 100000, 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for metric in ("euclidean", "mod_canberra", "cosine"):
    ms = t(lambda: core.knn(q, r, k, metric, 0.25, mode="exact"))
    print("exact %-13s %d x %d g=%d k=%d: %.2f ms  %.3e pairs/s  %.0f queries/s" % (metric, n, m, g, k, ms, n * m / ms * 1e3, n / ms * 1e3))
idx, _ = core.knn(q, r, k, "euclidean", mode="exact")
rk, _ = core.knn(r[:20000], r, k, "euclidean", drop_first=True, mode="exact")
rk = torch.randint(0, m, (m, k), dtype=torch.int32, device="cuda")
ms = t(lambda: core.snn_weights(idx, rk, k)); print("snn %d: %.3f ms" % (n, ms))
cnt, w = core.snn_weights(idx, rk, k)
ms = t(lambda: core.mapping_scores(idx, cnt, m, k)); print("scores: %.3f ms" % ms)
x = torch.rand(50000, 2000, device="cuda") ; x = (x * 3).floor().float()
gi = torch.arange(2000, dtype=torch.int32, device="cuda")
sf = torch.rand(50000, device="cuda") + 0.5
mu = torch.rand(2000, dtype=torch.float64, device="cuda"); sg = mu + 0.5
C_ = torch.randn(50, 2000, dtype=torch.float64, device="cuda"); mean = torch.randn(2000, dtype=torch.float64, device="cuda")
ms = t(lambda: core.project(x, gi, sf, mu, sg, C_, mean)); print("project dense 50000x2000x50: %.3f ms  %.2f TFLOP/s fp64" % (ms, 2*50000*2000*50/ms/1e9))

"""Device timing of the fast engine at config-2 scale (development probe)."""
import sys
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

g, k = 50, 30
for n, m in ((100000, 100000), (200000, 1250000)):
    q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
    r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
    for metric in ("euclidean", "cosine"):
        ms = t(lambda: core.knn(q, r, k, metric, mode="fast"))
        _, _, st = core.knn(q, r, k, metric, mode="fast", return_stats=True)
        print("fast %-10s %d x %d: %.2f ms  %.3e pairs/s  %.3e queries/s | %s" % (metric, n, m, ms, n * m / ms * 1e3, n / ms * 1e3, st))
        tf = 2.0 * g * n * m / (st["main_kernel_ms"] * 1e-3) / 1e12
        print("   candidate kernel: %.2f ms -> %.1f useful TFLOP/s (2*g per pair), %.1f executed TFLOP/s (K=160)" % (st["main_kernel_ms"], tf, tf * 160 / 50))
n, m = 100000, 100000
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
ms = t(lambda: core.knn(q, r, k, "mod_canberra", 0.25, mode="fast"), reps=2)
_, _, st = core.knn(q, r, k, "mod_canberra", 0.25, mode="fast", return_stats=True)
print("fast mod_canberra %d x %d: %.2f ms  %.3e pairs/s  %.3e queries/s | %s" % (n, m, ms, n * m / ms * 1e3, n / ms * 1e3, st))

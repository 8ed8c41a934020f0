"""SNN weights + integer score accumulation at config 2 (development probe)."""
import sys
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n = m = 100000; g, k = 50, 30
ref = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
tgt = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
rk, _ = core.knn(ref, ref, k, "euclidean", drop_first=True)
ti, _ = core.knn(tgt, ref, k, "euclidean")
cnt = torch.empty((n, k), dtype=torch.uint8, device="cuda"); w = torch.empty((n, k), dtype=torch.float64, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("snn_weights %d x k=%d: %.3f ms" % (n, k, t(lambda: core.snn_weights(ti, rk, k, out=(cnt, w)))))
acc = torch.zeros(m, dtype=torch.int64, device="cuda")
print("score_accumulate + finalize: %.3f ms" % t(lambda: core.scores_finalize(core.score_accumulate(ti, cnt, m, k, acc=acc.zero_()), n)))
print("mean snn count %.2f" % cnt.float().mean().item())

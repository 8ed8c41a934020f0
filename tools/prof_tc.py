"""One candidate-pass launch at a single-wave shape, for ncu (development probe)."""
import sys
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
import os
n, m, g, k = 148 * 384, int(os.environ.get("PROF_M", "131072")), 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for _ in range(2):
    core.knn_candidates(q, r, k, "euclidean")
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); core.knn_candidates(q, r, k, "euclidean"); b.record(); torch.cuda.synchronize()
print("candidates %d x %d: %.3f ms  %.3e pairs/s" % (n, m, a.elapsed_time(b), n * m / a.elapsed_time(b) * 1e3))

"""Two candidate-pass launches at the config-2 shape (100 000 x 100 000), for ncu (development probe)."""
import os, sys
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n, m, g, k = int(os.environ.get("PROF_N", "100000")), int(os.environ.get("PROF_M", "100000")), 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for _ in range(2):
    core.knn_candidates(q, r, k, "euclidean")
torch.cuda.synchronize()
print("ok")

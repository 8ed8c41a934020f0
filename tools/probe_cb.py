"""Modified-Canberra kNN at config 2 (development probe)."""
import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n = m = 100000; g, k = 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda(); r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for _ in range(2): core.knn(q, r, k, "mod_canberra", 0.25)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): core.knn(q, r, k, "mod_canberra", 0.25)
b.record(); torch.cuda.synchronize()
st = core.knn(q, r, k, "mod_canberra", 0.25, return_stats=True)[2]
print("mod_canberra %d x %d: %.2f ms per call; kernel %.2f ms, re-rank %.2f ms" % (n, m, a.elapsed_time(b) / 3, st["main_kernel_ms"], st["rerank_ms"]))

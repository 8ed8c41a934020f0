"""Modified-Canberra kNN at small target counts (reference split over CTAs), development probe."""
import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
g, k = 50, 30
r = torch.from_numpy(synth.pc_mixture(100000, g, 1)).cuda()
for n in (1000, 3000, 5000, 10000, 20000, 40000):
    q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
    for _ in range(2): core.knn(q, r, k, "mod_canberra", 0.25, mode="fast")
    torch.cuda.synchronize()
    st = core.knn(q, r, k, "mod_canberra", 0.25, mode="fast", return_stats=True)[2]
    print("n=%6d: main %.3f ms, rerank %.3f ms, fallback rows %d" % (n, st["main_kernel_ms"], st["rerank_ms"], st["rows_exact_fallback"]))

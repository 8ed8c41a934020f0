import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
g, k = 50, 30
r = torch.from_numpy(synth.pc_mixture(100000, g, 1)).cuda()
for n in (2000, 5000, 20000, 30000, 50000):
    q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
    for _ in range(2): core.knn(q, r, k, "euclidean", mode="fast")
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): core.knn(q, r, k, "euclidean", mode="fast")
    b.record(); torch.cuda.synchronize()
    st = core.knn(q, r, k, "euclidean", mode="fast", return_stats=True)[2]
    print("n=%6d: %.3f ms per call  (kernel %.3f, rerank %.3f, fallback rows %d, cand/row %d)" % (n, a.elapsed_time(b) / 5, st["main_kernel_ms"], st["rerank_ms"], st["rows_exact_fallback"], st["candidates_per_row"]))

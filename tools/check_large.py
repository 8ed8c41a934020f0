import sys, torch, numpy as np
sys.path.insert(0, ".")
from nabo_b200 import core, synth
g, k = 50, 30
r = torch.from_numpy(synth.pc_mixture(200000, g, 1)).cuda()
for n, metric in ((50000, "euclidean"), (20000, "euclidean"), (50000, "cosine"), (30000, "mod_canberra"), (3000, "mod_canberra")):
    q = torch.from_numpy(synth.pc_mixture(n, g, 101 + n)).cuda()
    fi, fd, st = core.knn(q, r, k, metric, 0.25, mode="fast", return_stats=True)
    ei, ed = core.knn(q, r, k, metric, 0.25, mode="exact")
    same = bool(torch.equal(fi, ei)) and bool(torch.equal(fd, ed))
    print("%-13s n=%6d x 200000: fast == exact: %s  (fallback rows %d, main %.2f ms)" % (metric, n, same, st["rows_exact_fallback"], st["main_kernel_ms"]))
# self-kNN with drop_first at 300k
ref = torch.from_numpy(synth.pc_mixture(300000, g, 7)).cuda()
fi, fd, st = core.knn(ref, ref, 15, "euclidean", drop_first=True, mode="fast", return_stats=True)
ei, ed = core.knn(ref[:40000], ref, 15, "euclidean", drop_first=True, mode="exact")
print("self-kNN 300k: first 40k rows fast == exact:", bool(torch.equal(fi[:40000], ei)) and bool(torch.equal(fd[:40000], ed)), "fallback", st["rows_exact_fallback"])

import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
m, g, k = 100000, 50, 30
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for w in (2, 4, 6, 8, 10, 12):
    n = 148 * 32 * w
    q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
    for _ in range(2): core.knn(q, r, k, "mod_canberra", 0.25, mode="fast")
    torch.cuda.synchronize()
    st = core.knn(q, r, k, "mod_canberra", 0.25, mode="fast", return_stats=True)[2]
    print("warps/CTA %2d: n=%6d  main %.3f ms  -> %.3e pairs/s" % (w, n, st["main_kernel_ms"], n * m / st["main_kernel_ms"] * 1e3))

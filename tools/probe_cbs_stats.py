"""Development probe: per-level pair counts of the bit-sliced Canberra pass (build with
NABO_NVCC_EXTRA=-DNABO_CBS_STATS python -m nabo_b200.build --force)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth, _lib
n, m, g, k = 148 * 384, 100000, 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
lib = _lib.lib()
out = (ctypes.c_ulonglong * 8)()
lib.nabo_debug_cbs_stats(out, 1)
core.knn(q, r, k, "mod_canberra", 0.25, mode="fast")
lib.nabo_debug_cbs_stats(out, 1)
names = ["level-1 survivors", "level-2 passes", "appends", "compactions", "level-3 rounds"]
for nm, v in zip(names, out):
    print("%-18s %12d  per query %9.1f  frac of pairs %.5f" % (nm, v, v / n, v / (n * m)))

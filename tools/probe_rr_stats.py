import ctypes as C, sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth, _lib
L = _lib.lib(); g, k = 50, 30
for n, m in ((100000, 100000), (227328, 1250000)):
    q = synth.pc_mixture_device(n, g, 101, "cuda"); r = synth.pc_mixture_device(m, g, 1, "cuda")
    out = (C.c_ulonglong * 8)(); L.nabo_dbg_rr_stats(out, 1)
    core.knn(q, r, k, "euclidean"); L.nabo_dbg_rr_stats(out, 1)
    rows = max(out[0], 1)
    print(n, m, "rows", out[0], "sort path %.2f %%" % (100.0 * out[1] / rows), "cut-bin size %.1f" % (out[2] / rows), "buffer count %.1f" % (out[3] / rows), "ranked away %.1f" % (out[4] / rows))

"""Cycle accounting of the sliced Canberra sweep (needs the NABO_CBS_PROF variant: tools/build_variant.py cbprof
-DNABO_CBS_PROF, then NABO_B200_LIB=tools/_variants/lib_cbprof.so).  Development probe."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth, _lib
g, k = 50, 30
L = _lib.lib()
names = ["tile waits", "count bound", "work list", "FP32 evaluation", "compaction", "stage release", "round set-up + final selection"]
for n, m in ((100000, 100000), (20000, 100000)):
    q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
    r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
    core.knn(q, r, k, "mod_canberra", 0.25)
    out = (C.c_ulonglong * 12)()
    L.nabo_dbg_cbs_prof(out, 1)
    st = core.knn(q, r, k, "mod_canberra", 0.25, return_stats=True)[2]
    L.nabo_dbg_cbs_prof(out, 1)
    tot = sum(out[i] for i in range(7))
    tiles = max(out[7], 1)
    print("%d x %d: kernel %.2f ms; busy warp-tiles %d, %.0f cycles per warp-tile; survivors per warp-tile %.1f (%.2f %% of pairs), "
          "evaluation passes %.2f, compactions %.3f per warp-tile" % (n, m, st["main_kernel_ms"], tiles, tot / tiles, out[10] / tiles,
                                                                      100.0 * out[10] / tiles / 4096 * (128 / 128), out[8] / tiles, out[9] / tiles))
    print("   " + "; ".join("%s %.1f %% (%.0f clk)" % (names[i], 100.0 * out[i] / tot, out[i] / tiles) for i in range(7)))
    print("   per evaluation pass %.0f clk, per compaction %.0f clk" % (out[3] / max(out[8], 1), out[4] / max(out[9], 1)))

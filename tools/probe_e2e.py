"""Host-buffer path (core.map_cells_host) at config 2 for different numbers of pipeline pieces (development probe)."""
import sys
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n = m = 100000; g, k = 50, 30
ref = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
tgt = torch.from_numpy(synth.pc_mixture(n, g, 101)).pin_memory()
rk, _ = core.knn(ref, ref, k, "euclidean", drop_first=True)
import nabo_b200.core as C
for chunks, floor in ((1, 32768), (2, 32768), (3, 32768), (4, 16384), (5, 16384)):
    def run():
        return core.map_cells_host(tgt, ref, rk, k, metric="euclidean", chunks=chunks, min_piece=floor)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    print("chunks=%d: %.3f ms per step -> %.2f M cells/s" % (chunks, a.elapsed_time(b) / 10, n / (a.elapsed_time(b) / 10) / 1e3))

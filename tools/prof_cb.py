"""One Canberra fast-engine call for ncu (development probe)."""
import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n, m, g, k = 148 * 384, 100000, 50, 30
q = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
r = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda()
for _ in range(2):
    core.knn(q, r, k, "mod_canberra", 0.25, mode="fast")
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); core.knn(q, r, k, "mod_canberra", 0.25, mode="fast"); b.record(); torch.cuda.synchronize()
print("canberra fast %d x %d: %.3f ms  %.3e pairs/s" % (n, m, a.elapsed_time(b), n * m / a.elapsed_time(b) * 1e3))
st = core.knn(q, r, k, "mod_canberra", 0.25, mode="fast", return_stats=True)[2]
print({kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in st.items()})

import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
dev = torch.device("cuda", 0)
ref = synth.pc_mixture_device(1250000, 50, seed=1001, device=dev)
for b in range(3):
    q = synth.pc_mixture_device(227328, 50, seed=101 + b, device=dev)
    r = core.knn(q, ref, 30, "euclidean", mode="fast", return_stats=True)
    print(b, r[2])

"""Stage times of the config-2 step with CUDA events between the calls (development probe)."""
import sys, torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth
n = m = 100000; g, k = 50, 30
ref = torch.from_numpy(synth.pc_mixture(m, g, 1)).cuda(); tgt = torch.from_numpy(synth.pc_mixture(n, g, 101)).cuda()
ref_knn, _ = core.knn(ref, ref, k, "euclidean", drop_first=True)
acc = torch.zeros(m, dtype=torch.int64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
names = ["knn", "snn", "zero", "accumulate", "finalize"]
tot = [0.0] * 5; steps = 20; st_sum = {}
for it in range(steps + 3):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    idx, dst = core.knn(tgt, ref, k, "euclidean"); ev[1].record()
    cnt, w = core.snn_weights(idx, ref_knn, k); ev[2].record()
    acc.zero_(); ev[3].record()
    core.score_accumulate(idx, cnt, m, k, acc=acc); ev[4].record()
    sc = core.scores_finalize(acc, n); ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i in range(5): tot[i] += ev[i].elapsed_time(ev[i + 1])
print("step %.3f ms: " % (sum(tot) / steps) + ", ".join("%s %.3f" % (names[i], tot[i] / steps) for i in range(5)))
for it in range(5):
    flush.zero_()
    st = core.knn(tgt, ref, k, "euclidean", return_stats=True)[2]
    for kk, v in st.items():
        if isinstance(v, (int, float)): st_sum[kk] = st_sum.get(kk, 0.0) + v / 5
print({kk: round(v, 3) for kk, v in st_sum.items()})

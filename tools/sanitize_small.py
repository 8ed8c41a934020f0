"""Small shapes of every hand-rolled pipeline (tcgen05/TMA/mbarrier candidate pass, bit-sliced Canberra pass with its
cp.async.bulk ring, DMMA projection, routed re-rank, parts merge, integer scores) for compute-sanitizer.  Development probe:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nabo_b200 import core, synth

g, k = 20, 7
ref = torch.from_numpy(synth.pc_mixture(3000, g, seed=1)).cuda()
tgt = torch.from_numpy(synth.pc_mixture(900, g, seed=101)).cuda()
ri, _ = core.knn(ref, ref, k, "euclidean", drop_first=True, mode="fast")          # tcgen05 candidates (unsplit + split)
ei, ed = core.knn(tgt, ref, k, "euclidean", mode="fast")
xi, xd = core.knn(tgt, ref, k, "euclidean", mode="exact")
assert torch.equal(ei, xi) and torch.equal(ed, xd)
ci, cd = core.knn(tgt, ref, k, "mod_canberra", 0.25, mode="fast")                 # bit-sliced Canberra pass
yi, yd = core.knn(tgt, ref, k, "mod_canberra", 0.25, mode="exact")
assert torch.equal(ci, yi) and torch.equal(cd, yd)
oi, od = core.knn(tgt, ref, k, "cosine", mode="fast")
bounds = [0, 300, 900]
bi = [torch.empty((bounds[p + 1] - bounds[p], k), dtype=torch.int32, device="cuda") for p in range(2)]
bd = [torch.empty((bounds[p + 1] - bounds[p], k), dtype=torch.float64, device="cuda") for p in range(2)]
core.knn(tgt, ref, k, "euclidean", out_parts=(bounds, [b.data_ptr() for b in bi], [b.data_ptr() for b in bd]))
assert torch.equal(torch.cat(bi), ei)
halves = [core.knn(tgt, ref[a:b].contiguous(), k, "euclidean", idx_offset=a) for a, b in ((0, 1400), (1400, 3000))]
mi, md = core.merge_topk_parts([h[0].data_ptr() for h in halves], [h[1].data_ptr() for h in halves], 900, k, "cuda")
assert torch.equal(mi, ei) and torch.equal(md, ed)
cnt, w = core.snn_weights(ei, ri, k)
acc = core.score_accumulate(ei, cnt, 3000, k)
sc = core.scores_finalize(acc, 900)
sc2 = core.mapping_scores(ei, cnt, 3000, k)
np.testing.assert_allclose(sc.cpu().numpy(), sc2.cpu().numpy(), rtol=1e-12)
counts = torch.from_numpy(synth.nb_counts(500, 300, seed=3).astype(np.float32)).cuda()
sf = torch.from_numpy(synth.size_factors(counts.cpu().numpy())).cuda()
rng = np.random.default_rng(0)
gi = rng.choice(300, 150, replace=False).astype(np.int32)
mu, sg = rng.random(150) + 0.1, rng.random(150) + 0.5
comps, mean = rng.normal(size=(g, 150)), rng.normal(size=150)
p1 = core.project(counts, gi, sf, mu, sg, comps, mean, engine="mma")
p2 = core.project(counts, gi, sf, mu, sg, comps, mean, engine="simple")
np.testing.assert_allclose(p1.cpu().numpy(), p2.cpu().numpy(), rtol=0, atol=1e-11 * float(p2.abs().max()))
torch.cuda.synchronize()
print("sanitize_small OK")

import os, sys, tempfile, cProfile, pstats
import numpy as np
sys.path.insert(0, ".")
from nabo_b200 import Mapping, Graph, store, synth, build
build.build()
n_ref = n_tgt = 100000; g, k = 50, 30
d = tempfile.mkdtemp()
ref_fn, tgt_fn, map_fn = (os.path.join(d, x) for x in ("ref.h5", "tgt.h5", "map.h5"))
rn, tn = synth.cell_names(n_ref, "R"), synth.cell_names(n_tgt, "T")
for fn, names, mat in ((ref_fn, rn, synth.pc_mixture(n_ref, g, 1)), (tgt_fn, tn, synth.pc_mixture(n_tgt, g, 101))):
    h = store.File(fn, "w"); h.create_row_group("data", names, mat); h.close()
import torch; torch.zeros(1).cuda(); torch.cuda.synchronize()      # CUDA context + torch lazy init outside the profile
m = Mapping(map_fn, "REF", ref_fn, "data", overwrite=True)
m.set_parameters(g, k, 0.25, 1000)
pr = cProfile.Profile(); pr.enable(); m.make_ref_graph(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
pr = cProfile.Profile(); pr.enable(); m.map_target("TGT", tgt_fn, "data"); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

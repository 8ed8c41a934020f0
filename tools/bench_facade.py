"""Drop-in surface timed end to end (development probe): Mapping.make_ref_graph + map_target + Graph loads +
get_mapping_score on config-2-sized inputs, files included.  Usage: python tools/bench_facade.py [n_ref] [n_tgt]"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
from nabo_b200 import Mapping, Graph, store, synth, build

build.build()
n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n_tgt = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g, k = 50, 30
d = tempfile.mkdtemp()
ref_fn, tgt_fn, map_fn = (os.path.join(d, x) for x in ("ref.h5", "tgt.h5", "map.h5"))
rn, tn = synth.cell_names(n_ref, "R"), synth.cell_names(n_tgt, "T")
for fn, names, mat in ((ref_fn, rn, synth.pc_mixture(n_ref, g, 1)), (tgt_fn, tn, synth.pc_mixture(n_tgt, g, 101))):
    h = store.File(fn, "w"); h.create_row_group("data", names, mat); h.close()

def lap(msg, t0):
    t1 = time.perf_counter(); print("%-46s %8.3f s" % (msg, t1 - t0), flush=True); return t1

t = time.perf_counter(); t_all = t
m = Mapping(map_fn, "REF", ref_fn, "data", overwrite=True)
m.set_parameters(g, k, 0.25, 1000)
t = lap("Mapping() + set_parameters", t)
m.make_ref_graph()
t = lap("make_ref_graph (%d cells, Euclidean self-kNN)" % n_ref, t)
m.map_target("TGT", tgt_fn, "data")
t = lap("map_target (%d cells, modified Canberra)" % n_tgt, t)
gph = Graph()
gph.load_from_h5(map_fn, "REF", "reference")
t = lap("Graph.load_from_h5 reference", t)
gph.load_from_h5(map_fn, "TGT", "target")
t = lap("Graph.load_from_h5 target", t)
sc = gph.get_mapping_score("TGT")
t = lap("get_mapping_score", t)
sp = gph.get_mapping_specificity("TGT")
t = lap("get_mapping_specificity (GPU multi-source BFS)", t)
print("total %.3f s; %d scores, max %.3f" % (t - t_all, len(sc), max(sc.values())))

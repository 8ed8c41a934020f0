"""One-off check (needs /root/reference, so it is not a test): a `<uid>_graph` group written by
nabo_b200's `graph_layout='reference'` writer is read by the UNMODIFIED reference `nabo.Graph.load_from_h5`
and gives the reference's own mapping scores.  Run: python tests/golden/verify_reference_reader.py"""
import os, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as MG
from oracle import nabo_oracle as O

nabo = MG.install_shims()
from nabo_b200 import store, synth
from nabo_b200.mapping import _write_reference_layout

g = np.load(os.path.join(HERE, "mapping_small.npz"))
k = int(g["k"])
rn, tn = synth.cell_names(len(g["ref"]), "R"), synth.cell_names(len(g["tgt"]), "T")
rk = g["ref_sorted_full"][:, :k].astype(np.int32)
tk = g["tgt_sorted_full"][:, :k].astype(np.int32)
cnt_t, _ = O.snn_weights(tk, rk, k)
cnt_r, _ = O.snn_weights(rk, rk, k)
# repair edges = golden reference edges that the SNN table does not explain
snn_pairs = {frozenset((int(a), int(b))) for a, b in zip(*np.nonzero(cnt_r > 0)) for b in [rk[a, b]]}
fix = np.array([(int(a), int(b)) for a, b in zip(g["ref_edge_a"], g["ref_edge_b"])
                if frozenset((int(a), int(b))) not in snn_pairs and a < b], dtype=np.int64).reshape(-1, 2)
with tempfile.TemporaryDirectory() as tmp:
    fn = os.path.join(tmp, "map.h5")
    h = store.File(fn, "w")
    h.create_dataset("name_stash/ref_name", data=np.array([b"REF", b"r" * 30]))
    h.create_dataset("name_stash/target_names", data=np.array([[b"TGT", b"t" * 30]]))
    h.create_dataset("ref_cells/ref_cells", data=np.array([x.encode() for x in rn]))
    from nabo_b200 import core
    fw = 0.5 / ((2 * (k - 1)) - 0.5)
    _write_reference_layout(h.create_group("r" * 30 + "_graph"), rn, "REF", rn, "REF", rk, cnt_r, k, fix, fw, True)
    _write_reference_layout(h.create_group("t" * 30 + "_graph"), tn, "TGT", rn, "REF", tk, cnt_t, k,
                            np.zeros((0, 2), np.int64), None, False)
    h.close()
    G = nabo.Graph()
    G.load_from_h5(fn, "REF", "reference")
    G.load_from_h5(fn, "TGT", "target")
    sc = G.get_mapping_score("TGT")
    got = np.array([sc[c + "_REF"] for c in sorted(rn)])
    print("reference Graph read %d nodes, %d edges" % (G.number_of_nodes(), G.number_of_edges()))
    print("scores equal to the reference's own run:", bool(np.allclose(got, g["score_default"], rtol=1e-12, atol=0)))
    exp_ref = {frozenset((int(a), int(b))) for a, b in zip(g["ref_edge_a"], g["ref_edge_b"])}
    got_ref = {frozenset((rn.index(a[:-4]), rn.index(b[:-4]))) for a, b in G.refG.edges()}
    print("reference-graph edges equal:", got_ref == exp_ref)

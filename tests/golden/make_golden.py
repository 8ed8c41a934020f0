#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (nabo 0.4.1).

Runs only in the build container (needs /root/reference).  The reference has no
tests or golden vectors of its own, so these fixtures are the
pin for ``oracle/`` and, through it, for the CUDA path.  Nothing from the
reference is copied: it is imported from where it lies and only its *outputs*
are stored.

Three shims make `import nabo` work here:
  1. ``h5py``  -> ``nabo_b200.store`` in memory-only mode (h5py is not installed)
  2. ``numpy.float = float`` (alias removed from NumPy; nabo/_mapping.py:118)
  3. ``geneStats`` cast to float64 after ``set_gene_stats`` (pandas 3 keeps object
     dtype; nabo/_dataset.py:631-635, 828)
plus empty stand-ins for matplotlib / seaborn / natsort (imported at module top
of the plotting modules, never called here).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import warnings
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nabo_b200 import store, synth  # noqa: E402


def install_shims():
    store.set_memory_only(True)
    sys.modules["h5py"] = store
    np.float = float  # noqa
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "natsort"):
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["natsort"].natsorted = sorted

    class _Anything:                      # plt.style.use(...), plt.rcParams[...] = ... at import
        def __getattr__(self, n):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __setitem__(self, k, v):
            pass

        def __getitem__(self, k):
            return _Anything()
    for attr in ("style", "rcParams", "cm", "rc"):
        setattr(sys.modules["matplotlib.pyplot"], attr, _Anything())
    sys.modules["matplotlib"].rcParams = _Anything()
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference")
    import nabo  # noqa
    return nabo


def write_pca_file(fn, grp, names, mat):
    h5 = store.File(fn, mode="a")
    if grp in h5:
        del h5[grp]
    g = h5.create_group(grp)
    for n, v in zip(names, mat):
        g.create_dataset(n, data=np.array(v, dtype=np.float64))
    h5.close()


def write_dataset_file(fn, counts, cells, genes):
    """nabo dataset layout (nabo/_io.py:103-115, 186-235)."""
    h5 = store.File(fn, mode="w")
    ng = h5.create_group("names")
    ng.create_dataset("genes", data=[g.encode() for g in genes])
    ng.create_dataset("cells", data=[c.encode() for c in cells])
    dt = np.dtype([("idx", np.uint32), ("val", counts.dtype)])
    cg = h5.create_group("cell_data")
    for i, c in enumerate(cells):
        nz = np.nonzero(counts[i])[0]
        a = np.zeros(len(nz), dtype=dt)
        a["idx"], a["val"] = nz, counts[i, nz]
        cg.create_dataset(c, data=a)
    gg = h5.create_group("gene_data")
    for j, g in enumerate(genes):
        nz = np.nonzero(counts[:, j])[0]
        a = np.zeros(len(nz), dtype=dt)
        a["idx"], a["val"] = nz, counts[nz, j]
        gg.create_dataset(g, data=a)
    h5.close()


def read_rows(h5, grp, names, k=None):
    return np.array([h5[grp][n][:k] if k else h5[grp][n][:] for n in names])


# ---------------------------------------------------------------- A: kernels
def golden_kernels(nabo, out):
    from nabo import _mapping as rm
    r = np.random.default_rng(11)
    x = r.normal(size=(37, 13)) * 3
    y = r.normal(size=(53, 13)) * 3
    y[5] = x[3]                     # exact duplicate -> zero distance
    y[6] = y[7]                     # duplicate references -> tie
    x[4, 2] = 0.0                   # |x| = 0 -> Canberra term always 1 in that dim
    x[9] = 0.0
    y[11] = 0.0
    x[12, 5] = np.nan               # NaN in target -> term 1 (Canberra), NaN (Euclid)
    x[20] = y[21] * (1 + 1e-9)      # near-threshold values
    res = {"x": x, "y": y}
    d = np.empty((37, 53), dtype=np.float64)
    rm._euclidean_dist(x, y, d)
    res["euclidean"] = d.copy()
    for f in (0.25, 0.6, 2.0):
        d = np.empty((37, 53), dtype=np.float64)
        rm._mod_canberra_dist(x, y, d, f)
        res["canberra_%s" % str(f).replace(".", "p")] = d.copy()
    np.savez_compressed(os.path.join(out, "kernels.npz"), **res)
    print("kernels.npz", {k: v.shape for k, v in res.items()})


# ---------------------------------------------------------------- B: mapping
def run_mapping(nabo, tmp, ref, tgt, ref_names, tgt_names, use_comps, k, f, chunk,
                ignore=None, tag="m", specificity=True):
    ref_fn = os.path.join(tmp, tag + "_ref_pca.h5")
    tgt_fn = os.path.join(tmp, tag + "_tgt_pca.h5")
    map_fn = os.path.join(tmp, tag + "_map.h5")
    write_pca_file(ref_fn, "data", ref_names, ref)
    write_pca_file(tgt_fn, "data", tgt_names, tgt)
    random.seed(5)
    m = nabo.Mapping(map_fn, "REF", ref_fn, "data", overwrite=True)
    m.set_parameters(use_comps, k, f, chunk)
    m.make_ref_graph()
    m.map_target("TGT", tgt_fn, "data", ignore_ref_cells=ignore)
    h5 = store.File(map_fn, mode="r")
    ruid = h5["name_stash/ref_name"][1].decode()
    tuid = [i[1].decode() for i in h5["name_stash/target_names"] if i[0].decode() == "TGT"][0]
    sref = sorted(ref_names)
    stgt = sorted(tgt_names)
    res = {}
    res["ref_sorted_full"] = read_rows(h5, ruid + "_sortedDist", sref).astype(np.int32)
    res["tgt_sorted_full"] = read_rows(h5, tuid + "_sortedDist", stgt).astype(np.int32)
    res["ref_dist_full"] = read_rows(h5, ruid + "_dist", sref)
    res["tgt_dist_full"] = read_rows(h5, tuid + "_dist", stgt)
    ridx = {n + "_REF": i for i, n in enumerate(sref)}

    def edges(uid, names, suffix):
        rows, cols, ws = [], [], []
        for i, n in enumerate(names):
            for e in h5[uid + "_graph"][n + "_" + suffix]:
                rows.append(i)
                cols.append(ridx[e[0].decode()])
                ws.append(float(e[1].decode()))
        return np.array(rows, np.int32), np.array(cols, np.int32), np.array(ws, np.float64)

    res["tgt_edge_t"], res["tgt_edge_r"], res["tgt_edge_w"] = edges(tuid, stgt, "TGT")
    res["ref_edge_a"], res["ref_edge_b"], res["ref_edge_w"] = edges(ruid, sref, "REF")
    res["graph_dtype"] = np.array(str(h5[tuid + "_graph"][stgt[0] + "_TGT"].dtype))
    h5.close()

    g = nabo.Graph()
    g.load_from_h5(map_fn, "REF", "reference")
    g.load_from_h5(map_fn, "TGT", "target")
    for name, kw in (("score_default", {}),
                     ("score_minw", dict(min_weight=0.12)),
                     ("score_unweighted", dict(weighted=False)),
                     ("score_minscore", dict(min_score=2.0))):
        sc = g.get_mapping_score("TGT", **kw)
        res[name] = np.array([sc[n + "_REF"] for n in sref], dtype=np.float64)
    if not specificity:
        return res
    # mapping specificity (nabo/_graph.py:794-857): raw (NaN kept), NaN-filled, and folded back on the reference
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp_raw = g.get_mapping_specificity("TGT", fill_na=False)
        sp_fill = g.get_mapping_specificity("TGT", fill_na=True)
    res["specificity_raw"] = np.array([sp_raw[n + "_TGT"] for n in stgt], dtype=np.float64)
    res["specificity_filled"] = np.array([sp_fill[n + "_TGT"] for n in stgt], dtype=np.float64)
    rs = g.get_ref_specificity("TGT", sp_fill)
    res["ref_specificity"] = np.array([rs.get(n + "_REF", np.nan) for n in sref], dtype=np.float64)
    rs0 = g.get_ref_specificity("TGT", sp_fill, incl_unmapped=True)
    res["ref_specificity_unmapped0"] = np.array([rs0[n + "_REF"] for n in sref], dtype=np.float64)
    return res


def golden_mapping(nabo, out, tmp):
    ref = synth.pc_mixture(300, 20, seed=1, n_clusters=6)
    tgt = synth.pc_mixture(200, 20, seed=101, n_clusters=6)
    tgt[7] = ref[3]                                  # a target identical to a reference cell
    ref[10] = ref[11]                                # duplicate reference cells
    rn, tn = synth.cell_names(300, "R"), synth.cell_names(200, "T")
    base = {"ref": ref, "tgt": tgt}
    res = run_mapping(nabo, tmp, ref, tgt, rn, tn, use_comps=15, k=11, f=0.25, chunk=64, tag="b1")
    np.savez_compressed(os.path.join(out, "mapping_small.npz"), use_comps=15, k=11, f=0.25, **base, **res)
    print("mapping_small.npz written")
    ign = [rn[i] for i in range(0, 300, 7)]
    res = run_mapping(nabo, tmp, ref, tgt, rn, tn, use_comps=20, k=5, f=0.5, chunk=97, ignore=ign, tag="b2")
    mask = np.zeros(300, dtype=bool)
    mask[::7] = True
    np.savez_compressed(os.path.join(out, "mapping_ignore.npz"), use_comps=20, k=5, f=0.5, mask=mask,
                        **base, **res)
    print("mapping_ignore.npz written")


def golden_mapping_edge(nabo, out, tmp):
    """Edge semantics of the reference itself: more neighbours asked for than there are un-ignored reference
    cells, a target with a NaN coordinate, targets far outside the reference (every Canberra term saturates:
    d == use_comps for all references, one big tie class), an all-zero target."""
    ref = synth.pc_mixture(60, 12, seed=3, n_clusters=3)
    tgt = synth.pc_mixture(40, 12, seed=103, n_clusters=3)
    tgt[3, 5] = np.nan
    tgt[4] = 1e6 * (1.0 + np.arange(12))             # saturates every dimension against every reference cell
    tgt[5] = -tgt[4]
    tgt[6] = 0.0                                     # |x - y| < f|x| is never true for x = 0
    rn, tn = synth.cell_names(60, "R"), synth.cell_names(40, "T")
    ign = [rn[i] for i in range(60) if i % 10 != 0]  # 6 reference cells left, k = 8 asked for
    res = run_mapping(nabo, tmp, ref, tgt, rn, tn, use_comps=10, k=8, f=0.25, chunk=16, ignore=ign, tag="e1",
                      specificity=False)
    mask = np.ones(60, dtype=bool)
    mask[::10] = False
    np.savez_compressed(os.path.join(out, "mapping_edge.npz"), use_comps=10, k=8, f=0.25, mask=mask, ref=ref, tgt=tgt,
                        **res)
    print("mapping_edge.npz written")


# ---------------------------------------------------------------- C: dataset
def golden_dataset(nabo, out, tmp):
    ng, nr, nt, nc = 600, 400, 300, 20
    cr = synth.nb_counts(nr, ng, seed=1)
    ct = synth.nb_counts(nt, ng, seed=101)
    genes = ["G%04d" % i for i in range(ng)]
    rn, tn = synth.cell_names(nr, "R"), synth.cell_names(nt, "T")
    rfn, tfn = os.path.join(tmp, "c_ref.h5"), os.path.join(tmp, "c_tgt.h5")
    write_dataset_file(rfn, cr, rn, genes)
    write_dataset_file(tfn, ct, tn, genes)

    def prep(fn):
        d = nabo.Dataset(fn, force_recalc=True)
        d.set_sf()
        d.set_gene_stats()
        gs = d.geneStats
        for c in ("m", "nzm", "variance", "ncells"):
            gs[c] = gs[c].astype(np.float64)
        gs["valid_gene"] = gs["valid_gene"].astype(bool)
        return d

    dr, dt = prep(rfn), prep(tfn)
    valid = dr.geneStats[dr.geneStats.valid_gene]
    disp = (valid.variance / valid.m).sort_values(ascending=False)
    hvg = [g for g in disp.index[:250] if dt.geneStats.valid_gene[g]]
    dr.fit_ipca(hvg, n_comps=nc, disable_tqdm=True)
    sp = dr.get_scaling_params(hvg)
    pr_fn, pt_fn = os.path.join(tmp, "c_ref_pca.h5"), os.path.join(tmp, "c_tgt_pca.h5")
    dr.transform_pca(pr_fn, "data", dr.ipca, sp, disable_tqdm=True)
    dt.transform_pca(pt_fn, "data", dr.ipca, sp, disable_tqdm=True)
    h = store.File(pr_fn, "r")
    pr = np.array([h["data"][n][:] for n in rn])
    h.close()
    h = store.File(pt_fn, "r")
    pt = np.array([h["data"][n][:] for n in tn])
    h.close()
    scaled_t = np.array([a for _, a in dt.get_scaled_values(sp, disable_tqdm=True)])
    gidx = np.array([genes.index(g) for g in sp.index], dtype=np.int32)
    np.savez_compressed(
        os.path.join(out, "dataset_small.npz"),
        counts_ref=cr.astype(np.int16), counts_tgt=ct.astype(np.int16), gene_idx=gidx,
        sf_ref=dr.sf, sf_tgt=dt.sf, mu=sp["mu"].values.astype(np.float64),
        sigma=sp["sigma"].values.astype(np.float64),
        gene_m_ref=dr.geneStats.loc[genes, "m"].values.astype(np.float64),
        gene_var_ref=dr.geneStats.loc[genes, "variance"].values.astype(np.float64),
        components=dr.ipca.components_, mean=dr.ipca.mean_,
        pca_ref=pr, pca_tgt=pt, scaled_tgt=scaled_t)
    print("dataset_small.npz: hvg %d, pca_ref %s, pca_tgt %s" % (len(hvg), pr.shape, pt.shape))


# ---------------------------------------------------------------- D: config 1 scale
def golden_c1(nabo, out, tmp):
    n, g, k = 5000, 25, 10
    ref = synth.pc_mixture(n, g, seed=1)
    tgt = synth.pc_mixture(n, g, seed=101)
    rn, tn = synth.cell_names(n, "R"), synth.cell_names(n, "T")
    res = run_mapping(nabo, tmp, ref, tgt, rn, tn, use_comps=g, k=k, f=0.25, chunk=1000, tag="d")
    rs, ts = res["ref_sorted_full"][:, :k], res["tgt_sorted_full"][:, :k]
    np.savez_compressed(
        os.path.join(out, "mapping_c1.npz"), n=n, g=g, k=k, f=0.25,
        input_sha=np.array(synth.sha256_of(ref, tgt)),
        ref_knn=rs.astype(np.uint16), tgt_knn=ts.astype(np.uint16),
        ref_knn_dist=np.take_along_axis(res["ref_dist_full"], rs.astype(np.int64), 1),
        tgt_knn_dist=np.take_along_axis(res["tgt_dist_full"], ts.astype(np.int64), 1),
        tgt_edge_t=res["tgt_edge_t"].astype(np.uint16), tgt_edge_r=res["tgt_edge_r"].astype(np.uint16),
        tgt_edge_w=res["tgt_edge_w"].astype(np.float32), score_default=res["score_default"])
    print("mapping_c1.npz written")


# ---------------------------------------------------------------- E: config 1 through Dataset -> Mapping -> Graph
def golden_c1_chain(nabo, out, tmp):
    """BASELINE config 1 at its stated shape through the whole reference chain: 5 000 reference + 5 000 target
    cells, 2 000 HVGs (of 2 400 genes), IncrementalPCA with 25 components, k = 10: Dataset.set_sf / set_gene_stats /
    fit_ipca / transform_pca -> Mapping.make_ref_graph / map_target -> Graph.get_mapping_score."""
    ng, n, nhvg, nc, k = 2400, 5000, 2000, 25, 10
    cr = synth.nb_counts(n, ng, seed=1)
    ct = synth.nb_counts(n, ng, seed=101)
    genes = ["G%04d" % i for i in range(ng)]
    rn, tn = synth.cell_names(n, "R"), synth.cell_names(n, "T")
    rfn, tfn = os.path.join(tmp, "e_ref.h5"), os.path.join(tmp, "e_tgt.h5")
    write_dataset_file(rfn, cr, rn, genes)
    write_dataset_file(tfn, ct, tn, genes)

    def prep(fn):
        d = nabo.Dataset(fn, force_recalc=True)
        d.set_sf()
        d.set_gene_stats()
        gs = d.geneStats
        for c in ("m", "nzm", "variance", "ncells"):
            gs[c] = gs[c].astype(np.float64)
        gs["valid_gene"] = gs["valid_gene"].astype(bool)
        return d

    dr, dt = prep(rfn), prep(tfn)
    valid = dr.geneStats[dr.geneStats.valid_gene]
    disp = (valid.variance / valid.m).sort_values(ascending=False)
    hvg = [g for g in disp.index if dt.geneStats.valid_gene[g]][:nhvg]
    dr.fit_ipca(hvg, n_comps=nc, disable_tqdm=True)
    sp = dr.get_scaling_params(hvg)
    pr_fn, pt_fn = os.path.join(tmp, "e_ref_pca.h5"), os.path.join(tmp, "e_tgt_pca.h5")
    dr.transform_pca(pr_fn, "data", dr.ipca, sp, disable_tqdm=True)
    dt.transform_pca(pt_fn, "data", dr.ipca, sp, disable_tqdm=True)
    h = store.File(pr_fn, "r")
    pr = np.array([h["data"][c][:] for c in rn])
    h.close()
    h = store.File(pt_fn, "r")
    pt = np.array([h["data"][c][:] for c in tn])
    h.close()
    res = run_mapping(nabo, tmp, pr, pt, rn, tn, use_comps=nc, k=k, f=0.25, chunk=1000, tag="e", specificity=False)
    rs, ts = res["ref_sorted_full"][:, :k + 1], res["tgt_sorted_full"][:, :k + 1]      # one rank past k: tie margin
    gidx = np.array([genes.index(g) for g in sp.index], dtype=np.int16)
    np.savez_compressed(
        os.path.join(out, "chain_c1.npz"), n=n, n_genes=ng, n_hvg=len(hvg), n_comps=nc, k=k, f=0.25,
        counts_sha=np.array(synth.sha256_of(cr, ct)), gene_idx=gidx,
        sf_ref=dr.sf, sf_tgt=dt.sf, mu=sp["mu"].values.astype(np.float64), sigma=sp["sigma"].values.astype(np.float64),
        components=dr.ipca.components_, mean=dr.ipca.mean_,
        pca_ref=pr, pca_tgt=pt, ref_knn=rs.astype(np.uint16), tgt_knn=ts.astype(np.uint16),
        ref_knn_dist=np.take_along_axis(res["ref_dist_full"], rs.astype(np.int64), 1),
        tgt_knn_dist=np.take_along_axis(res["tgt_dist_full"], ts.astype(np.int64), 1),
        tgt_edge_t=res["tgt_edge_t"].astype(np.uint16), tgt_edge_r=res["tgt_edge_r"].astype(np.uint16),
        tgt_edge_w=res["tgt_edge_w"].astype(np.float32), score_default=res["score_default"])
    print("chain_c1.npz: hvg %d, pca %s" % (len(hvg), pr.shape))


if __name__ == "__main__":
    nabo = install_shims()
    which = sys.argv[1:] or ["kernels", "mapping", "edge", "dataset", "c1", "chain"]
    with tempfile.TemporaryDirectory() as tmp:
        if "kernels" in which:
            golden_kernels(nabo, HERE)
        if "mapping" in which:
            golden_mapping(nabo, HERE, tmp)
        if "edge" in which:
            golden_mapping_edge(nabo, HERE, tmp)
        if "dataset" in which:
            golden_dataset(nabo, HERE, tmp)
        if "c1" in which:
            golden_c1(nabo, HERE, tmp)
        if "chain" in which:
            golden_c1_chain(nabo, HERE, tmp)

"""Pin the oracle: every restated function against outputs of the unmodified
reference (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import nabo_oracle as O


def bit_equal(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64)) or \
        bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def test_euclidean_bit_exact(golden):
    g = golden("kernels")
    assert bit_equal(O.euclidean_dist(g["x"], g["y"]), g["euclidean"])


@pytest.mark.parametrize("f,key", [(0.25, "canberra_0p25"), (0.6, "canberra_0p6"), (2.0, "canberra_2p0")])
def test_canberra_bit_exact(golden, f, key):
    g = golden("kernels")
    assert bit_equal(O.mod_canberra_dist(g["x"], g["y"], f), g[key])


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_full_distance_rows(golden, name):
    g = golden(name)
    uc = int(g["use_comps"])
    ref, tgt = g["ref"][:, :uc], g["tgt"][:, :uc]
    assert bit_equal(O.euclidean_dist(ref, ref), g["ref_dist_full"])
    assert bit_equal(O.mod_canberra_dist(tgt, ref, float(g["f"])), g["tgt_dist_full"])


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_knn_tie_classes(golden, name):
    g = golden(name)
    uc, k = int(g["use_comps"]), int(g["k"])
    ref, tgt = g["ref"][:, :uc], g["tgt"][:, :uc]
    mask = g["mask"] if "mask" in g.files else None
    # reference <-> reference never masks (make_ref_graph passes [] , _mapping.py:538-539)
    idx, dst = O.knn(ref, ref, k, "euclidean", drop_first=True)
    ridx = g["ref_sorted_full"][:, :k].astype(np.int64)
    rdst = np.take_along_axis(g["ref_dist_full"], ridx, 1)
    assert O.tie_classes_equal(idx, dst, ridx, rdst, head_truncated=True)
    idx, dst = O.knn(tgt, ref, k, "mod_canberra", float(g["f"]), mask=mask)
    tidx = g["tgt_sorted_full"][:, :k].astype(np.int64)
    tdst = np.take_along_axis(g["tgt_dist_full"], tidx, 1)
    if mask is not None:
        assert not mask[tidx].any()          # ignored cells never appear while k <= #unmasked
    assert O.tie_classes_equal(idx, dst, tidx, tdst)


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_full_sorted_rows_masked_last(golden, name):
    g = golden(name)
    mask = g["mask"] if "mask" in g.files else None
    full = O.sorted_neighbours(g["tgt_dist_full"], mask)
    ref_full = g["tgt_sorted_full"].astype(np.int64)
    assert full.shape == ref_full.shape
    d_a = np.take_along_axis(g["tgt_dist_full"], full, 1)
    d_b = np.take_along_axis(g["tgt_dist_full"], ref_full, 1)
    if mask is not None:
        nm = int(mask.sum())
        assert mask[ref_full[:, -nm:]].all() and mask[full[:, -nm:]].all()
        d_a, d_b = d_a[:, :-nm], d_b[:, :-nm]
    assert np.array_equal(d_a, d_b)
    # reference rows drop the first sorted element (_mapping.py:141-142)
    rfull = O.sorted_neighbours(g["ref_dist_full"], None, drop_first=True)
    assert rfull.shape == g["ref_sorted_full"].shape
    assert np.array_equal(np.take_along_axis(g["ref_dist_full"], rfull, 1),
                          np.take_along_axis(g["ref_dist_full"], g["ref_sorted_full"].astype(np.int64), 1))


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_snn_edges_and_weights(golden, name):
    """Edges/weights computed from the REFERENCE's own top-k lists (so tie order
    cannot interfere) must equal the reference graph dump exactly."""
    g = golden(name)
    k = int(g["k"])
    tk = g["tgt_sorted_full"][:, :k].astype(np.int64)
    rk = g["ref_sorted_full"][:, :k].astype(np.int64)
    cnt, w = O.snn_weights(tk, rk, k)
    got = {(t, int(tk[t, j])): w[t, j] for t in range(tk.shape[0]) for j in range(k) if cnt[t, j] > 0}
    exp = {(int(t), int(r)): float(x) for t, r, x in zip(g["tgt_edge_t"], g["tgt_edge_r"], g["tgt_edge_w"])}
    assert got == exp
    assert str(g["graph_dtype"]).startswith("|S")


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_ref_graph_edges(golden, name):
    """Reference graph = SNN edges among reference cells (+ repair edges of
    weight fix_weight, _mapping.py:478-479, which only the facade adds)."""
    g = golden(name)
    k = int(g["k"])
    rk = g["ref_sorted_full"][:, :k].astype(np.int64)
    cnt, w = O.snn_weights(rk, rk, k)
    got = {}
    for t in range(rk.shape[0]):
        for j in range(k):
            if cnt[t, j] > 0:
                got[(t, int(rk[t, j]))] = w[t, j]
    exp = {(int(a), int(b)): float(x) for a, b, x in zip(g["ref_edge_a"], g["ref_edge_b"], g["ref_edge_w"])}
    fw = O.fix_weight(k)
    # nx.Graph is undirected: the dump of node a lists b if either direction made the edge;
    # the later add_edge overwrites the weight.
    und = {}
    for (a, b), x in got.items():
        und.setdefault(frozenset((a, b)), []).append(x)
    for (a, b), x in exp.items():
        key = frozenset((a, b))
        if key in und:
            assert x in und[key]
        else:
            assert abs(x - fw) < 1e-15, (a, b, x)
    assert {frozenset(e) for e in exp} >= set(und)


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_mapping_scores(golden, name):
    g = golden(name)
    k = int(g["k"])
    tk = g["tgt_sorted_full"][:, :k].astype(np.int64)
    rk = g["ref_sorted_full"][:, :k].astype(np.int64)
    _, w = O.snn_weights(tk, rk, k)
    m = rk.shape[0]
    np.testing.assert_allclose(O.mapping_scores(tk, w, m), g["score_default"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(O.mapping_scores(tk, w, m, min_weight=0.12), g["score_minw"], rtol=1e-12)
    np.testing.assert_allclose(O.mapping_scores(tk, w, m, weighted=False), g["score_unweighted"], rtol=1e-12)
    np.testing.assert_allclose(O.mapping_scores(tk, w, m, min_score=2.0), g["score_minscore"], rtol=1e-12)


def test_projection(golden):
    g = golden("dataset_small")
    gi = g["gene_idx"]
    z = O.scale_counts(g["counts_tgt"][:, gi], g["sf_tgt"], g["mu"], g["sigma"])
    assert bit_equal(z, g["scaled_tgt"])
    p = O.pca_transform(z, g["components"], g["mean"])
    np.testing.assert_allclose(p, g["pca_tgt"], rtol=0, atol=1e-12 * np.abs(g["pca_tgt"]).max())
    pr = O.project(g["counts_ref"][:, gi], g["sf_ref"], g["mu"], g["sigma"], g["components"], g["mean"])
    np.testing.assert_allclose(pr, g["pca_ref"], rtol=0, atol=1e-12 * np.abs(g["pca_ref"]).max())


def test_scaling_params_from_counts(golden):
    """mu / sigma restatement (nabo/_dataset.py:612-623, 826-830): float32 normalised
    values, population variance."""
    g = golden("dataset_small")
    x = g["counts_ref"].astype(np.float32) * g["sf_ref"].astype(np.float32)[:, None]
    gi = g["gene_idx"]
    # per gene on a contiguous float32 vector, exactly as the reference does
    m = np.array([np.ascontiguousarray(x[:, j]).mean() for j in gi], dtype=np.float64)
    v = np.array([np.ascontiguousarray(x[:, j]).var() for j in gi], dtype=np.float64)
    assert np.array_equal(m, g["mu"])
    assert np.array_equal(np.sqrt(v), g["sigma"])


def test_c1_scale(golden):
    from nabo_b200 import synth
    g = golden("mapping_c1")
    n, gg, k = int(g["n"]), int(g["g"]), int(g["k"])
    ref, tgt = synth.pc_mixture(n, gg, seed=1), synth.pc_mixture(n, gg, seed=101)
    assert synth.sha256_of(ref, tgt) == str(g["input_sha"])
    idx, dst = O.knn(ref, ref, k, "euclidean", drop_first=True)
    assert O.tie_classes_equal(idx, dst, g["ref_knn"].astype(np.int64), g["ref_knn_dist"], head_truncated=True)
    tidx, tdst = O.knn(tgt, ref, k, "mod_canberra", 0.25)
    assert O.tie_classes_equal(tidx, tdst, g["tgt_knn"].astype(np.int64), g["tgt_knn_dist"])
    _, w = O.snn_weights(g["tgt_knn"].astype(np.int64), g["ref_knn"].astype(np.int64), k)
    np.testing.assert_allclose(O.mapping_scores(g["tgt_knn"].astype(np.int64), w, n), g["score_default"], rtol=1e-12)


def test_c_port_matches_golden_and_numpy(golden):
    """The plain-C port (timed as the CPU baseline) is bit-identical to the reference kernels."""
    from oracle import c_port
    g = golden("kernels")
    assert bit_equal(c_port.dist(g["x"], g["y"], "euclidean"), g["euclidean"])
    assert bit_equal(c_port.dist(g["x"], g["y"], "mod_canberra", 0.25), g["canberra_0p25"])
    m = golden("mapping_ignore")
    uc, k = int(m["use_comps"]), int(m["k"])
    ref, tgt = m["ref"][:, :uc], m["tgt"][:, :uc]
    i1, d1 = c_port.knn(tgt, ref, k, "mod_canberra", float(m["f"]), mask=m["mask"], nthreads=2)
    i2, d2 = O.knn(tgt, ref, k, "mod_canberra", float(m["f"]), mask=m["mask"])
    assert np.array_equal(i1, i2) and bit_equal(d1, d2)
    i1, d1 = c_port.knn(ref, ref, k, "euclidean", drop_first=True)
    i2, d2 = O.knn(ref, ref, k, "euclidean", drop_first=True)
    assert np.array_equal(i1, i2) and bit_equal(d1, d2)
    cnt = c_port.snn_counts(i2.astype(np.int32), i2.astype(np.int32))
    assert np.array_equal(cnt, O.snn_counts(i2, i2))
    lut = O.snn_weight_lut(k)
    np.testing.assert_allclose(c_port.scores(i2, cnt, lut, ref.shape[0]), O.mapping_scores(i2, lut[cnt], ref.shape[0]),
                               rtol=1e-13)


def _mapped_arrays(g):
    """(N, kmax) reference indices per target (-1 padded) + 0/1 counts from the golden target edge list."""
    n = len(g["tgt"])
    t, r = g["tgt_edge_t"].astype(int), g["tgt_edge_r"].astype(int)
    kmax = max(1, int(np.bincount(t, minlength=n).max()))
    knn = np.full((n, kmax), -1, dtype=np.int32)
    cnt = np.zeros((n, kmax), dtype=np.uint8)
    fill = np.zeros(n, dtype=int)
    for a, b in zip(t, r):
        knn[a, fill[a]] = b
        cnt[a, fill[a]] = 1
        fill[a] += 1
    return knn, cnt


@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_mapping_specificity_matches_reference(golden, name):
    """Graph.get_mapping_specificity of the unmodified reference (networkx BFS per pair) == BFS restatement."""
    g = golden(name)
    knn, cnt = _mapped_arrays(g)
    got = O.mapping_specificity(g["ref_edge_a"], g["ref_edge_b"], len(g["ref"]), knn, cnt)
    exp = g["specificity_raw"]
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert np.array_equal(got[~np.isnan(exp)], exp[~np.isnan(exp)])
    assert np.isnan(exp).sum() == (cnt.sum(1) < 2).sum()


def test_edge_semantics_of_the_reference(golden):
    """Oracle kNN on the reference's edge cases (golden mapping_edge): ignored cells fill the tail of a row when
    k exceeds the live cells, NaN coordinates count as saturated terms, fully saturated rows tie at use_comps."""
    g = golden("mapping_edge")
    uc, k, f, mask = int(g["use_comps"]), int(g["k"]), float(g["f"]), g["mask"]
    n_live = int((~mask).sum())
    oi, od = O.knn(g["tgt"][:, :uc], g["ref"][:, :uc], k, "mod_canberra", f, mask=mask)
    gt = g["tgt_sorted_full"][:, :n_live].astype(np.int64)
    assert O.tie_classes_equal(oi[:, :n_live], od[:, :n_live], gt, np.take_along_axis(g["tgt_dist_full"], gt, 1))
    assert mask[oi[:, n_live:]].all() and np.isnan(od[:, n_live:]).all()
    assert (od[[4, 5, 6], :n_live] == uc).all()
    full = O.mod_canberra_dist(g["tgt"][:, :uc], g["ref"][:, :uc], f)
    assert np.array_equal(full, g["tgt_dist_full"])          # incl. the NaN-coordinate target and the zero target


def test_config1_chain_golden(golden):
    """BASELINE config 1 at its stated shape (5 000 + 5 000 cells, 2 000 HVGs, 25 PCs, k = 10) through the whole
    reference chain Dataset -> Mapping -> Graph (tests/golden/make_golden.py::golden_c1_chain): the oracle, fed the
    regenerated counts and the reference's model, reproduces the projection, both neighbour tables, the target
    edges with their weights and the mapping scores."""
    from nabo_b200 import synth
    g = golden("chain_c1")
    n, ng, k = int(g["n"]), int(g["n_genes"]), int(g["k"])
    cr, ct = synth.nb_counts(n, ng, seed=1), synth.nb_counts(n, ng, seed=101)
    assert synth.sha256_of(cr, ct) == str(g["counts_sha"])
    gi = g["gene_idx"].astype(np.int64)
    for counts, sf, exp in ((cr, g["sf_ref"], g["pca_ref"]), (ct, g["sf_tgt"], g["pca_tgt"])):
        p = O.project(counts[:, gi], sf, g["mu"], g["sigma"], g["components"], g["mean"])
        np.testing.assert_allclose(p, exp, rtol=0, atol=1e-12 * np.abs(exp).max())
    ri, rd = O.knn(g["pca_ref"], g["pca_ref"], k, "euclidean", drop_first=True)
    ti, td = O.knn(g["pca_tgt"], g["pca_ref"], k, "mod_canberra", float(g["f"]))
    assert np.array_equal(ri, g["ref_knn"][:, :k]) and np.array_equal(rd, g["ref_knn_dist"][:, :k])
    assert np.array_equal(ti, g["tgt_knn"][:, :k]) and np.array_equal(td, g["tgt_knn_dist"][:, :k])
    cnt, w = O.snn_weights(ti, ri, k)
    t, j = np.nonzero(cnt > 0)
    got = sorted(zip(t.tolist(), ti[t, j].tolist(), w[t, j].astype(np.float32).tolist()))
    exp = sorted(zip(g["tgt_edge_t"].tolist(), g["tgt_edge_r"].tolist(), g["tgt_edge_w"].tolist()))
    assert got == exp
    np.testing.assert_allclose(O.mapping_scores(ti, w, n, counts=cnt), g["score_default"], rtol=1e-12)

"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors
made by the unmodified reference.  Bit-exact for distances / indices / SNN counts /
weights; stated tolerances for projection and scores."""
import numpy as np
import pytest

from oracle import nabo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core():
    from nabo_b200 import build, core as c
    build.build()
    return c


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def rand(m, g, seed, scale=3.0):
    return np.random.default_rng(seed).normal(size=(m, g)) * scale


# ------------------------------------------------------------------ distance tiles
def test_dist_tiles_golden(core, golden):
    g = golden("kernels")
    x, y = g["x"], g["y"]
    d = np.empty((x.shape[0], y.shape[0]))
    assert core.euclidean_dist(x, y, d) is d            # caller-allocated, filled in place
    assert same_bits(d, g["euclidean"])
    for f, key in ((0.25, "canberra_0p25"), (0.6, "canberra_0p6"), (2.0, "canberra_2p0")):
        assert same_bits(core.mod_canberra_dist(x, y, None, f), g[key])


@pytest.mark.parametrize("m,n,g", [(1, 1, 1), (64, 64, 50), (65, 130, 25), (200, 333, 50), (17, 700, 100)])
def test_dist_tiles_oracle(core, m, n, g):
    x, y = rand(m, g, 1), rand(n, g, 2)
    assert same_bits(core.euclidean_dist(x, y), O.euclidean_dist(x, y))
    assert same_bits(core.mod_canberra_dist(x, y, None, 0.25), O.mod_canberra_dist(x, y, 0.25))
    assert same_bits(core.cosine_dist(x, y), O.cosine_dist(x, y))


def test_dist_empty_and_errors(core):
    assert core.euclidean_dist(np.zeros((0, 5)), np.zeros((3, 5))).shape == (0, 3)
    with pytest.raises(ValueError):
        core.mod_canberra_dist(np.zeros((2, 5)), np.zeros((3, 5)), None, 0.0)
    with pytest.raises(ValueError):
        core.euclidean_dist(np.zeros((2, 5)), np.zeros((3, 4)))


# ------------------------------------------------------------------ kNN, exact engine
@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("metric", ["euclidean", "mod_canberra", "cosine"])
@pytest.mark.parametrize("n,m,g,k", [(100, 257, 15, 11), (300, 1000, 50, 30), (5, 70, 3, 64), (70, 64, 25, 1)])
def test_knn_matches_oracle(core, mode, metric, n, m, g, k):
    q, r = rand(n, g, 3), rand(m, g, 4)
    if metric == "mod_canberra":
        q, r = q + 4, r + 4        # keep a useful share of unsaturated dimensions
    idx, dst = core.knn(q, r, k, metric, 0.25, mode=mode)
    oi, od = O.knn(q, r, k, metric, 0.25)
    assert idx.dtype == np.int32 and dst.dtype == np.float64
    assert same_bits(dst, od)
    assert np.array_equal(idx, oi)             # same (distance, index) order as the oracle


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_knn_ties_mask_dropfirst(core, mode):
    r = rand(400, 20, 5)
    r[10] = r[11]
    r[50:60] = r[50]
    q = r.copy()
    # reference <-> reference: drop the first sorted element
    idx, dst = core.knn(q, r, 15, "euclidean", drop_first=True, mode=mode)
    oi, od = O.knn(q, r, 15, "euclidean", drop_first=True)
    assert same_bits(dst, od) and np.array_equal(idx, oi)
    # ignore_ref_cells: masked cells sort last; with k > #unmasked they fill the tail as NaN
    mask = np.ones(400, bool)
    mask[::9] = False
    k = int((~mask).sum()) + 3
    idx, dst = core.knn(q[:50], r, k, "mod_canberra", 0.4, ref_mask=mask, mode=mode)
    oi, od = O.knn(q[:50], r, k, "mod_canberra", 0.4, mask=mask)
    assert same_bits(dst, od) and np.array_equal(idx, oi)
    assert np.isnan(dst[:, -3:]).all() and mask[idx[:, -3:]].all()
    # index offset (reference-sharded mode)
    idx2, _ = core.knn(q[:50], r, 7, "euclidean", idx_offset=1000, mode=mode)
    idx1, _ = core.knn(q[:50], r, 7, "euclidean", mode=mode)
    assert np.array_equal(idx2, idx1 + 1000)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_knn_nan_and_saturated(core, mode):
    q, r = rand(40, 12, 6), rand(90, 12, 7)
    q[3, 4] = np.nan
    q[5] = 0.0                  # |x| = 0: every Canberra term is 1 -> all distances tie at g
    for metric in ("euclidean", "mod_canberra"):
        idx, dst = core.knn(q, r, 9, metric, 0.25, mode=mode)
        oi, od = O.knn(q, r, 9, metric, 0.25)
        assert same_bits(dst, od)
        assert np.array_equal(idx, oi)


def test_knn_argument_errors(core):
    q, r = rand(4, 3, 1), rand(5, 3, 2)
    with pytest.raises(ValueError):
        core.knn(q, r, 6)
    with pytest.raises(ValueError):
        core.knn(q, r, 5, drop_first=True)
    with pytest.raises(ValueError):
        core.knn(q, r, 2, metric="manhattan")
    with pytest.raises(ValueError):
        core.knn(q, r[:, :2], 2)


def test_rerank_exact(core):
    q, r = rand(64, 30, 8), rand(500, 30, 9)
    oi, od = O.knn(q, r, 10, "euclidean")
    rng = np.random.default_rng(0)
    cand = np.concatenate([oi, rng.integers(0, 500, size=(64, 20))], axis=1).astype(np.int32)
    for row in cand:                      # candidates must be distinct
        seen = set()
        for j in range(len(row)):
            if row[j] in seen:
                row[j] = -1
            seen.add(row[j])
    perm = rng.permutation(cand.shape[1])
    idx, dst = core.rerank_exact(q, r, cand[:, perm], 10, "euclidean")
    assert same_bits(dst, od) and np.array_equal(idx, oi)


# ------------------------------------------------------------------ golden mapping (reference outputs)
@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("name", ["mapping_small", "mapping_ignore"])
def test_golden_mapping(core, golden, name, mode):
    g = golden(name)
    uc, k, f = int(g["use_comps"]), int(g["k"]), float(g["f"])
    ref = np.ascontiguousarray(g["ref"][:, :uc])
    tgt = np.ascontiguousarray(g["tgt"][:, :uc])
    mask = g["mask"] if "mask" in g.files else None
    ridx, rdst = core.knn(ref, ref, k, "euclidean", drop_first=True, mode=mode)
    gi = g["ref_sorted_full"][:, :k].astype(np.int64)
    assert O.tie_classes_equal(ridx, rdst, gi, np.take_along_axis(g["ref_dist_full"], gi, 1), head_truncated=True)
    tidx, tdst = core.knn(tgt, ref, k, "mod_canberra", f, ref_mask=mask, mode=mode)
    gt = g["tgt_sorted_full"][:, :k].astype(np.int64)
    assert O.tie_classes_equal(tidx, tdst, gt, np.take_along_axis(g["tgt_dist_full"], gt, 1))
    # weights / edges / scores from the reference's own neighbour lists
    cnt, w = core.snn_weights(gt.astype(np.int32), gi.astype(np.int32), k)
    got = {(t, int(gt[t, j])): w[t, j] for t in range(gt.shape[0]) for j in range(k) if cnt[t, j] > 0}
    exp = {(int(t), int(r)): float(x) for t, r, x in zip(g["tgt_edge_t"], g["tgt_edge_r"], g["tgt_edge_w"])}
    assert got == exp
    m = ref.shape[0]
    np.testing.assert_allclose(core.mapping_scores(gt.astype(np.int32), cnt, m, k), g["score_default"], rtol=1e-12)
    np.testing.assert_allclose(core.mapping_scores(gt.astype(np.int32), cnt, m, k, min_weight=0.12),
                               g["score_minw"], rtol=1e-12)
    np.testing.assert_allclose(core.mapping_scores(gt.astype(np.int32), cnt, m, k, weighted=False),
                               g["score_unweighted"], rtol=1e-12)
    np.testing.assert_allclose(core.mapping_scores(gt.astype(np.int32), cnt, m, k, min_score=2.0),
                               g["score_minscore"], rtol=1e-12)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_golden_c1(core, golden, mode):
    """Config 1 shape (5k x 5k, 25 PCs, k=10) against the reference's own output."""
    from nabo_b200 import synth
    g = golden("mapping_c1")
    n, gg, k = int(g["n"]), int(g["g"]), int(g["k"])
    ref, tgt = synth.pc_mixture(n, gg, seed=1), synth.pc_mixture(n, gg, seed=101)
    assert synth.sha256_of(ref, tgt) == str(g["input_sha"])
    ridx, rdst = core.knn(ref, ref, k, "euclidean", drop_first=True, mode=mode)
    assert O.tie_classes_equal(ridx, rdst, g["ref_knn"].astype(np.int64), g["ref_knn_dist"], head_truncated=True)
    assert np.array_equal(ridx, g["ref_knn"].astype(np.int32))      # no ties in this draw: exact
    res = core.map_cells(tgt, ref, ridx, k, dist_factor=0.25, mode=mode)
    assert O.tie_classes_equal(res["idx"], res["dist"], g["tgt_knn"].astype(np.int64), g["tgt_knn_dist"])
    # scores from the reference's lists (tie order cannot interfere)
    cnt, _ = core.snn_weights(g["tgt_knn"].astype(np.int32), g["ref_knn"].astype(np.int32), k)
    np.testing.assert_allclose(core.mapping_scores(g["tgt_knn"].astype(np.int32), cnt, n, k),
                               g["score_default"], rtol=1e-12)
    if np.array_equal(res["idx"], g["tgt_knn"].astype(np.int32)):
        np.testing.assert_allclose(res["scores"], g["score_default"], rtol=1e-12)


# ------------------------------------------------------------------ SNN / scores / classify / merge
# k = 130 / 200: the 1 024-slot table at 13 % / 20 % load, keys displaced by two and more slots (the generic probe path)
@pytest.mark.parametrize("n,m,k", [(1, 40, 5), (333, 500, 30), (1000, 200, 11), (64, 300, 130), (200, 400, 200)])
def test_snn_and_scores_oracle(core, n, m, k):
    rng = np.random.default_rng(n + m + k)
    ref_knn = np.array([rng.choice(m, size=k, replace=False) for _ in range(m)], dtype=np.int32)
    tgt_knn = np.array([rng.choice(m // 2, size=k, replace=False) for _ in range(n)], dtype=np.int32)
    cnt, w = core.snn_weights(tgt_knn, ref_knn, k)
    oc, ow = O.snn_weights(tgt_knn, ref_knn, k)
    assert np.array_equal(cnt, oc) and np.array_equal(w, ow)
    for kw in ({}, dict(min_weight=0.1), dict(weighted=False), dict(min_score=5.0), dict(score_multiplier=1.0)):
        got = core.mapping_scores(tgt_knn, cnt, m, k, **kw)
        exp = O.mapping_scores(tgt_knn, ow, m, **kw)
        np.testing.assert_allclose(got, exp, rtol=1e-13, atol=0)
    inc = rng.random(n) < 0.5
    if inc.any():
        got = core.mapping_scores(tgt_knn, cnt, m, k, include=inc)
        exp = O.mapping_scores(tgt_knn, ow, m, include=np.nonzero(inc)[0])
        np.testing.assert_allclose(got, exp, rtol=1e-13)


def test_scores_are_deterministic_and_hub_safe(core):
    rng = np.random.default_rng(3)
    m, n, k = 3000, 20000, 15
    ref_knn = np.array([rng.choice(m, size=k, replace=False) for _ in range(m)], dtype=np.int32)
    tgt_knn = np.array([rng.choice(40, size=k, replace=False) for _ in range(n)], dtype=np.int32)  # 40 hubs
    cnt, w = core.snn_weights(tgt_knn, ref_knn, k)
    a = core.mapping_scores(tgt_knn, cnt, m, k)
    b = core.mapping_scores(tgt_knn, cnt, m, k)
    assert np.array_equal(a, b)
    np.testing.assert_allclose(a, O.mapping_scores(tgt_knn, w, m), rtol=1e-13)
    # property: sum of scores = 1000 * sum of weights / N
    np.testing.assert_allclose(a.sum(), 1000.0 * w.sum() / n, rtol=1e-12)


def test_classify_targets(core):
    rng = np.random.default_rng(5)
    m, n, k, nl = 400, 700, 11, 6
    ref_knn = np.array([rng.choice(m, size=k, replace=False) for _ in range(m)], dtype=np.int32)
    tgt_knn = np.array([rng.choice(m // 3, size=k, replace=False) for _ in range(n)], dtype=np.int32)
    labels = rng.integers(-1, nl, size=m).astype(np.int32)
    cnt, w = core.snn_weights(tgt_knn, ref_knn, k)
    for kw in (dict(), dict(weight_frac=0.3, min_degree=1, min_weight=0.0), dict(min_degree=6)):
        got = core.classify_targets(tgt_knn, cnt, labels, nl, k, **kw)
        okw = dict(weight_frac=0.5, min_degree=2, min_weight=0.1)
        okw.update(kw)
        exp = O.classify_targets(tgt_knn, w, labels, nl, **okw)
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("s,n,k", [(2, 100, 30), (8, 257, 30), (3, 10, 1)])
def test_merge_topk(core, s, n, k):
    rng = np.random.default_rng(s * n)
    idx = rng.permutation(s * n * k).reshape(s, n, k).astype(np.int32)
    dist = np.sort(rng.random((s, n, k)), axis=2)
    dist[0, :, -1] = np.nan
    dist[:, 0, :] = 0.5                     # full tie row: index order decides
    oi, od = O.merge_topk(list(idx), list(dist), k)
    gi, gd = core.merge_topk(idx, dist)
    assert np.array_equal(gi, oi) and same_bits(gd, od)


def test_sharded_knn_equals_unsharded(core):
    """reference-sharded mode: local top-k per shard + merge == single-shard top-k, bit for bit."""
    q, r = rand(200, 25, 11), rand(1003, 25, 12)
    r[500] = r[3]
    full_i, full_d = core.knn(q, r, 20, "euclidean", mode="exact")
    bounds = [0, 251, 502, 760, 1003]
    parts = [core.knn(q, r[a:b], 20, "euclidean", idx_offset=a, mode="exact") for a, b in zip(bounds, bounds[1:])]
    mi, md = core.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(mi, full_i) and same_bits(md, full_d)


# ------------------------------------------------------------------ projection
def test_projection_golden(core, golden):
    g = golden("dataset_small")
    gi = g["gene_idx"].astype(np.int32)
    for counts, sf, exp in ((g["counts_tgt"], g["sf_tgt"], g["pca_tgt"]), (g["counts_ref"], g["sf_ref"], g["pca_ref"])):
        p = core.project(counts.astype(np.float32), gi, sf, g["mu"], g["sigma"], g["components"], g["mean"])
        tol = 1e-11 * np.abs(exp).max()     # FP64 with a different summation order than BLAS
        np.testing.assert_allclose(p, exp, rtol=0, atol=tol)
        # CSR form over all genes of the dataset
        import scipy.sparse as sp
        csr = sp.csr_matrix(counts.astype(np.float32))
        pos = np.full(counts.shape[1], -1, np.int32)
        pos[gi] = np.arange(len(gi), dtype=np.int32)
        p2 = core.project_csr(csr.indptr.astype(np.int64), csr.indices.astype(np.int32), csr.data, pos, sf,
                              g["mu"], g["sigma"], g["components"], g["mean"])
        np.testing.assert_allclose(p2, exp, rtol=0, atol=tol)


def test_projection_missing_genes_and_oracle(core):
    from nabo_b200 import synth
    counts = synth.nb_counts(70, 300, seed=4).astype(np.float32)
    sf = synth.size_factors(counts)
    rng = np.random.default_rng(1)
    G, nc = 120, 50
    gi = rng.choice(300, size=G, replace=False).astype(np.int32)
    gi[::10] = -1                               # fill_missing=True: gene absent -> value 0
    mu, sigma = rng.random(G) + 0.1, rng.random(G) + 0.5
    comps, mean = rng.normal(size=(nc, G)), rng.normal(size=G)
    dense = np.where(gi[None, :] >= 0, counts[:, np.maximum(gi, 0)], 0.0)
    exp = O.project(dense, sf, mu, sigma, comps, mean)
    got = core.project(counts, gi, sf, mu, sigma, comps, mean)
    np.testing.assert_allclose(got, exp, rtol=0, atol=1e-11 * np.abs(exp).max())


def test_map_cells_host_pipeline_equals_device_path(core):
    import torch
    from nabo_b200 import synth
    n, m, g, k = 70000, 20000, 20, 10
    ref = torch.from_numpy(synth.pc_mixture(m, g, seed=1)).cuda()
    tgt = synth.pc_mixture(n, g, seed=101)
    rk, _ = core.knn(ref, ref, k, "euclidean", drop_first=True)
    a = core.map_cells(torch.from_numpy(tgt).cuda(), ref, rk, k, metric="euclidean")
    h = core.map_cells_host(torch.from_numpy(tgt).pin_memory(), ref, rk, k, metric="euclidean")
    assert torch.equal(h["idx"], a["idx"].cpu()) and torch.equal(h["dist"], a["dist"].cpu())
    assert torch.equal(h["weights"], a["weights"].cpu())
    # the pipeline adds up integer weight sums piece by piece (same bits for any piece order / GPU count);
    # map_cells sums the FP64 weights in target order like the reference: equal to rounding
    whole = core.scores_finalize(core.score_accumulate(a["idx"], a["counts"], m, k), n).cpu()
    assert torch.equal(h["scores"], whole)
    np.testing.assert_allclose(h["scores"].numpy(), a["scores"].cpu().numpy(), rtol=1e-12)
    h2 = core.map_cells_host(torch.from_numpy(tgt).pin_memory(), ref, rk, k, metric="euclidean", chunks=1)
    assert torch.equal(h2["scores"], whole)


@pytest.mark.parametrize("m,n,k,deg,comps", [(500, 300, 8, 3, 1), (4000, 600, 30, 2, 1), (2000, 400, 12, 2, 3),
                                             (300, 100, 64, 4, 1)])
def test_mapping_specificity_matches_oracle(core, m, n, k, deg, comps):
    """Bit-parallel multi-source BFS == one BFS per source (oracle) on sparse random graphs: long paths,
    targets with 0 / 1 / many mapped cells, padding (-1) entries, and - with several components - pairs
    without a path, which must be flagged (the reference raises there)."""
    rng = np.random.default_rng(m + k)
    comp = rng.integers(0, comps, size=m)
    a, b = [], []
    for c in range(comps):
        nodes = np.nonzero(comp == c)[0]
        ring = np.roll(nodes, 1)                                   # connected inside a component
        a += list(nodes); b += list(ring)
        for _ in range(deg - 1):
            a += list(nodes); b += list(rng.permutation(nodes))
    a, b = np.array(a), np.array(b)
    import scipy.sparse as sp
    keep = a != b
    adj = sp.coo_matrix((np.ones(2 * keep.sum(), np.int8), (np.r_[a[keep], b[keep]], np.r_[b[keep], a[keep]])),
                        shape=(m, m)).tocsr()
    adj.sum_duplicates()
    knn = np.stack([rng.choice(m, k, replace=False) for _ in range(n)]).astype(np.int32)
    cnt = (rng.random((n, k)) < 0.6).astype(np.uint8)
    cnt[0] = 0                                                      # nothing mapped
    cnt[1] = 0; cnt[1, 3] = 1                                       # one cell mapped
    knn[2, ::2] = -1                                                # padding entries are not edges
    if comps > 1:
        same = np.array([len(set(comp[r] for r, c in zip(knn[t], cnt[t]) if c and r >= 0)) <= 1 for t in range(n)])
    else:
        same = np.ones(n, dtype=bool)
    mean, connected = core.mapping_specificity(adj.indptr.astype(np.int64), adj.indices.astype(np.int32), knn, cnt)
    assert np.array_equal(connected, same)
    sel = np.nonzero(same)[0]
    exp = O.mapping_specificity(a, b, m, knn[sel], cnt[sel])
    assert np.array_equal(np.isnan(mean[sel]), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.array_equal(mean[sel][ok], exp[ok])
    assert np.isnan(mean[0]) and np.isnan(mean[1])


@pytest.mark.parametrize("n,deg,seed", [(1, 0, 0), (50, 0, 1), (1000, 1, 2), (5000, 2, 3), (100000, 3, 4), (3000, 40, 5)])
def test_connected_components_match_scipy(core, n, deg, seed):
    """GPU union-find labels (= smallest node id of the component) against scipy's connected components on
    random sparse graphs: isolated nodes, chains, dense blobs, self loops, duplicate and out-of-range edges."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    rng = np.random.default_rng(seed)
    e = n * deg // 2
    a = rng.integers(0, n, size=e).astype(np.int32)
    b = rng.integers(0, n, size=e).astype(np.int32)
    if n >= 1000:                                              # a long chain: worst case for label propagation
        chain = np.arange(0, n // 4 - 1, dtype=np.int32)
        a, b = np.concatenate([a, chain]), np.concatenate([b, chain + 1])
    if e:
        a[:3], b[:3] = a[3:6], a[3:6]                          # self loops
    lab = core.connected_components(a, b, n)
    keep = a != b
    ncomp, ref = connected_components(sp.coo_matrix((np.ones(keep.sum(), np.int8), (a[keep], b[keep])), shape=(n, n)),
                                      directed=False)
    assert len(np.unique(lab)) == ncomp
    # same partition, and every label is the smallest member of its class
    first = {}
    for i, (x, y) in enumerate(zip(lab, ref)):
        assert first.setdefault(y, x) == x
    assert all(lab[x] == x for x in np.unique(lab)) and (lab <= np.arange(n)).all()


@pytest.mark.parametrize("shape", ["path", "cycle", "star", "tree", "grid", "clique"])
def test_mapping_specificity_structured_graphs(core, shape):
    """Odd and even pair distances, long paths, hubs: the lock-step meeting rule (2L - 1 when a bit joins a node
    that already holds the other, 2L when both arrive together) must give networkx's distances exactly."""
    import scipy.sparse as sp
    rng = np.random.default_rng(len(shape))
    if shape == "path":
        m = 90; a = np.arange(m - 1); b = a + 1
    elif shape == "cycle":
        m = 77; a = np.arange(m); b = (a + 1) % m
    elif shape == "star":
        m = 200; a = np.zeros(m - 1, dtype=int); b = np.arange(1, m)
    elif shape == "tree":
        m = 255; b = np.arange(1, m); a = (b - 1) // 2
    elif shape == "grid":
        side = 17; m = side * side
        idx = np.arange(m).reshape(side, side)
        a = np.r_[idx[:, :-1].ravel(), idx[:-1, :].ravel()]; b = np.r_[idx[:, 1:].ravel(), idx[1:, :].ravel()]
    else:
        m = 40; a, b = np.triu_indices(m, 1)
    adj = sp.coo_matrix((np.ones(2 * len(a), np.int8), (np.r_[a, b], np.r_[b, a])), shape=(m, m)).tocsr()
    adj.sum_duplicates()
    n, k = 120, 9
    knn = np.stack([rng.choice(m, k, replace=False) for _ in range(n)]).astype(np.int32)
    knn[0] = np.array([0, m - 1, m // 2, 1, m - 2, m // 3, 2, m - 3, m // 4])[:k]      # the extremes together
    cnt = (rng.random((n, k)) < 0.8).astype(np.uint8)
    cnt[0] = 1
    mean, connected = core.mapping_specificity(adj.indptr.astype(np.int64), adj.indices.astype(np.int32), knn, cnt)
    assert connected.all()
    exp = O.mapping_specificity(a, b, m, knn, cnt)
    assert np.array_equal(mean, exp, equal_nan=True)

"""Randomised parity sweep: seeded random shapes, metrics, masks, NaNs, duplicates, k, offsets.
The fast engine must equal the exact engine bit for bit everywhere, and the exact engine the oracle."""
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


def _case(seed):
    rng = np.random.default_rng(seed)
    g = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 25, 31, 33, 50, 56, 57, 64, 65, 100]))
    m = int(rng.choice([40, 127, 128, 129, 300, 1000, 2049, 5000]))
    n = int(rng.choice([1, 31, 32, 33, 100, 383, 385, 700]))
    k = int(min(rng.choice([1, 2, 5, 10, 15, 30, 31, 32, 33, 60, 90, 120]), m - 2))
    metric = str(rng.choice(["euclidean", "mod_canberra", "cosine"]))
    f = float(rng.choice([0.1, 0.25, 0.6, 1.0, 2.0]))
    scale = 10.0 ** rng.uniform(-3, 3, size=g) if rng.random() < 0.5 else np.ones(g)
    centers = rng.normal(size=(4, g)) * 3
    r = (centers[rng.integers(0, 4, m)] + rng.normal(size=(m, g))) * scale
    q = (centers[rng.integers(0, 4, n)] + rng.normal(size=(n, g))) * scale
    if rng.random() < 0.5:                                   # exact duplicates: ties across and inside the top-k
        r[rng.integers(0, m, m // 10)] = r[rng.integers(0, m)]
        q[rng.integers(0, n, max(1, n // 10))] = r[rng.integers(0, m)]
    if rng.random() < 0.3:
        r[rng.integers(0, m, 3), rng.integers(0, g, 3)] = np.nan
        q[rng.integers(0, n, 2), rng.integers(0, g, 2)] = np.nan
    if rng.random() < 0.3:
        q[rng.integers(0, n), :] = 0.0
        r[rng.integers(0, m), :] = 0.0
    mask = (rng.random(m) < rng.choice([0.05, 0.5])) if rng.random() < 0.4 else None
    if mask is not None and (~mask).sum() < k + 2:
        mask = None
    off = int(rng.choice([0, 0, 1000]))
    return dict(q=q, r=r, k=k, metric=metric, f=f, mask=mask, off=off)


@pytest.mark.parametrize("seed", range(160))
def test_random_knn_case(core, seed):
    c = _case(seed)
    kw = dict(ref_mask=c["mask"], idx_offset=c["off"])
    fi, fd = core.knn(c["q"], c["r"], c["k"], c["metric"], c["f"], mode="fast", **kw)
    ei, ed = core.knn(c["q"], c["r"], c["k"], c["metric"], c["f"], mode="exact", **kw)
    assert same_bits(fd, ed) and np.array_equal(fi, ei), "fast != exact"
    if len(c["q"]) * len(c["r"]) <= 4_000_000:               # oracle on the cases it finishes quickly
        oi, od = O.knn(c["q"], c["r"], c["k"], c["metric"], c["f"], mask=c["mask"])
        assert same_bits(ed, od) and np.array_equal(ei, np.where(oi >= 0, oi + c["off"], oi)), "exact != oracle"


@pytest.mark.parametrize("seed", range(30))
def test_random_self_knn_case(core, seed):
    """Reference <-> reference mode (drop_first): the first sorted element is dropped, whatever it is."""
    c = _case(1000 + seed)
    r = c["r"]
    k = min(c["k"], len(r) - 2)
    metric = "euclidean" if seed % 2 == 0 else c["metric"]
    fi, fd = core.knn(r, r, k, metric, c["f"], drop_first=True, mode="fast")
    ei, ed = core.knn(r, r, k, metric, c["f"], drop_first=True, mode="exact")
    assert same_bits(fd, ed) and np.array_equal(fi, ei)
    if len(r) <= 1000:
        oi, od = O.knn(r, r, k, metric, c["f"], drop_first=True)
        assert same_bits(ed, od) and np.array_equal(ei, oi)


@pytest.mark.parametrize("seed", range(24))
def test_random_graph_ops(core, seed):
    """SNN counts / weights, mapping scores (all keyword variants), cluster vote and shard merge on random
    index tables of random shape against the oracle."""
    rng = np.random.default_rng(5000 + seed)
    m = int(rng.choice([40, 257, 1000, 5003]))
    n = int(rng.choice([1, 33, 500, 2000]))
    k = int(min(rng.choice([1, 3, 10, 30, 31, 33, 64, 100]), m - 1))
    if k == 2:
        k = 3                                                   # k = 2 raises ZeroDivisionError upstream
    k_ref = int(min(rng.choice([k, k, max(1, k - 2), k + 5]), m - 1))
    near = rng.random() < 0.5                                   # neighbours drawn from a window: many shared ones
    def table(rows, width):
        out = np.empty((rows, width), dtype=np.int32)
        for i in range(rows):
            pool = (np.arange(width * 3) + rng.integers(0, m)) % m if near else np.arange(m)
            out[i] = rng.choice(pool, width, replace=False)
        return out
    tk, rk = table(n, k), table(m, k_ref)
    cnt, w = core.snn_weights(tk, rk, k)
    ocnt, ow = O.snn_weights(tk, rk[:, :min(k, k_ref)], k)
    assert np.array_equal(cnt, ocnt) and np.array_equal(w, ow)
    for kw in (dict(), dict(min_weight=float(np.median(w[w > 0])) if (w > 0).any() else 0.0), dict(weighted=False),
               dict(min_score=5.0), dict(score_multiplier=1.0),
               dict(include=np.sort(rng.choice(n, max(1, n // 3), replace=False)))):
        got = core.mapping_scores(tk, cnt, m, k, **kw)
        exp = O.mapping_scores(tk, w, m, counts=cnt, **kw)
        np.testing.assert_allclose(got, exp, rtol=1e-12, atol=0)
    n_labels = int(rng.choice([1, 3, 17]))
    labels = rng.integers(-1, n_labels, size=m).astype(np.int32)
    for kw in (dict(), dict(weight_frac=0.2, min_degree=1, min_weight=0.0), dict(min_degree=5, min_weight=float(w.max()) / 2)):
        got = core.classify_targets(tk, cnt, labels, n_labels, k, **kw)
        full = dict(weight_frac=0.5, min_degree=2, min_weight=0.1)
        full.update(kw)
        exp = O.classify_targets(tk, w, labels, n_labels, counts=cnt, **full)
        assert np.array_equal(got, exp)
    shards = int(rng.choice([1, 2, 3, 8]))
    si = rng.integers(0, 10 * m, size=(shards, n, k)).astype(np.int32)
    sd = np.round(rng.random((shards, n, k)) * 4, 1)            # coarse values: ties across shards
    sd[rng.random(sd.shape) < 0.05] = np.nan
    order = np.argsort(np.where(np.isnan(sd), np.inf, sd), axis=2, kind="stable")      # shards arrive sorted
    si, sd = np.take_along_axis(si, order, 2), np.take_along_axis(sd, order, 2)
    mi, md = core.merge_topk(si, sd)
    oi, od = O.merge_topk(list(si), list(sd), k)
    assert np.array_equal(md, od, equal_nan=True)
    ok = ~np.isnan(od)
    assert np.array_equal(mi[ok], oi[ok])


@pytest.mark.parametrize("seed", range(10))
def test_random_projection(core, seed):
    """Scaling + projection from dense and CSR counts with missing genes, random sizes."""
    import scipy.sparse as sp
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([1, 33, 700]))
    n_genes = int(rng.choice([50, 333, 2000]))
    G = int(min(rng.choice([7, 64, 500]), n_genes))
    C = int(min(rng.choice([1, 25, 50, 100]), G))
    counts = (rng.gamma(0.3, 8.0, size=(n, n_genes)) * (rng.random((n, n_genes)) < 0.2)).astype(np.int64).astype(np.float32)
    sf = (1000.0 / np.maximum(counts.sum(1), 1)).astype(np.float32)
    gi = rng.choice(n_genes, G, replace=False).astype(np.int32)
    missing = rng.random(G) < 0.1
    mu, sg = rng.random(G) + 0.05, rng.random(G) + 0.3
    comps, mean = rng.normal(size=(C, G)), rng.normal(size=G)
    gi_m = np.where(missing, -1, gi).astype(np.int32)
    dense = counts[:, gi].copy()
    dense[:, missing] = 0.0
    exp = O.project(dense, sf, mu, sg, comps, mean)
    tol = 1e-11 * max(1.0, np.abs(exp).max())
    np.testing.assert_allclose(core.project(counts, gi_m, sf, mu, sg, comps, mean), exp, rtol=0, atol=tol)
    csr = sp.csr_matrix(counts)
    csr.sort_indices()
    pos = np.full(n_genes, -1, dtype=np.int32)
    pos[gi[~missing]] = np.nonzero(~missing)[0].astype(np.int32)
    got = core.project_csr(csr.indptr.astype(np.int64), csr.indices.astype(np.int32), csr.data.astype(np.float32), pos,
                           sf, mu, sg, comps, mean)
    np.testing.assert_allclose(got, exp, rtol=0, atol=tol)

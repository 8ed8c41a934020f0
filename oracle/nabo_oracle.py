"""CPU oracle for the nabo cell-projection hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain NumPy (FP64) restatement of
the reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``nabo_b200`` never imports anything under ``oracle/``.

Pinning: the reference ships no tests and no golden vectors, so
this restatement is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports the
unmodified reference from /root/reference behind an in-memory h5py stand-in) and
committed under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
every function below against those fixtures (bit-exact for distances / indices /
SNN counts / weights, 1e-12 for projection and scores).

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "euclidean_dist", "mod_canberra_dist", "cosine_dist", "dist_matrix",
    "sorted_neighbours", "knn", "snn_weight_lut", "snn_counts", "snn_weights",
    "mapping_scores", "scale_counts", "pca_transform", "project", "fix_weight",
    "merge_topk", "classify_targets",
]


# ----------------------------------------------------------------------------
# distances
# ----------------------------------------------------------------------------
def euclidean_dist(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """nabo/_mapping.py:16-26 (_euclidean_dist).

    d[i, j] = sqrt(sum_k (x[i,k]-y[j,k])**2), accumulated over k in ascending
    order with a separate multiply and add (the numba loop is not FMA
    contracted), sqrt last.  Looping over k with whole-matrix operands gives the
    same per-element operation order, hence bit-identical results.
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m, g = x.shape
    n = y.shape[0]
    td = np.zeros((m, n), dtype=np.float64)
    for k in range(g):
        t = x[:, k][:, None] - y[:, k][None, :]
        td = td + t * t
    return np.sqrt(td)


def mod_canberra_dist(x: np.ndarray, y: np.ndarray, f: float) -> np.ndarray:
    """nabo/_mapping.py:29-45 (_mod_canberra_dist); x = target, y = reference.

    term_k = |x-y| / (|x| + |y| + 0.01)  if |x-y| < f*|x|  else 1
    (a NaN operand makes the comparison false, i.e. the term is 1).
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = float(f)
    m, g = x.shape
    n = y.shape[0]
    d = np.zeros((m, n), dtype=np.float64)
    with np.errstate(invalid="ignore"):
        for k in range(g):
            absx = np.abs(x[:, k])[:, None]
            num = np.abs(x[:, k][:, None] - y[:, k][None, :])
            absy = np.abs(y[:, k])[None, :]
            den = (absx + absy) + 0.01
            t = np.where(num < f * absx, num / den, 1.0)
            d = d + t
    return d


def cosine_dist(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Extension metric (BASELINE.json asks for it; the reference has none, so
    parity is pinned only by this definition): 1 - x.y / (||x|| * ||y||), dot
    product and squared norms accumulated over k in ascending order, separate
    multiply and add."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m, g = x.shape
    n = y.shape[0]
    dot = np.zeros((m, n), dtype=np.float64)
    nx = np.zeros(m, dtype=np.float64)
    ny = np.zeros(n, dtype=np.float64)
    for k in range(g):
        dot = dot + x[:, k][:, None] * y[:, k][None, :]
        nx = nx + x[:, k] * x[:, k]
        ny = ny + y[:, k] * y[:, k]
    with np.errstate(invalid="ignore", divide="ignore"):
        return 1.0 - dot / (np.sqrt(nx)[:, None] * np.sqrt(ny)[None, :])


def dist_matrix(x, y, metric: str, dist_factor: float = 0.25) -> np.ndarray:
    """Metric dispatch of nabo/_mapping.py:119-124 plus the two extensions."""
    if metric == "euclidean":
        return euclidean_dist(x, y)
    if metric == "mod_canberra":
        return mod_canberra_dist(x, y, dist_factor)
    if metric == "cosine":
        return cosine_dist(x, y)
    raise ValueError("unknown metric %r" % (metric,))


# ----------------------------------------------------------------------------
# sort / top-k
# ----------------------------------------------------------------------------
def sorted_neighbours(dist: np.ndarray, mask: np.ndarray | None = None,
                      drop_first: bool = False) -> np.ndarray:
    """nabo/_mapping.py:135-146: per row, argsort of the row with ignored
    reference cells masked (numpy.ma fills them with NaN, so they sort last),
    dropping the first sorted element for reference<->reference rows.

    The reference uses NumPy's default *unstable* sort, so the order inside a
    run of equal distances is unspecified there; this oracle fixes it to
    ascending index (stable sort), which is also what the CUDA path does.
    """
    d = np.array(dist, dtype=np.float64, copy=True)
    if mask is not None:
        d[:, np.asarray(mask, dtype=bool)] = np.nan
    order = np.argsort(d, axis=1, kind="stable")
    if drop_first:
        order = order[:, 1:]
    return order


def knn(x, y, k: int, metric: str, dist_factor: float = 0.25, mask=None,
        drop_first: bool = False, chunk: int = 512):
    """Top-k restatement of _calc_dist (nabo/_mapping.py:48-148) restricted to
    what _calc_snn consumes (``[:k]`` of each sorted row, :190, :193).

    Returns (idx int64 (N,k), dist float64 (N,k)); masked entries that make it
    into the top-k (k > number of unmasked references) carry dist = NaN.
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    m = y.shape[0]
    kk = min(k, m - (1 if drop_first else 0))
    idx = np.empty((n, kk), dtype=np.int64)
    dst = np.empty((n, kk), dtype=np.float64)
    for s in range(0, n, chunk):
        d = dist_matrix(x[s:s + chunk], y, metric, dist_factor)
        if mask is not None:
            d[:, np.asarray(mask, dtype=bool)] = np.nan
        o = np.argsort(d, axis=1, kind="stable")
        if drop_first:
            o = o[:, 1:]
        o = o[:, :kk]
        idx[s:s + chunk] = o
        dst[s:s + chunk] = np.take_along_axis(d, o, axis=1)
    return idx, dst


def tie_classes_equal(idx_a, dist_a, idx_b, dist_b, head_truncated: bool = False) -> bool:
    """True when two top-k results agree up to the order inside runs of exactly
    equal distance (and the membership of the last, possibly truncated, run).
    ``head_truncated``: the first run may also differ in membership - the
    reference drops "the first sorted element" of reference rows
    (nabo/_mapping.py:141-142), which for duplicate cells is an arbitrary member
    of the zero-distance run."""
    idx_a, idx_b = np.asarray(idx_a), np.asarray(idx_b)
    da, db = np.asarray(dist_a, dtype=np.float64), np.asarray(dist_b, dtype=np.float64)
    if idx_a.shape != idx_b.shape:
        return False
    same_d = (da == db) | (np.isnan(da) & np.isnan(db))
    if not same_d.all():
        return False
    for r in np.nonzero((idx_a != idx_b).any(axis=1))[0]:
        d = da[r]
        last = d[-1]
        for v in np.unique(d[~np.isnan(d)]):
            if v == last or (head_truncated and v == d[0]):
                continue  # truncated run: membership may differ
            sel = d == v
            if set(idx_a[r][sel]) != set(idx_b[r][sel]):
                return False
    return True


# ----------------------------------------------------------------------------
# SNN weights
# ----------------------------------------------------------------------------
def snn_weight_lut(k: int) -> np.ndarray:
    """nabo/_mapping.py:185, 194: weight = round(snn / (2*(k-1) - snn), 2) using
    Python's round() on a double.  Entry 0 is 0.0 (no edge, :195)."""
    factor = 2 * (k - 1)
    lut = np.zeros(k + 1, dtype=np.float64)
    for snn in range(1, k + 1):
        lut[snn] = round(snn / (factor - snn), 2)
    return lut


def fix_weight(k: int) -> float:
    """nabo/_mapping.py:478-479."""
    return 0.5 / ((2 * (k - 1)) - 0.5)


def snn_counts(target_knn: np.ndarray, ref_knn: np.ndarray) -> np.ndarray:
    """nabo/_mapping.py:186-198: for query t with A = set(target_knn[t]) and each
    j in A, snn = |A ∩ set(ref_knn[j])|.  Returns (N,k) counts aligned with
    target_knn (the reference iterates the set; alignment is ours)."""
    target_knn = np.asarray(target_knn)
    ref_knn = np.asarray(ref_knn)
    n, k = target_knn.shape
    out = np.zeros((n, k), dtype=np.uint8)
    for t in range(n):
        a = target_knn[t]
        b = ref_knn[a]                                   # (k, k_ref)
        out[t] = (b[:, :, None] == a[None, None, :]).any(axis=2).sum(axis=1)
    return out


def snn_weights(target_knn, ref_knn, k: int | None = None):
    """Counts pushed through the weight LUT; weight 0.0 means "no edge"."""
    cnt = snn_counts(target_knn, ref_knn)
    kk = target_knn.shape[1] if k is None else k
    return cnt, snn_weight_lut(kk)[cnt]


# ----------------------------------------------------------------------------
# mapping score
# ----------------------------------------------------------------------------
def mapping_scores(target_knn, weights, n_ref: int, n_targets: int | None = None,
                   min_weight: float = 0.0, min_score: float = 0.0,
                   weighted: bool = True, score_multiplier: float = 1000.0,
                   include: np.ndarray | None = None, counts=None) -> np.ndarray:
    """nabo/_graph.py:643-653, 690-693.

    score[r] = score_multiplier * sum_{t in include, edge (r,t), w > min_weight} w
               / len(include)          (weighted)
             = score_multiplier * #edges / len(include)   (unweighted)
    then zeroed below min_score.  An edge exists where snn > 0 (_mapping.py:195):
    pass ``counts`` to say so exactly; without it a non-zero weight stands for an
    edge (wrong only where round() gives 0.0 for snn > 0, k >= 102; note that k = 1
    yields the weight -1).  Summation follows target order (adjacency insertion
    order of the reference)."""
    target_knn = np.asarray(target_knn)
    weights = np.asarray(weights, dtype=np.float64)
    edge = (np.asarray(counts) > 0) if counts is not None else (weights != 0)
    n = target_knn.shape[0]
    rows = np.arange(n) if include is None else np.asarray(include)
    denom = len(rows) if n_targets is None else n_targets
    acc = np.zeros(n_ref, dtype=np.float64)
    for t in rows:
        for j, w, e in zip(target_knn[t], weights[t], edge[t]):
            if e:                           # an edge exists
                if weighted:
                    if w > min_weight:
                        acc[j] += w
                else:
                    acc[j] += 1
    score = score_multiplier * acc / denom
    return np.where(score >= min_score, score, 0.0)


def classify_targets(target_knn, weights, ref_labels, n_labels: int,
                     weight_frac: float = 0.5, min_degree: int = 2,
                     min_weight: float = 0.0, counts=None):
    """nabo/_graph.py:722-792 (classify_target), array form.  For each target
    node: degree (= number of edges, weight > 0) < min_degree -> -1 (na_label);
    edges with weight > min_weight vote their weight for the reference node's
    cluster (ref_labels[j] < 0 = not in cluster_dict), while *every* edge adds to
    the total (:768-771); the best cluster wins iff its sum > weight_frac*total,
    else -1.  The reference's tie between equal-weight clusters follows set
    iteration order (unspecified); here the lowest label wins."""
    target_knn = np.asarray(target_knn)
    weights = np.asarray(weights, dtype=np.float64)
    edge = (np.asarray(counts) > 0) if counts is not None else (weights != 0)     # as in mapping_scores
    n = target_knn.shape[0]
    out = np.full(n, -1, dtype=np.int64)
    for t in range(n):
        votes = np.zeros(n_labels, dtype=np.float64)
        deg = 0
        tot = 0.0
        for j, w, e in zip(target_knn[t], weights[t], edge[t]):
            if e:
                deg += 1
                if w > min_weight and ref_labels[j] >= 0:
                    votes[ref_labels[j]] += w
                tot += w
        if deg < min_degree:
            continue
        best = int(np.argmax(votes))
        if votes[best] > weight_frac * tot:
            out[t] = best
    return out


# ----------------------------------------------------------------------------
# scaling + PCA projection
# ----------------------------------------------------------------------------
def scale_counts(counts, sf, mu, sigma) -> np.ndarray:
    """nabo/_dataset.py:905-913 (get_scaled_values core):
    a = zeros(float32); a[idx] = val; a = a[goi]; a = ((a * sf[i]) - mu) / sigma
    `a` and `sf` are float32 (the product is rounded to float32), mu/sigma are
    float64 (:826-830), so the subtraction and division happen in float64."""
    a = np.asarray(counts, dtype=np.float32)
    sf = np.asarray(sf, dtype=np.float32)
    prod = a * sf[:, None]                               # float32 product
    return (prod.astype(np.float64) - np.asarray(mu, np.float64)[None, :]) / \
        np.asarray(sigma, np.float64)[None, :]


def pca_transform(z, components, mean) -> np.ndarray:
    """nabo/_dataset.py:1028 -> sklearn 1.9.0 _BasePCA.transform (un-vendored,
    unpinned in requirements.txt:6): X @ components_.T - mean_ @ components_.T,
    no whitening (_dataset.py:957-960)."""
    z = np.asarray(z, dtype=np.float64)
    c = np.asarray(components, dtype=np.float64)
    return z @ c.T - (np.asarray(mean, np.float64).reshape(1, -1) @ c.T)


def project(counts, sf, mu, sigma, components, mean) -> np.ndarray:
    """Dataset.transform_pca (nabo/_dataset.py:985-1033) on a dense count block."""
    return pca_transform(scale_counts(counts, sf, mu, sigma), components, mean)


# ----------------------------------------------------------------------------
# multi-shard merge (no reference counterpart: defines the expected result of
# the reference-sharded mode = the single-shard result)
# ----------------------------------------------------------------------------
def merge_topk(idx_shards, dist_shards, k: int):
    """Merge per-shard (N,k_s) candidates (global indices) by (dist, idx), NaN last."""
    idx = np.concatenate(idx_shards, axis=1)
    dst = np.concatenate(dist_shards, axis=1)
    key = np.where(np.isnan(dst), np.inf, dst)
    order = np.lexsort((idx, key), axis=1)[:, :k]
    return np.take_along_axis(idx, order, 1), np.take_along_axis(dst, order, 1)


def mapping_specificity(ref_edges_a, ref_edges_b, n_ref, tgt_knn, tgt_counts):
    """Graph.get_mapping_specificity (nabo/_graph.py:794-824), array form: for every target the mean
    unweighted shortest-path length (hops) in the undirected reference graph given by the edge list
    (a[i], b[i]) between all pairs i < j of reference cells it has an edge to (counts > 0).
    NaN with fewer than two mapped cells (np.mean of an empty list upstream); raises ValueError when
    a pair is not connected (networkx.NetworkXNoPath upstream).  Plain BFS per source, small cases only."""
    adj = [[] for _ in range(n_ref)]
    for x, y in zip(ref_edges_a, ref_edges_b):
        x, y = int(x), int(y)
        if x != y:
            adj[x].append(y)
            adj[y].append(x)
    out = np.full(len(tgt_knn), np.nan)
    for t in range(len(tgt_knn)):
        mapped = [int(r) for r, c in zip(tgt_knn[t], tgt_counts[t]) if c > 0 and r >= 0]
        if len(mapped) < 2:
            continue
        spls = []
        for i, src in enumerate(mapped):
            dist = {src: 0}
            frontier = [src]
            while frontier:
                nxt = []
                for u in frontier:
                    for v in adj[u]:
                        if v not in dist:
                            dist[v] = dist[u] + 1
                            nxt.append(v)
                frontier = nxt
            for dst in mapped[i + 1:]:
                if dst not in dist:
                    raise ValueError("no path between reference cells %d and %d" % (src, dst))
                spls.append(dist[dst])
        out[t] = float(np.mean(spls))
    return out


def size_factor_sums(indptr, idx, val, n_cols, keep_cols):
    """Dataset.set_sf (nabo/_dataset.py:573-583): float32 `temp[keepGenesIdx].sum()` of every cell's densified
    count vector (NumPy's own float32 pairwise reduction).  CSR over all cells / all genes."""
    out = np.zeros(len(indptr) - 1, dtype=np.float32)
    for i in range(len(out)):
        temp = np.zeros(n_cols, dtype=np.float32)
        temp[idx[indptr[i]:indptr[i + 1]]] = val[indptr[i]:indptr[i + 1]]
        out[i] = temp[keep_cols].sum()
    return out


def gene_stats(indptr, idx, val, n_cells, keep_cells, sf):
    """Dataset.set_gene_stats (nabo/_dataset.py:609-622) for the genes of a gene-major CSR (= CSC of the count
    matrix): temp = densified counts of the kept cells times their size factors (float32), then
    m = temp.mean(), nzm = temp[temp > 0].mean(), variance = temp.var(), ncells = (temp > 0).sum()."""
    n = len(indptr) - 1
    m, nzm, var = (np.zeros(n, np.float32) for _ in range(3))
    nc = np.zeros(n, np.int64)
    sfk = np.asarray(sf, np.float32)[keep_cells]
    for i in range(n):
        temp = np.zeros(n_cells, dtype=np.float32)
        temp[idx[indptr[i]:indptr[i + 1]]] = val[indptr[i]:indptr[i + 1]]
        temp = temp[keep_cells] * sfk
        pos = temp > 0
        nc[i] = pos.sum()
        if nc[i]:
            m[i], nzm[i], var[i] = temp.mean(), temp[pos].mean(), temp.var()
    return m, nzm, var, nc

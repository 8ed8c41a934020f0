"""BASELINE config-2 size (100 000 x 100 000 cells, 50 PCs, k = 30) through size-independent
properties: the oracle cannot run here (it would need an 80 GB distance matrix), so the full-size
result is checked against the exact engine on sampled rows, against itself under reference
sharding, and through invariants (sortedness, uniqueness, score mass)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N = M = 100_000
G, K = 50, 30


@pytest.fixture(scope="module")
def data():
    from nabo_b200 import build, core, synth
    build.build()
    ref = torch.from_numpy(synth.pc_mixture(M, G, seed=1)).cuda()
    tgt = torch.from_numpy(synth.pc_mixture(N, G, seed=101)).cuda()
    return core, ref, tgt


@pytest.mark.parametrize("metric", ["euclidean", "mod_canberra", "cosine"])
def test_full_size_knn_properties(data, metric):
    core, ref, tgt = data
    idx, dst, st = core.knn(tgt, ref, K, metric, 0.25, mode="fast", return_stats=True)
    assert st["rows_exact_fallback"] < N // 100
    d = dst.cpu().numpy()
    i = idx.cpu().numpy()
    assert np.isfinite(d).all() and (np.diff(d, axis=1) >= 0).all()            # sorted rows
    assert i.min() >= 0 and i.max() < M
    s = np.sort(i, axis=1)
    assert (np.diff(s, axis=1) > 0).all()                                      # no duplicate neighbour
    ties = (np.diff(d, axis=1) == 0)
    assert (np.diff(i, axis=1)[ties] > 0).all()                                # ties broken by index
    # sampled rows against the exact FP64 brute-force engine: bit-identical
    rows = np.random.default_rng(0).choice(N, 768, replace=False)
    ei, ed = core.knn(tgt[torch.from_numpy(rows).cuda()].contiguous(), ref, K, metric, 0.25, mode="exact")
    assert np.array_equal(ei.cpu().numpy(), i[rows]) and np.array_equal(ed.cpu().numpy(), d[rows])
    # ... and against the ORACLE (C port of the reference loops, independent of every kernel of this build) on 96
    # of them: all 100 000 distances per row, full-row sort, first k
    if metric != "cosine":                     # the reference has no cosine metric: the C port has none either
        from oracle import c_port
        c_port.build()
        sub = rows[:96]
        oi, od = c_port.knn(tgt[torch.from_numpy(sub).cuda()].cpu().numpy(), ref.cpu().numpy(), K, metric, 0.25,
                            nthreads=4)
        assert np.array_equal(oi, i[sub]) and np.array_equal(od, d[sub])


def test_full_size_reference_sharding_and_scores(data):
    core, ref, tgt = data
    full_i, full_d = core.knn(tgt, ref, K, "euclidean", mode="fast")
    bounds = [0, 25_000, 50_001, 74_999, M]
    parts = [core.knn(tgt, ref[a:b].contiguous(), K, "euclidean", idx_offset=a, mode="fast")
             for a, b in zip(bounds, bounds[1:])]
    mi, md = core.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, full_i) and torch.equal(md, full_d)                 # sharded == unsharded, bit for bit
    # self-kNN of the reference (config-3 shape of the call), then weights and scores
    rk, rd = core.knn(ref, ref, K, "euclidean", drop_first=True, mode="fast")
    assert (rk != torch.arange(M, device="cuda", dtype=torch.int32)[:, None]).all()   # self dropped (no duplicates here)
    cnt, w = core.snn_weights(full_i, rk, K)
    sc = core.mapping_scores(full_i, cnt, M, K)
    np.testing.assert_allclose(sc.sum().item(), 1000.0 * w.sum().item() / N, rtol=1e-11)
    assert torch.equal(sc, core.mapping_scores(full_i, cnt, M, K))             # deterministic
    lut = torch.from_numpy(core.snn_weight_lut(K)).cuda()
    assert torch.equal(w, lut[cnt.long()])
    assert int(cnt.max()) <= K and (w[cnt == 0] == 0).all()
    # a reference cell's own row: |A ∩ A| = K -> snn count of every neighbour >= 1 when mapped onto itself
    c2, _ = core.snn_weights(rk, rk, K)
    assert c2.shape == (M, K)


def test_config3_self_knn_one_million_cells():
    """BASELINE config 3: reference self-kNN for 1 M cells, 50 PCs, k = 15 (the `make_ref_graph` call).
    Sampled rows must equal the exact FP64 engine bit for bit; every row drops itself."""
    from nabo_b200 import core, synth
    m, k = 1_000_000, 15
    ref = torch.from_numpy(synth.pc_mixture(m, G, seed=3, n_clusters=64)).cuda()
    idx, dst, st = core.knn(ref, ref, k, "euclidean", drop_first=True, mode="fast", return_stats=True)
    assert st["rows_exact_fallback"] < m // 1000
    assert (idx != torch.arange(m, device="cuda", dtype=torch.int32)[:, None]).all()
    assert (dst[:, 1:] >= dst[:, :-1]).all() and (dst > 0).all()
    rows = torch.from_numpy(np.random.default_rng(5).choice(m, 256, replace=False)).cuda()
    # exact engine on the sampled rows: k+1 neighbours without the self drop, then drop the first (= self, d = 0)
    ei, ed = core.knn(ref[rows].contiguous(), ref, k + 1, "euclidean", mode="exact")
    assert torch.equal(ei[:, 0].long(), rows) and (ed[:, 0] == 0).all()
    assert torch.equal(ei[:, 1:], idx[rows]) and torch.equal(ed[:, 1:], dst[rows])
    # symmetric-ness sanity of the SNN step on the same table
    cnt, w = core.snn_weights(idx[:100_000].contiguous(), idx, k)
    assert int(cnt.max()) <= k and float(w.max()) <= core.snn_weight_lut(k).max()


def test_config5_cosine_projection_fused_path():
    """BASELINE config 5 in miniature: raw counts -> projection into the reference PCA space -> cosine
    kNN -> SNN weights -> mapping scores through the one-call API, against the oracle."""
    from nabo_b200 import core, synth
    from oracle import nabo_oracle as O
    rng = np.random.default_rng(0)
    n_genes, G_, nc, k = 900, 400, 50, 12
    counts_r = synth.nb_counts(3000, n_genes, seed=1)
    counts_t = synth.nb_counts(1500, n_genes, seed=101)
    gi = np.sort(rng.choice(n_genes, G_, replace=False)).astype(np.int32)
    sf_r, sf_t = synth.size_factors(counts_r), synth.size_factors(counts_t)
    x = counts_r[:, gi].astype(np.float32) * sf_r[:, None]
    mu, sigma = x.mean(0).astype(np.float64), x.std(0).astype(np.float64) + 1e-3
    z = O.scale_counts(counts_r[:, gi], sf_r, mu, sigma)
    mean = z.mean(0)
    _, _, vt = np.linalg.svd(z - mean, full_matrices=False)
    comps = vt[:nc]
    model = lambda sf: dict(gene_idx=gi, sf=sf, mu=mu, sigma=sigma, components=comps, mean=mean)
    ref_pca = core.project(counts_r.astype(np.float32), **model(sf_r))
    np.testing.assert_allclose(ref_pca, O.project(counts_r[:, gi], sf_r, mu, sigma, comps, mean), rtol=0,
                               atol=1e-11 * np.abs(ref_pca).max())
    ref_knn, _ = core.knn(ref_pca, ref_pca, k, "cosine", drop_first=True)
    res = core.map_cells(counts_t.astype(np.float32), ref_pca, ref_knn, k, metric="cosine", pca_model=model(sf_t))
    oi, od = O.knn(res["pca"], ref_pca, k, "cosine")
    assert np.array_equal(res["idx"], oi) and np.array_equal(res["dist"], od)
    ork, _ = O.knn(ref_pca, ref_pca, k, "cosine", drop_first=True)
    assert np.array_equal(ref_knn, ork)
    cnt, w = O.snn_weights(oi, ork, k)
    assert np.array_equal(res["weights"], w)
    np.testing.assert_allclose(res["scores"], O.mapping_scores(oi, w, len(ref_pca)), rtol=1e-12)

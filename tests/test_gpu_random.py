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

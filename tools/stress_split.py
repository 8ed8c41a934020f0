"""Randomised fast == exact sweep at shapes that trigger the reference-split paths (few queries, >= 8 k references)."""
import sys, numpy as np
sys.path.insert(0, ".")
from nabo_b200 import build, core
build.build()
bad = 0
for seed in range(90):
    rng = np.random.default_rng(9000 + seed)
    g = int(rng.choice([3, 8, 17, 25, 40, 50, 56, 64, 90]))
    m = int(rng.choice([8200, 9000, 20000, 33333, 70000]))
    n = int(rng.choice([1, 5, 40, 333, 1500, 4000, 9000, 15000]))
    k = int(rng.choice([1, 5, 15, 30, 50, 80]))
    metric = str(rng.choice(["euclidean", "mod_canberra", "cosine"]))
    f = float(rng.choice([0.1, 0.25, 0.7]))
    centers = rng.normal(size=(6, g)) * 3
    r = centers[rng.integers(0, 6, m)] + rng.normal(size=(m, g))
    q = centers[rng.integers(0, 6, n)] + rng.normal(size=(n, g))
    if rng.random() < 0.5:
        r[rng.integers(0, m, 200)] = r[rng.integers(0, m)]
        q[rng.integers(0, n, max(1, n // 20))] = r[rng.integers(0, m)]
    mask = (rng.random(m) < 0.2) if rng.random() < 0.4 else None
    drop = bool(rng.random() < 0.3)
    kw = dict(ref_mask=mask, drop_first=drop, idx_offset=int(rng.choice([0, 77])))
    fi, fd, st = core.knn(q, r, k, metric, f, mode="fast", return_stats=True, **kw)
    ei, ed = core.knn(q, r, k, metric, f, mode="exact", **kw)
    ok = np.array_equal(fi, ei) and bool(((fd == ed) | (np.isnan(fd) & np.isnan(ed))).all())
    if not ok:
        bad += 1
        print("FAIL seed", seed, dict(g=g, m=m, n=n, k=k, metric=metric, f=f, mask=mask is not None, drop=drop))
print("done, failures:", bad)

import sys, numpy as np, traceback
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_random as T
from nabo_b200 import build, core
build.build()
bad = 0
for seed in range(160, 1400):
    try:
        T.test_random_knn_case(core, seed)
    except Exception as e:
        bad += 1
        c = T._case(seed)
        print("FAIL seed", seed, type(e).__name__, str(e)[:200], {k: (v.shape if hasattr(v, "shape") else v) for k, v in c.items()})
for seed in range(30, 300):
    try:
        T.test_random_self_knn_case(core, seed)
    except Exception as e:
        bad += 1
        print("FAIL self seed", seed, type(e).__name__, str(e)[:200])
for seed in range(24, 200):
    try:
        T.test_random_graph_ops(core, seed)
    except Exception as e:
        bad += 1
        print("FAIL graph seed", seed, type(e).__name__, str(e)[:300])
print("done, failures:", bad)

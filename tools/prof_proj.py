"""Three launches of the FP64 tensor-core projection at the config-5 slice shape, for ncu (development probe)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
r = bench.projection_block(torch.device("cuda", 0), steps=1, warmup=1, shapes=(("config5_slice", 262144, 2000, 50),))
print({k: round(v["mma"]["ms"], 3) for k, v in r.items()})

"""Sharded modes with the CUDA engine: world_size 1 always; 2 ranks over NCCL when 2 GPUs exist."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import nabo_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_cuda_engine_matches_oracle():
    from nabo_b200 import build, parallel as P, synth
    build.build()
    ref = synth.pc_mixture(2000, 20, seed=1)
    tgt = synth.pc_mixture(900, 20, seed=101)
    k = 12
    rt, tt = torch.from_numpy(ref).cuda(), torch.from_numpy(tgt).cuda()
    lo, hi, ri, rd = P.knn_reference_sharded(rt, rt, 0, k, "euclidean", drop_first=True)
    oi, od = O.knn(ref, ref, k, "euclidean", drop_first=True)
    assert (lo, hi) == (0, 2000) and np.array_equal(ri.cpu().numpy(), oi) and np.array_equal(rd.cpu().numpy(), od)
    a = P.map_targets_sharded(tt, rt, ri, k, len(tgt))
    b = P.map_reference_sharded(tt, rt, 0, len(ref), ri, k)
    ti, td = O.knn(tgt, ref, k, "mod_canberra", 0.25)
    _, w = O.snn_weights(ti, oi, k)
    sc = O.mapping_scores(ti, w, len(ref))
    for r in (a, b):
        assert np.array_equal(r["idx"].cpu().numpy(), ti) and np.array_equal(r["dist"].cpu().numpy(), td)
        assert np.array_equal(r["weights"].cpu().numpy(), w)
        np.testing.assert_allclose(r["scores"].cpu().numpy(), sc, rtol=1e-12)


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from nabo_b200 import parallel as P, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ref = synth.pc_mixture(5001, 30, seed=1); tgt = synth.pc_mixture(2503, 30, seed=101)
ref[40] = ref[4000]; tgt[5] = ref[40]
k = 15
rt, tt = torch.from_numpy(ref).cuda(), torch.from_numpy(tgt).cuda()
lo, hi = P.shard_bounds(len(ref), world, rank)
qlo, qhi, ri, rd = P.knn_reference_sharded(rt, rt[lo:hi].contiguous(), lo, k, "euclidean", drop_first=True, merge_slice=False)
tlo, thi = P.shard_bounds(len(tgt), world, rank)
a = P.map_targets_sharded(tt[tlo:thi].contiguous(), rt, ri, k, len(tgt), metric="euclidean")
b = P.map_reference_sharded(tt, rt[lo:hi].contiguous(), lo, len(ref), ri, k, metric="euclidean")
c = P.map_reference_sharded(tt, rt[lo:hi].contiguous(), lo, len(ref), ri, k, metric="mod_canberra", dist_factor=0.25)
np.savez(os.path.join(sys.argv[2], "r%d.npz" % rank), ref_knn=ri.cpu().numpy(), ref_dst=rd.cpu().numpy(),
         a_idx=a["idx"].cpu().numpy(), a_dist=a["dist"].cpu().numpy(), a_sc=a["scores"].cpu().numpy(),
         b_lo=b["lo"], b_idx=b["idx"].cpu().numpy(), b_dist=b["dist"].cpu().numpy(), b_sc=b["scores"].cpu().numpy(),
         c_idx=c["idx"].cpu().numpy(), c_dist=c["dist"].cpu().numpy(), c_sc=c["scores"].cpu().numpy())
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_nccl(tmp_path):
    from nabo_b200 import core, synth
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT, str(tmp_path)],
                   check=True, timeout=600)
    ref = synth.pc_mixture(5001, 30, seed=1)
    tgt = synth.pc_mixture(2503, 30, seed=101)
    ref[40] = ref[4000]
    tgt[5] = ref[40]
    k = 15
    ri, rd = core.knn(ref, ref, k, "euclidean", drop_first=True, mode="exact")
    ti, td = core.knn(tgt, ref, k, "euclidean", mode="exact")
    cnt, _ = core.snn_weights(ti, ri, k)
    sc = core.mapping_scores(ti, cnt, len(ref), k)
    res = [np.load(str(tmp_path / ("r%d.npz" % r))) for r in range(2)]
    for r in res:
        assert np.array_equal(r["ref_knn"], ri) and np.array_equal(r["ref_dst"], rd)
        np.testing.assert_allclose(r["a_sc"], sc, rtol=1e-12)
        np.testing.assert_allclose(r["b_sc"], sc, rtol=1e-12)
    assert np.array_equal(np.concatenate([r["a_idx"] for r in res]), ti)
    assert np.array_equal(np.concatenate([r["b_idx"] for r in res]), ti)
    assert np.array_equal(np.concatenate([r["b_dist"] for r in res]), td)
    # modified Canberra (bit-sliced candidate pass per shard), reference-sharded == unsharded exact engine
    ci, cd = core.knn(tgt, ref, k, "mod_canberra", 0.25, mode="exact")
    ccnt, _ = core.snn_weights(ci, ri, k)
    csc = core.mapping_scores(ci, ccnt, len(ref), k)
    assert np.array_equal(np.concatenate([r["c_idx"] for r in res]), ci)
    assert np.array_equal(np.concatenate([r["c_dist"] for r in res]), cd)
    for r in res:
        np.testing.assert_allclose(r["c_sc"], csc, rtol=1e-12)

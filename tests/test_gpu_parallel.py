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
    assert torch.equal(a["score_acc"], b["score_acc"]) and torch.equal(a["scores"], b["scores"])


def test_routed_rows_merge_parts_and_integer_scores():
    """nabo_knn_routed delivers row t to its part; nabo_merge_topk_parts merges blocks at arbitrary addresses
    (with the global drop_first); nabo_score_accumulate / nabo_scores_finalize equal the oracle's score."""
    from nabo_b200 import build, core, synth
    build.build()
    ref = synth.pc_mixture(3001, 24, seed=1)
    tgt = synth.pc_mixture(1203, 24, seed=101)
    ref[17] = ref[2500]
    tgt[3] = ref[17]
    k = 13
    rt, tt = torch.from_numpy(ref).cuda(), torch.from_numpy(tgt).cuda()
    for metric in ("euclidean", "mod_canberra"):
        for mode in ("fast", "exact"):
            pi, pd = core.knn(tt, rt, k, metric, 0.25, mode=mode, idx_offset=7)
            bounds = [0, 100, 100, 777, 1203]                       # an empty part in the middle
            bi = [torch.full((bounds[p + 1] - bounds[p], k), -9, dtype=torch.int32, device="cuda") for p in range(4)]
            bd = [torch.full((bounds[p + 1] - bounds[p], k), -9.0, dtype=torch.float64, device="cuda") for p in range(4)]
            core.knn(tt, rt, k, metric, 0.25, mode=mode, idx_offset=7,
                     out_parts=(bounds, [b.data_ptr() for b in bi], [b.data_ptr() for b in bd]))
            assert torch.equal(torch.cat(bi), pi) and torch.equal(torch.cat(bd).view(torch.int64), pd.view(torch.int64))
    # reference split in 3 shards, self-kNN with the global "drop first": merge of blocks in place
    cuts = [0, 900, 2100, 3001]
    li, ld = [], []
    for s in range(3):
        i, d = core.knn(rt, rt[cuts[s]:cuts[s + 1]].contiguous(), k + 1, "euclidean", idx_offset=cuts[s])
        li.append(i)
        ld.append(d)
    mi, md = core.merge_topk_parts([x.data_ptr() for x in li], [x.data_ptr() for x in ld], len(ref), k + 1, "cuda",
                                   drop_first=True)
    oi, od = O.knn(ref, ref, k, "euclidean", drop_first=True)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(md.cpu().numpy(), od)
    # integer score accumulation, in two batches, against the oracle's sequential FP64 sum
    ti, _ = O.knn(tgt, ref, k, "mod_canberra", 0.25)
    cnt, w = O.snn_weights(ti, oi, k)
    tid, cd = torch.from_numpy(ti.astype(np.int32)).cuda(), torch.from_numpy(cnt).cuda()
    acc = core.score_accumulate(tid[:500], cd[:500], len(ref), k)
    acc = core.score_accumulate(tid[500:], cd[500:], len(ref), k, acc=acc)
    whole = core.score_accumulate(tid, cd, len(ref), k)
    assert torch.equal(acc, whole)
    sc = core.scores_finalize(acc, len(tgt)).cpu().numpy()
    np.testing.assert_allclose(sc, O.mapping_scores(ti, w, len(ref)), rtol=1e-12)
    assert np.array_equal(sc, 1000.0 * (acc.cpu().numpy() / 100.0) / len(tgt))
    np.testing.assert_allclose(sc, core.mapping_scores(tid, cd, len(ref), k).cpu().numpy(), rtol=1e-12)
    # unweighted / min_weight variants agree with the sorted reduction
    for kw in (dict(weighted=False), dict(min_weight=0.3)):
        a1 = core.scores_finalize(core.score_accumulate(tid, cd, len(ref), k, **kw), len(tgt),
                                  weighted=kw.get("weighted", True)).cpu().numpy()
        a2 = core.mapping_scores(tid, cd, len(ref), k, **kw).cpu().numpy()
        np.testing.assert_allclose(a1, a2, rtol=1e-12)


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
assert P._exchange_for(len(tgt), k, tt.device, "auto").transport == os.environ.get("NABO_EXCHANGE", "p2p")
np.savez(os.path.join(sys.argv[2], "r%d.npz" % rank), ref_knn=ri.cpu().numpy(), ref_dst=rd.cpu().numpy(),
         a_idx=a["idx"].cpu().numpy(), a_dist=a["dist"].cpu().numpy(), a_sc=a["scores"].cpu().numpy(),
         b_lo=b["lo"], b_idx=b["idx"].cpu().numpy(), b_dist=b["dist"].cpu().numpy(), b_sc=b["scores"].cpu().numpy(),
         c_idx=c["idx"].cpu().numpy(), c_dist=c["dist"].cpu().numpy(), c_sc=c["scores"].cpu().numpy(),
         a_acc=a["score_acc"].cpu().numpy(), b_acc=b["score_acc"].cpu().numpy())
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["p2p", "a2a"])
def test_two_ranks_nccl(tmp_path, transport):
    """p2p: rerank_kernel stores result rows into the peer GPU's receive buffer over NVLink (symmetric memory);
    a2a: into a local send buffer moved by one NCCL all_to_all_single.  Both must equal the unsharded result."""
    from nabo_b200 import core, synth
    os.environ["NABO_EXCHANGE"] = transport
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    try:
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT, str(tmp_path)],
                       check=True, timeout=600)
    finally:
        os.environ.pop("NABO_EXCHANGE", None)
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
    # integer weight sums: the 2-rank scores have the bits of the 1-GPU integer path, in both sharded modes
    acc1 = core.score_accumulate(ti, cnt, len(ref), k)
    sc1 = core.scores_finalize(acc1, len(tgt))
    acc1, sc1 = acc1.cpu().numpy(), np.asarray(sc1.cpu() if hasattr(sc1, "cpu") else sc1)
    for r in res:
        assert np.array_equal(r["a_acc"], acc1) and np.array_equal(r["b_acc"], acc1)
        assert np.array_equal(r["a_sc"], sc1) and np.array_equal(r["b_sc"], sc1)
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

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, global index offsets,
all-gather layout, per-rank merge slices, score all-reduce.  The compute engine is the oracle here
(test infrastructure); on GPUs the same code runs with the CUDA engine (tests/test_gpu_parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nabo_oracle as O


class OracleEngine:
    routed = False         # no device addresses on CPU: parallel.py copies the rows into the send blocks

    def knn(self, q, r, k, metric, f, ref_mask, drop_first, idx_offset, mode, out_parts=None):
        assert out_parts is None
        m = None if ref_mask is None else np.asarray(ref_mask)
        i, d = O.knn(q.numpy(), r.numpy(), k, metric, f, mask=m, drop_first=drop_first)
        return torch.from_numpy((i + idx_offset).astype(np.int32)), torch.from_numpy(d)

    def merge_topk(self, idx, dst):
        i, d = O.merge_topk(list(idx.numpy().astype(np.int64)), list(dst.numpy()), idx.shape[2])
        i = np.where(np.isnan(d) & (i < 0), -1, i)
        return torch.from_numpy(i.astype(np.int32)), torch.from_numpy(d)

    def merge_parts(self, idx_blocks, dist_blocks, n_rows, k, drop_first):
        i, d = self.merge_topk(torch.stack(idx_blocks), torch.stack(dist_blocks))
        return (i[:, 1:].contiguous(), d[:, 1:].contiguous()) if drop_first else (i, d)

    def score_accumulate(self, tgt_knn, counts, n_ref, k):
        # integer weight sums per reference cell (hundredths; every SNN weight is round(w, 2))
        iw = np.rint(O.snn_weight_lut(k) * 100).astype(np.int64)
        acc = np.zeros(n_ref, dtype=np.int64)
        t, c = tgt_knn.numpy().astype(np.int64), counts.numpy().astype(np.int64)
        ok = (c > 0) & (t >= 0)
        np.add.at(acc, t[ok], iw[c[ok]])
        return torch.from_numpy(acc)

    def scores_finalize(self, acc, n_total):
        return torch.from_numpy(1000.0 * (acc.numpy().astype(np.float64) / 100.0) / n_total)

    def snn_weights(self, tgt_knn, ref_knn, k):
        c, w = O.snn_weights(tgt_knn.numpy().astype(np.int64), ref_knn.numpy().astype(np.int64), k)
        return torch.from_numpy(c), torch.from_numpy(w)



def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    from nabo_b200 import synth
    ref = synth.pc_mixture(301, 12, seed=1, n_clusters=5)
    tgt = synth.pc_mixture(157, 12, seed=101, n_clusters=5)
    ref[40] = ref[200]                      # duplicates across shards: merge must break ties by global index
    tgt[5] = ref[40]
    return ref, tgt


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nabo_b200 import parallel as P
    eng = OracleEngine()
    ref, tgt = _data()
    k = 9
    rt, tt = torch.from_numpy(ref), torch.from_numpy(tgt)
    # reference self-kNN, reference-sharded, with the global "drop first"
    lo, hi = P.shard_bounds(len(ref), world, rank)
    qlo, qhi, ri, rd = P.knn_reference_sharded(rt, rt[lo:hi], lo, k, "euclidean", drop_first=True, engine=eng)
    gathered = [None] * world
    dist.all_gather_object(gathered, (qlo, qhi, ri.numpy(), rd.numpy()))
    ref_knn = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
    ref_dst = np.concatenate([g[3] for g in sorted(gathered, key=lambda g: g[0])])
    # the all-gathered form returns the same table on every rank
    alo, ahi, ai, ad = P.knn_reference_sharded(rt, rt[lo:hi], lo, k, "euclidean", drop_first=True, engine=eng,
                                               merge_slice=False)
    assert (alo, ahi) == (0, len(ref)) and np.array_equal(ai.numpy(), ref_knn) and np.array_equal(ad.numpy(), ref_dst)
    rk = torch.from_numpy(ref_knn)
    # (a) target-sharded
    tlo, thi = P.shard_bounds(len(tgt), world, rank)
    a = P.map_targets_sharded(tt[tlo:thi], rt, rk, k, len(tgt), metric="mod_canberra", engine=eng)
    # (b) reference-sharded
    b = P.map_reference_sharded(tt, rt[lo:hi], lo, len(ref), rk, k, metric="mod_canberra", engine=eng)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), ref_knn=ref_knn, ref_dst=ref_dst, tlo=tlo, thi=thi,
             a_idx=a["idx"].numpy(), a_dist=a["dist"].numpy(), a_w=a["weights"].numpy(), a_sc=a["scores"].numpy(),
             b_lo=b["lo"], b_hi=b["hi"], b_idx=b["idx"].numpy(), b_dist=b["dist"].numpy(), b_w=b["weights"].numpy(),
             b_sc=b["scores"].numpy(), a_acc=a["score_acc"].numpy(), b_acc=b["score_acc"].numpy())
    dist.destroy_process_group()


def test_sharded_modes_equal_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref, tgt = _data()
    k = 9
    ri, rd = O.knn(ref, ref, k, "euclidean", drop_first=True)
    ti, td = O.knn(tgt, ref, k, "mod_canberra", 0.25)
    cnt, w = O.snn_weights(ti, ri, k)
    sc = O.mapping_scores(ti, w, len(ref))
    res = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    for r in res:
        assert np.array_equal(r["ref_knn"], ri) and np.array_equal(r["ref_dst"], rd)
        np.testing.assert_allclose(r["a_sc"], sc, rtol=1e-13)
        np.testing.assert_allclose(r["b_sc"], sc, rtol=1e-13)
        # integer weight sums: identical in both modes and on both ranks, hence identical score bits
        assert np.array_equal(r["a_acc"], res[0]["b_acc"]) and np.array_equal(r["a_sc"], res[0]["b_sc"])
    assert np.array_equal(np.concatenate([r["a_idx"] for r in res]), ti)
    assert np.array_equal(np.concatenate([r["a_dist"] for r in res]), td)
    assert np.array_equal(np.concatenate([r["a_w"] for r in res]), w)
    assert [(int(r["b_lo"]), int(r["b_hi"])) for r in res] == [(0, 79), (79, 157)]
    assert np.array_equal(np.concatenate([r["b_idx"] for r in res]), ti)
    assert np.array_equal(np.concatenate([r["b_dist"] for r in res]), td)
    assert np.array_equal(np.concatenate([r["b_w"] for r in res]), w)


def test_shard_bounds():
    from nabo_b200.parallel import shard_bounds
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1

"""Multi-GPU execution of the hot path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch) for the plumbing.  The reference has no parallelism at all;
the path shards because every target cell is an independent query (nabo/_mapping.py:98-130,
186-198).  Two layouts:

* **target-sharded** - reference, its self-kNN table and the PCA model are replicated, each rank
  maps a contiguous block of targets.  No data-path collective: neighbour lists and weights stay
  on the rank that made them; only the per-reference scores (a sum over targets) need one
  all-reduce of M doubles.
* **reference-sharded** - the reference rows are split, every rank sees all targets, produces a
  local top-k with GLOBAL indices (``idx_offset``), candidates are exchanged with one
  all-gather of N*k*(4+8) bytes per rank and merged by (distance, index) - bit-identical to the
  single-GPU result.  Each rank merges only its own slice of the targets.

The compute engine is a parameter so that the host-side logic (bounds, offsets, gather layout,
merge slices) is testable on CPU with ``gloo``; the default engine is the CUDA library and
raises without a GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "CudaEngine", "map_targets_sharded", "knn_reference_sharded",
           "map_reference_sharded"]


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block [lo, hi) of `n` rows for `rank` (name-sorted order is preserved)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CudaEngine:
    """Default engine: thin pass-through to nabo_b200.core on the current CUDA device."""

    def knn(self, q, r, k, metric, dist_factor, ref_mask, drop_first, idx_offset, mode):
        from . import core
        return core.knn(q, r, k, metric, dist_factor, ref_mask=ref_mask, drop_first=drop_first,
                        idx_offset=idx_offset, mode=mode)

    def merge_topk(self, idx, dst):
        from . import core
        return core.merge_topk(idx, dst)

    def snn_weights(self, tgt_knn, ref_knn, k):
        from . import core
        return core.snn_weights(tgt_knn, ref_knn, k)

    def mapping_scores(self, tgt_knn, counts, n_ref, k, n_total):
        from . import core
        return core.mapping_scores(tgt_knn, counts, n_ref, k, n_targets_total=n_total)


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def map_targets_sharded(target_shard, ref, ref_knn, k: int, n_targets_total: int, metric: Optional[str] = None,
                        dist_factor: float = 0.25, ref_mask=None, mode: str = "fast", engine=None,
                        scores: bool = True) -> Dict[str, torch.Tensor]:
    """Target-sharded mapping of this rank's block of target cells.

    Returns this rank's idx / dist / counts / weights and (if ``scores``) the GLOBAL per-reference
    mapping scores (identical on every rank after the all-reduce)."""
    engine = engine or CudaEngine()
    rank, world = _world()
    metric = metric or "mod_canberra"
    idx, dst = engine.knn(target_shard, ref, k, metric, dist_factor, ref_mask, False, 0, mode)
    cnt, w = engine.snn_weights(idx, ref_knn, k)
    out = {"idx": idx, "dist": dst, "counts": cnt, "weights": w}
    if scores:
        part = engine.mapping_scores(idx, cnt, ref.shape[0], k, n_targets_total)
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        out["scores"] = part
    return out


def knn_reference_sharded(q, ref_shard, ref_offset: int, k: int, metric: str, dist_factor: float = 0.25,
                          ref_mask_shard=None, drop_first: bool = False, mode: str = "fast", engine=None,
                          merge_slice: bool = True):
    """kNN of ALL queries against a row-sharded reference.

    Each rank computes a local top-(k [+1]) with global indices, the candidates are all-gathered
    (shard-major, the layout nabo_merge_topk expects) and merged.  With ``merge_slice`` every rank
    merges only its own block of queries and returns (lo, hi, idx, dist) for that block; otherwise
    all queries are merged on every rank."""
    engine = engine or CudaEngine()
    rank, world = _world()
    kk = k + (1 if drop_first else 0)                  # the dropped "first" element is global, decide after merge
    n = q.shape[0]
    kk_local = min(kk, ref_shard.shape[0])
    idx, dst = engine.knn(q, ref_shard, kk_local, metric, dist_factor, ref_mask_shard, False, ref_offset, mode)
    if kk_local < kk:                                  # tiny shard: pad with "missing"
        pad_i = torch.full((n, kk - kk_local), -1, dtype=idx.dtype, device=idx.device)
        pad_d = torch.full((n, kk - kk_local), float("nan"), dtype=dst.dtype, device=dst.device)
        idx, dst = torch.cat([idx, pad_i], 1), torch.cat([dst, pad_d], 1)
    if world > 1:
        # output = ranks concatenated along dim 0 (the form both NCCL and gloo accept) = shard-major
        gi = torch.empty((world * n, kk), dtype=idx.dtype, device=idx.device)
        gd = torch.empty((world * n, kk), dtype=dst.dtype, device=dst.device)
        dist.all_gather_into_tensor(gi, idx.contiguous())
        dist.all_gather_into_tensor(gd, dst.contiguous())
        gi, gd = gi.view(world, n, kk), gd.view(world, n, kk)
    else:
        gi, gd = idx.unsqueeze(0), dst.unsqueeze(0)
    lo, hi = shard_bounds(n, world, rank) if merge_slice else (0, n)
    mi, md = engine.merge_topk(gi[:, lo:hi].contiguous(), gd[:, lo:hi].contiguous())
    if drop_first:
        mi, md = mi[:, 1:].contiguous(), md[:, 1:].contiguous()
    return lo, hi, mi, md


def map_reference_sharded(targets, ref_shard, ref_offset: int, n_ref_total: int, ref_knn, k: int,
                          metric: Optional[str] = None, dist_factor: float = 0.25, ref_mask_shard=None,
                          mode: str = "fast", engine=None, scores: bool = True) -> Dict[str, object]:
    """Reference-sharded mapping (BASELINE config 4): all targets against this rank's reference rows,
    NCCL candidate merge, then SNN weights with the replicated reference kNN table (global indices)
    and an all-reduce of the per-reference scores.  Returns results for this rank's target block."""
    engine = engine or CudaEngine()
    rank, world = _world()
    metric = metric or "mod_canberra"
    lo, hi, idx, dst = knn_reference_sharded(targets, ref_shard, ref_offset, k, metric, dist_factor, ref_mask_shard,
                                             False, mode, engine, merge_slice=True)
    cnt, w = engine.snn_weights(idx, ref_knn, k)
    out = {"lo": lo, "hi": hi, "idx": idx, "dist": dst, "counts": cnt, "weights": w}
    if scores:
        part = engine.mapping_scores(idx, cnt, n_ref_total, k, targets.shape[0])
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        out["scores"] = part
    return out

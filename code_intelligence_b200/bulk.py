"""Multi-GPU bulk per-repo encode (BASELINE.json config 4).

Issues are independent (the encoder state is reset per call, inference.py:56), so the path shards with no data-path
collective: global argsort by length (py/code_intelligence/inference.py:192-194), sorted position j goes to rank
``j mod G`` (every rank sees the same length distribution), each rank encodes its shard with its own weight replica,
and exactly ONE exchange step follows -- an all-gather of the (ceil(N/G), 2400) float32 outputs (NCCL over NVLink on
GPUs; gloo in the CPU tests) -- before the inverse permutation restores input order (inference.py:226).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np


def encode_sorted_batches(docs: List[np.ndarray], encode_padded: Callable, pad_idx: int, out_dim: int, bs: int = 100,
                          max_bs: int = 256, min_batches_rule: bool = True, coalesce: bool = False) -> np.ndarray:
    """The single-device bulk loop of ``df_to_embedding`` from the numericalised docs on
    (py/code_intelligence/inference.py:171-229): ``bs = min(bs, N//20 + 1)`` (at least 20 batches so that length
    sorting pays), argsort by length, batches of ``bs`` consecutive sorted docs right-padded to the batch's own max
    with ``pad_idx`` (pad_sequence, :207), ``encode_padded(ids[B,T], lengths[B]) -> (B, out_dim)``, on
    ``RuntimeError`` (CUDA OOM, :214-223) halve ``bs`` and retry the same position -- re-raised as ``Exception`` at
    bs == 1 -- and finally unsort with argsort(argsort) (:226).

    ``coalesce=True``: consecutive sorted batches are merged into device calls of ``max_bs`` rows.  On the B200 path a
    row's result does not depend on its batch mates or on the padded length (bit-exact, tests/test_gpu_parity.py), so
    ``bs`` -- a memory knob of the reference, default 100 -- only decides how many rows ride one launch; merging keeps
    the results and lets a caller with the reference's default arguments reach the 768-row kernels."""
    n = len(docs)
    if n == 0:
        return np.empty((0, out_dim), dtype=np.float32)
    if min_batches_rule:
        bs = min(bs, (n // 20) + 1)
    bs = max(1, min(bs, max_bs))
    if coalesce:
        bs = max_bs
    length_arr = np.array([len(d) for d in docs])
    if (length_arr < 1).any():
        raise ValueError("empty token sequence")
    len_mask = length_arr.argsort(kind="stable")
    len_mask_reversed = len_mask.argsort()
    ordered_lengths = length_arr[len_mask]
    pooled = np.empty((n, out_dim), dtype=np.float32)
    i = 0
    while i < n:
        try:
            idx = len_mask[i:i + bs]
            T = int(ordered_lengths[i + len(idx) - 1])
            bp = np.full((len(idx), T), pad_idx, dtype=np.int64)
            for r, j in enumerate(idx):
                bp[r, :length_arr[j]] = docs[j]
            pooled[i:i + len(idx)] = encode_padded(bp, ordered_lengths[i:i + len(idx)].astype(np.int32))
            i += bs
        except RuntimeError as e:
            if bs == 1:
                raise Exception(e)
            bs = max(1, min(bs, n - i) // 2)     # halve what was actually attempted (the tail may be shorter than bs)
    assert pooled.shape[0] == length_arr.shape[0]
    return pooled[len_mask_reversed, :]


def shard_plan(lengths: np.ndarray, world: int):
    """-> (order, shards): order = stable argsort by length; shards[r] = input indices of rank r (sorted order)."""
    order = np.asarray(lengths).argsort(kind="stable")
    return order, [order[r::world] for r in range(world)]


def gather_rows(local, n_total: int, world: int, rank: int, group=None):
    """local: torch tensor (n_r, D) of this rank's rows in sorted-position order r, r+G, r+2G, ...
    Returns (n_total, D) in sorted order on every rank.  One all_gather."""
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    D = local.shape[1]
    buf = torch.zeros((per, D), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    if world == 1:
        allr = buf[None]
    else:
        allr = torch.empty((world, per, D), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(allr.view(world * per, D), buf, group=group)
    # sorted position j = k*world + r  <-  allr[r, k]
    return allr.transpose(0, 1).reshape(per * world, D)[:n_total]


def encode_bulk_distributed(docs: List[np.ndarray], encode_local: Callable[[List[np.ndarray]], "np.ndarray"],
                            device: Optional[str] = None, group=None) -> np.ndarray:
    """Every rank passes the same ``docs`` (the reference's per-repo list) and gets the full (N, D) float32 array in
    input order.  ``encode_local(list_of_id_arrays) -> (n, D)`` is the per-rank encoder, e.g.
    ``IssueEncoder.encode_id_list`` with ``min_batches_rule=False``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(docs)
    lengths = np.array([len(d) for d in docs])
    order, shards = shard_plan(lengths, world)
    mine = [docs[i] for i in shards[rank]]
    local = encode_local(mine) if len(mine) else None
    if local is None:
        local = np.zeros((0, 1), dtype=np.float32)
    lt = torch.as_tensor(np.ascontiguousarray(local, dtype=np.float32))
    # every rank must agree on D even when its shard is empty
    d_t = torch.tensor([lt.shape[1] if lt.shape[0] else 0], dtype=torch.int64)
    if device is not None:
        lt, d_t = lt.to(device), d_t.to(device)
    if world > 1:
        dist.all_reduce(d_t, op=dist.ReduceOp.MAX, group=group)
    D = int(d_t.item())
    if lt.shape[0] == 0:
        lt = torch.zeros((0, D), dtype=torch.float32, device=lt.device)
    sorted_rows = gather_rows(lt, n, world, rank, group)
    inv = order.argsort()
    return sorted_rows.cpu().numpy()[inv]

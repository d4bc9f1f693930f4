"""Sharding by utterance across the GPUs of one box: utterance u -> rank (u mod G).

Utterances are independent (reset_segmentation @B25053 resets all state), so there is NO data-path
collective: each rank runs its own stream of files through its own Engine; only the feature rows are
gathered on the host (torch.distributed gather of small tensors over a host-side group).
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_utt: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, n_utt, world, dtype=np.int64)


def pack_rows(utt_ids, rows_per_utt, width: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """(keys[int64 n,2] = (utt_id, row_in_utt), rows[float64 n,width]) for a host-side gather.  width: 53 (levels 5 / 13),
    264 (level 11); taken from the rows when not given."""
    keys, rows = [], []
    for u, r in zip(utt_ids, rows_per_utt):
        r = np.asarray(r, np.float64)
        if width is None:
            width = r.shape[-1] if r.ndim == 2 else 53
        r = r.reshape(-1, width)
        keys.append(np.stack([np.full(len(r), u, np.int64), np.arange(len(r), dtype=np.int64)], axis=1))
        rows.append(r)
    if not keys:
        return np.zeros((0, 2), np.int64), np.zeros((0, width or 53), np.float64)
    return np.concatenate(keys), np.concatenate(rows)


def keys_from_counts(utt_ids: np.ndarray, rows_per_utt: np.ndarray) -> np.ndarray:
    """keys[int64 n,2] = (utt_id, row_in_utt) for a dense row table laid out utterance after utterance (Engine.result(None))."""
    utt_ids = np.asarray(utt_ids, np.int64)
    rows_per_utt = np.asarray(rows_per_utt, np.int64)
    n = int(rows_per_utt.sum())
    keys = np.empty((n, 2), np.int64)
    keys[:, 0] = np.repeat(utt_ids, rows_per_utt)
    starts = np.cumsum(rows_per_utt) - rows_per_utt
    keys[:, 1] = np.arange(n, dtype=np.int64) - np.repeat(starts, rows_per_utt)
    return keys


def gather_rows(keys: np.ndarray, rows: np.ndarray, dst: int = 0, group=None, sort: bool = True):
    """Gather (keys, rows) from every rank to `dst`, ordered by (utt_id, row).  Rows of any width (53, 264, ...).
    `group`: the process group to gather over -- a host-side (gloo) group keeps the rows on the host, as north_star asks;
    with the default group of an NCCL job the rows go through the device.  Returns None on other ranks."""
    import torch
    import torch.distributed as dist

    keys = np.ascontiguousarray(keys, np.int64).reshape(-1, 2)
    rows = np.ascontiguousarray(rows, np.float64)
    width = rows.shape[1] if rows.ndim == 2 else 53
    rows = rows.reshape(-1, width)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if not sort:
            return keys, rows
        order = np.lexsort((keys[:, 1], keys[:, 0]))
        return keys[order], rows[order]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    meta = torch.tensor([keys.shape[0], width], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0].item()) for m in metas]
    if any(int(m[1].item()) != width for m in metas):
        raise ValueError("gather_rows: ranks disagree on the row width")
    mx = max(counts)
    # one padded buffer per rank: keys (2 int64 bit patterns) and the row, side by side as float64 words
    buf = torch.zeros((mx, 2 + width), dtype=torch.float64, device=dev)
    if keys.shape[0]:
        buf[: keys.shape[0], :2] = torch.from_numpy(keys.view(np.float64)).to(dev)
        buf[: rows.shape[0], 2:] = torch.from_numpy(rows).to(dev)
    gl = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gl, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    if rank != dst:
        return None
    parts = [gl[i][: counts[i]].cpu().numpy() for i in range(world)]
    allb = np.concatenate(parts) if parts else np.zeros((0, 2 + width))
    K = np.ascontiguousarray(allb[:, :2]).view(np.int64)
    R = np.ascontiguousarray(allb[:, 2:])
    if not sort:
        return K, R
    order = np.lexsort((K[:, 1], K[:, 0]))
    return K[order], R[order]

"""Sharding by utterance across the GPUs of one box: utterance u -> rank (u mod G).

Utterances are independent (reset_segmentation @B25053 resets all state), so there is NO data-path
collective: each rank runs its own stream of files through its own Engine; only the feature rows are
gathered on the host (torch.distributed gather of small tensors, gloo or NCCL).
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_utt: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, n_utt, world, dtype=np.int64)


def pack_rows(utt_ids, rows_per_utt) -> tuple[np.ndarray, np.ndarray]:
    """(keys[int64 n,2] = (utt_id, row_in_utt), rows[float64 n,53]) for a host-side gather."""
    keys, rows = [], []
    for u, r in zip(utt_ids, rows_per_utt):
        r = np.asarray(r, np.float64).reshape(-1, 53)
        keys.append(np.stack([np.full(len(r), u, np.int64), np.arange(len(r), dtype=np.int64)], axis=1))
        rows.append(r)
    if not keys:
        return np.zeros((0, 2), np.int64), np.zeros((0, 53), np.float64)
    return np.concatenate(keys), np.concatenate(rows)


def gather_rows(keys: np.ndarray, rows: np.ndarray, dst: int = 0):
    """Gather (keys, rows) from every rank to `dst`, ordered by (utt_id, row).  Returns None on other ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        order = np.lexsort((keys[:, 1], keys[:, 0]))
        return keys[order], rows[order]
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    mx = int(max(int(c.item()) for c in counts))
    kp = torch.zeros((mx, 2), dtype=torch.int64, device=dev)
    rp = torch.zeros((mx, 53), dtype=torch.float64, device=dev)
    kp[: keys.shape[0]] = torch.from_numpy(keys).to(dev)
    rp[: rows.shape[0]] = torch.from_numpy(rows).to(dev)
    kl = [torch.zeros_like(kp) for _ in range(world)] if rank == dst else None
    rl = [torch.zeros_like(rp) for _ in range(world)] if rank == dst else None
    dist.gather(kp, kl, dst=dst)
    dist.gather(rp, rl, dst=dst)
    if rank != dst:
        return None
    K = np.concatenate([kl[i][: int(counts[i].item())].cpu().numpy() for i in range(world)])
    R = np.concatenate([rl[i][: int(counts[i].item())].cpu().numpy() for i in range(world)])
    order = np.lexsort((K[:, 1], K[:, 0]))
    return K[order], R[order]

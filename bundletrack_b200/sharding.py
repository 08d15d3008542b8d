"""Multi-GPU plumbing for the pose-graph path: tracking windows are independent units, so they are SHARDED over ranks
(one process per GPU) with no data-path collective; the only exchange is the final gather of the optimised poses
(N x 16 floats per window) — SURVEY.md §8e.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin: window w -> rank w mod world (balanced to within one window for any n)."""
    return list(range(rank, n_items, world))


def gather_poses(local_poses: Sequence[np.ndarray], n_items: int, frames_per_item: Sequence[int], rank: int, world: int, device=None):
    """All ranks contribute the poses of their shard; every rank gets the full list back in window order.
    local_poses[k] is the [N_k,4,4] result of the k-th window of this rank's shard (shard_indices order)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [np.asarray(p, np.float32) for p in local_poses]
    counts = [sum(frames_per_item[i] for i in shard_indices(n_items, r, world)) * 16 for r in range(world)]
    mx = max(counts)
    buf = torch.zeros(mx, dtype=torch.float32, device=device)
    if local_poses:
        flat = np.concatenate([np.asarray(p, np.float32).reshape(-1) for p in local_poses])
        buf[: flat.size] = torch.from_numpy(flat).to(buf.device)
    out = [torch.zeros(mx, dtype=torch.float32, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    result = [None] * n_items
    for r in range(world):
        host = out[r].cpu().numpy()
        o = 0
        for i in shard_indices(n_items, r, world):
            n = frames_per_item[i]
            result[i] = host[o:o + n * 16].reshape(n, 4, 4).copy()
            o += n * 16
    return result

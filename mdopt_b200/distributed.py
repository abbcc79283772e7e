"""Multi-GPU plumbing: syndromes are independent, so ranks take contiguous shards of the error list and the
only collective is the reduction of per-shard counts (reference: the Python mean over Pool.starmap results,
examples/decoding/quantum_surface.py:242-246, 290)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` of `world` (first n % world ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_counts(dist_rows: np.ndarray, success: np.ndarray) -> np.ndarray:
    """int64[3] = (decoded, successes, NaN rows) of one shard."""
    nan_rows = int(np.isnan(np.asarray(dist_rows, dtype=float)).any(axis=1).sum()) if len(dist_rows) else 0
    return np.array([len(success), int(np.sum(success)), nan_rows], dtype=np.int64)


def reduce_counts(counts: np.ndarray, device=None) -> np.ndarray:
    """Sum the per-shard counts over the default process group (NCCL on GPUs, gloo on CPU); identity when
    torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(counts, dtype=np.int64)
    t = torch.as_tensor(np.asarray(counts, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_rows(rows: np.ndarray, device=None):
    """Gather variable-length float64 row blocks from every rank onto every rank, in rank order."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(rows)
    world = dist.get_world_size()
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n_local = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    width = rows.shape[1] if rows.ndim == 2 else 1
    pad = torch.zeros((max(sizes), width), dtype=torch.float64, device=device)
    pad[: rows.shape[0]] = torch.as_tensor(rows.reshape(-1, width), device=device)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return np.concatenate([o[:n].cpu().numpy() for o, n in zip(out, sizes)], axis=0)

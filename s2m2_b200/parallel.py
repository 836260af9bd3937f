"""Multi-GPU plumbing for the S2M2 forward: stereo pairs are independent units, so the batch
is sharded across one-process-per-GPU ranks with NO collective in the forward.  The only exchange is a single
broadcast of the flat weight blob from rank 0 at start-up (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` stereo pairs owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flat_weights(model) -> torch.Tensor:
    """All parameters of an S2M2 module as one contiguous fp32 blob (reference key order)."""
    return torch.cat([p.data.reshape(-1).float() for p in model._p.values()])


def load_flat_weights(model, flat: torch.Tensor) -> None:
    off = 0
    for p in model._p.values():
        n = p.numel()
        p.data.copy_(flat[off:off + n].view_as(p))
        off += n
    if off != flat.numel():
        raise ValueError(f"flat weight blob has {flat.numel()} elements, model needs {off}")
    model._packed = None      # repack (fp16 K-major, fused QKV...) on next forward


def broadcast_weights(model, src: int = 0) -> int:
    """One collective: broadcast the flat blob from `src` to every rank.  Returns the number of bytes moved."""
    import torch.distributed as dist
    flat = flat_weights(model)
    dist.broadcast(flat, src)
    load_flat_weights(model, flat)
    return flat.numel() * 4

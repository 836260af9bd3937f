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
    model._graphs = {}        # captured forwards reference the old packed tensors


def broadcast_weights(model, src: int = 0) -> int:
    """One collective: broadcast the flat blob from `src` to every rank.  Returns the number of bytes moved."""
    import torch.distributed as dist
    flat = flat_weights(model)
    dist.broadcast(flat, src)
    load_flat_weights(model, flat)
    return flat.numel() * 4


# ---- packed blob ("one ncclBroadcast of the packed weight blob, L: 361 MB fp16") ----------------------------
def _packed_leaves(node, out):
    """Deterministic walk over S2M2._packed (nested dicts / lists of ConvW, tensors and scalars): every tensor leaf as
    (holder, key) so that it can be read and overwritten in place.  Every rank builds the same structure from the same
    (C, NTR, precision), so the order is identical across ranks."""
    from .model import ConvW
    if isinstance(node, dict):
        for k in node:
            v = node[k]
            if torch.is_tensor(v):
                out.append((node, k))
            else:
                _packed_leaves(v, out)
    elif isinstance(node, (list, tuple)):
        for i, v in enumerate(node):
            if torch.is_tensor(v):
                out.append((node, i))
            else:
                _packed_leaves(v, out)
    elif isinstance(node, ConvW):
        for k in ("w", "b"):
            if torch.is_tensor(getattr(node, k)):
                out.append((node, k))
    return out


def _get(holder, key):
    return getattr(holder, key) if isinstance(key, str) and not isinstance(holder, dict) else holder[key]


def broadcast_packed(model, device, src: int = 0) -> int:
    """The start-up collective of the data-parallel path: rank `src` packs its weights once (fp16 K-major, fused QKV / gate+fusion /
    GRU z+r, folded constants — model.S2M2._pack) and ONE broadcast ships the packed blob; the other ranks never see the fp32
    state-dict and never re-pack real weights (they pack their zero-initialised parameters once, only to allocate the identical
    structure the blob is scattered into).  Returns the number of bytes broadcast (L fp16: ~365 MB instead of 723 MB fp32).
    Note: on ranks != src `state_dict()` keeps returning the zero-initialised fp32 parameters — they serve, they do not export."""
    import torch.distributed as dist
    device = torch.device(device)
    model._ensure_packed(device)
    leaves = _packed_leaves(model._packed, [])
    offs, total = [], 0
    for holder, key in leaves:
        t = _get(holder, key)
        offs.append(total)
        total += (t.numel() * t.element_size() + 15) // 16 * 16
    flat = torch.zeros(total, dtype=torch.uint8, device=device)
    rank = dist.get_rank()
    if rank == src:
        for (holder, key), off in zip(leaves, offs):
            t = _get(holder, key)
            n = t.numel() * t.element_size()
            flat[off:off + n].copy_(t.detach().contiguous().reshape(-1).view(torch.uint8).to(device))
    dist.broadcast(flat, src)
    if rank != src:
        for (holder, key), off in zip(leaves, offs):
            t = _get(holder, key)
            n = t.numel() * t.element_size()
            src_t = flat[off:off + n].view(t.dtype).reshape(t.shape)
            t.copy_(src_t if t.is_cuda == src_t.is_cuda else src_t.to(t.device))
    return total

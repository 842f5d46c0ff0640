"""theta-batch sharding across the GPUs of one box.

The reference evaluates one theta per call in one process; the batched engine
shards a [B, P] theta matrix over ranks (one process per GPU, torch.distributed):
rank r evaluates the contiguous slice [r*ceil(B/G), (r+1)*ceil(B/G)) on its own
GPU against its own replica of the observations, and one all-gather of B/G
doubles per rank reassembles logL.  No reduction is involved, so every value is
bit-identical to the single-GPU result.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world: int, rank: int):
    """Contiguous slice of rank `rank`: (start, stop, per_rank) with per_rank = ceil(B / world);
    trailing ranks may get a short or empty slice."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} for world size {world}")
    per = -(-B // world) if B > 0 else 0
    start = min(B, rank * per)
    stop = min(B, start + per)
    return start, stop, per


def evaluate_sharded(evaluate, theta, group=None):
    """Evaluate ``evaluate(theta_slice) -> logl_slice`` on this rank's slice of `theta`
    ([B, P], identical on every rank) and all-gather the full [B] result.

    `theta` may be a numpy array (gathered through the group's CPU backend, e.g. gloo) or a
    torch CUDA tensor (NCCL).  Without an initialised process group this is a plain call."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return evaluate(theta)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = theta.shape[0]
    start, stop, per = shard_bounds(B, world, rank)
    is_torch = isinstance(theta, torch.Tensor)
    local = evaluate(theta[start:stop]) if stop > start else None
    if is_torch:
        buf = torch.full((per,), float("nan"), dtype=torch.float64, device=theta.device)
        if local is not None:
            buf[:stop - start] = local
        out = torch.empty(per * world, dtype=torch.float64, device=theta.device)
        dist.all_gather_into_tensor(out, buf, group=group)
        return out[:B] if per * world == B else torch.cat([out[r * per:r * per + max(0, min(B, (r + 1) * per) - r * per)]
                                                             for r in range(world)])
    buf = torch.full((per,), float("nan"), dtype=torch.float64)
    if local is not None:
        buf[:stop - start] = torch.as_tensor(np.asarray(local, dtype=np.float64))
    parts = [torch.empty(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return np.concatenate([parts[r][:max(0, min(B, (r + 1) * per) - r * per)].numpy() for r in range(world)])

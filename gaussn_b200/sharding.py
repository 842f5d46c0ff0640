"""theta-batch sharding across the GPUs of one box.

The reference evaluates one theta per call in one process; the batched engine
shards a [B, P] theta matrix over ranks (one process per GPU, torch.distributed):
rank r evaluates the contiguous slice [r*ceil(B/G), (r+1)*ceil(B/G)) on its own
GPU against its own replica of the observations, and one all-gather of B/G
doubles per rank reassembles logL.  No reduction is involved, so every value is
bit-identical to the single-GPU result.

Two ways to reassemble: ``evaluate_sharded`` (evaluate, then one NCCL / gloo all-gather) and ``FusedGather``
(NVLink-native: the kernel's epilogue stores every logL straight into the gathered buffer of every rank through
peer-mapped symmetric memory, ``gsn_logl_batch_gather``; only a barrier follows).
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


class FusedGather:
    """Gathered logL buffer in symmetric (peer-mapped) memory for the fused kernel epilogue.

    Every rank allocates ``world * per_rank`` doubles with torch's symmetric-memory allocator and exchanges
    the mappings once (``rendezvous``).  ``run`` launches this rank's shard with the pointers of ALL ranks'
    buffers, so each result is written locally and to the seven NVLink peers by the kernel itself; a device-side
    barrier on the same stream then makes the full vector valid on every rank.  No data-path collective is issued."""

    def __init__(self, per_rank: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("the fused gather addresses at most 16 peers")
        self.per_rank = int(per_rank)
        self.buf = symm_mem.empty(self.world * self.per_rank, dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def run(self, problem, theta, flags=0):
        """theta: this rank's [B_local <= per_rank, P] CUDA tensor.  Returns the gathered [world * per_rank] tensor
        (valid on the current stream after the call)."""
        import torch
        B = theta.shape[0]
        if B > self.per_rank:
            raise ValueError(f"shard of {B} rows exceeds the buffer's {self.per_rank} per rank")
        stream = torch.cuda.current_stream(theta.device).cuda_stream
        self.handle.barrier(channel=0)      # peers have consumed the previous contents
        problem.logl_device_gather(theta.data_ptr(), B, theta.shape[1], self.ptrs, self.rank * self.per_rank, 0, flags, stream)
        self.handle.barrier(channel=1)      # every rank's kernel (and its peer stores) is complete
        return self.buf

"""Frame-batch sharding across the GPUs of one box.

Frames are independent, so rank r of R processes frames [r*B/R, (r+1)*B/R) with
replicated weights and tables; the only collective is one all-gather of the
(B/R, J, 3) poses (180 B per frame).  Backend: NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (total % world) ranks get one extra frame."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_poses(local_kp: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather (n_local, J, 3) poses into (total, J, 3) in frame order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local_kp.shape[0] == total
        return local_kp
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(local_kp.shape[1:]), dtype=local_kp.dtype, device=local_kp.device)
    pad[: local_kp.shape[0]] = local_kp
    out = torch.empty((world * nmax,) + tuple(local_kp.shape[1:]), dtype=local_kp.dtype, device=local_kp.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view((world, nmax) + tuple(local_kp.shape[1:]))
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)

"""Frame-batch sharding across the GPUs of one box.

Frames are independent, so rank r of R processes frames [r*B/R, (r+1)*B/R) with
replicated weights and tables; the only collective is one all-gather of the
(B/R, J, 3) poses (180 B per frame).  Backend: NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of the PCIe root the GPU hangs off (sysfs), or None when the platform does not say."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the GPU's NUMA node BEFORE it allocates pinned host buffers: first-touch then
    places the staging memory next to the GPU's PCIe root, so with 8 ranks the host->device copies do not all cross
    the socket interconnect (round 1: 55 -> 23 GB/s per GPU from 1 to 8 ranks).  Returns the node, or None if the
    topology is unknown (nothing is changed then)."""
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        torch.set_num_threads(max(1, min(len(cpus), torch.get_num_threads())))
        return node
    except Exception:
        return None


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

"""CPU: the N>1 path (frame sharding + one all-gather of poses) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sceneego_b200.parallel import gather_poses, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    frames = torch.arange(total, dtype=torch.float32)
    local = (frames[lo:hi, None, None] * 10 + torch.arange(15)[None, :, None] + torch.arange(3)[None, None, :] * 0.1)
    out = gather_poses(local, total)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_poses_world2_ragged():
    for total in (7, 8):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
        [p.start() for p in procs]
        res = dict(q.get(timeout=120) for _ in range(2))
        [p.join(timeout=60) for p in procs]
        frames = torch.arange(total, dtype=torch.float32)
        ref = frames[:, None, None] * 10 + torch.arange(15)[None, :, None] + torch.arange(3)[None, None, :] * 0.1
        for r in range(2):
            assert res[r].shape == (total, 15, 3)
            assert torch.equal(res[r], ref)


def _worker_steps(rank, world, port, total, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    frames = torch.arange(total, dtype=torch.float32)
    # what bench.py / HostStagePipeline gather ONCE at the end of a run: (local frames, steps, J, 3)
    local = (frames[lo:hi, None, None, None] * 100 + torch.arange(steps)[None, :, None, None] * 10
             + torch.arange(15)[None, None, :, None] + torch.arange(3)[None, None, None, :] * 0.1)
    q.put((rank, gather_poses(local, total)))
    dist.barrier()
    dist.destroy_process_group()


def test_one_final_gather_of_all_steps_world2():
    """The poses of every step of a run are exchanged in ONE collective at the end (north_star: "NCCL only for the final
    gather of poses"): the trailing dimensions ride along, frame order is global."""
    total, steps = 6, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_steps, args=(r, 2, port, total, steps, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    frames = torch.arange(total, dtype=torch.float32)
    ref = (frames[:, None, None, None] * 100 + torch.arange(steps)[None, :, None, None] * 10
           + torch.arange(15)[None, None, :, None] + torch.arange(3)[None, None, None, :] * 0.1)
    for r in range(2):
        assert res[r].shape == (total, steps, 15, 3) and torch.equal(res[r], ref)


def test_numa_helpers_do_not_need_a_gpu():
    from sceneego_b200 import parallel
    assert parallel._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parallel._parse_cpulist("") == []
    assert parallel.gpu_numa_node(0) is None or isinstance(parallel.gpu_numa_node(0), int)   # no GPU here: None

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

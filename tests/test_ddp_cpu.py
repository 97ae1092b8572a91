"""CPU (gloo, world_size 2): host-side logic of the data-parallel path — sharding and the flat-buffer all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realise_b200.ddp import allreduce_sum_, shard_examples


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    w = allreduce_sum_(flat)
    ok = w == world and torch.equal(flat, torch.arange(10, dtype=torch.float32) * 3)
    # mean-of-ranks semantics after the optimizer's division by grad_div = W (DDP averages gradients)
    ok = ok and torch.allclose(flat / w, torch.arange(10, dtype=torch.float32) * 1.5)
    shard = shard_examples(list(range(11)), rank, world)
    ok = ok and shard == ([0, 2, 4, 6, 8] if rank == 0 else [1, 3, 5, 7, 9])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gloo_allreduce_and_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_shard_examples_drops_tail_like_reference():
    assert shard_examples(list(range(10)), 1, 4) == [1, 5]        # 10 // 4 = 2 per rank (src/run.py:131-137)
    assert shard_examples(list(range(3)), 0, 1) == [0, 1, 2]
    assert allreduce_sum_(torch.ones(3)) == 1                      # no process group: identity

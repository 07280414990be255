"""world_size-2 gloo test of the multi-GPU plumbing: images shard over ranks, ONE all-gather of the
final logits; rank order == image order.  Runs on CPU (gloo) with the same helper the NCCL path uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from odise_b200.pipeline import gather_logits
    B, Q, K1 = 3, 5, 7
    local = torch.full((B, Q, K1), float(rank)) + torch.arange(B).view(B, 1, 1) * 0.1
    out = gather_logits(local)
    q.put((rank, out.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_logits_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for r in range(world):
        out = res[r]
        assert out.shape == (6, 5, 7)
        for src in range(world):
            want = torch.full((3, 5, 7), float(src)) + torch.arange(3).view(3, 1, 1) * 0.1
            assert torch.equal(out[src * 3:(src + 1) * 3], want)


def test_gather_is_identity_without_process_group():
    from odise_b200.pipeline import gather_logits
    x = torch.randn(2, 3, 4)
    assert gather_logits(x) is x

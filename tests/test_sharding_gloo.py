"""CPU, world_size 2, gloo: the N>1 host path -- contiguous batch shards, replicated seeded
parameters, all_gather of predictions in batch order, max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp_

from monkey_pose_b200.sharding import PredictionGatherer, gather_predictions, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, rank, world)
    # stand-in for the per-rank forward: prediction row i is a function of the global frame index
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 69)
    full = gather_predictions(local, n_total)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the asynchronous, double-buffered form of the same gather over a stream of steps (equal shards only)
    if n_total % world == 0:
        g = PredictionGatherer(hi - lo, 69, depth=2)
        slots = []
        for step in range(5):
            slots.append(g.submit(local + 100.0 * step))
        g.drain()
        last, prev = g.result().clone(), g.result(slots[-2]).clone()
        try:
            g.submit(local[:-1])
            ragged_error = False
        except ValueError:
            ragged_error = True
        if rank == 0:
            ret["stream_last"], ret["stream_prev"], ret["ragged_error"] = last, prev, ragged_error
    if rank == 0:
        ret["full"] = full.clone()
        ret["tmax"] = float(t)
    dist.destroy_process_group()


def _run(n_total, world=2):
    mgr = mp_.Manager()
    ret = mgr.dict()
    mp_.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    return ret


def test_gather_predictions_even_shards():
    ret = _run(8)
    assert ret["full"].shape == (8, 69)
    assert torch.equal(ret["full"][:, 0], torch.arange(8, dtype=torch.float32))
    assert ret["tmax"] == 2.0
    # steps 4 and 3 of the stream, in batch order, from the two buffers
    assert torch.equal(ret["stream_last"][:, 3], torch.arange(8, dtype=torch.float32) + 400.0)
    assert torch.equal(ret["stream_prev"][:, 68], torch.arange(8, dtype=torch.float32) + 300.0)
    assert ret["ragged_error"]


def test_gather_predictions_ragged_shards():
    ret = _run(7)
    assert ret["full"].shape == (7, 69)
    assert torch.equal(ret["full"][:, 5], torch.arange(7, dtype=torch.float32))

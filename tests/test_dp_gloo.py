"""CPU, world_size 2 over gloo: the host-side data-parallel logic (camera sharding, single flat-gradient all-reduce,
replica consistency) with an oracle-backed stand-in for the per-rank frame gradient.  The CUDA kernels are not
involved here (they are covered by -m gpu); what is tested is exactly the code path trainer.py runs around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gms_b200.trainer import shard_cameras


def test_shard_cameras_partitions_every_step():
    for world in (1, 2, 4, 8):
        for step in range(5):
            got = [shard_cameras(16, step, r, world) for r in range(world)]
            assert len(set(got)) == world            # distinct cameras within a step
        seen = [shard_cameras(16, s, r, world) for s in range(16 // world) for r in range(world)]
        assert sorted(seen) == list(range(16))       # one sweep covers every camera exactly once


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # replica of the shared parameters, gradients in ONE flat buffer (what FlatAdam exposes as flat_grad)
    shapes = [(30, 3), (20, 2, 3), (40, 16, 3), (40, 1), (40, 1)]
    n = sum(int(np.prod(s)) for s in shapes)
    flat = torch.zeros(n)
    # per-rank "frame gradient": deterministic function of the camera this rank renders
    cam = shard_cameras(8, 3, rank, world)
    g = torch.Generator().manual_seed(100 + cam)
    flat.copy_(torch.randn(n, generator=g))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    out[rank] = (cam, flat.clone())
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    cams = [out[r][0] for r in range(world)]
    assert cams == [shard_cameras(8, 3, r, world) for r in range(world)] and cams[0] != cams[1]
    expect = sum(torch.randn(out[0][1].numel(), generator=torch.Generator().manual_seed(100 + c)) for c in cams) / world
    for r in range(world):
        torch.testing.assert_close(out[r][1], expect)      # every replica ends with the same averaged gradient

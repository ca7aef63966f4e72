"""CPU, world_size 2 over gloo: the host-side data-parallel logic -- camera sharding and the SHARDED FlatAdam
(gms_b200/optim.py: gradient exchange -> Adam on this rank's slice -> all-gather of the parameters) -- driven through the
product's own FlatAdam class with the CUDA kernel replaced by a float64 stub (`kernel=` hook) and a deterministic stand-in
for the per-rank frame gradient.  The kernels themselves are covered by -m gpu (tests/test_gpu_step.py, tests/test_gpu_dist.py)."""
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


def _adam_stub(d):
    """Float64 restatement of gms_adam_step on the descriptor FlatAdam hands to its kernel (per-segment learning rates,
    DC/rest phase of the packed SH segment, shard offset)."""
    n, off = d["n"], d["offset"]
    idx = torch.arange(off, off + n)
    lr = torch.zeros(n, dtype=torch.float64)
    start = 0
    for end, lr0, lr1, inner, period in zip(d["seg_end"], d["lr0"], d["lr1"], d["inner"], d["period"]):
        sel = (idx >= start) & (idx < end)
        if period > 0:
            phase = ((idx - start) // inner) % period
            lr[sel] = torch.where(phase[sel] == 0, lr0, lr1).double() if sel.any() else lr[sel]
        else:
            lr[sel] = lr0
        start = end
    g = d["g"][:n].double()
    m = d["beta1"] * d["m"][:n].double() + (1 - d["beta1"]) * g
    v = d["beta2"] * d["v"][:n].double() + (1 - d["beta2"]) * g * g
    step = lr / (1 - d["beta1"] ** d["step"])
    p = d["p"][:n].double() - step * m / (v.sqrt() / (1 - d["beta2"] ** d["step"]) ** 0.5 + d["eps"])
    d["p"][:n].copy_(p.float()); d["m"][:n].copy_(m.float()); d["v"][:n].copy_(v.float())
    if d["zero_grad"] == 1:
        d["g"][:n].zero_()
    elif d["zero_grad"] == 2:
        k = max(0, min(n, d["zero_end"] - off))
        d["g"][:k].zero_()


def _make_groups(seed=0):
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g))
    return [dict(param=mk(30, 3), lr=1e-4, name="vertices"), dict(param=mk(20, 2, 3), lr=1e-3, name="alpha"),
            dict(param=mk(40, 16, 3), lr0=2.5e-3, lr1=1.25e-4, inner=3, period=16, name="features"),
            dict(param=mk(40, 1), lr=5e-2, name="opacity"), dict(param=mk(40, 1), lr=5e-3, name="scaling")]


def _frame_gradient(n, cam, step):
    return torch.randn(n, generator=torch.Generator().manual_seed(1000 * step + cam))


def _flat_adam_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gms_b200.optim import FlatAdam
    opt = FlatAdam(_make_groups(), world=world, rank=rank, kernel=_adam_stub)
    for step in range(3):
        cam = shard_cameras(8, step, rank, world)
        opt.g.copy_(_frame_gradient(opt.n, cam, step))
        opt.step(zero_end=opt.ends[0])
    out[rank] = (opt.p.clone(), opt.g.clone(), opt.shard, opt.n)
    dist.destroy_process_group()


def test_sharded_flat_adam_world2_equals_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_flat_adam_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    from gms_b200.optim import FlatAdam
    ref = FlatAdam(_make_groups(), world=1, rank=0, kernel=_adam_stub)
    n = out[0][3]
    assert out[0][2] * world == n and out[0][2] % 64 == 0            # equal, 256-byte aligned shards
    pad = n - ref.n
    for step in range(3):
        cams = [shard_cameras(8, step, r, world) for r in range(world)]
        gavg = sum(_frame_gradient(n, c, step) for c in cams) / world
        ref.g.copy_(gavg[:ref.n] if pad else gavg)
        ref.step(zero_end=ref.ends[0])
    for r in range(world):
        torch.testing.assert_close(out[r][0][:ref.n], ref.p, rtol=1e-6, atol=1e-7)     # every replica == the single-process optimiser
        assert torch.equal(out[r][0], out[0][0])                                         # replicas identical bit for bit
        assert float(out[r][1][:ref.ends[0]].abs().max()) == 0.0                         # the vertex-gradient segment was re-zeroed


def _groups_features_last(seed=0):
    g = _make_groups(seed)
    return [g[0], g[1], g[3], g[4], g[2]]          # mesh_model_groups(features_last=True): the packed SH tensor closes the list


def _factored_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gms_b200.optim import FlatAdam
    opt = FlatAdam(_groups_features_last(), world=world, rank=rank, kernel=_adam_stub, sh_factored=True)
    seen = []
    opt._adam_sh = lambda sh: seen.append(sh["exchange"].clone())          # (the SH update itself is a CUDA kernel: -m gpu tests)
    P, slot = 40, 192
    ex = torch.zeros(world, slot)
    for step in range(3):
        cam = shard_cameras(8, step, rank, world)
        opt.g.copy_(_frame_gradient(opt.n, cam, step))
        ex.zero_()
        ex[rank, :3 * P + 3] = torch.arange(3 * P + 3, dtype=torch.float32) + 1000.0 * cam      # this rank's colour gradients + camera centre
        opt.step(zero_end=opt.ends[0], sh=dict(xyz=0, exchange=ex, degree=3, event=None))
    out[rank] = (opt.p.clone(), opt.g.clone(), opt.n, opt.ends[-2], [t.clone() for t in seen], opt.m.numel())
    dist.destroy_process_group()


def test_factored_flat_adam_world2_exchange_and_replicated_update():
    """FlatAdam(sh_factored=True) under gloo: the colour-gradient slots of BOTH ranks reach every rank before the SH update
    is called, the other groups' gradients are averaged and updated by the (replicated) k_adam pass, the SH segment is left
    to the factored kernel, and the moments are kept in full on every rank."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_factored_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    from gms_b200.optim import FlatAdam
    ref = FlatAdam(_groups_features_last(), world=1, rank=0, kernel=_adam_stub, sh_factored=True)
    ref._adam_sh = lambda sh: None
    n, prefix = out[0][2], out[0][3]
    assert prefix == ref.ends[-2] and out[0][5] == n                # full-length moments: replicated optimizer
    for step in range(3):
        cams = [shard_cameras(8, step, r, world) for r in range(world)]
        gavg = sum(_frame_gradient(n, c, step) for c in cams) / world
        ref.g.copy_(gavg[:ref.n])
        ref.step(zero_end=ref.ends[0], sh=dict(xyz=0, exchange=torch.zeros(1, 192), degree=3, event=None))
        for r in range(world):
            got = out[r][4][step]
            for q, c in enumerate(cams):                            # slot q holds rank q's frame on every rank
                assert float(got[q, 0]) == 1000.0 * c + 0.0 and float(got[q, 122]) == 1000.0 * c + 122.0
    init = _groups_features_last()[-1]["param"].detach().reshape(-1)
    for r in range(world):
        torch.testing.assert_close(out[r][0][:prefix], ref.p[:prefix], rtol=1e-6, atol=1e-7)
        assert torch.equal(out[r][0], out[0][0])
        assert torch.equal(out[r][0][prefix:prefix + init.numel()], init)      # the SH parameters were not touched by k_adam
        assert float(out[r][1][:ref.ends[0]].abs().max()) == 0.0

"""FlatAdam -- the reference's Adam (gaussian_mesh_model.py:171-183, train.py:146-148) as ONE kernel launch.

All learnable tensors, their gradients and both Adam moments live in four flat fp32 buffers; `param.data` and
`param.grad` are views into them.  One launch of gms_adam_step updates everything, applies the per-group learning rates
(vertices / alpha / f_dc / f_rest / opacity / scaling, arguments_games/__init__.py:17-30) and zeroes the gradient in the
same pass.  The flat gradient buffer is also what the data-parallel all-reduce sends (trainer.py)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from . import _lib

# name -> learning rate (OptimizationParamsMesh, arguments_games/__init__.py:17-30)
REFERENCE_LRS = dict(vertices=0.0, alpha=0.001, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005)


class FlatAdam:
    def __init__(self, groups: Sequence[dict], betas=(0.9, 0.999), eps: float = 1e-15, world: int = 1, rank: int = 0):
        """groups: dicts with `param` and either `lr`, or (`lr0`, `lr1`, `inner`, `period`) for the packed SH tensor.
        world > 1: SHARDED optimizer (ZeRO-1 style).  The gradient exchange is a reduce-scatter, every rank keeps Adam
        moments for and updates only its 1/world slice of the flat buffer, and an all-gather brings the updated
        parameters back -- the same bytes on NVLink as an all-reduce, but the 28 B/parameter Adam pass shrinks by `world`."""
        self.groups = list(groups)
        self.world, self.rank = int(world), int(rank)
        assert 1 <= len(self.groups) <= 8
        params = [g["param"] for g in self.groups]
        dev = params[0].device
        pad = lambda k: (k + 63) // 64 * 64     # every segment starts 256-byte aligned (the kernels use 128-bit accesses)
        n = sum(pad(p.numel()) for p in params)
        n = (n + 64 * self.world - 1) // (64 * self.world) * (64 * self.world)     # equal, 256-byte aligned shards
        self.n = n
        self.shard = n // self.world
        self.p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)      # moments: this rank's slice only
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.g_shard = torch.zeros(self.shard, dtype=torch.float32, device=dev) if self.world > 1 else None
        off = 0
        self.ends = []
        for p in params:
            k = p.numel()
            self.p[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.p[off:off + k].view(p.shape)
            p.grad = self.g[off:off + k].view(p.shape)
            off += pad(k)
            self.ends.append(off)
        self.betas, self.eps, self.t = betas, eps, 0

    @property
    def flat_grad(self) -> torch.Tensor:
        return self.g

    def zero_grad(self):
        self.g.zero_()

    def step(self, zero_end=None):
        """world == 1: one launch over the whole flat buffer (gradient zeroed in the same pass).
        world > 1: reduce-scatter(mean) -> Adam on the local slice -> all-gather of the parameters; the full gradient
        buffer is re-zeroed with one memset.
        zero_end: when the producer of the gradients OVERWRITES everything at flat indices >= zero_end each frame
        (gms_train_frame: all but the atomically accumulated vertex gradients), only [0, zero_end) is zeroed."""
        import torch.distributed as dist
        self.t += 1
        a = _lib.AdamArgs()
        if self.world > 1:
            # NCCL averages inside the collective (exact for power-of-two world sizes): no separate scaling pass
            dist.reduce_scatter_tensor(self.g_shard, self.g, op=dist.ReduceOp.AVG)
            off = self.rank * self.shard
            p_local = self.p[off:off + self.shard]
            a.n, a.offset = self.shard, off
            a.p, a.g, a.m, a.v = p_local.data_ptr(), self.g_shard.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
            a.zero_grad = 0            # g_shard is overwritten by the next reduce-scatter
        else:
            a.n, a.offset = self.n, 0
            a.p, a.g, a.m, a.v = self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
            a.zero_grad, a.zero_end = (1, 0) if zero_end is None else (2, int(zero_end))
        a.nseg = len(self.groups)
        for i, g in enumerate(self.groups):
            a.seg_end[i] = self.ends[i]
            a.lr0[i] = float(g.get("lr0", g.get("lr", 0.0)))
            a.lr1[i] = float(g.get("lr1", g.get("lr", 0.0)))
            a.inner[i] = int(g.get("inner", 1))
            a.period[i] = int(g.get("period", 0))
        a.beta1, a.beta2, a.eps, a.step = self.betas[0], self.betas[1], self.eps, self.t
        dev = self.p.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gms_adam_step(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "gms_adam_step")
        if self.world > 1:
            dist.all_gather_into_tensor(self.p, p_local)
            (self.g if zero_end is None else self.g[:int(zero_end)]).zero_()

    def zero_grad_partial(self, zero_end=None):
        (self.g if zero_end is None else self.g[:int(zero_end)]).zero_()


def mesh_model_groups(model, lrs=REFERENCE_LRS) -> List[dict]:
    """Parameter groups of a MeshGaussianModel in the reference's order and learning rates."""
    g = [dict(param=model.vertices, lr=lrs["vertices"], name="vertices"), dict(param=model._alpha, lr=lrs["alpha"], name="alpha")]
    if model._features is not None:
        M = model._features.shape[1]
        g.append(dict(param=model._features, lr0=lrs["f_dc"], lr1=lrs["f_rest"], inner=3, period=M, name="features"))
    else:
        g += [dict(param=model._features_dc, lr=lrs["f_dc"], name="f_dc"), dict(param=model._features_rest, lr=lrs["f_rest"], name="f_rest")]
    g += [dict(param=model._opacity, lr=lrs["opacity"], name="opacity"), dict(param=model._scale, lr=lrs["scaling"], name="scaling")]
    return g

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
    def __init__(self, groups: Sequence[dict], betas=(0.9, 0.999), eps: float = 1e-15):
        """groups: dicts with `param` and either `lr`, or (`lr0`, `lr1`, `inner`, `period`) for the packed SH tensor."""
        self.groups = list(groups)
        assert 1 <= len(self.groups) <= 8
        params = [g["param"] for g in self.groups]
        dev = params[0].device
        pad = lambda k: (k + 63) // 64 * 64     # every segment starts 256-byte aligned (the kernels use 128-bit accesses)
        n = sum(pad(p.numel()) for p in params)
        self.n = n
        self.p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
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

    def step(self):
        self.t += 1
        a = _lib.AdamArgs()
        a.n, a.p, a.g, a.m, a.v = self.n, self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        a.nseg = len(self.groups)
        for i, g in enumerate(self.groups):
            a.seg_end[i] = self.ends[i]
            a.lr0[i] = float(g.get("lr0", g.get("lr", 0.0)))
            a.lr1[i] = float(g.get("lr1", g.get("lr", 0.0)))
            a.inner[i] = int(g.get("inner", 1))
            a.period[i] = int(g.get("period", 0))
        a.beta1, a.beta2, a.eps, a.step, a.zero_grad = self.betas[0], self.betas[1], self.eps, self.t, 1
        dev = self.p.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gms_adam_step(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "gms_adam_step")


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

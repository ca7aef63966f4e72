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
    def __init__(self, groups: Sequence[dict], betas=(0.9, 0.999), eps: float = 1e-15, world: int = 1, rank: int = 0, kernel=None,
                 sh_factored: bool = False):
        """groups: dicts with `param` and either `lr`, or (`lr0`, `lr1`, `inner`, `period`) for the packed SH tensor.
        world > 1: SHARDED optimizer (ZeRO-1 style).  The gradient exchange is a reduce-scatter, every rank keeps Adam
        moments for and updates only its 1/world slice of the flat buffer, and an all-gather brings the updated
        parameters back -- the same bytes on NVLink as an all-reduce, but the 28 B/parameter Adam pass shrinks by `world`."""
        self.groups = list(groups)
        self.world, self.rank = int(world), int(rank)
        # sh_factored: the LAST group is the packed SH tensor [P,16,3] and its gradient arrives as factors (step(sh=...)):
        # the optimizer is then REPLICATED (full moments on every rank), the exchange is an all-reduce of the other groups'
        # gradients (32 B/Gaussian) plus an all-gather of the colour gradients (12 B/Gaussian/rank), and no parameter
        # all-gather follows.  Dense mode (default): sharded optimizer, reduce-scatter + all-gather of everything.
        self.sh_factored = bool(sh_factored)
        if self.sh_factored and not ("period" in self.groups[-1] and self.groups[-1]["param"].dim() == 3):
            raise ValueError("sh_factored needs the packed SH tensor as the last group (mesh_model_groups(features_last=True))")
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
        nm = n if self.sh_factored else self.shard
        self.m = torch.zeros(nm, dtype=torch.float32, device=dev)      # moments: this rank's slice only (dense mode)
        self.v = torch.zeros(nm, dtype=torch.float32, device=dev)
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
        self._kernel = kernel or self._cuda_kernel       # `kernel`: test hook (a callable taking the _adam_desc dict)
        self._comm = None                                # side stream of the factored exchange (created on first use)

    @property
    def flat_grad(self) -> torch.Tensor:
        return self.g

    def zero_grad(self):
        self.g.zero_()

    # ---- the three stages of a step; `step()` strings them together (tests drive them under gloo with a stub kernel)
    def _exchange_gradient(self):
        """world > 1: average the flat gradient over the ranks and return (this rank's slice of it, parameter slice, flat
        offset).  NCCL: one reduce-scatter that averages inside the collective (exact for power-of-two world sizes).
        Backends without reduce-scatter / AVG (gloo: the CPU tests): all-reduce(SUM), scale, slice."""
        import torch.distributed as dist
        off = self.rank * self.shard
        p_local = self.p[off:off + self.shard]
        if self.world <= 1:
            return self.g, self.p, 0
        if dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(self.g_shard, self.g, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.g, op=dist.ReduceOp.SUM)
            self.g_shard.copy_(self.g[off:off + self.shard]).mul_(1.0 / self.world)
        return self.g_shard, p_local, off

    def _adam_desc(self, n, offset, p, g, zero_grad, zero_end):
        """Everything gms_adam_step needs, as plain Python (the stub kernel of the CPU tests reads the same dict)."""
        return dict(n=int(n), offset=int(offset), p=p, g=g, m=self.m, v=self.v, seg_end=list(self.ends),
                    lr0=[float(g_.get("lr0", g_.get("lr", 0.0))) for g_ in self.groups],
                    lr1=[float(g_.get("lr1", g_.get("lr", 0.0))) for g_ in self.groups],
                    inner=[int(g_.get("inner", 1)) for g_ in self.groups], period=[int(g_.get("period", 0)) for g_ in self.groups],
                    beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, step=self.t, zero_grad=int(zero_grad), zero_end=int(zero_end))

    def _cuda_kernel(self, d):
        a = _lib.AdamArgs()
        a.n, a.offset = d["n"], d["offset"]
        a.p, a.g, a.m, a.v = d["p"].data_ptr(), d["g"].data_ptr(), d["m"].data_ptr(), d["v"].data_ptr()
        a.nseg = len(d["seg_end"])
        for i in range(a.nseg):
            a.seg_end[i], a.lr0[i], a.lr1[i], a.inner[i], a.period[i] = d["seg_end"][i], d["lr0"][i], d["lr1"][i], d["inner"][i], d["period"][i]
        a.beta1, a.beta2, a.eps, a.step, a.zero_grad, a.zero_end = d["beta1"], d["beta2"], d["eps"], d["step"], d["zero_grad"], d["zero_end"]
        dev = d["p"].device
        if not d["p"].is_cuda:
            raise RuntimeError("FlatAdam: CUDA tensors required (no CPU path in the product)")
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gms_adam_step(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "gms_adam_step")

    def _publish_parameters(self, p_local, zero_end):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.p, p_local)
            (self.g if zero_end is None else self.g[:int(zero_end)]).zero_()

    def _step_factored(self, zero_end, sh):
        """sh = dict(xyz=<device pointer>, exchange=<[R, slot] float tensor: slot r = colour gradients + camera centre of rank r's frame>,
        degree=<active SH degree>)."""
        import torch.distributed as dist
        prefix = self.ends[-2]                       # everything but the SH group
        ex = sh["exchange"]
        gathered = reduced = None
        if self.world > 1:
            dev = self.p.device
            avg = dist.get_backend() == "nccl"

            def reduce_rest():
                if avg:
                    dist.all_reduce(self.g[:prefix], op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(self.g[:prefix], op=dist.ReduceOp.SUM); self.g[:prefix].mul_(1.0 / self.world)

            if self.g.is_cuda:
                # Both collectives run from a side stream: the all-gather of the colour gradients starts right after the
                # preprocess backward (event recorded inside gms_train_frame) and overlaps the opacity / expansion backward; the
                # all-reduce of the other gradients starts when the frame is complete and overlaps k_adam_sh, which needs the
                # gather only.  k_adam (the non-SH parameters) waits for the all-reduce.
                main = torch.cuda.current_stream(dev)
                if self._comm is None:
                    self._comm = torch.cuda.Stream(dev)
                comm = self._comm
                comm.wait_event(sh["event"] if sh.get("event") is not None else main.record_event())
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(ex.view(-1), ex[self.rank])
                    gathered = comm.record_event()
                comm.wait_event(main.record_event())
                with torch.cuda.stream(comm):
                    reduce_rest()
                    reduced = comm.record_event()
            else:
                dist.all_gather_into_tensor(ex.view(-1), ex[self.rank])
                reduce_rest()
        if gathered is not None:
            torch.cuda.current_stream(self.p.device).wait_event(gathered)
        self._adam_sh(sh)
        if reduced is not None:
            torch.cuda.current_stream(self.p.device).wait_event(reduced)
        d = self._adam_desc(prefix, 0, self.p, self.g, 1 if zero_end is None else 2, 0 if zero_end is None else zero_end)
        for k in ("seg_end", "lr0", "lr1", "inner", "period"):
            d[k] = d[k][:-1]
        self._kernel(d)

    def _adam_sh(self, sh):
        ex = sh["exchange"]
        gsh = self.groups[-1]
        f = gsh["param"]
        off = self.ends[-2]
        a = _lib.AdamShArgs()
        a.P, a.M, a.sh_degree, a.R = f.shape[0], f.shape[1], int(sh["degree"]), ex.shape[0]
        a.xyz, a.exchange, a.slot_floats, a.grad_scale = int(sh["xyz"]), ex.data_ptr(), ex.shape[1], 1.0 / ex.shape[0]
        a.p, a.m, a.v = f.data_ptr(), self.m[off:].data_ptr(), self.v[off:].data_ptr()
        a.lr_dc, a.lr_rest = float(gsh["lr0"]), float(gsh["lr1"])
        a.beta1, a.beta2, a.eps, a.step = self.betas[0], self.betas[1], self.eps, self.t
        dev = f.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gms_adam_sh_factored(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "gms_adam_sh_factored")

    def step(self, zero_end=None, sh=None):
        """sh given (sh_factored optimizers): see _step_factored.  Otherwise:
        world == 1: one launch over the whole flat buffer (gradient zeroed in the same pass).
        world > 1: reduce-scatter(mean) -> Adam on the local slice -> all-gather of the parameters; the full gradient
        buffer is re-zeroed with one memset.
        zero_end: when the producer of the gradients OVERWRITES everything at flat indices >= zero_end each frame
        (gms_train_frame: all but the atomically accumulated vertex gradients), only [0, zero_end) is zeroed."""
        self.t += 1
        if self.sh_factored:
            if sh is None:
                raise ValueError("this FlatAdam was built with sh_factored=True: step() needs the SH gradient factors (sh=...)")
            return self._step_factored(zero_end, sh)
        g, p_local, off = self._exchange_gradient()
        if self.world > 1:
            d = self._adam_desc(self.shard, off, p_local, g, 0, 0)          # g_shard is overwritten by the next exchange
        else:
            d = self._adam_desc(self.n, 0, self.p, self.g, 1 if zero_end is None else 2, 0 if zero_end is None else zero_end)
        self._kernel(d)
        self._publish_parameters(p_local, zero_end)

    def zero_grad_partial(self, zero_end=None):
        (self.g if zero_end is None else self.g[:int(zero_end)]).zero_()


def mesh_model_groups(model, lrs=REFERENCE_LRS, features_last: bool = False) -> List[dict]:
    """Parameter groups of a MeshGaussianModel with the reference's learning rates; the reference's order (vertices, alpha,
    f_dc, f_rest, opacity, scaling -- gaussian_mesh_model.py:174-181), or with the packed SH tensor moved to the end
    (features_last: what FlatAdam(sh_factored=True) needs; the order of Adam groups has no numerical meaning)."""
    g = [dict(param=model.vertices, lr=lrs["vertices"], name="vertices"), dict(param=model._alpha, lr=lrs["alpha"], name="alpha")]
    feats = []
    if model._features is not None:
        M = model._features.shape[1]
        feats.append(dict(param=model._features, lr0=lrs["f_dc"], lr1=lrs["f_rest"], inner=3, period=M, name="features"))
    else:
        feats += [dict(param=model._features_dc, lr=lrs["f_dc"], name="f_dc"), dict(param=model._features_rest, lr=lrs["f_rest"], name="f_rest")]
    rest = [dict(param=model._opacity, lr=lrs["opacity"], name="opacity"), dict(param=model._scale, lr=lrs["scaling"], name="scaling")]
    return g + (rest + feats if features_last else feats + rest)

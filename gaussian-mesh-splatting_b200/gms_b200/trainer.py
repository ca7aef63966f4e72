"""One optimisation step of gs_mesh on the B200-native path, single GPU or frame-sharded data parallel.

Step structure = train.py:89-157 of the reference: pick a camera, render (expansion + rasterizer), loss
(train.py:105-107), backward (train.py:108), optimizer step (train.py:146-148), re-expand (train.py:154-157 -- here the
expansion is simply the first thing the next step's render does, in one fused launch).

Multi-GPU (the reference has none): one process per GPU, every rank holds a replica of the mesh-Gaussian
parameters, rank r renders camera `step*world + r` of the schedule, and ONE NCCL all-reduce per step averages the
flattened gradient of the shared parameters (vertices, _alpha, _scale, _features_dc, _features_rest, _opacity).
All gradients live in ONE contiguous buffer (param.grad are views into it), so the collective is a single call.
"""
from __future__ import annotations

import collections

import math
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

import diff_gaussian_rasterization as dgr

from .losses import fused_training_loss
from .model import MeshGaussianModel
from .optim import FlatAdam, mesh_model_groups
from .scenes import Camera


def shard_cameras(n_cameras: int, step: int, rank: int, world: int) -> int:
    """Camera index rank `rank` renders at `step`: consecutive cameras of the schedule go to consecutive ranks."""
    return (step * world + rank) % n_cameras


def render_frame(model: MeshGaussianModel, cam: Camera, bg: torch.Tensor, fused: bool = True, antialiasing: bool = False):
    """Expansion + rasterizer forward for one camera.  Returns (image, radii, invdepth)."""
    if fused:
        xyz, scales, rots = model.expand_fused(activated=True)
    else:   # the reference's two-step protocol + getters
        model.update_alpha(); model.prepare_scaling_rot()
        xyz, scales, rots = model.get_xyz, model.get_scaling, model.get_rotation
    rs = dgr.GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=model.active_sh_degree, campos=cam.camera_center, prefiltered=False, debug=False, antialiasing=antialiasing)
    means2D = torch.zeros_like(xyz, requires_grad=xyz.requires_grad)
    return dgr.GaussianRasterizer(raster_settings=rs)(means3D=xyz, means2D=means2D, opacities=model.get_opacity,
                                                      shs=model.get_features, scales=scales, rotations=rots)


class NativeFrame:
    """One training frame through gms_train_frame: expansion, rasterizer, loss and both backward passes issued from ONE C
    call on the current stream -- no autograd graph, no per-op Python.  Gradients land in the parameters' preallocated
    .grad views (FlatAdam's flat buffer); intermediates live in a persistent workspace, rasterizer scratch in grow-only
    buffers served through the allocation callback.

    sync_free (default): only the FIRST frame learns N through the stock-style 4-byte read-back; from then on the binning
    region is capacity-sized, N stays on the device and is mirrored by the range kernel into a ring of mapped pinned host
    slots (`n_host`: one (N, overflow flag) pair per in-flight frame) that the host polls WITHOUT synchronising.
    The tile sort runs over the capacity, so the capacity is predicted PER VIEW: 1.08x the N this camera had at its last
    visit (more when it moved a lot between its last two visits) + 64k; a camera seen for the first time gets 1.25x the
    largest N seen so far + 256k.  A frame whose N exceeds its capacity renders the background with zero gradients; the
    host notices when it harvests that frame's slot, counts it in `overflows`, and the camera's next visit is sized from
    the true N."""

    RING = 64       # mapped (N, flag) slots = frames the host may run ahead of the device before it waits

    def __init__(self, model: MeshGaussianModel, width: int, height: int, lambda_dssim: float = 0.2, sync_free: bool = True,
                 world: int = 1, rank: int = 0):
        import ctypes as C
        from . import _lib
        assert model._features is not None, "NativeFrame needs packed SH features"
        self.model, self.W, self.H, self.lam = model, int(width), int(height), float(lambda_dssim)
        dev = model.vertices.device
        self.dev = dev
        P = model._scale.shape[0]
        self.ws = torch.empty(int(_lib.lib().gms_frame_workspace_bytes(P, self.W, self.H)), dtype=torch.uint8, device=dev)
        self.loss = torch.zeros(3, dtype=torch.float32, device=dev)
        self.n_rendered = C.c_int64(0)
        self.sync_free = bool(sync_free)
        self.capacity = 0                       # duplicates the most recent frame's binning region was sized for (0: not known yet)
        self.capacity_override = None           # tests: force the next frames' capacity
        self.n_host = torch.zeros(self.RING, 2, dtype=torch.int32).pin_memory()   # slot i: (N, overflow flag) of an in-flight frame
        self._n_np = self.n_host.numpy()
        self._pending = collections.deque()     # (slot, view key, capacity) of frames whose N has not been harvested yet
        self._view_n = {}                       # view key -> (N at the last visit, N at the visit before)
        self._n_max, self._n_last, self._frame_no = 0, 0, 0
        self.overflows = 0
        # factored SH gradient (run(..., factored=True)): slot r of `exchange` = [3P colour gradients | camera centre | pad] of
        # rank r's frame; this rank's frame writes slot `rank`, FlatAdam(sh_factored=True) all-gathers and consumes the rest
        self.world, self.rank = int(world), int(rank)
        self.slot = (3 * P + 3 + 63) // 64 * 64
        self.exchange = None
        self.ev_loss = None         # recorded by gms_train_frame right after the loss kernels (read_loss_async)
        self._loss_read, self._side = None, None
        self.ev_sh = None           # recorded by gms_train_frame right after the preprocess backward (world > 1: the exchange of the
                                    # colour gradients starts there, on the optimizer's communication stream)
        self._check_model()
        scratch = {}
        self._scratch = scratch

        def _alloc(user, which, nbytes):        # grow-only, persistent across frames: no allocator traffic in steady state
            t = scratch.get(int(which))
            if t is None or t.numel() < nbytes:
                try:
                    t = torch.empty(int(nbytes * 1.25) + (1 << 20), dtype=torch.uint8, device=dev)
                except Exception:
                    return 0
                scratch[int(which)] = t
            return t.data_ptr()

        self._cb = _lib.ALLOC_FN(_alloc)        # closure captures `scratch`/`dev` only (no reference cycle through self)

    def _check_model(self):
        """Raw pointers go straight to CUDA kernels: dtype / device / layout are checked here, once, instead of failing
        as an illegal address later."""
        m = self.model
        if m.faces.dtype != torch.int64 or not m.faces.is_contiguous():
            m.faces = m.faces.long().contiguous()
        for name in ("vertices", "_alpha", "_scale", "_features", "_opacity"):
            t = getattr(m, name)
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == self.dev):
                raise RuntimeError(f"NativeFrame: model.{name} must be a contiguous float32 CUDA tensor on {self.dev}")
            if t.grad is None or not t.grad.is_contiguous() or t.grad.device != self.dev:
                raise RuntimeError(f"NativeFrame: model.{name}.grad must be a preallocated contiguous buffer (FlatAdam provides it)")
        if m.faces.device != self.dev:
            raise RuntimeError("NativeFrame: model.faces must live on the model's device")

    @property
    def last_num_rendered(self) -> int:
        """N of the most recent frame whose range kernel has run (no synchronisation: may lag behind the queue)."""
        if self.sync_free and self.capacity > 0:
            self._harvest()
            return self._n_last
        return int(self.n_rendered.value)

    def _harvest(self) -> None:
        """Collect (N, overflow) of the frames the device has finished binning; their slots become reusable."""
        while self._pending:
            slot, key, cap = self._pending[0]
            n = int(self._n_np[slot, 0])
            if n < 0:                       # that frame's k_tile_ranges has not run yet
                break
            self._pending.popleft()
            self._note(key, n)
            if n > cap:
                self.overflows += 1

    def _note(self, key, n: int) -> None:
        prev = self._view_n.get(key)
        self._view_n[key] = (n, prev[0] if prev else 0)
        self._n_max, self._n_last = max(self._n_max, n), n

    def _predict_capacity(self, key) -> int:
        if self.capacity_override is not None:
            return int(self.capacity_override)
        known = self._view_n.get(key)
        if known is None:
            return int(self._n_max * 1.25) + (1 << 18)
        n1, n0 = known
        drift = abs(n1 - n0) / max(n1, 1) if n0 else 0.0
        return int(n1 * (1.0 + max(0.08, 3.0 * drift))) + (1 << 16)

    def read_loss_async(self, loss_host: torch.Tensor, loss_ready: Optional[torch.cuda.Event] = None) -> None:
        """Copy the last frame's loss into pinned host memory from a side stream as soon as the LOSS kernels are done (the
        event gms_train_frame records ~40 % into the frame), not after the whole frame: a caller that waits for `loss_ready`
        gets the value while the backward pass is still running and has that time to queue its next step."""
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        self._side.wait_event(self.ev_loss)
        with torch.cuda.stream(self._side):
            loss_host.copy_(self.loss[:1].reshape(loss_host.shape), non_blocking=True)
            if loss_ready is not None:
                loss_ready.record(self._side)
            self._loss_read = torch.cuda.Event()
            self._loss_read.record(self._side)

    def sh_factors(self) -> dict:
        """What FlatAdam.step(sh=...) needs after a factored frame."""
        import ctypes as C
        from . import _lib
        v = _lib.FrameView()
        _lib.check(_lib.lib().gms_frame_views(self.ws.data_ptr(), self.model._scale.shape[0], self.W, self.H, C.byref(v)), "gms_frame_views")
        return dict(xyz=v.xyz, exchange=self.exchange, degree=self.model.active_sh_degree, event=self.ev_sh)

    def run(self, cam: Camera, gt: torch.Tensor, bg: torch.Tensor, factored: bool = False) -> torch.Tensor:
        import ctypes as C
        from . import _lib
        for t, what in ((gt, "gt"), (bg, "bg"), (cam.world_view_transform, "camera matrices"), (cam.full_proj_transform, "camera matrices"),
                        (cam.camera_center, "camera centre")):
            if not t.is_cuda or t.device != self.dev or t.dtype != torch.float32:
                raise RuntimeError(f"NativeFrame.run: {what} must be float32 on {self.dev}")
        if int(cam.image_width) != self.W or int(cam.image_height) != self.H or tuple(gt.shape[-2:]) != (self.H, self.W):
            raise ValueError(f"NativeFrame was sized for {self.W}x{self.H}; got a {cam.image_width}x{cam.image_height} camera")
        if not gt.is_contiguous() or gt.dtype != torch.float32:
            gt = gt.contiguous().float()
        m = self.model
        a = _lib.FrameArgs()
        a.V, a.F, a.K, a.M = m.vertices.shape[0], m._alpha.shape[0], m._alpha.shape[1], m._features.shape[1]
        a.vertices, a.faces, a.alpha_raw, a.scale_raw = m.vertices.data_ptr(), m.faces.data_ptr(), m._alpha.data_ptr(), m._scale.data_ptr()
        a.features, a.opacity_raw, a.eps = m._features.data_ptr(), m._opacity.data_ptr(), m.eps_s0
        a.d_vertices, a.d_alpha_raw, a.d_scale_raw = m.vertices.grad.data_ptr(), m._alpha.grad.data_ptr(), m._scale.grad.data_ptr()
        a.d_features, a.d_opacity_raw = m._features.grad.data_ptr(), m._opacity.grad.data_ptr()
        if factored:        # no SH gradient rows: the colour gradient + camera centre go to this rank's exchange slot
            if self.exchange is None:
                self.exchange = torch.zeros(self.world, self.slot, dtype=torch.float32, device=self.dev)
            a.d_features, a.d_color_sh = None, self.exchange[self.rank].data_ptr()
            if self.world > 1:
                if self.ev_sh is None:
                    self.ev_sh = torch.cuda.Event()
                    self.ev_sh.record(torch.cuda.current_stream(self.dev))      # (creates the underlying cudaEvent_t)
                a.event_sh_ready = self.ev_sh.cuda_event
        if self.ev_loss is None:
            self.ev_loss = torch.cuda.Event()
            self.ev_loss.record(torch.cuda.current_stream(self.dev))            # (creates the underlying cudaEvent_t)
        a.event_loss_ready = self.ev_loss.cuda_event
        if self._loss_read is not None:         # an asynchronous read-back of the previous frame's loss (read_loss_async) must be
            torch.cuda.current_stream(self.dev).wait_event(self._loss_read)     # done before this frame's loss kernels overwrite it
            self._loss_read = None
        s = a.settings
        s.image_height, s.image_width, s.tanfovx, s.tanfovy = self.H, self.W, cam.tanfovx, cam.tanfovy
        s.bg, s.scale_modifier = bg.data_ptr(), 1.0
        s.viewmatrix, s.projmatrix, s.campos = cam.world_view_transform.data_ptr(), cam.full_proj_transform.data_ptr(), cam.camera_center.data_ptr()
        s.sh_degree, s.prefiltered, s.debug, s.antialiasing = m.active_sh_degree, 0, 0, 0
        a.gt, a.lambda_dssim, a.loss = gt.data_ptr(), self.lam, self.loss.data_ptr()
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        a.num_rendered = C.pointer(self.n_rendered)
        key = getattr(cam, "uid", None)
        if key is None:
            key = id(cam)
        first = self.capacity == 0
        if self.sync_free and not first:
            self._harvest()
            if len(self._pending) >= self.RING - 1:         # the host is a whole ring ahead of the device: wait for the oldest frames
                torch.cuda.current_stream(self.dev).synchronize()
                self._harvest()
            slot = self._frame_no % self.RING
            self._frame_no += 1
            self.capacity = self._predict_capacity(key)
            self._n_np[slot, 0], self._n_np[slot, 1] = -1, 0
            self._pending.append((slot, key, self.capacity))
            a.binning_capacity, a.n_host_mapped = self.capacity, self.n_host.data_ptr() + 8 * slot
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().gms_train_frame(C.byref(a), self._cb, None, torch.cuda.current_stream(self.dev).cuda_stream),
                       "gms_train_frame")
        if self.sync_free and first:           # the one synchronising frame told us N
            n = max(int(self.n_rendered.value), 0)
            self._note(key, n)
            self.capacity = n
        return self.loss[0]


class MeshTrainer:
    """fwd + loss + bwd (+ gradient all-reduce) + Adam for one frame per rank.

    fast=True  : fused expansion launch, packed SH features (zero-copy get_features), fused L1+SSIM loss kernels,
                 FlatAdam (one launch, zeroes the gradient) -- everything on the library's kernels except sigmoid.
    fast=False : the reference's op sequence (two-step expansion + getters, `loss_fn`, torch.optim.Adam) for A/B runs."""

    def __init__(self, model: MeshGaussianModel, bg: torch.Tensor, lambda_dssim: float = 0.2, world: int = 1,
                 rank: int = 0, optimizer_step: bool = True, fast: bool = True, native: bool = False, sync_free: bool = True,
                 loss_fn=None, sh_factored: bool = True):
        self.model, self.bg, self.lambda_dssim = model, bg, lambda_dssim
        self.world, self.rank = world, rank
        self.optimizer_step = optimizer_step
        self.fast = fast
        self.native = native and fast       # native: the whole frame is one C call (NativeFrame), no autograd
        self.sync_free = sync_free          # native frames after the first never synchronise with the host (NativeFrame)
        self.loss_fn = loss_fn or fused_training_loss    # fast=False A/B arm: callers may pass an ATen loss (tests/aten_reference.py)
        self._frame = None
        self.sh_factored = False
        if fast:
            # native frames hand the SH gradient over as factors (12 B instead of 192 B per Gaussian; replicated optimizer);
            # the autograd-driven path materialises it (dense mode: sharded optimizer when world > 1)
            self.sh_factored = bool(sh_factored) and self.native and model._features is not None and model._features.shape[1] == 16
            self.opt = FlatAdam(mesh_model_groups(model, features_last=self.sh_factored), world=world, rank=rank, sh_factored=self.sh_factored)
            self.flat_grad = self.opt.flat_grad
        else:
            self.opt = model.training_setup()
            self.flat_grad = None

    def _all_reduce(self):
        if self.world <= 1:
            return
        if self.fast and self.optimizer_step:
            return      # FlatAdam.step() does reduce-scatter / sharded update / all-gather itself
        if self.flat_grad is not None:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.mul_(1.0 / self.world)
        else:
            for p in self.model.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                    p.grad.mul_(1.0 / self.world)

    def step(self, cam: Camera, gt: torch.Tensor, loss_host: Optional[torch.Tensor] = None,
             loss_ready: Optional[torch.cuda.Event] = None) -> torch.Tensor:
        """One optimisation step.  If `loss_host` (pinned) / `loss_ready` are given, the loss is copied to the host from a side
        stream as soon as the loss kernels have run (native frames: NativeFrame.read_loss_async; otherwise right after the
        backward pass, before the optimizer kernels are queued): a caller that logs the loss every step gets it while the
        backward pass is still running and queues its next step in that time."""
        if self.native:
            if self._frame is None:
                self._frame = NativeFrame(self.model, cam.image_width, cam.image_height, self.lambda_dssim, sync_free=self.sync_free,
                                          world=self.world, rank=self.rank)
            factored = self.sh_factored and self.optimizer_step      # without an optimizer step the full gradient is materialised
            loss = self._frame.run(cam, gt, self.bg, factored=factored)
            from . import rasterizer as _r
            _r.last_num_rendered = self._frame.last_num_rendered
            if loss_host is not None:
                self._frame.read_loss_async(loss_host, loss_ready)
            self._all_reduce()
            # gms_train_frame overwrites every gradient except the atomically accumulated vertex segment (group 0)
            if self.optimizer_step:
                self.opt.step(zero_end=self.opt.ends[0], sh=self._frame.sh_factors() if factored else None)
            else:
                self.opt.zero_grad_partial(self.opt.ends[0])
            return loss
        from . import rasterizer as _r
        prev = _r.DIRECT_SH_GRAD
        _r.DIRECT_SH_GRAD = self.fast      # FlatAdam keeps .grad preallocated and zeroed: write dL/dshs in place
        try:
            image, radii, _ = render_frame(self.model, cam, self.bg, fused=self.fast)
            loss = fused_training_loss(image, gt, self.lambda_dssim) if self.fast else self.loss_fn(image, gt, self.lambda_dssim)
            loss.backward()
        finally:
            _r.DIRECT_SH_GRAD = prev       # never leak the in-place mode to other users of the rasterizer
        if loss_host is not None:
            loss_host.copy_(loss.detach().reshape(loss_host.shape), non_blocking=True)
            if loss_ready is not None:
                loss_ready.record(torch.cuda.current_stream(loss.device))
        self._all_reduce()
        if self.optimizer_step:
            self.opt.step()
            if not self.fast:
                self.opt.zero_grad(set_to_none=False)
        elif self.fast:
            self.opt.zero_grad()
        else:
            for p in self.model.parameters():
                p.grad = None
        return loss.detach()

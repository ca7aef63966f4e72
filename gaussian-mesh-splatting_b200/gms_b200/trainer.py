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

import math
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

import diff_gaussian_rasterization as dgr

from .losses import training_loss
from .model import MeshGaussianModel
from .scenes import Camera


def shard_cameras(n_cameras: int, step: int, rank: int, world: int) -> int:
    """Camera index rank `rank` renders at `step`: consecutive cameras of the schedule go to consecutive ranks."""
    return (step * world + rank) % n_cameras


class FlatGrads:
    """Points every parameter's .grad into one flat fp32 buffer (single-collective gradient exchange)."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, world: int, group=None):
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)


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


class MeshTrainer:
    def __init__(self, model: MeshGaussianModel, bg: torch.Tensor, lambda_dssim: float = 0.2, world: int = 1,
                 rank: int = 0, optimizer_step: bool = True, fused_expansion: bool = True):
        self.model, self.bg, self.lambda_dssim = model, bg, lambda_dssim
        self.world, self.rank = world, rank
        self.optimizer_step = optimizer_step
        self.fused = fused_expansion
        self.opt = model.training_setup() if optimizer_step else None
        self.grads = FlatGrads(model.parameters())

    def step(self, cam: Camera, gt: torch.Tensor) -> torch.Tensor:
        """fwd + loss + bwd (+ gradient all-reduce) (+ Adam).  Returns the (device) loss scalar."""
        self.grads.zero_()
        image, radii, _ = render_frame(self.model, cam, self.bg, self.fused)
        loss = training_loss(image, gt, self.lambda_dssim)
        loss.backward()
        self.grads.all_reduce_mean(self.world)
        if self.opt is not None:
            self.opt.step()
        return loss.detach()

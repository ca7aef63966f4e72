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

from .losses import training_loss, fused_training_loss
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


class MeshTrainer:
    """fwd + loss + bwd (+ gradient all-reduce) + Adam for one frame per rank.

    fast=True  : fused expansion launch, packed SH features (zero-copy get_features), fused L1+SSIM loss kernels,
                 FlatAdam (one launch, zeroes the gradient) -- everything on the library's kernels except sigmoid.
    fast=False : the reference's op sequence (two-step expansion + getters, ATen loss, torch.optim.Adam)."""

    def __init__(self, model: MeshGaussianModel, bg: torch.Tensor, lambda_dssim: float = 0.2, world: int = 1,
                 rank: int = 0, optimizer_step: bool = True, fast: bool = True):
        self.model, self.bg, self.lambda_dssim = model, bg, lambda_dssim
        self.world, self.rank = world, rank
        self.optimizer_step = optimizer_step
        self.fast = fast
        if fast:
            self.opt = FlatAdam(mesh_model_groups(model), world=world, rank=rank)   # sharded over the ranks when world > 1
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

    def step(self, cam: Camera, gt: torch.Tensor) -> torch.Tensor:
        from . import rasterizer as _r
        _r.DIRECT_SH_GRAD = self.fast      # FlatAdam keeps .grad preallocated and zeroed: write dL/dshs in place
        image, radii, _ = render_frame(self.model, cam, self.bg, fused=self.fast)
        loss = fused_training_loss(image, gt, self.lambda_dssim) if self.fast else training_loss(image, gt, self.lambda_dssim)
        loss.backward()
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

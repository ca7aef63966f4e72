"""Shared builders for the parity tests (oracle side)."""
import numpy as np
import torch

from gms_b200 import scenes
from oracle import raster


def settings_from_camera(cam, sh_degree=3, bg=(0.0, 0.0, 0.0), scale_modifier=1.0, antialiasing=False):
    return raster.Settings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                           np.asarray(bg, np.float32), scale_modifier, cam.world_view_transform.numpy(),
                           cam.full_proj_transform.numpy(), sh_degree, cam.camera_center.numpy(),
                           False, False, antialiasing)


def random_gaussians(P, seed=0, extent=1.0, scale_mu=-2.5, flat_frac=0.3):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    scales = torch.exp(scale_mu + 0.5 * torch.randn(P, 3, generator=g))
    nflat = int(P * flat_frac)
    scales[:nflat, 0] = 2e-8
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(1.0 + 1.5 * torch.randn(P, 1, generator=g))
    dc = (torch.rand(P, 1, 3, generator=g) - 0.5) / scenes.SH_C0
    rest = 0.1 * torch.randn(P, 15, 3, generator=g)
    return dict(means3D=xyz, scales=scales, rotations=q, opacities=opac,
                shs=torch.cat([dc, rest], 1).contiguous())


def mesh_scene(level=3, K=3, seed=0, trained_like=True):
    v, f = scenes.icosphere(level)
    return scenes.init_mesh_gaussians(v, f, K, seed=seed, trained_like=trained_like)

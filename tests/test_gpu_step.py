"""-m gpu: the training-step glue kernels (fused L1+SSIM loss, FlatAdam) against their PyTorch fp32 references, and the
fast trainer against the reference-ordered op sequence."""
import os

import numpy as np
import pytest
import torch

from gms_b200 import losses, scenes
from gms_b200.model import MeshGaussianModel
from gms_b200.optim import FlatAdam, mesh_model_groups, REFERENCE_LRS
from gms_b200.trainer import MeshTrainer, render_frame

pytestmark = pytest.mark.gpu


def test_fused_loss_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    a = torch.tensor(g["img"], device="cuda", requires_grad=True); b = torch.tensor(g["gt"], device="cuda")
    loss = losses.fused_training_loss(a, b, 0.2)
    assert abs(loss.item() - float(g["loss"])) < 2e-6
    loss.backward()
    ref = g["grad"]
    assert np.abs(a.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()


@pytest.mark.parametrize("H,W", [(1080, 1920), (201, 333), (32, 32), (7, 5)])
def test_fused_loss_vs_aten_reference(H, W):
    gen = torch.Generator().manual_seed(H)
    a = torch.rand(3, H, W, generator=gen).cuda().requires_grad_(True)
    b = (a.detach().cpu() + 0.1 * torch.randn(3, H, W, generator=gen)).clamp(0, 1).cuda()
    a2 = a.detach().clone().requires_grad_(True)
    l1 = losses.fused_training_loss(a, b, 0.2); l2 = losses.training_loss(a2, b, 0.2)
    assert abs(l1.item() - l2.item()) <= 2e-6 * max(1.0, abs(l2.item()))
    (l1 * 3.0).backward(); (l2 * 3.0).backward()
    ref = a2.grad
    assert (a.grad - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_flat_adam_matches_torch_adam():
    p = scenes.init_mesh_gaussians(*scenes.icosphere(2), K=3, seed=3)
    m1 = MeshGaussianModel.from_params(p, "cuda", packed_features=True)
    m2 = MeshGaussianModel.from_params(p, "cuda", packed_features=False)
    lrs = dict(REFERENCE_LRS); lrs["vertices"] = 1e-4
    opt1 = FlatAdam(mesh_model_groups(m1, lrs))
    opt2 = torch.optim.Adam([{"params": [m2.vertices], "lr": lrs["vertices"]}, {"params": [m2._alpha], "lr": lrs["alpha"]},
                             {"params": [m2._features_dc], "lr": lrs["f_dc"]}, {"params": [m2._features_rest], "lr": lrs["f_rest"]},
                             {"params": [m2._opacity], "lr": lrs["opacity"]}, {"params": [m2._scale], "lr": lrs["scaling"]}],
                            lr=0.0, eps=1e-15)
    gen = torch.Generator().manual_seed(0)
    for it in range(5):
        gs = {k: torch.randn(getattr(m2, k).shape, generator=gen).cuda() * (10.0 ** (it - 2))
              for k in ("vertices", "_alpha", "_features_dc", "_features_rest", "_opacity", "_scale")}
        for k, v in gs.items():
            getattr(m2, k).grad = v.clone()
        m1.vertices.grad.copy_(gs["vertices"]); m1._alpha.grad.copy_(gs["_alpha"]); m1._opacity.grad.copy_(gs["_opacity"])
        m1._scale.grad.copy_(gs["_scale"]); m1._features.grad.copy_(torch.cat((gs["_features_dc"], gs["_features_rest"]), 1))
        opt1.step(); opt2.step()
        assert opt1.flat_grad.abs().max().item() == 0.0     # consumed + zeroed in the same pass
    for k in ("vertices", "_alpha", "_opacity", "_scale"):
        np.testing.assert_allclose(getattr(m1, k).detach().cpu().numpy(), getattr(m2, k).detach().cpu().numpy(), rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(m1._features.detach().cpu().numpy(),
                               torch.cat((m2._features_dc, m2._features_rest), 1).detach().cpu().numpy(), rtol=2e-5, atol=2e-7)


def test_fast_trainer_equals_reference_ordered_step():
    """One optimisation step: fused expansion + packed SH + fused loss + FlatAdam  ==  two-step expansion + getters +
    ATen loss + torch.optim.Adam (the reference's op sequence, train.py:89-157) on the same rasterizer."""
    p = scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=5)
    cam = scenes.look_at_camera((2.4, 0.5, 0.9), (0, 0, 0), 320, 240).to("cuda")
    bg = torch.ones(3, device="cuda")
    gt_model = MeshGaussianModel.from_params(scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=77), "cuda")
    with torch.no_grad():
        gt = render_frame(gt_model, cam, bg)[0].clamp(0, 1).contiguous()
    ma = MeshGaussianModel.from_params(p, "cuda", packed_features=True)
    mb = MeshGaussianModel.from_params(p, "cuda", packed_features=False)
    ta = MeshTrainer(ma, bg, fast=True); tb = MeshTrainer(mb, bg, fast=False)
    la = [ta.step(cam, gt).item() for _ in range(3)]
    lb = [tb.step(cam, gt).item() for _ in range(3)]
    np.testing.assert_allclose(la, lb, rtol=2e-4)
    assert la[2] < la[0]
    np.testing.assert_allclose(ma._opacity.detach().cpu().numpy(), mb._opacity.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(ma._features_dc.detach().cpu().numpy(), mb._features_dc.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)


def test_native_frame_equals_autograd_path():
    """gms_train_frame (one C call for the whole frame) == the autograd fast path: same loss, same gradients, same
    parameters after a few optimizer steps."""
    p = scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=6)
    cams = [c.to("cuda") for c in scenes.ring_cameras(3, 2.5, 352, 256)]
    bg = torch.ones(3, device="cuda")
    gt_model = MeshGaussianModel.from_params(scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=78), "cuda")
    with torch.no_grad():
        gts = [render_frame(gt_model, c, bg)[0].clamp(0, 1).contiguous() for c in cams]
    ma = MeshGaussianModel.from_params(p, "cuda", packed_features=True)
    mb = MeshGaussianModel.from_params(p, "cuda", packed_features=True)
    ta = MeshTrainer(ma, bg, fast=True, native=True, optimizer_step=False)
    tb = MeshTrainer(mb, bg, fast=True, native=False, optimizer_step=False)
    # one frame, gradients compared before they are cleared
    from gms_b200.trainer import NativeFrame
    fr = NativeFrame(ma, 352, 256)
    la = fr.run(cams[0], gts[0], bg).item()
    ga = ta.opt.flat_grad.clone(); ta.opt.zero_grad()
    from gms_b200 import rasterizer
    rasterizer.DIRECT_SH_GRAD = True
    image, _, _ = render_frame(mb, cams[0], bg)
    lb = losses.fused_training_loss(image, gts[0], 0.2); lb.backward()
    gb = tb.opt.flat_grad.clone(); tb.opt.zero_grad()
    assert abs(la - lb.item()) <= 1e-6 * max(1.0, abs(lb.item()))
    assert (ga - gb).abs().max().item() <= 2e-3 * gb.abs().max().item()
    # a few full steps
    ta = MeshTrainer(ma, bg, fast=True, native=True); tb = MeshTrainer(mb, bg, fast=True, native=False)
    for s in range(4):
        l1 = ta.step(cams[s % 3], gts[s % 3]).item(); l2 = tb.step(cams[s % 3], gts[s % 3]).item()
        assert abs(l1 - l2) <= 2e-4 * abs(l2)
    np.testing.assert_allclose(ma._opacity.detach().cpu().numpy(), mb._opacity.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(ma._features.detach().cpu().numpy(), mb._features.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)

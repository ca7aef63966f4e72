"""-m gpu: the training-step glue kernels (fused L1+SSIM loss, FlatAdam) against their PyTorch fp32 references, and the
fast trainer against the reference-ordered op sequence."""
import math
import os

import numpy as np
import pytest
import torch

import aten_reference
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
    l1 = losses.fused_training_loss(a, b, 0.2); l2 = aten_reference.training_loss(a2, b, 0.2)
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


def test_adam_abi_unaligned_segments_offsets_and_partial_zero():
    """gms_adam_step through the raw C ABI: segment ends that are NOT multiples of 4 (per-element path), a shard offset, the
    DC/rest learning-rate phase of a packed segment, and zero_grad mode 2 (only flat indices < zero_end are cleared);
    against a float64 restatement of torch.optim.Adam's update (gaussian_mesh_model.py:183)."""
    import ctypes as C
    from gms_b200 import _lib
    gen = torch.Generator().manual_seed(5)
    n_total = 64 * 37
    ends = [101, 101 + 48 * 19 + 2, n_total]            # packed segment in the middle: inner 3, period 16
    lr0, lr1, inner, period = [1e-2, 3e-3, 5e-2], [1e-2, 2e-4, 5e-2], [1, 3, 1], [0, 16, 0]
    p0 = torch.randn(n_total, generator=gen); g0 = torch.randn(n_total, generator=gen) * 0.1
    m0 = torch.randn(n_total, generator=gen) * 0.01; v0 = torch.rand(n_total, generator=gen) * 1e-3
    idx = torch.arange(n_total)
    lr = torch.full((n_total,), lr0[2], dtype=torch.float64)
    lr[idx < ends[1]] = torch.where(((idx[idx < ends[1]] - ends[0]) // 3) % 16 == 0, lr0[1], lr1[1]).double()
    lr[idx < ends[0]] = lr0[0]
    step, b1, b2, eps = 7, 0.9, 0.999, 1e-15
    m_ref = b1 * m0.double() + (1 - b1) * g0.double()
    v_ref = b2 * v0.double() + (1 - b2) * g0.double() ** 2
    p_ref = p0.double() - lr / (1 - b1 ** step) * m_ref / (v_ref.sqrt() / (1 - b2 ** step) ** 0.5 + eps)
    for off, n, zero_mode, zero_end in ((0, n_total, 2, 150), (64 * 5, 64 * 20, 1, 0), (4, n_total - 4 - 3, 0, 0)):
        p, g, m, v = (t.clone().cuda() for t in (p0, g0, m0, v0))
        a = _lib.AdamArgs()
        a.n, a.offset = n, off
        a.p, a.g, a.m, a.v = (t[off:].data_ptr() for t in (p, g, m, v))
        a.nseg = 3
        for i in range(3):
            a.seg_end[i], a.lr0[i], a.lr1[i], a.inner[i], a.period[i] = ends[i], lr0[i], lr1[i], inner[i], period[i]
        a.beta1, a.beta2, a.eps, a.step, a.zero_grad, a.zero_end = b1, b2, eps, step, zero_mode, zero_end
        _lib.check(_lib.lib().gms_adam_step(C.byref(a), torch.cuda.current_stream().cuda_stream), "gms_adam_step")
        sl = slice(off, off + n)
        np.testing.assert_allclose(p.cpu()[sl].numpy(), p_ref[sl].float().numpy(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(m.cpu()[sl].numpy(), m_ref[sl].float().numpy(), rtol=2e-6, atol=2e-8)   # fp32 ulp of 0.9*m + 0.1*g
        np.testing.assert_allclose(v.cpu()[sl].numpy(), v_ref[sl].float().numpy(), rtol=2e-6, atol=1e-12)
        untouched = torch.ones(n_total, dtype=torch.bool); untouched[sl] = False
        assert torch.equal(p.cpu()[untouched], p0[untouched]) and torch.equal(g.cpu()[untouched], g0[untouched])
        want_g = g0.clone()
        if zero_mode == 1:
            want_g[sl] = 0
        elif zero_mode == 2:
            want_g[off:min(off + n, zero_end)] = 0
        assert torch.equal(g.cpu(), want_g)


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
    ta = MeshTrainer(ma, bg, fast=True); tb = MeshTrainer(mb, bg, fast=False, loss_fn=aten_reference.training_loss)
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
    ga = [q.grad.clone() for q in ma.parameters()]; ta.opt.zero_grad()       # (the two optimisers order their flat buffers differently)
    from gms_b200 import rasterizer
    rasterizer.DIRECT_SH_GRAD = True
    image, _, _ = render_frame(mb, cams[0], bg)
    lb = losses.fused_training_loss(image, gts[0], 0.2); lb.backward()
    gb = [q.grad.clone() for q in mb.parameters()]; tb.opt.zero_grad()
    assert abs(la - lb.item()) <= 1e-6 * max(1.0, abs(lb.item()))
    for x, y in zip(ga, gb):
        assert (x - y).abs().max().item() <= 2e-3 * y.abs().max().item()
    # a few full steps
    ta = MeshTrainer(ma, bg, fast=True, native=True); tb = MeshTrainer(mb, bg, fast=True, native=False)
    for s in range(4):
        l1 = ta.step(cams[s % 3], gts[s % 3]).item(); l2 = tb.step(cams[s % 3], gts[s % 3]).item()
        assert abs(l1 - l2) <= 2e-4 * abs(l2)
    np.testing.assert_allclose(ma._opacity.detach().cpu().numpy(), mb._opacity.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(ma._features.detach().cpu().numpy(), mb._features.detach().cpu().numpy(), rtol=1e-3, atol=2e-3)


def test_sync_free_native_frame_equals_synchronising_frame():
    """NativeFrame(sync_free=True): after the first (synchronising) frame no host sync happens, N is polled from mapped
    pinned memory, and losses / gradients / parameters are bit-identical to the stock-style frame."""
    from gms_b200.trainer import NativeFrame
    p = scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=8)
    cams = [c.to("cuda") for c in scenes.ring_cameras(5, 2.5, 352, 256)]
    bg = torch.ones(3, device="cuda")
    gt_model = MeshGaussianModel.from_params(scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=79), "cuda")
    with torch.no_grad():
        gts = [render_frame(gt_model, c, bg)[0].clamp(0, 1).contiguous() for c in cams]
    res = []
    for sync_free in (True, False):
        m = MeshGaussianModel.from_params(p, "cuda", packed_features=True)
        opt = FlatAdam(mesh_model_groups(m))
        fr = NativeFrame(m, 352, 256, sync_free=sync_free)
        losses_, ns = [], []
        for s in range(7):
            losses_.append(fr.run(cams[s % 5], gts[s % 5], bg).clone())
            torch.cuda.synchronize()
            ns.append(fr.last_num_rendered)
            opt.step(zero_end=opt.ends[0])
        res.append((torch.stack(losses_).cpu(), ns, opt.p.clone().cpu(), fr))
    (la, na, pa, fa), (lb, nb, pb, fb) = res
    assert fa.capacity > 0 and fa.overflows == 0 and fb.capacity == 0
    assert na == nb and min(na) > 0
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=0)        # (the fp32 atomics of the backward make runs differ in the last bits)
    np.testing.assert_allclose(pa.numpy(), pb.numpy(), rtol=1e-3, atol=2e-4)    # Adam's first steps amplify the atomics' last-bit noise
    # forced overflow: capacity below N -> flag, background frame, then automatic growth
    # per-view capacity: the second visit of a camera is sized from its own N, not from the largest N seen
    assert fa.capacity <= int(1.3 * max(na)) + (1 << 16)
    # forced overflow: capacity below N -> background frame with zero gradients, noticed when the slot is harvested; the
    # camera's next visit is sized from the true N
    fa.capacity_override = max(na) // 3
    loss_over = fa.run(cams[0], gts[0], bg).item(); torch.cuda.synchronize()
    assert fa.last_num_rendered > fa.capacity and fa.overflows == 1
    assert all(float(g.abs().max()) == 0.0 for g in (fa.model._alpha.grad, fa.model._scale.grad, fa.model._opacity.grad))
    fa.capacity_override = None
    loss_ok = fa.run(cams[0], gts[0], bg).item(); torch.cuda.synchronize()
    assert fa.overflows == 1 and fa.capacity >= fa.last_num_rendered > 0 and loss_ok < loss_over
    assert abs(fa.last_num_rendered - na[0]) <= 0.05 * na[0]        # (same camera; the parameters moved by seven Adam steps)


def test_adam_sh_factored_abi_survives_denormal_second_moments():
    """gms_adam_sh_factored through the raw C ABI with colour gradients from 1e-25 (their squares underflow to denormals /
    zero in fp32) to 1e-2 and exact zeros, degree 0 (gradient of coefficient 0 = 0.2820948 * colour gradient, the other 15
    rows zero): finite everywhere and equal to torch.optim.Adam (fp32, eps 1e-15 as gaussian_mesh_model.py:183) over three steps."""
    import ctypes as C
    from gms_b200 import _lib
    P, M = 4099, 16
    gen = torch.Generator().manual_seed(5)
    mag = 10.0 ** (torch.rand(P, 3, generator=gen) * 23.0 - 25.0)
    gcol = (mag * torch.sign(torch.randn(P, 3, generator=gen))).float()
    gcol[::7] = 0.0
    xyz = torch.randn(P, 3, generator=gen).float().cuda()
    slot = (3 * P + 3 + 63) // 64 * 64
    ex = torch.zeros(1, slot, device="cuda")
    ex[0, :3 * P] = gcol.reshape(-1).cuda()
    ex[0, 3 * P:3 * P + 3] = torch.tensor([0.3, -2.0, 4.0])
    p0 = torch.randn(P, M, 3, generator=gen).float()
    p = p0.clone().cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    lr_dc, lr_rest, b1, b2, eps = 2.5e-3, 1.25e-4, 0.9, 0.999, 1e-15
    r_dc = p0[:, :1, :].clone().cuda().requires_grad_(True); r_rest = p0[:, 1:, :].clone().cuda().requires_grad_(True)
    ref = torch.optim.Adam([{"params": [r_dc], "lr": lr_dc}, {"params": [r_rest], "lr": lr_rest}], lr=0.0, betas=(b1, b2), eps=eps)
    g_dc = (gcol.cuda() * np.float32(0.28209479177387814)).reshape(P, 1, 3)
    for step in (1, 2, 3):
        a = _lib.AdamShArgs()
        a.P, a.M, a.sh_degree, a.R = P, M, 0, 1
        a.xyz, a.exchange, a.slot_floats, a.grad_scale = xyz.data_ptr(), ex.data_ptr(), slot, 1.0
        a.p, a.m, a.v = p.data_ptr(), m.data_ptr(), v.data_ptr()
        a.lr_dc, a.lr_rest, a.beta1, a.beta2, a.eps, a.step = lr_dc, lr_rest, b1, b2, eps, step
        _lib.check(_lib.lib().gms_adam_sh_factored(C.byref(a), torch.cuda.current_stream().cuda_stream), "gms_adam_sh_factored")
        r_dc.grad = g_dc.clone(); r_rest.grad = torch.zeros_like(r_rest)
        ref.step()
    torch.cuda.synchronize()
    for t_ in (p, m, v):
        assert torch.isfinite(t_).all()
    # the branch-free division / square root are the compiler's own fast-path sequences: bit-identical to sqrtf and `/`
    p2 = p0.clone().cuda(); m2 = torch.zeros_like(p2); v2 = torch.zeros_like(p2)
    old = _lib.set_option("adam_sh_ieee", 1)
    try:
        for step in (1, 2, 3):
            a = _lib.AdamShArgs()
            a.P, a.M, a.sh_degree, a.R = P, M, 0, 1
            a.xyz, a.exchange, a.slot_floats, a.grad_scale = xyz.data_ptr(), ex.data_ptr(), slot, 1.0
            a.p, a.m, a.v = p2.data_ptr(), m2.data_ptr(), v2.data_ptr()
            a.lr_dc, a.lr_rest, a.beta1, a.beta2, a.eps, a.step = lr_dc, lr_rest, b1, b2, eps, step
            _lib.check(_lib.lib().gms_adam_sh_factored(C.byref(a), torch.cuda.current_stream().cuda_stream), "gms_adam_sh_factored")
    finally:
        _lib.set_option("adam_sh_ieee", old)
    torch.cuda.synchronize()
    nd = int((p != p2).sum()), int((m != m2).sum()), int((v != v2).sum())
    print(f"[adam_sh] elements differing from the sqrtf / division build after 3 steps: p {nd[0]}, m {nd[1]}, v {nd[2]} of {p.numel()}; "
          f"max |dp| {float((p - p2).abs().max()):.3e}")
    assert nd[1] == 0 and nd[2] == 0 and float((p - p2).abs().max()) <= 2.4e-7
    err = float((p[:, :1, :] - r_dc.detach()).abs().max())
    assert err <= 5e-7, err                                         # (|p| ~ 1..4: a couple of ulps; the update itself is ~2.5e-3 per step)
    assert torch.equal(p.cpu()[:, 1:, :], p0[:, 1:, :])            # rows above the active degree: zero gradient, untouched
    moved = (p[:, 0, :].cpu() - p0[:, 0, :]).abs()
    big = gcol.abs() > 1e-12
    assert float(moved[big].min()) > 1e-3 and float(moved[gcol == 0].max()) == 0.0


def test_factored_sh_gradient_equals_dense_path():
    """The native trainer hands the SH gradient over as factors (colour gradient x SH basis rebuilt inside k_adam_sh,
    gms_adam_sh_factored) instead of 192 B/Gaussian of rows: same parameters as the dense path (gradient rows written by
    k_preprocess_bwd, read by k_adam) after a few steps, for every active SH degree."""
    for degree in (3, 1, 0):
        p = scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=9)
        cams = [c.to("cuda") for c in scenes.ring_cameras(4, 2.5, 352, 256)]
        bg = torch.ones(3, device="cuda")
        gt_model = MeshGaussianModel.from_params(scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=80), "cuda")
        with torch.no_grad():
            gts = [render_frame(gt_model, c, bg)[0].clamp(0, 1).contiguous() for c in cams]
        out = []
        for factored in (True, False):
            m = MeshGaussianModel.from_params(p, "cuda", packed_features=True, active_sh_degree=degree)
            tr = MeshTrainer(m, bg, fast=True, native=True, sh_factored=factored)
            losses_ = [tr.step(cams[s % 4], gts[s % 4]).item() for s in range(5)]
            assert tr.sh_factored == factored
            out.append((losses_, m._features.detach().clone(), m._opacity.detach().clone(), m._alpha.detach().clone()))
        (la, fa, oa, aa), (lb, fb, ob, ab) = out
        np.testing.assert_allclose(la, lb, rtol=1e-5)
        assert la[-1] < la[0]
        ref = float(fb.abs().max())
        assert float((fa - fb).abs().max()) <= 2e-5 * ref, (degree, float((fa - fb).abs().max()), ref)
        np.testing.assert_allclose(oa.cpu().numpy(), ob.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(aa.cpu().numpy(), ab.cpu().numpy(), rtol=1e-4, atol=1e-5)
        if degree < 3:      # coefficients above the active degree get a zero gradient: untouched by either path
            nc = (degree + 1) ** 2
            init = torch.cat((p._features_dc, p._features_rest), 1).cuda()
            assert torch.equal(fa[:, nc:], init[:, nc:]) and torch.equal(fb[:, nc:], init[:, nc:])

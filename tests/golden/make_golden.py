"""Generate the golden fixtures in this directory by IMPORTING THE REFERENCE's own Python
(/root/reference) -- run in the build container only; the fixtures are committed, the reference is
never read at test time.

    python tests/golden/make_golden.py

What is pinned (everything the reference ships in Python on or beside the hot path):
  expansion_mesh.npz   GaussianMeshModel.update_alpha / prepare_scaling_rot + getters, and the
                       reference's own autograd gradients of a fixed scalar loss
                       (games/mesh_splatting/scene/gaussian_mesh_model.py:86-169, utils/general_utils.py:19-96,
                        scene/gaussian_model.py:95-115)
  expansion_multi.npz  GaussianMultiMeshModel (games/multi_mesh_splatting/scene/gaussian_multi_mesh_model.py:99-199)
  sh_colors.npz        eval_sh + 0.5 clamp (utils/sh_utils.py:57-112; renderer/gaussian_renderer/__init__.py:82-87)
  cov3d.npz            build_scaling_rotation / strip_symmetric (utils/general_utils.py:144-190,
                       scene/gaussian_model.py:27-31)  == --compute_cov3D_python
  points_model.npz     PointsGaussianModel.prepare_vertices / prepare_scaling_rot / get_scaling
                       (games/flat_splatting/scene/points_gaussian_model.py:28-109)
  loss.npz             l1_loss / ssim / 0.8*L1 + 0.2*(1-SSIM) and its autograd gradient (utils/loss_utils.py:17-64, train.py:105-107)
  camera.npz           getWorld2View2 / getProjectionMatrix / Camera matrix algebra
                       (utils/graphics_utils.py:22-71, scene/cameras.py:54-57), geom_transform_points
The reference hard-codes device="cuda" in a few helpers; the generator temporarily maps those
allocations to the CPU (the arithmetic is untouched).
"""
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "gaussian-mesh-splatting_b200"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_stub("plyfile", PlyData=object, PlyElement=object)
_stub("simple_knn")
_stub("simple_knn._C", distCUDA2=None)
_stub("trimesh")
_stub("smplx")
_stub("smplx.lbs", lbs=None, batch_rodrigues=None, vertices2landmarks=None, find_dynamic_lmk_idx_and_bcoords=None)
_stub("smplx.utils", Struct=object, to_tensor=None, to_np=None, rot_mat_to_euler=None)
_stub("diff_gaussian_rasterization", GaussianRasterizationSettings=object, GaussianRasterizer=object)

# device="cuda" -> cpu for allocation helpers used by the reference utilities
_orig_zeros = torch.zeros


def _zeros_cpu(*a, **k):
    if k.get("device", None) in ("cuda", torch.device("cuda")):
        k["device"] = "cpu"
    return _orig_zeros(*a, **k)


torch.zeros = _zeros_cpu

from games.mesh_splatting.scene.gaussian_mesh_model import GaussianMeshModel  # noqa: E402
from games.multi_mesh_splatting.scene.gaussian_multi_mesh_model import GaussianMultiMeshModel  # noqa: E402
from utils.sh_utils import eval_sh  # noqa: E402
from utils.general_utils import build_scaling_rotation, strip_symmetric, rot_to_quat_batch  # noqa: E402
from utils.graphics_utils import getWorld2View2, getProjectionMatrix, geom_transform_points  # noqa: E402

from utils.loss_utils import l1_loss as ref_l1, ssim as ref_ssim  # noqa: E402

from gms_b200 import scenes  # noqa: E402


def expansion_mesh():
    torch.manual_seed(0)
    verts, faces = scenes.icosphere(1)
    # perturb so faces are not regular; include one degenerate-ish sliver and a zero/negative alpha entry
    verts = verts + 0.05 * np.random.RandomState(1).randn(*verts.shape).astype(np.float32)
    F, K = faces.shape[0], 3
    m = GaussianMeshModel(3)
    m.vertices = torch.nn.Parameter(torch.tensor(verts))
    m.faces = torch.tensor(faces)
    a = torch.rand(F, K, 3)
    a[0, 0, 0] = -0.3
    a[1, 1, :] = 0.0
    m._alpha = torch.nn.Parameter(a)
    sc = 0.5 + torch.rand(F * K, 1)
    sc[5, 0] = -0.2   # relu branch
    m._scale = torch.nn.Parameter(sc)
    P = F * K
    m._opacity = torch.nn.Parameter(torch.randn(P, 1))
    m._features_dc = torch.nn.Parameter(torch.randn(P, 1, 3))
    m._features_rest = torch.nn.Parameter(torch.randn(P, 15, 3))
    m.update_alpha()
    m.prepare_scaling_rot()
    xyz, scaling, rotation = m.get_xyz, m.get_scaling, m.get_rotation
    opacity, feats = m.get_opacity, m.get_features
    g = torch.Generator().manual_seed(7)
    wx, ws, wr = torch.randn(P, 3, generator=g), torch.randn(P, 3, generator=g), torch.randn(P, 4, generator=g)
    # loss on log-scaling (not exp) keeps s0-column gradients finite and comparable
    loss = (xyz * wx).sum() + (m._scaling * ws).sum() + (rotation * wr).sum()
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "expansion_mesh.npz"),
                        vertices=verts, faces=faces, _alpha=a.numpy(), _scale=sc.numpy(),
                        _opacity=m._opacity.detach().numpy(), _features_dc=m._features_dc.detach().numpy(),
                        _features_rest=m._features_rest.detach().numpy(),
                        alpha=m.alpha.detach().numpy(), triangles=m.triangles.detach().numpy(),
                        xyz=xyz.detach().numpy(), _scaling=m._scaling.detach().numpy(),
                        _rotation=m._rotation.detach().numpy(), get_scaling=scaling.detach().numpy(),
                        get_rotation=rotation.detach().numpy(), get_opacity=opacity.detach().numpy(),
                        get_features=feats.detach().numpy(), wx=wx.numpy(), ws=ws.numpy(), wr=wr.numpy(),
                        g_vertices=m.vertices.grad.numpy(), g_alpha=m._alpha.grad.numpy(),
                        g_scale=m._scale.grad.numpy())


def expansion_multi():
    torch.manual_seed(1)
    m = GaussianMultiMeshModel(3)
    vs, fs, als, scs = [], [], [], []
    for k, (lvl, K) in enumerate([(0, 2), (1, 3)]):
        v, f = scenes.icosphere(lvl, radius=0.5 + 0.3 * k)
        v = v + np.float32([1.5 * k, 0, 0])
        vs.append(torch.nn.Parameter(torch.tensor(v)))
        fs.append(torch.tensor(f))
        als.append(torch.nn.Parameter(torch.rand(f.shape[0], K, 3)))
        scs.append(torch.nn.Parameter(0.5 + torch.rand(f.shape[0] * K, 1)))
    m.vertices, m.faces, m._alpha, m._scale = vs, fs, als, scs
    m.update_alpha()
    m.prepare_scaling_rot()
    out = dict(xyz=m.get_xyz.detach().numpy(), _scaling=m._scaling.detach().numpy(),
               _rotation=m._rotation.detach().numpy(), n_mesh=np.int64(2))
    for k in range(2):
        out[f"vertices{k}"] = vs[k].detach().numpy(); out[f"faces{k}"] = fs[k].numpy()
        out[f"_alpha{k}"] = als[k].detach().numpy(); out[f"_scale{k}"] = scs[k].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "expansion_multi.npz"), **out)


def sh_colors():
    g = torch.Generator().manual_seed(3)
    P = 64
    shs = torch.randn(P, 16, 3, generator=g) * 0.4
    xyz = torch.randn(P, 3, generator=g)
    campos = torch.tensor([0.3, -2.0, 1.1])
    out = dict(shs=shs.numpy(), xyz=xyz.numpy(), campos=campos.numpy())
    for deg in range(4):
        shs_view = shs.transpose(1, 2).view(-1, 3, 16)
        d = xyz - campos.repeat(P, 1)
        d = d / d.norm(dim=1, keepdim=True)
        out[f"rgb_deg{deg}"] = torch.clamp_min(eval_sh(deg, shs_view, d) + 0.5, 0.0).numpy()
    np.savez_compressed(os.path.join(HERE, "sh_colors.npz"), **out)


def cov3d():
    g = torch.Generator().manual_seed(4)
    P = 64
    s = torch.exp(torch.randn(P, 3, generator=g) - 2)
    s[:8, 0] = 2e-8  # flat mesh Gaussians
    q = torch.randn(P, 4, generator=g)       # un-normalised on purpose (build_rotation normalises)
    mod = 1.3
    L = build_scaling_rotation(mod * s, q)
    cov = strip_symmetric(L @ L.transpose(1, 2))
    qn = q / q.norm(dim=1, keepdim=True)
    np.savez_compressed(os.path.join(HERE, "cov3d.npz"), scales=s.numpy(), rotations_raw=q.numpy(),
                        rotations_unit=qn.numpy(), scale_modifier=np.float32(mod), cov3D=cov.numpy())


def camera():
    out = {}
    eye = np.array([2.5, -1.7, 1.9])
    cam = scenes.look_at_camera(eye, (0.1, 0.0, -0.2), 200, 120)
    # recover the (R, T) the reference's Camera would be built from and run ITS functions
    w2c = cam.world_view_transform.t().numpy()
    R = w2c[:3, :3].T
    T = w2c[:3, 3]
    wvt = torch.tensor(getWorld2View2(R, T)).transpose(0, 1)
    proj = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=cam.FoVx, fovY=cam.FoVy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = wvt.inverse()[3, :3]
    pts = torch.randn(32, 3, generator=torch.Generator().manual_seed(5))
    ndc = geom_transform_points(pts, full)
    out.update(eye=eye, target=np.array([0.1, 0.0, -0.2]), width=np.int64(200), height=np.int64(120),
               fovx=np.float64(cam.FoVx), fovy=np.float64(cam.FoVy), world_view_transform=wvt.numpy(),
               full_proj_transform=full.numpy(), camera_center=center.numpy(), points=pts.numpy(),
               ndc=ndc.numpy())
    np.savez_compressed(os.path.join(HERE, "camera.npz"), **out)


def quat():
    g = torch.Generator().manual_seed(6)
    q = torch.randn(256, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    out = rot_to_quat_batch(R)
    np.savez_compressed(os.path.join(HERE, "rot_to_quat.npz"), R=R.numpy(), quat=out.numpy())


def points():
    """PointsGaussianModel.prepare_scaling_rot + get_scaling (games/flat_splatting/scene/points_gaussian_model.py:61-109)."""
    from games.flat_splatting.scene.points_gaussian_model import PointsGaussianModel
    g = torch.Generator().manual_seed(11)
    P = 200
    tri = torch.randn(P, 3, 3, generator=g)
    tri[:, 1] = tri[:, 0] + 0.2 * torch.randn(P, 3, generator=g)
    tri[:, 2] = tri[:, 0] + 0.2 * torch.randn(P, 3, generator=g)
    m = PointsGaussianModel(3)
    _ones = torch.ones
    torch.ones = lambda *a, **k: _ones(*a, **{kk: vv for kk, vv in k.items()})
    m.prepare_scaling_rot(tri)
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # get_scaling builds s0 with .cuda()
    try:
        gs = m.get_scaling
    finally:
        torch.Tensor.cuda = _cuda
        torch.ones = _ones
    # prepare_vertices (:28-59) on an independent flat-Gaussian state: raw (unnormalised) quaternions, both orders of the
    # two in-plane scales, one tie (mask = s_2 > s_3 is False on equality)
    v = PointsGaussianModel(3)
    v._xyz = torch.randn(P, 3, generator=g)
    sl = -2.0 + 0.7 * torch.randn(P, 2, generator=g)
    sl[3, 1] = sl[3, 0]
    v._scaling = sl
    v._rotation = torch.randn(P, 4, generator=g) * (0.5 + torch.rand(P, 1, generator=g))
    torch.ones = lambda *a, **k: _ones(*a, **{kk: vv for kk, vv in k.items()})
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        v.prepare_vertices()
    finally:
        torch.Tensor.cuda = _cuda
        torch.ones = _ones
    np.savez_compressed(os.path.join(HERE, "points_model.npz"), triangles=tri.numpy(), _scaling=m._scaling.numpy(),
                        _rotation=m._rotation.numpy(), get_scaling=gs.numpy(),
                        get_rotation=torch.nn.functional.normalize(m._rotation).numpy(),
                        pv_xyz=v._xyz.numpy(), pv_scaling=v._scaling.numpy(), pv_rotation=v._rotation.numpy(),
                        pv_triangles=v.triangles.numpy())


def loss():
    g = torch.Generator().manual_seed(8)
    a = torch.rand(3, 45, 70, generator=g, requires_grad=True)
    b = (a.detach() + 0.15 * torch.randn(3, 45, 70, generator=g)).clamp(0, 1)
    l1 = ref_l1(a, b); ss = ref_ssim(a, b)
    total = (1.0 - 0.2) * l1 + 0.2 * (1.0 - ss)     # train.py:106-107, lambda_dssim = 0.2
    total.backward()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), img=a.detach().numpy(), gt=b.numpy(), l1=np.float64(l1.item()),
                        ssim=np.float64(ss.item()), loss=np.float64(total.item()), grad=a.grad.numpy())


if __name__ == "__main__":
    expansion_mesh(); expansion_multi(); sh_colors(); cov3d(); camera(); quat(); loss(); points()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))

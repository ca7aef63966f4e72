"""Pin the oracles against golden vectors produced by the REFERENCE's own Python
(tests/golden/make_golden.py imported /root/reference to make them)."""
import os

import numpy as np
import torch

from gms_b200 import scenes
from oracle import expansion, raster
from helpers import settings_from_camera


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_expansion_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "expansion_mesh.npz")
    v = torch.tensor(g["vertices"]); f = torch.tensor(g["faces"])
    xyz, _scaling, _rotation, alpha, tri = expansion.expand(v, f, torch.tensor(g["_alpha"]), torch.tensor(g["_scale"]))
    np.testing.assert_allclose(alpha.numpy(), g["alpha"], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(tri.numpy(), g["triangles"])
    np.testing.assert_allclose(xyz.numpy(), g["xyz"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(_scaling.numpy(), g["_scaling"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_rotation.numpy(), g["_rotation"], rtol=0, atol=1e-6)
    sc, rot, op, feat = expansion.activate(_scaling, _rotation, torch.tensor(g["_opacity"]),
                                           torch.tensor(g["_features_dc"]), torch.tensor(g["_features_rest"]))
    np.testing.assert_allclose(sc.numpy(), g["get_scaling"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(rot.numpy(), g["get_rotation"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(op.numpy(), g["get_opacity"], rtol=1e-6, atol=0)
    np.testing.assert_array_equal(feat.numpy(), g["get_features"])


def test_expansion_backward_matches_reference_autograd(golden_dir):
    g = _load(golden_dir, "expansion_mesh.npz")
    v = torch.tensor(g["vertices"], requires_grad=True)
    a = torch.tensor(g["_alpha"], requires_grad=True)
    s = torch.tensor(g["_scale"], requires_grad=True)
    xyz, _scaling, _rotation, _, _ = expansion.expand(v, torch.tensor(g["faces"]), a, s)
    rot = torch.nn.functional.normalize(_rotation)
    loss = (xyz * torch.tensor(g["wx"])).sum() + (_scaling * torch.tensor(g["ws"])).sum() + (rot * torch.tensor(g["wr"])).sum()
    loss.backward()
    np.testing.assert_allclose(v.grad.numpy(), g["g_vertices"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(a.grad.numpy(), g["g_alpha"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s.grad.numpy(), g["g_scale"], rtol=1e-4, atol=1e-5)


def test_expansion_multi_mesh_matches_reference(golden_dir):
    g = _load(golden_dir, "expansion_multi.npz")
    n = int(g["n_mesh"])
    xyz, sc, rot = expansion.expand_multi([torch.tensor(g[f"vertices{k}"]) for k in range(n)],
                                          [torch.tensor(g[f"faces{k}"]) for k in range(n)],
                                          [torch.tensor(g[f"_alpha{k}"]) for k in range(n)],
                                          [torch.tensor(g[f"_scale{k}"]) for k in range(n)])
    np.testing.assert_allclose(xyz.numpy(), g["xyz"], atol=1e-6)
    np.testing.assert_allclose(sc.numpy(), g["_scaling"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rot.numpy(), g["_rotation"], atol=1e-6)


def test_rot_to_quat_matches_reference(golden_dir):
    g = _load(golden_dir, "rot_to_quat.npz")
    out = expansion.rot_to_quat(torch.tensor(g["R"]))
    np.testing.assert_allclose(out.numpy(), g["quat"], atol=1e-6)


def test_camera_matrices_match_reference(golden_dir):
    g = _load(golden_dir, "camera.npz")
    cam = scenes.look_at_camera(g["eye"], g["target"], int(g["width"]), int(g["height"]))
    np.testing.assert_allclose(cam.world_view_transform.numpy(), g["world_view_transform"], atol=1e-6)
    np.testing.assert_allclose(cam.full_proj_transform.numpy(), g["full_proj_transform"], atol=1e-5)
    np.testing.assert_allclose(cam.camera_center.numpy(), g["camera_center"], atol=1e-5)
    assert abs(cam.FoVy - float(g["fovy"])) < 1e-12


def _all_visible_state(means3D, shs=None, colors=None, scales=None, rots=None, mod=1.0, deg=3, campos=None):
    """Camera far back on -z looking at +z so every point is in front and on screen."""
    cam = scenes.look_at_camera((0.0, 0.0, -12.0), (0.0, 0.0, 0.0), 160, 160, up=(0.0, 1.0, 0.0))
    S = settings_from_camera(cam, sh_degree=deg, scale_modifier=mod)
    if campos is not None:
        S.campos = campos
    P = means3D.shape[0]
    if scales is None:
        scales = np.full((P, 3), 0.05, np.float32)
        rots = np.tile(np.float32([1, 0, 0, 0]), (P, 1))
    return raster.preprocess(S, means3D, np.full((P, 1), 0.5, np.float32), shs=shs, colors_precomp=colors,
                             scales=scales, rotations=rots), S


def test_sh_colors_match_reference_eval_sh(golden_dir):
    g = _load(golden_dir, "sh_colors.npz")
    for deg in range(4):
        st, _ = _all_visible_state(g["xyz"], shs=g["shs"], deg=deg, campos=g["campos"])
        assert (st.radii > 0).all()
        np.testing.assert_allclose(st.rgb, g[f"rgb_deg{deg}"], rtol=1e-5, atol=2e-6)
        # clamp mask == (unclamped value < 0)
        assert ((st.rgb == 0) >= (st.clamped > 0)).all()


def test_cov3d_matches_reference_python_path(golden_dir):
    g = _load(golden_dir, "cov3d.npz")
    P = g["scales"].shape[0]
    xyz = np.random.RandomState(0).uniform(-1, 1, (P, 3)).astype(np.float32)
    st, _ = _all_visible_state(xyz, colors=np.ones((P, 3), np.float32), scales=g["scales"],
                               rots=g["rotations_unit"], mod=float(g["scale_modifier"]))
    assert (st.radii > 0).all()
    np.testing.assert_allclose(st.cov3Ds, g["cov3D"], rtol=2e-5, atol=1e-9)


def test_projection_matches_reference_geom_transform(golden_dir):
    g = _load(golden_dir, "camera.npz")
    cam = scenes.look_at_camera(g["eye"], g["target"], int(g["width"]), int(g["height"]))
    S = settings_from_camera(cam)
    P = g["points"].shape[0]
    st = raster.preprocess(S, g["points"], np.full((P, 1), 0.5, np.float32), colors_precomp=np.ones((P, 3), np.float32),
                           scales=np.full((P, 3), 0.01, np.float32), rotations=np.tile(np.float32([1, 0, 0, 0]), (P, 1)))
    vis = st.radii > 0
    assert vis.sum() >= 8
    W, H = int(g["width"]), int(g["height"])
    px = ((g["ndc"][:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((g["ndc"][:, 1] + 1.0) * H - 1.0) * 0.5
    np.testing.assert_allclose(st.means2D[vis, 0], px[vis], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(st.means2D[vis, 1], py[vis], rtol=1e-5, atol=2e-4)


def test_loss_restatement_matches_reference_loss_utils(golden_dir):
    """tests/aten_reference.training_loss (the fp32 reference of the fused loss kernel) == utils/loss_utils.py."""
    import aten_reference as losses
    g = _load(golden_dir, "loss.npz")
    a = torch.tensor(g["img"], requires_grad=True); b = torch.tensor(g["gt"])
    assert abs(losses.l1_loss(a, b).item() - float(g["l1"])) < 1e-7
    assert abs(losses.ssim(a, b).item() - float(g["ssim"])) < 1e-6
    total = losses.training_loss(a, b, 0.2)
    assert abs(total.item() - float(g["loss"])) < 1e-6
    total.backward()
    np.testing.assert_allclose(a.grad.numpy(), g["grad"], rtol=1e-4, atol=1e-9)


def test_points_pseudomesh_oracle_matches_reference(golden_dir):
    """oracle/expansion.py points_* against PointsGaussianModel.prepare_vertices / prepare_scaling_rot / get_scaling
    (games/flat_splatting/scene/points_gaussian_model.py:28-109)."""
    g = _load(golden_dir, "points_model.npz")
    tri = expansion.points_prepare_vertices(torch.tensor(g["pv_xyz"]), torch.tensor(g["pv_scaling"]),
                                            torch.tensor(g["pv_rotation"]))
    np.testing.assert_allclose(tri.numpy(), g["pv_triangles"], rtol=0, atol=1e-6)
    sl, rot = expansion.points_prepare_scaling_rot(torch.tensor(g["triangles"]))
    np.testing.assert_allclose(sl.numpy(), g["_scaling"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rot.numpy(), g["_rotation"], atol=1e-6)
    np.testing.assert_allclose(expansion.points_get_scaling(sl).numpy(), g["get_scaling"], rtol=1e-6, atol=1e-12)
    # round trip: the triangle of a flat Gaussian maps back to the same in-plane scales (longer first) and the same frame
    sl2, rot2 = expansion.points_prepare_scaling_rot(tri)
    want = np.sort(g["pv_scaling"], axis=1)[:, ::-1]
    np.testing.assert_allclose(sl2.numpy(), want, atol=2e-5)


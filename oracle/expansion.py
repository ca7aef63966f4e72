"""CPU oracle for the mesh -> Gaussian expansion.

TEST INFRASTRUCTURE ONLY.  PyTorch-CPU float32 restatement (autograd gives the backward) of
  E1  GaussianMeshModel.update_alpha / _calc_xyz   games/mesh_splatting/scene/gaussian_mesh_model.py:153-169, 86-101
  E2  GaussianMeshModel.prepare_scaling_rot        games/mesh_splatting/scene/gaussian_mesh_model.py:103-151
  E3  rot_to_quat_batch (+_sqrt_positive_part, standardize_quaternion)   utils/general_utils.py:19-96
  E4  activation getters exp / normalize / sigmoid / cat                 scene/gaussian_model.py:95-115
and of the multi-mesh variant (loop over meshes + cat)
      games/multi_mesh_splatting/scene/gaussian_multi_mesh_model.py:99-119, 121-174, 176-199.
Parity is PINNED: tests/golden/expansion_*.npz hold outputs of the reference's own functions
(imported from /root/reference by tests/golden/make_golden.py) and tests/test_oracle_golden.py checks
this file against them.
"""
from __future__ import annotations

import torch

EPS_S0 = 1e-8  # gaussian_mesh_model.py:43


def update_alpha(_alpha, vertices, faces):
    """-> alpha [F,K,3], triangles [F,3,3], xyz [F*K,3]."""
    alpha = torch.relu(_alpha) + 1e-8
    alpha = alpha / alpha.sum(dim=-1, keepdim=True)
    triangles = vertices[faces]
    xyz = torch.matmul(alpha, triangles)
    return alpha, triangles, xyz.reshape(-1, 3)


def _rowdot(a, b):
    return torch.sum(a * b, dim=-1, keepdim=True)


def _unit(v, eps):
    """v / (|v| + eps) and the norm itself."""
    n = torch.linalg.vector_norm(v, dim=-1, keepdim=True)
    return v / (n + eps), n


def face_frames(triangles, eps=EPS_S0):
    """Per-face orthonormal frame and in-plane half extents.
    -> rows [F,3,3] = (unit normal, unit centroid->corner1, Gram-Schmidt of centroid->corner2), scales [F,3] = (eps, s1, s2)."""
    p0, p1, p2 = triangles.unbind(dim=1)
    normal_hat, _ = _unit(torch.linalg.cross(p1 - p0, p2 - p0, dim=1), eps)
    centroid = triangles.mean(dim=1)
    arm1 = p1 - centroid
    len1 = torch.linalg.vector_norm(arm1, dim=-1, keepdim=True) + eps
    axis1 = arm1 / len1
    arm2 = p2 - centroid
    resid = arm2 - _rowdot(arm2, normal_hat) * normal_hat - _rowdot(arm2, axis1) * axis1
    axis2, _ = _unit(resid, eps)
    half1 = len1 / 2.0
    half2 = _rowdot(arm2, axis2) / 2.0
    thin = torch.full_like(half1, eps)
    return torch.stack((normal_hat, axis1, axis2), dim=1), torch.cat((thin, half1, half2), dim=1)


# candidate numerators of the quaternion, one row per choice of the dominant component (w, x, y, z);
# entries: ('d', k) -> squared magnitude k, (+1/-1, (a, b), (c, d)) -> m[a][b] +/- m[c][d]
_CAND = (
    (('d', 0), (-1, (2, 1), (1, 2)), (-1, (0, 2), (2, 0)), (-1, (1, 0), (0, 1))),
    ((-1, (2, 1), (1, 2)), ('d', 1), (+1, (1, 0), (0, 1)), (+1, (0, 2), (2, 0))),
    ((-1, (0, 2), (2, 0)), (+1, (1, 0), (0, 1)), ('d', 2), (+1, (1, 2), (2, 1))),
    ((-1, (1, 0), (0, 1)), (+1, (2, 0), (0, 2)), (+1, (2, 1), (1, 2)), ('d', 3)),
)


def rot_to_quat(rot):
    """Rotation matrices [...,3,3] -> quaternions (w,x,y,z), w >= 0: the pytorch3d construction the reference uses
    (utils/general_utils.py:43-96): four candidate magnitudes sqrt(max(0, 1 +- m00 +- m11 +- m22)), take the largest,
    divide its numerator row by 2*max(|q|, 0.1)."""
    m = rot.reshape(-1, 3, 3)
    d0, d1, d2 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    arg = torch.stack((1.0 + d0 + d1 + d2, 1.0 + d0 - d1 - d2, 1.0 - d0 + d1 - d2, 1.0 - d0 - d1 + d2), dim=-1)
    mag = torch.zeros_like(arg)
    pos = arg > 0
    mag[pos] = torch.sqrt(arg[pos])
    rows = []
    for spec in _CAND:
        ent = []
        for e in spec:
            if e[0] == 'd':
                ent.append(mag[:, e[1]] ** 2)
            else:
                sgn, (a, b), (c, d) = e
                ent.append(m[:, a, b] + m[:, c, d] if sgn > 0 else m[:, a, b] - m[:, c, d])
        rows.append(torch.stack(ent, dim=-1))
    table = torch.stack(rows, dim=-2)                                           # [F,4,4]
    floor = torch.tensor(0.1, dtype=mag.dtype)
    table = table / (2.0 * mag[..., None].max(floor))
    pick = mag.argmax(dim=-1)
    q = table[torch.arange(table.shape[0]), pick]
    return torch.where(q[:, :1] < 0, -q, q)


def prepare_scaling_rot(triangles, _scale, K, eps=EPS_S0):
    """-> _scaling [F*K,3] (log-space), _rotation [F*K,4] (w,x,y,z; un-normalised output of rot_to_quat)."""
    rows, s = face_frames(triangles, eps)
    F = triangles.shape[0]
    scales = s.unsqueeze(1).broadcast_to(F, K, 3).flatten(0, 1)
    _scaling = torch.log(torch.relu(_scale * scales) + eps)
    rotation = rows.unsqueeze(1).broadcast_to(F, K, 3, 3).flatten(0, 1).transpose(-2, -1)
    return _scaling, rot_to_quat(rotation)


def expand(vertices, faces, _alpha, _scale, eps=EPS_S0):
    """E1+E2+E3: -> xyz [P,3], _scaling [P,3], _rotation [P,4], alpha, triangles."""
    alpha, triangles, xyz = update_alpha(_alpha, vertices, faces)
    _scaling, _rotation = prepare_scaling_rot(triangles, _scale, _alpha.shape[1], eps)
    return xyz, _scaling, _rotation, alpha, triangles


def activate(_scaling, _rotation, _opacity, _features_dc, _features_rest):
    """E4: scene/gaussian_model.py:95-115."""
    return (torch.exp(_scaling), torch.nn.functional.normalize(_rotation), torch.sigmoid(_opacity),
            torch.cat((_features_dc, _features_rest), dim=1))


def expand_multi(vertices_list, faces_list, alpha_list, scale_list, eps=EPS_S0):
    """gaussian_multi_mesh_model.py:99-119,121-174: per-mesh expansion, concatenated."""
    outs = [expand(v, f, a, s, eps)[:3] for v, f, a, s in zip(vertices_list, faces_list, alpha_list, scale_list)]
    return tuple(torch.cat([o[i] for o in outs]) for i in range(3))


# ---- gs_points pseudo-mesh (games/flat_splatting/scene/points_gaussian_model.py) ----

def points_prepare_vertices(xyz, _scaling, _rotation):
    """Flat Gaussians -> their pseudo-mesh triangles [P,3,3]  (points_gaussian_model.py:28-59; get_scaling :106-109;
    build_rotation utils/general_utils.py:158-179)."""
    ext = torch.exp(_scaling[:, -2:])
    q = _rotation / torch.sqrt((_rotation * _rotation).sum(dim=1, keepdim=True))
    w, x, y, z = q.unbind(dim=1)
    axis_b = torch.stack((2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)), dim=1)   # column 1 of R
    axis_c = torch.stack((2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)), dim=1)   # column 2 of R
    tip_b = xyz + ext[:, :1] * axis_b
    tip_c = xyz + ext[:, 1:] * axis_c
    first = (ext[:, 0] > ext[:, 1])[:, None]
    return torch.stack((xyz, torch.where(first, tip_b, tip_c), torch.where(first, tip_c, tip_b)), dim=1)


def points_prepare_scaling_rot(triangles, eps=1e-8):
    """Pseudo-mesh triangles -> (_scaling [P,2], _rotation [P,4])  (points_gaussian_model.py:61-104)."""
    apex, b, c = triangles.unbind(dim=1)
    e_b, e_c = b - apex, c - apex
    n_hat, _ = _unit(torch.linalg.cross(e_b, e_c, dim=1), eps)
    len_b = torch.linalg.vector_norm(e_b, dim=-1, keepdim=True) + eps
    u_b = e_b / len_b
    u_c, _ = _unit(e_c - _rowdot(e_c, n_hat) * n_hat - _rowdot(e_c, u_b) * u_b, eps)
    len_c = _rowdot(e_c, u_c)
    scaling = torch.log(torch.cat((len_b, len_c), dim=1).abs())
    rot = torch.stack((n_hat, u_b, u_c), dim=1).transpose(-2, -1)
    return scaling, rot_to_quat(rot)


def points_get_scaling(_scaling, eps=1e-8):
    """get_scaling = (eps, exp(last two log-scales))  (points_gaussian_model.py:106-109)."""
    return torch.cat((torch.full_like(_scaling[:, :1], eps), torch.exp(_scaling[:, -2:])), dim=1)


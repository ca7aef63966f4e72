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


def _dot(v, u):
    return (v * u).sum(dim=-1, keepdim=True)


def face_frames(triangles, eps=EPS_S0):
    """-> per-face rotation rows (v0,v1,v2) [F,3,3] and scales (s0,s1,s2) [F,3]."""
    t0, t1, t2 = triangles[:, 0], triangles[:, 1], triangles[:, 2]
    normals = torch.linalg.cross(t1 - t0, t2 - t0, dim=1)
    v0 = normals / (torch.linalg.vector_norm(normals, dim=-1, keepdim=True) + eps)
    means = torch.mean(triangles, dim=1)
    v1 = t1 - means
    v1_norm = torch.linalg.vector_norm(v1, dim=-1, keepdim=True) + eps
    v1 = v1 / v1_norm
    v2_init = t2 - means
    v2 = v2_init - _dot(v2_init, v0) * v0 - _dot(v2_init, v1) * v1
    v2 = v2 / (torch.linalg.vector_norm(v2, dim=-1, keepdim=True) + eps)
    s1 = v1_norm / 2.0
    s2 = _dot(v2_init, v2) / 2.0
    s0 = eps * torch.ones_like(s1)
    return torch.stack((v0, v1, v2), dim=1), torch.cat((s0, s1, s2), dim=1)


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    m = x > 0
    ret[m] = torch.sqrt(x[m])
    return ret


def rot_to_quat(rot):
    """pytorch3d matrix_to_quaternion as restated in utils/general_utils.py:43-96; rot [...,3,3]."""
    m = rot.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[:, i] for i in range(9)]
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                             1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1, dtype=q_abs.dtype)
    cand = cand / (2.0 * q_abs[..., None].max(flr))
    idx = q_abs.argmax(dim=-1)
    out = cand[torch.arange(cand.shape[0]), idx]
    return torch.where(out[:, 0:1] < 0, -out, out)


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

"""MeshGaussianModel -- the attribute surface renderer/*/__init__.py reads from a `pc`, backed by the fused ops.

Mirrors games/mesh_splatting/scene/gaussian_mesh_model.py (GaussianMeshModel) + the getters of
scene/gaussian_model.py:95-122 closely enough that the reference's `render(viewpoint_camera, pc, pipe, bg)`
(renderer/gaussian_renderer/__init__.py:25) and its animated variant run on it unchanged.  It exists so that
bench.py / tests do not need /root/reference at run time; a reference GaussianMeshModel instance can instead be
patched in place with expansion.patch_mesh_model().
"""
from __future__ import annotations

import torch
from torch import nn

from . import expansion
from .scenes import MeshGaussianParams


class MeshGaussianModel:
    def __init__(self, sh_degree: int = 3):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.eps_s0 = expansion.EPS_S0
        self.vertices = self.faces = self._alpha = self._scale = None
        self._features = None
        self._opacity = None
        self.alpha = self.triangles = self._xyz = self._scaling = self._rotation = None
        self.optimizer = None

    @classmethod
    def from_params(cls, p: MeshGaussianParams, device="cuda", sh_degree: int = 3, active_sh_degree: int = 3,
                    packed_features: bool = False):
        """packed_features: keep SH coefficients in ONE [P,16,3] parameter (`_features`); `_features_dc` / `_features_rest`
        become views of it and get_features is zero-copy (the reference's torch.cat re-materialises 192 B/Gaussian
        every frame, scene/gaussian_model.py:107-111).  Needs FlatAdam for the DC / rest learning rates."""
        m = cls(sh_degree)
        m.active_sh_degree = active_sh_degree
        m.faces = p.faces.to(device)
        mk = lambda t: nn.Parameter(t.to(device).float().contiguous().requires_grad_(True))
        for k in ("vertices", "_alpha", "_scale", "_opacity"):
            setattr(m, k, mk(getattr(p, k)))
        if packed_features:
            m._features = mk(torch.cat((p._features_dc, p._features_rest), dim=1))
        else:
            m._features = None
            m._features_dc, m._features_rest = mk(p._features_dc), mk(p._features_rest)
        m.update_alpha()
        m.prepare_scaling_rot()
        return m

    def __getattr__(self, name):   # only called when normal lookup fails: packed-mode views
        if name in ("_features_dc", "_features_rest") and self.__dict__.get("_features") is not None:
            f = self.__dict__["_features"]
            return f[:, :1] if name == "_features_dc" else f[:, 1:]
        raise AttributeError(name)

    # -- the two hooks train.py:154-157 calls every step
    def update_alpha(self):
        self.alpha, self.triangles, self._xyz = expansion.update_alpha_op(self.vertices, self.faces, self._alpha)

    def prepare_scaling_rot(self):
        self._scaling, self._rotation = expansion.prepare_scaling_rot_op(self.triangles, self._scale,
                                                                         self._alpha.shape[1], self.eps_s0)

    def expand_fused(self, activated: bool = True):
        """Fast path: one launch for E1-E4; returns (xyz, scaling, rotation) and refreshes alpha/triangles."""
        xyz, sc, rot, self.alpha, self.triangles = expansion.expand(self.vertices, self.faces, self._alpha, self._scale,
                                                                    self.eps_s0, activated)
        if not activated:
            self._xyz, self._scaling, self._rotation = xyz, sc, rot
        return xyz, sc, rot

    # -- getters (scene/gaussian_model.py:95-118)
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        if self._features is not None:
            return self._features
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    def parameters(self):
        if self._features is not None:
            return [self.vertices, self._alpha, self._features, self._opacity, self._scale]
        return [self.vertices, self._alpha, self._features_dc, self._features_rest, self._opacity, self._scale]

    def training_setup(self, vertices_lr=0.0, alpha_lr=0.001, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005):
        """Adam groups of gaussian_mesh_model.py:171-183 (lrs: arguments_games/__init__.py:17-30)."""
        if self._features is not None:
            raise RuntimeError("packed features need gms_b200.optim.FlatAdam (per-coefficient learning rates)")
        groups = [{"params": [self.vertices], "lr": vertices_lr, "name": "vertices"},
                  {"params": [self._alpha], "lr": alpha_lr, "name": "alpha"},
                  {"params": [self._features_dc], "lr": feature_lr, "name": "f_dc"},
                  {"params": [self._features_rest], "lr": feature_lr / 20.0, "name": "f_rest"},
                  {"params": [self._opacity], "lr": opacity_lr, "name": "opacity"},
                  {"params": [self._scale], "lr": scaling_lr, "name": "scaling"}]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        return self.optimizer


class MultiMeshGaussianModel(MeshGaussianModel):
    """gs_multi_mesh: several meshes, one Gaussian set (games/multi_mesh_splatting/scene/gaussian_multi_mesh_model.py).

    The reference keeps per-mesh lists (vertices / faces / _alpha / _scale) and loops `expand -> torch.cat`
    (:99-119, :121-174, :176-199).  Per-face work is independent of which mesh a face belongs to, so when every mesh
    uses the same K the meshes are merged ONCE at construction (faces re-indexed into one vertex array) and the whole
    set runs through the single-mesh kernels: one launch instead of a loop, no concatenation.  Meshes with different K
    keep the reference's loop (`expand_per_mesh`)."""

    @classmethod
    def from_mesh_params(cls, plist, device="cuda", sh_degree: int = 3, active_sh_degree: int = 3, packed_features: bool = False):
        Ks = {p._alpha.shape[1] for p in plist}
        if len(Ks) != 1:
            raise ValueError("merged fast path needs one K; use expand_per_mesh for heterogeneous K")
        off, faces = 0, []
        for p in plist:
            faces.append(p.faces + off)
            off += p.vertices.shape[0]
        merged = MeshGaussianParams(torch.cat([p.vertices for p in plist]), torch.cat(faces),
                                    torch.cat([p._alpha for p in plist]), torch.cat([p._scale for p in plist]),
                                    torch.cat([p._features_dc for p in plist]), torch.cat([p._features_rest for p in plist]),
                                    torch.cat([p._opacity for p in plist]))
        m = cls.from_params(merged, device, sh_degree, active_sh_degree, packed_features)
        m.mesh_face_counts = [p.faces.shape[0] for p in plist]
        return m

    @staticmethod
    def expand_per_mesh(vertices_list, faces_list, alpha_list, scale_list, eps: float = expansion.EPS_S0):
        """Reference-shaped path for heterogeneous K: per-mesh fused launch, then cat (-> xyz, _scaling, _rotation)."""
        outs = [expansion.expand(v, f, a, s, eps, activated=False)[:3] for v, f, a, s in
                zip(vertices_list, faces_list, alpha_list, scale_list)]
        return tuple(torch.cat([o[i] for o in outs]) for i in range(3))

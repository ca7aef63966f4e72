"""`point_cloud.ply` / `model_params.pt` IO without the `plyfile` dependency.

Layout = what the reference writes in GaussianModel._save_ply (scene/gaussian_model.py:177-216): one `vertex` element,
all properties float32, in the order  x y z nx ny nz f_dc_0..2 f_rest_0..(3*(M-1)-1) opacity scale_0..2 rot_0..3 ;
`f_dc` / `f_rest` are stored channel-major ([P,3,M-1] flattened) and transposed back to [P,M-1,3] on load
(scene/gaussian_model.py:224-262).  Mesh models add `model_params.pt` next to the PLY with `_alpha`, `_scale`, `vertices`,
`faces`, `triangles` (games/mesh_splatting/scene/gaussian_mesh_model.py:189-225).
Reads binary_little_endian (plyfile's default) and ascii PLY; writes binary_little_endian."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np
import torch

from .scenes import MeshGaussianParams

_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
          "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4",
          "uint": "<u4", "uint32": "<u4"}


def read_ply_vertices(path: str) -> Tuple[np.ndarray, List[str]]:
    """-> (structured array of the `vertex` element, property names)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties on the vertex element are not supported")
                props.append((tok[2], _TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        dtype = np.dtype(props)
        if fmt == "binary_little_endian":
            data = np.fromfile(f, dtype=dtype, count=count)
        elif fmt == "ascii":
            raw = np.loadtxt(f, max_rows=count, ndmin=2)
            data = np.empty(count, dtype=dtype)
            for i, (n, _) in enumerate(props):
                data[n] = raw[:, i]
        else:
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
    if data.shape[0] != count:
        raise ValueError(f"{path}: expected {count} vertices, found {data.shape[0]}")
    return data, [n for n, _ in props]


def _sorted_cols(data, names, prefix):
    cols = sorted([n for n in names if n.startswith(prefix)], key=lambda x: int(x.split("_")[-1]))
    return np.stack([np.asarray(data[n], np.float32) for n in cols], axis=1) if cols else np.zeros((data.shape[0], 0), np.float32)


def load_gaussian_ply(path: str) -> Dict[str, torch.Tensor]:
    """-> dict(_xyz [P,3], _features_dc [P,1,3], _features_rest [P,M-1,3], _opacity [P,1], _scaling [P,S], _rotation [P,4])
    exactly as GaussianModel._load_ply builds them (scene/gaussian_model.py:224-262)."""
    data, names = read_ply_vertices(path)
    P = data.shape[0]
    xyz = np.stack([data["x"], data["y"], data["z"]], axis=1).astype(np.float32)
    fdc = _sorted_cols(data, names, "f_dc_").reshape(P, 3, 1)
    frest = _sorted_cols(data, names, "f_rest_")
    frest = frest.reshape(P, 3, frest.shape[1] // 3) if frest.shape[1] else frest.reshape(P, 3, 0)
    out = dict(_xyz=torch.tensor(xyz), _features_dc=torch.tensor(fdc).transpose(1, 2).contiguous(),
               _features_rest=torch.tensor(frest).transpose(1, 2).contiguous(),
               _opacity=torch.tensor(np.asarray(data["opacity"], np.float32)[:, None]),
               _scaling=torch.tensor(_sorted_cols(data, names, "scale_")), _rotation=torch.tensor(_sorted_cols(data, names, "rot")))
    return out


def save_gaussian_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation, eps_s0: float = 1e-8) -> None:
    """Writer with the reference's property order / channel-major SH layout (scene/gaussian_model.py:177-216).  A two-column
    `scaling` (flat Gaussians, gs_flat / gs_points) gets the constant log(eps_s0) column prepended, as _save_ply does (:196-199),
    so the file always carries scale_0..2."""
    t = lambda a: a.detach().cpu().float() if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=torch.float32)
    xyz, fdc, frest, op, sc, rot = map(t, (xyz, features_dc, features_rest, opacity, scaling, rotation))
    P = xyz.shape[0]
    if sc.dim() == 2 and sc.shape[1] == 2:
        sc = torch.cat([torch.log(torch.ones(P, 1) * eps_s0), sc], dim=1)
    cols = [xyz, torch.zeros_like(xyz), fdc.transpose(1, 2).reshape(P, -1), frest.transpose(1, 2).reshape(P, -1), op.reshape(P, 1),
            sc.reshape(P, -1), rot.reshape(P, -1)]
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(cols[2].shape[1])] + \
            [f"f_rest_{i}" for i in range(cols[3].shape[1])] + ["opacity"] + [f"scale_{i}" for i in range(cols[5].shape[1])] + \
            [f"rot_{i}" for i in range(cols[6].shape[1])]
    arr = np.ascontiguousarray(torch.cat(cols, dim=1).numpy().astype("<f4"))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P).encode())
        for n in names:
            f.write(f"property float {n}\n".encode())
        f.write(b"end_header\n")
        arr.tofile(f)


def load_mesh_model(ply_path: str) -> MeshGaussianParams:
    """point_cloud.ply + model_params.pt of a trained gs_mesh run -> raw parameters for MeshGaussianModel.from_params
    (GaussianMeshModel.load_ply, games/mesh_splatting/scene/gaussian_mesh_model.py:211-225)."""
    g = load_gaussian_ply(ply_path)
    params = torch.load(ply_path.replace("point_cloud.ply", "model_params.pt"), map_location="cpu", weights_only=False)
    d = lambda x: (x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(x)).cpu()
    verts = d(params["vertices"]).float()
    faces = d(params["faces"]).long()
    return MeshGaussianParams(verts, faces, d(params["_alpha"]).float(), d(params["_scale"]).float(), g["_features_dc"],
                              g["_features_rest"], g["_opacity"])


def save_mesh_model(ply_path: str, model) -> None:
    """Counterpart of GaussianMeshModel.save_ply (gaussian_mesh_model.py:189-209) for a gms_b200 MeshGaussianModel.
    model_params.pt holds what the reference's writer holds -- `_alpha`, `_scale`, `vertices` as nn.Parameters and
    `triangles`, `faces` as tensors, all ON THE MODEL'S DEVICE, plus the `point_cloud` key -- because the reference's
    load_ply (:211-225) uses them where they are (no .cuda()): a checkpoint written here loads in scripts/render.py."""
    model.update_alpha(); model.prepare_scaling_rot()
    save_gaussian_ply(ply_path, model._xyz, model._features_dc, model._features_rest, model._opacity, model._scaling, model._rotation,
                      getattr(model, "eps_s0", 1e-8))
    par = lambda t: torch.nn.Parameter(t.detach().clone().contiguous(), requires_grad=True)
    torch.save({"_alpha": par(model._alpha), "_scale": par(model._scale), "point_cloud": None,
                "triangles": model.triangles.detach().clone(), "vertices": par(model.vertices), "faces": model.faces.detach().clone()},
               ply_path.replace("point_cloud.ply", "model_params.pt"))

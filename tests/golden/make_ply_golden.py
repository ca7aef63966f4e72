"""Generate tests/golden/ply/{point_cloud.ply, model_params.pt, expected.npz} by running the REFERENCE's own
GaussianMeshModel.save_ply (games/mesh_splatting/scene/gaussian_mesh_model.py:189-209 -> GaussianModel._save_ply,
scene/gaussian_model.py:185-216) in the build container.  Property order, the channel-major SH flattening and the
model_params.pt keys therefore come from the reference's code; only the PLY *container* is written by the small
`plyfile` stand-in below (plyfile itself is not installed here), following the PLY specification exactly as plyfile's
binary_little_endian writer does: one `element vertex N` with `property float <name>` lines, then the packed records.

    python tests/golden/make_ply_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ply")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "gaussian-mesh-splatting_b200"))

_PLY_TYPES = {"f4": "float", "f8": "double", "u1": "uchar", "i4": "int", "u4": "uint", "i2": "short", "u2": "ushort", "i1": "char"}


class PlyElement:
    def __init__(self, data, name):
        self.data, self.name = data, name

    @staticmethod
    def describe(data, name):
        return PlyElement(np.asarray(data), name)


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        with open(path, "wb") as f:
            f.write(b"ply\nformat binary_little_endian 1.0\n")
            for el in self.elements:
                f.write(f"element {el.name} {len(el.data)}\n".encode("ascii"))
                for n in el.data.dtype.names:
                    dt = el.data.dtype[n]
                    f.write(f"property {_PLY_TYPES[dt.str[1:]]} {n}\n".encode("ascii"))
            f.write(b"end_header\n")
            for el in self.elements:
                f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


_stub("plyfile", PlyData=PlyData, PlyElement=PlyElement)
_stub("simple_knn"); _stub("simple_knn._C", distCUDA2=None); _stub("trimesh"); _stub("smplx")
_stub("smplx.lbs", lbs=None, batch_rodrigues=None, vertices2landmarks=None, find_dynamic_lmk_idx_and_bcoords=None)
_stub("smplx.utils", Struct=object, to_tensor=None, to_np=None, rot_mat_to_euler=None)
_stub("diff_gaussian_rasterization", GaussianRasterizationSettings=object, GaussianRasterizer=object)

from games.mesh_splatting.scene.gaussian_mesh_model import GaussianMeshModel  # noqa: E402
from gms_b200 import scenes  # noqa: E402


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(5)
    verts, faces = scenes.icosphere(1)
    verts = verts + 0.03 * np.random.RandomState(2).randn(*verts.shape).astype(np.float32)
    F, K = faces.shape[0], 2
    P = F * K
    m = GaussianMeshModel(3)
    m.vertices = torch.nn.Parameter(torch.tensor(verts))
    m.faces = torch.tensor(faces)
    m._alpha = torch.nn.Parameter(torch.rand(F, K, 3))
    m._scale = torch.nn.Parameter(0.5 + torch.rand(P, 1))
    m._opacity = torch.nn.Parameter(torch.randn(P, 1))
    m._features_dc = torch.nn.Parameter(torch.randn(P, 1, 3))
    m._features_rest = torch.nn.Parameter(torch.randn(P, 15, 3))
    ply = os.path.join(OUT, "point_cloud.ply")
    m.save_ply(ply)                       # the reference's writer: update_alpha, prepare_scaling_rot, _save_ply, torch.save
    # model_params.pt as the reference wrote it holds nn.Parameters and a `point_cloud: None` entry; keep it verbatim
    np.savez_compressed(os.path.join(OUT, "expected.npz"), vertices=verts, faces=faces, _alpha=m._alpha.detach().numpy(),
                        _scale=m._scale.detach().numpy(), _opacity=m._opacity.detach().numpy(),
                        _features_dc=m._features_dc.detach().numpy(), _features_rest=m._features_rest.detach().numpy(),
                        _xyz=m._xyz.detach().numpy(), _scaling=m._scaling.detach().numpy(), _rotation=m._rotation.detach().numpy())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

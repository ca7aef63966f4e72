"""Make the reference's own Python importable on the GPU box: `baseline/_ref` (git-ignored, travels with the gpurun snapshot).

`pip install /root/reference` has nothing to install (no setup.py / pyproject; the only native piece, the rasterizer, is an
empty submodule), so the "install" of the reference is a verbatim, unmodified snapshot of its Python packages made by
`__graft_entry__.build()` in the build container.  Nothing here is product source and nothing is tracked by git; the
`-m gpu` tests run the reference's `render()` / `GaussianMeshModel` from it, and bench.py times the reference's PyTorch
expansion on the GPU from it.  When neither /root/reference nor a previous snapshot exists the callers skip.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = os.path.join(HERE, "_ref")
SRC = "/root/reference"
PACKAGES = ["renderer", "games", "scene", "utils", "arguments", "arguments_games", "scripts"]


def make_snapshot(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        return SNAP if os.path.isdir(SNAP) else ""
    if force and os.path.isdir(SNAP):
        shutil.rmtree(SNAP)
    for pkg in PACKAGES:
        s, d = os.path.join(SRC, pkg), os.path.join(SNAP, pkg)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.obj", "*.ply", "*.png", "*.npz", "*.pkl"))
    return SNAP


def reference_root() -> str:
    """Directory to put on sys.path to import the reference ('' if unavailable)."""
    if os.path.isdir(os.path.join(SNAP, "renderer")):
        return SNAP
    if os.path.isdir(os.path.join(SRC, "renderer")):
        return SRC
    return ""


def install_stubs(rasterizer_module=None) -> None:
    """SURVEY.md 8(c) import recipe: empty stand-ins for the reference's absent third-party imports, and
    `diff_gaussian_rasterization` -> the given module (default: the product shim)."""
    stubs = [("plyfile", dict(PlyData=object, PlyElement=object)), ("simple_knn", {}), ("simple_knn._C", dict(distCUDA2=None)),
             ("trimesh", {}), ("smplx", {}),
             ("smplx.lbs", dict(lbs=None, batch_rodrigues=None, vertices2landmarks=None, find_dynamic_lmk_idx_and_bcoords=None)),
             ("smplx.utils", dict(Struct=object, to_tensor=None, to_np=None, rot_mat_to_euler=None))]
    for name, attrs in stubs:
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    if rasterizer_module is not None:
        sys.modules["diff_gaussian_rasterization"] = rasterizer_module


def import_reference(rasterizer_module=None):
    """Put the snapshot on sys.path (after the stubs) -> True if the reference can be imported."""
    root = reference_root()
    if not root:
        return False
    install_stubs(rasterizer_module)
    if root not in sys.path:
        sys.path.insert(0, root)
    return True


if __name__ == "__main__":
    print(make_snapshot(force="--force" in sys.argv))

"""torchrun --nproc-per-node 2 tests/dist_check_sharded.py   (needs 2 GPUs; not collected by pytest)

Frame-sharded data parallelism with the SHARDED FlatAdam (reduce-scatter -> local Adam slice -> all-gather) must give
every rank the same parameters as a single process that averages the two cameras' gradients itself and runs the
unsharded FlatAdam."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-mesh-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from gms_b200 import scenes
from gms_b200.model import MeshGaussianModel
from gms_b200.trainer import MeshTrainer, render_frame, shard_cameras
from gms_b200.losses import fused_training_loss


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p = scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=5)
    cams = [c.to(dev) for c in scenes.ring_cameras(4, 2.6, 320, 240)]
    bg = torch.ones(3, device=dev)
    gt_model = MeshGaussianModel.from_params(scenes.init_mesh_gaussians(*scenes.icosphere(4), K=3, seed=77), dev)
    with torch.no_grad():
        gts = [render_frame(gt_model, c, bg)[0].clamp(0, 1).contiguous() for c in cams]
    # distributed runs: autograd fast path (full gradient re-zeroing) and the one-call native frame (only the vertex
    # segment is re-zeroed, every other gradient is overwritten by the next frame)
    runs = {}
    for native in (False, True):
        m = MeshGaussianModel.from_params(p, dev, packed_features=True)
        tr = MeshTrainer(m, bg, world=world, rank=rank, fast=True, native=native)
        for s in range(3):
            ci = shard_cameras(len(cams), s, rank, world)
            tr.step(cams[ci], gts[ci])
        runs[native] = (m, tr)
    # single-process reference on every rank: average the per-camera gradients by hand
    r = MeshGaussianModel.from_params(p, dev, packed_features=True)
    rt = MeshTrainer(r, bg, world=1, rank=0, fast=True)
    from gms_b200 import rasterizer
    rasterizer.DIRECT_SH_GRAD = False      # two frames accumulate into one gradient here: plain autograd accumulation
    for s in range(3):
        for q in range(world):
            ci = shard_cameras(len(cams), s, q, world)
            image, _, _ = render_frame(r, cams[ci], bg)
            (fused_training_loss(image, gts[ci], 0.2) / world).backward()
        rt.opt.step()
    worst, same = 0.0, True
    for native, (m, tr) in runs.items():
        for a, b in zip(m.parameters(), r.parameters()):
            scale = b.detach().abs().max().item() + 1e-12
            worst = max(worst, (a.detach() - b.detach()).abs().max().item() / scale)
        # replicas agree bit for bit across ranks (all-gather of the same slices)
        flat = tr.opt.p.clone()
        other = flat.clone()
        dist.broadcast(other, src=0)
        same = same and bool(torch.equal(flat, other))
    t = torch.tensor([worst, 0.0 if same else 1.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"sharded-vs-single max rel diff {t[0].item():.3e}; replicas identical: {t[1].item() == 0.0}")
        assert t[0].item() < 2e-3 and t[1].item() == 0.0
        print("DIST_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

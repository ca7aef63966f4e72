"""Diagnose free-running vs per-step-synchronised step time (allocator / GC / run-ahead effects)."""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-mesh-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench
from gms_b200.model import MeshGaussianModel
from gms_b200.trainer import MeshTrainer, render_frame

dev = torch.device("cuda", 0)
params, cams, dims = bench.build_scene(sys.argv[1] if len(sys.argv) > 1 else "gs_mesh_1M_1080p")
model = MeshGaussianModel.from_params(params, dev, packed_features=True)
bg = torch.ones(3, device=dev)
cams = [c.to(dev) for c in cams]
with torch.no_grad():
    gts = [render_frame(model, c, bg)[0].clamp(0, 1).contiguous() for c in cams]
tr = MeshTrainer(model, bg)
for s in range(20):
    tr.step(cams[s % 16], gts[s % 16])
torch.cuda.synchronize()

def stats():
    m = torch.cuda.memory_stats()
    return m["num_device_alloc"], m["num_device_free"], m["reserved_bytes.all.current"] >> 20, m["allocated_bytes.all.current"] >> 20

def run(name, n, sync_each, item=False):
    a = stats(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(n):
        l = tr.step(cams[s % 16], gts[s % 16])
        if sync_each: torch.cuda.synchronize()
        if item: l.item()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    b = stats()
    print(f"{name:28s} {dt:7.2f} ms/step   cudaMalloc +{b[0]-a[0]}  cudaFree +{b[1]-a[1]}  reserved {b[2]} MiB allocated {b[3]} MiB", flush=True)

run("free-running", 50, False)
run("sync each step", 50, True)
run("loss.item() each step", 50, False, True)
run("free-running again", 50, False)
gc.disable(); run("free-running, gc disabled", 50, False); gc.enable()
gc.collect(); run("after gc.collect", 50, False)

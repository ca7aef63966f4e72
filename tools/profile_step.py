"""Where does a training step's device time go?  torch.profiler kernel table over a few steps (run under gpurun).
    python tools/profile_step.py [workload] > gpurun_out/step_profile.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-mesh-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity

import bench
from gms_b200.model import MeshGaussianModel
from gms_b200.trainer import MeshTrainer, render_frame

wl = sys.argv[1] if len(sys.argv) > 1 else "gs_mesh_1M_1080p"
dev = torch.device("cuda", 0)
params, cams, dims = bench.build_scene(wl)
model = MeshGaussianModel.from_params(params, dev, packed_features=True)
bg = torch.ones(3, device=dev)
cams = [c.to(dev) for c in cams]
with torch.no_grad():
    gts = [render_frame(model, c, bg)[0].clamp(0, 1).contiguous() for c in cams[:4]]
tr = MeshTrainer(model, bg)
for s in range(5):
    tr.step(cams[s % 4], gts[s % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(10):
        tr.step(cams[s % 4], gts[s % 4])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
import time
t0 = time.perf_counter()
for s in range(20):
    tr.step(cams[s % 4], gts[s % 4])
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"CPU-side enqueue time per step (includes the forward's host sync): {t_cpu / 20 * 1e3:.2f} ms; wall {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms")

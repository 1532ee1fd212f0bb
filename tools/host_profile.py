#!/usr/bin/env python
"""Where does the HOST spend its time enqueuing one train step?  cProfile over a few steps (device work is asynchronous, so
cumulative times are Python / launch overhead, not kernel time).

    python tools/host_profile.py [--config presight] [--steps 5]
"""
import argparse, cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from presight_b200 import synthetic
from presight_b200.cameras.rays import RayBundle
from presight_b200.model import VIDEO_ID

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="presight")
ap.add_argument("--rays", type=int, default=65536)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = bench.build_config(args.config, "b200")
torch.manual_seed(42)
host = synthetic.make_rays(args.rays, seed=42)
model = bench.build_model(args.config, cfg, host, dev).train()
params = [p for p in model.parameters() if p.requires_grad]
keys = ("origins", "directions", "camera_indices", "video_ids", "rgb", "features", "sky")
b = {k: host[k].to(dev) for k in keys}


def step():
    for p in params:
        p.grad = None
    rb = RayBundle(origins=b["origins"], directions=b["directions"], camera_indices=b["camera_indices"],
                   metadata={VIDEO_ID: b["video_ids"]})
    model.proposal_sampler._step = 0
    out = model(rb)
    loss = bench.step_loss(model, out, b)
    loss.backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(args.steps):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"# host enqueue {1e3 * (t1 - t0) / args.steps:.2f} ms/step, {len(params)} parameters")
pr = cProfile.Profile()
pr.enable()
for _ in range(args.steps):
    step()
pr.disable()
torch.cuda.synchronize()
for sort in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).strip_dirs().sort_stats(sort).print_stats(28)
    print(s.getvalue()[:6000])

#!/usr/bin/env python
"""Can the whole train step (forward + losses + backward, side streams included) be captured into ONE CUDA graph?
Captures it, replays it, compares loss / gradients with the eager step and times both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from presight_b200 import synthetic
from presight_b200.cameras.rays import RayBundle
from presight_b200.model import VIDEO_ID, NerfactoNuscMSModel

dev = torch.device("cuda", 0)
cfg = synthetic.config_c2("b200")
n = int(os.environ.get("RAYS", 65536))
torch.manual_seed(42)
host = synthetic.make_rays(n, seed=42)
model = NerfactoNuscMSModel(cfg, torch.zeros(1, 3), synthetic.tile_aabb(), host["n_cameras"], host["n_videos"]).to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
keys = ("origins", "directions", "camera_indices", "video_ids", "rgb", "features", "sky")
static = {k: host[k].to(dev) for k in keys}


def step():
    rb = RayBundle(origins=static["origins"], directions=static["directions"], camera_indices=static["camera_indices"],
                   metadata={VIDEO_ID: static["video_ids"]})
    model.proposal_sampler._step = 0
    out = model(rb)
    loss = bench.step_loss(model, out, static)
    loss.backward()
    return loss.detach()


def timed(fn, iters=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def eager():
    for p in params:
        p.grad = None
    return step()


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        eager()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print(f"eager: {timed(eager):.3f} ms/step", flush=True)
for p in params:
    p.grad = None
g = torch.cuda.CUDAGraph()
t0 = time.perf_counter()
with torch.cuda.graph(g):
    static_loss = step()
print(f"captured in {time.perf_counter() - t0:.2f} s", flush=True)
g.replay(); torch.cuda.synchronize()
print("graph loss", float(static_loss), flush=True)
print(f"graph: {timed(g.replay):.3f} ms/step", flush=True)
grads_graph = [p.grad.clone() for p in params if p.grad is not None]
print("grads", len(grads_graph), "finite", all(torch.isfinite(x).all() for x in grads_graph))

# ---- two half-batches on two streams inside one graph --------------------------------------------------------------
MICRO = int(os.environ.get("MICRO", 2))
streams = [torch.cuda.Stream() for _ in range(MICRO)]
halves = [{k: static[k][n * i // MICRO:n * (i + 1) // MICRO] for k in keys} for i in range(MICRO)]


def step_micro():
    main = torch.cuda.current_stream()
    losses_ = []
    for st, hb in zip(streams, halves):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            rb = RayBundle(origins=hb["origins"], directions=hb["directions"], camera_indices=hb["camera_indices"],
                           metadata={VIDEO_ID: hb["video_ids"]})
            model.proposal_sampler._step = 0
            out = model(rb)
            loss = bench.step_loss(model, out, hb) / MICRO
            loss.backward()
            losses_.append(loss.detach())
    for st in streams:
        main.wait_stream(st)
    return torch.stack(losses_).sum()


def eager_micro():
    for p in params:
        p.grad = None
    return step_micro()


with torch.cuda.stream(s):
    s.wait_stream(torch.cuda.current_stream())
    for _ in range(3):
        eager_micro()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
for p in params:
    p.grad = None
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    static_loss2 = step_micro()
g2.replay(); torch.cuda.synchronize()
print(f"micro={MICRO} graph loss", float(static_loss2), flush=True)
print(f"micro={MICRO} graph: {timed(g2.replay):.3f} ms/step", flush=True)

#!/usr/bin/env python
"""Kernel timeline of ONE train step (kineto / CUPTI through torch.profiler; nsys is not in the image):
prints every GPU kernel / memset / memcpy of the last profiled step with its stream, start offset and duration, plus
the idle gaps of the device (no kernel running on any stream).

    python tools/timeline.py [--rays 65536] > gpurun_out/timeline.txt
"""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from presight_b200 import synthetic
from presight_b200.cameras.rays import RayBundle
from presight_b200.model import VIDEO_ID, NerfactoNuscMSModel

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=65536)
ap.add_argument("--config", default="c2")
args = ap.parse_args()
# under torchrun (WORLD_SIZE > 1): data-parallel step with the gradient exchange; rank 0 prints its own timeline
import torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sync = None
if world > 1:
    from presight_b200.parallel import GradSynchronizer, init_nccl
    init_nccl(dev)
cfg = bench.build_config(args.config, "b200")
torch.manual_seed(42)
host = synthetic.make_rays(args.rays, seed=42)
model = bench.build_model(args.config, cfg, host, dev).train()
params = [p for p in model.parameters() if p.requires_grad]
if world > 1:
    sync = GradSynchronizer(params, overlap=True)
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
    if sync is not None:
        sync.finish()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        with torch.profiler.record_function("PS_STEP"):
            step()
        torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
path = os.path.join(tempfile.gettempdir(), "ps_trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))["traceEvents"]
steps = sorted([e for e in tr if e.get("name") == "PS_STEP" and e.get("ph") == "X" and e.get("cat") in ("user_annotation", "cpu_op")],
               key=lambda e: e["ts"])
t_lo = steps[-1]["ts"]
gpu = sorted([e for e in tr if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and e["ts"] >= t_lo],
             key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
print(f"# {len(gpu)} device activities; step host span {steps[-1]['dur'] / 1e3:.3f} ms; device span "
      f"{(max(e['ts'] + e['dur'] for e in gpu) - t0) / 1e3:.3f} ms")
print("# start_ms  dur_ms  stream  name")
end_all, idle = t0, 0.0
gaps = []
for e in gpu:
    if e["ts"] > end_all:
        idle += e["ts"] - end_all
        gaps.append((e["ts"] - end_all, (end_all - t0) / 1e3, e["name"][:60]))
    end_all = max(end_all, e["ts"] + e["dur"])
    print(f"{(e['ts'] - t0) / 1e3:8.3f} {e['dur'] / 1e3:7.3f}  s{e['args'].get('stream', '?'):<3} {e['name'][:90]}")
import collections
by = collections.Counter()
cnt = collections.Counter()
for e in gpu:
    by[e["name"][:70]] += e["dur"]
    cnt[e["name"][:70]] += 1
print("# device time by kernel (ms, launches):")
for k, v in by.most_common(25):
    print(f"#   {v / 1e3:8.3f} {cnt[k]:4d}  {k}")
print(f"# device idle inside the step: {idle / 1e3:.3f} ms in {len(gaps)} gaps; largest:")
for g, at, nm in sorted(gaps, reverse=True)[:15]:
    print(f"#   {g / 1e3:.3f} ms at {at:.3f} ms before {nm}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

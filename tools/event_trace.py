#!/usr/bin/env python
"""Kernel end times of ONE train step from CUDA events alone (no CUPTI: the profiler perturbs which kernels co-run):
every probed launch is bracketed by events on its own stream; printed relative to an event recorded at the step's start.
`start` = when the stream reached the launch (not when the kernel got its SMs), `end` = when the kernel finished.

    python tools/event_trace.py [--config c2] [--rays 65536]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from presight_b200 import ops, synthetic
from presight_b200.cameras.rays import RayBundle
from presight_b200.model import VIDEO_ID

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=65536)
ap.add_argument("--config", default="c2")
args = ap.parse_args()
# under torchrun: the data-parallel step with the gradient exchange (rank 0 prints)
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = bench.build_config(args.config, "b200")
torch.manual_seed(42)
host = synthetic.make_rays(args.rays, seed=42 + rank)
model = bench.build_model(args.config, cfg, host, dev).train()
params = [p for p in model.parameters() if p.requires_grad]
sync = None
if world > 1:
    import torch.distributed as dist
    from presight_b200 import fused
    from presight_b200.parallel import GradSynchronizer, init_nccl, level_groups
    init_nccl(dev)
    partial = []
    if os.environ.get("PS_PARTIAL_AR", "1") == "1":
        partial = [(m.hash_table, level_groups(m.num_levels, world=world)) for n, m in model.named_modules()
                   if n.endswith("mlp_base_grid") and hasattr(m, "hash_table") and "proposal" not in n]
    else:
        fused.set_overlap_prop_bwd(False)
    sync = GradSynchronizer(params, overlap=True, partial_tables=partial, peer=os.environ.get("PS_EXCHANGE", "nccl") == "peer")
keys = ("origins", "directions", "camera_indices", "video_ids", "rgb", "features", "sky")
b = {k: host[k].to(dev) for k in keys}


def step():
    for p in params:
        p.grad = None
    rb = RayBundle(origins=b["origins"], directions=b["directions"], camera_indices=b["camera_indices"],
                   metadata={VIDEO_ID: b["video_ids"]})
    model.proposal_sampler._step = 0
    loss = bench.step_loss(model, model(rb), b)
    loss.backward()
    step.bwd_done = torch.cuda.Event(enable_timing=True)
    step.bwd_done.record()
    if sync is not None:
        sync.finish()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
ops.PROBE = ops.KernelProbe()
base = torch.cuda.Event(enable_timing=True)
base.record()
step()
done = torch.cuda.Event(enable_timing=True)
done.record()
torch.cuda.synchronize()
rows = []
for name, pairs in ops.PROBE.events.items():
    for a, e in pairs:
        rows.append((base.elapsed_time(a), base.elapsed_time(e), name))
rows.sort()
if rank != 0:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
print(f"# step: {base.elapsed_time(done):.3f} ms (one step, host enqueue included); backward's last kernel on the main stream done "
      f"at {base.elapsed_time(step.bwd_done):.3f} ms" + (f"; gradient exchange complete at {base.elapsed_time(done):.3f} ms "
                                                       f"({world} ranks)" if world > 1 else ""))
print("# start_ms   end_ms   dur_ms  name")
for s, e, n in rows:
    print(f"{s:8.3f} {e:8.3f} {e - s:8.3f}  {n}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

#!/usr/bin/env python
"""BASELINE config 5: dense density/feature query on a 400x200x16 BEV voxel grid per tile.

    python tools/bench_prior_query.py [--tiles-per-gpu 1] [--iters 20]
    torchrun --nproc-per-node 8 tools/bench_prior_query.py      # tiles sharded over ranks, no communication

Prints one JSON line: points/s and algorithmic GB/s.  Algorithmic bytes per point: reference-faithful count
(main hash encoded twice) = 3 060 B; with the duplicate main encode elided (what runs) = 1 896 B."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from presight_b200 import synthetic  # noqa: E402
from presight_b200.model import NerfactoNuscMSModel  # noqa: E402
from presight_b200.parallel import shard_range  # noqa: E402


def tile_grid(tile_index: int, dev) -> torch.Tensor:
    """400 x 200 x 16 voxel centres = 100 m x 50 m x 8 m at 0.25/0.25/0.5 m, world metres x 0.05."""
    g = torch.Generator().manual_seed(tile_index)
    centre = (torch.rand(2, generator=g) - 0.5) * 300.0
    xs = torch.arange(400, dtype=torch.float32) * 0.25 - 50.0 + centre[0]
    ys = torch.arange(200, dtype=torch.float32) * 0.25 - 25.0 + centre[1]
    zs = torch.arange(16, dtype=torch.float32) * 0.5 - 2.0
    pts = torch.stack(torch.meshgrid(xs, ys, zs, indexing="ij"), dim=-1).reshape(-1, 3)
    return (pts * synthetic.POSE_SCALE).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=8, help="total tiles (Boston Seaport has 8)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.config_c2("b200+fp32" if args.fp32 else "b200")
    torch.manual_seed(42)
    model = NerfactoNuscMSModel(cfg, torch.zeros(1, 3), synthetic.tile_aabb()).to(dev).eval()
    lo, hi = shard_range(args.tiles, rank, world)
    grids = [tile_grid(t, dev) for t in range(lo, hi)]
    M = grids[0].shape[0] if grids else 0
    for gpts in grids:                       # warm-up
        model.query_priors(gpts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        for gpts in grids:
            mean, feats = model.query_priors(gpts)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / 1e3
    if world > 1:
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt)
    if rank == 0:
        total_pts = M * args.tiles * args.iters
        per_tile_ms = t / (args.iters * max(1, len(grids))) * 1e3
        faithful, elided = 3060, 2 * 300 + 1164 + 132
        print(json.dumps({"metric": "prior_query_points_per_s", "value": total_pts / t, "unit": "points/s",
                          "n_gpus": world, "tiles": args.tiles, "points_per_tile": M, "ms_per_tile": per_tile_ms,
                          "algorithmic_GBps_reference_count": faithful * total_pts / t / 1e9 / world,
                          "algorithmic_GBps_elided_count": elided * total_pts / t / 1e9 / world,
                          "scaling": "tiles sharded over ranks, no communication",
                          "dtype": "f32 hash + " + ("tf32x3" if args.fp32 else "bf16") + " MLP, fp16 features"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

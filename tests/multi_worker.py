"""Worker of tests/test_gpu_multi.py, launched with torchrun on >= 2 GPUs of one box (NCCL).

  grads     data-parallel parity: every rank runs the C2-shaped step on its shard of the rays and the
            gradients are averaged by GradSynchronizer (NCCL all-reduce); rank 0 then runs the CONCATENATED batch alone
            and the two sets of gradients must agree.
  partial   the same parity with the main hash table's gradient exchanged level group by level group from inside the backward
            (GradSynchronizer(partial_tables=...), proposal levels on their single-GPU side-stream schedule).
  peer      the partial exchange on the copy engines between IPC-mapped buffers (peer_exchange.py) instead of NCCL.
  peerk     the same exchange as two kernels with P2P stores (PS_PEER_MODE=kernel).
  sharded   ShardedFusedAdam (reduce-scatter -> Adam on the shard -> all-gather) against GradSynchronizer + FusedAdam:
            same parameters after several steps, on every rank.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make(n_rays, log2_main=18):
    from presight_b200 import synthetic
    from presight_b200.model import NerfactoNuscMSModel
    cfg = synthetic.config_c2("b200")
    cfg.log2_hashmap_size = log2_main
    torch.manual_seed(42)
    host = synthetic.make_rays(n_rays, seed=11)
    model = NerfactoNuscMSModel(cfg, torch.zeros(1, 3), synthetic.tile_aabb(), host["n_cameras"], host["n_videos"])
    with torch.no_grad():
        for f in model.field.fields:
            f.mlp_base_grid.hash_table.mul_(300.0)
        for p in model.proposal_networks:
            for f in p.fields:
                f.encoding.hash_table.mul_(300.0)
    return model, cfg, host


def step_grads(model, host, lo, hi, jit, dev):
    from presight_b200 import losses
    from presight_b200.cameras.rays import RayBundle
    from presight_b200.model import VIDEO_ID
    for p in model.parameters():
        p.grad = None
    sl = slice(lo, hi)
    rb = RayBundle(origins=host["origins"][sl].to(dev), directions=host["directions"][sl].to(dev),
                   camera_indices=host["camera_indices"][sl].to(dev), metadata={VIDEO_ID: host["video_ids"][sl].to(dev)})
    model.proposal_sampler._step = 0
    out = model(rb, jitters=[j[sl].to(dev) for j in jit])
    # per-ray means only (the expected depth is clipped to the BATCH's min / max sample distance, renderers.py:377-379,
    # which a shard cannot know — it is left out of this loss on purpose)
    loss = ((out["rgb"] - host["rgb"][sl].to(dev)) ** 2).mean() \
        + 0.5 * ((out["semantics"] - host["features"][sl].to(dev).clip(0, 1)) ** 2).mean() \
        + losses.z_anti_aliasing_interlevel_loss(out["weights_list"], [rs.sp_bins for rs in out["ray_samples_list"]]) \
        + 0.002 * losses.distortion_loss(out["weights_list"], [rs.sp_bins for rs in out["ray_samples_list"]])
    loss.backward()
    return loss


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    from presight_b200 import fused
    from presight_b200.parallel import GradSynchronizer, init_nccl, shard_range
    init_nccl(dev)
    fused.set_overlap_prop_bwd(mode in ("partial", "peer", "peerk"))
    n = 4096
    model, cfg, host = make(n)
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    g = torch.Generator().manual_seed(5)
    jit = [torch.rand(n, 1, generator=g) for _ in range(3)]
    lo, hi = shard_range(n, rank, world)
    ok = True
    if mode in ("grads", "partial", "peer", "peerk"):
        partial = []
        if mode == "peerk":
            os.environ["PS_PEER_MODE"] = "kernel"
        if mode in ("partial", "peer", "peerk"):
            from presight_b200.parallel import level_groups
            enc = model.field.fields[0].mlp_base_grid
            partial = [(enc.hash_table, level_groups(enc.num_levels, world=world))]
        sync = GradSynchronizer(params, overlap=True, partial_tables=partial, peer=mode in ("peer", "peerk"))
        step_grads(model, host, lo, hi, jit, dev)
        sync.finish()
        torch.cuda.synchronize()
        dp = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        sync.remove()
        dist.barrier()
        if rank == 0:
            step_grads(model, host, 0, n, jit, dev)
            worst = ("", 0.0)
            for k, p in model.named_parameters():
                if p.grad is None:
                    continue
                e = rel_l2(dp[k], p.grad)
                if e > worst[1]:
                    worst = (k, e)
            print(f"MULTI {mode} world={world} worst rel-L2 {worst[1]:.3e} ({worst[0]}) over {len(dp)} tensors", flush=True)
            ok = worst[1] < 1e-4
    elif mode == "sharded":
        from presight_b200.optim import FusedAdam, ShardedFusedAdam
        import copy
        model_b = copy.deepcopy(model)
        params_b = [p for p in model_b.parameters() if p.requires_grad]
        opt_a = FusedAdam(params, lr=1e-2, eps=1e-15, weight_decay=1e-5)
        sync = GradSynchronizer(params, overlap=False)
        opt_b = ShardedFusedAdam(params_b, lr=1e-2, eps=1e-15, weight_decay=1e-5)
        for it in range(3):
            step_grads(model, host, lo, hi, jit, dev)
            sync.finish()
            opt_a.step()
            step_grads(model_b, host, lo, hi, jit, dev)
            opt_b.step()
        torch.cuda.synchronize()
        worst = ("", 0.0)
        for (k, pa), pb in zip(model.named_parameters(), model_b.parameters()):
            e = rel_l2(pb.detach(), pa.detach())
            if e > worst[1]:
                worst = (k, e)
        # every rank must also hold the same parameters
        chk = torch.stack([p.detach().double().sum() for p in params_b])
        ref = chk.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(chk, ref))
        print(f"MULTI sharded world={world} rank={rank} worst rel-L2 vs all-reduce+Adam {worst[1]:.3e} ({worst[0]}) "
              f"replicas identical {same}", flush=True)
        ok = worst[1] < 1e-4 and same
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag) else 0)


if __name__ == "__main__":
    main()

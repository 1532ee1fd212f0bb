"""Host-side data-parallel logic on CPU: world_size-2 gloo processes (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from presight_b200.parallel import GradSynchronizer, shard_range
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.zeros(1 << 17))          # "hash table": async path
    small = [torch.nn.Parameter(torch.zeros(64, 8)), torch.nn.Parameter(torch.zeros(64))]
    unused = torch.nn.Parameter(torch.zeros(1 << 17))       # a sub-field no ray of this rank visited
    params = [big, *small, unused]
    sync = GradSynchronizer(params, overlap=overlap, min_async_numel=1 << 16)
    # every rank gets its own shard of "rays"
    lo, hi = shard_range(1000, rank, world)
    x = torch.arange(lo, hi, dtype=torch.float32)
    loss = (big[: x.numel()] * x).sum() + (small[0].sum() + small[1].sum()) * (rank + 1)
    if not overlap and rank == 1:
        loss = loss + unused.sum() * 3.0                    # only rank 1 touches the "unused" parameter
    loss.backward()
    sync.finish()
    out = {"big": big.grad.clone(), "s0": small[0].grad.clone(), "s1": small[1].grad.clone(),
           "unused": None if unused.grad is None else unused.grad.clone(), "range": (lo, hi)}
    torch.save(out, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_grad_synchronizer_gloo_world2(overlap, tmp_path):
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = {r: torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(2)}
    assert res[0]["range"] == (0, 500) and res[1]["range"] == (500, 1000)
    # mean over ranks of the per-rank gradients
    want_big = torch.zeros(1 << 17)
    want_big[:500] = (torch.arange(0, 500) + torch.arange(500, 1000)).float() / 2
    for r in range(2):
        assert torch.allclose(res[r]["big"], want_big)
        assert torch.allclose(res[r]["s0"], torch.full((64, 8), 1.5))
        assert torch.allclose(res[r]["s1"], torch.full((64,), 1.5))
        if not overlap:
            assert torch.allclose(res[r]["unused"], torch.full((1 << 17,), 1.5))   # zeros on rank 0, 3 on rank 1


def test_shard_range_covers_everything():
    from presight_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1000, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]


def _partial_worker(rank, world, port, outdir):
    """The partial (level group by level group) exchange of a table gradient, host logic only: the fused backward's calls into
    the registered sink are replayed by hand on CPU tensors under gloo."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from presight_b200 import fused
    from presight_b200.parallel import GradSynchronizer, level_groups
    L, T, F = 16, 256, 2
    table = torch.nn.Parameter(torch.zeros(L * T, F))
    other = torch.nn.Parameter(torch.zeros(1 << 16))
    groups = level_groups(L, world=world)
    sync = GradSynchronizer([table, other], overlap=True, min_async_numel=1 << 12, partial_tables=[(table, groups)])
    fn, reg_groups, alloc = fused._PARTIAL_SINKS[table.data_ptr()]
    assert reg_groups == groups and alloc is None
    # "backward": the scatter fills the gradient buffer group by group and hands every group to the sink at once
    g = torch.Generator().manual_seed(rank)
    dtable = torch.zeros(L * T, F)
    full = torch.randn(L * T, F, generator=g)
    for l0, l1 in groups:
        dtable[l0 * T:l1 * T] = full[l0 * T:l1 * T]
        fn(dtable, l0 * T, l1 * T)
    table.grad = dtable                       # autograd adopts the kernel's buffer (grad was None)
    sync._on_grad_ready(table)                # ... and the post-accumulate hook must not reduce it a second time
    other.grad = torch.full((1 << 16,), float(rank + 1))
    sync._on_grad_ready(other)
    sync.finish()
    torch.save({"table": table.grad.clone(), "other": other.grad.clone(), "local": full}, os.path.join(outdir, f"p{rank}.pt"))
    # a gradient that was accumulated into an older buffer cannot be exchanged piecewise: loud error
    err = None
    fn(dtable, 0, groups[0][1] * T)
    table.grad = dtable.clone()
    try:
        sync._on_grad_ready(table)
    except RuntimeError as e:
        err = str(e)
    sync.finish()
    assert err is not None and "None before backward" in err
    sync.remove()
    assert table.data_ptr() not in fused._PARTIAL_SINKS
    dist.barrier()
    dist.destroy_process_group()


def test_partial_table_exchange_gloo_world2(tmp_path):
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_partial_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = {r: torch.load(os.path.join(str(tmp_path), f"p{r}.pt")) for r in range(2)}
    want = (res[0]["local"] + res[1]["local"]) / 2
    for r in range(2):
        assert torch.allclose(res[r]["table"], want, atol=1e-7)
        assert torch.allclose(res[r]["other"], torch.full((1 << 16,), 1.5))


def test_level_groups_cover_all_levels():
    from presight_b200.parallel import level_groups
    for L in (1, 2, 5, 8, 10, 16, 24):
        for world in (2, 8):
            gs = level_groups(L, world=world)
            assert gs[0][0] == 0 and gs[-1][1] == L and all(a[1] == b[0] for a, b in zip(gs, gs[1:]))
            assert all(b > a for a, b in gs)
    assert level_groups(16, cuts=[4, 8]) == [(0, 4), (4, 8), (8, 16)]
    assert level_groups(16, world=8)[-1] == (15, 16)          # little left in flight when the backward ends

"""Data-parallel gradient exchange (reference: DDP wrap at pipelines/PreSight/my_pipeline.py:121-124).

Rays are sharded across ranks (data/PreSight/my_datamanager.py:206-212); the only exchange step is the
all-reduce of the dense fp32 gradients (hash tables + MLPs).  Instead of DDP's 25 MB buckets, every large
parameter's all-reduce is launched from a post-accumulate-grad hook the moment autograd has finished that
parameter, so the 512 MiB main-table reduction (finished first in backward) travels over NVLink while the proposal
networks' backward is still running.

Two modes:
  overlap=True   hooks + async all-reduce.  Requires that every rank produces gradients for the same set of
                 parameters in the same order (single sub-field models; update / non-update steps are decided by
                 the step counter, identically on all ranks).
  overlap=False  one fixed-order pass in `finish()`; parameters without a gradient on this rank contribute zeros —
                 the behaviour DDP's `find_unused_parameters=True` gives the reference when a rank's rays miss a
                 sub-field (fields/PreSight/ingp_field_ms.py:103).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import nn


def init_nccl(device: torch.device, high_priority: bool = True, **kw) -> None:
    """`dist.init_process_group("nccl")` with NCCL's kernels on a HIGH-PRIORITY stream.  The all-reduce of the main-table
    gradient is meant to travel under the proposal levels' backward; those are persistent kernels that fill every SM,
    and an NCCL CTA (up to 640 threads x ~96 registers) needs an almost empty SM.  At default priority the block
    scheduler refills retiring SMs with the next compute kernel and the all-reduce starts only when compute ends; at
    high priority NCCL's few CTAs get the first SMs that drain and the transfer overlaps the rest."""
    # Protocol: on the NVLink-only (no NVLS multicast) B200 boxes measured here NCCL's tuner moves the 512 MiB table gradient
    # with the Simple protocol; LL128 is 0.67 ms per step faster at 4 GPUs (13.20 vs 13.87 ms, tools/nccl_variants.sh).
    # An NCCL_PROTO already in the environment wins.
    # At 2 GPUs the default (LL) is as fast (12.44 vs 12.57 ms), so it is left alone there.
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 2:
        os.environ.setdefault("NCCL_PROTO", "LL128")
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=high_priority)
    dist.init_process_group("nccl", device_id=device, pg_options=opts, **kw)


class GradSynchronizer:
    """Average gradients across ranks.   sync = GradSynchronizer(params); loss.backward(); sync.finish()"""

    def __init__(self, params: Iterable[nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 overlap: bool = True, min_async_numel: int = 1 << 16, partial_tables=(), peer: bool = False) -> None:
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap
        self.min_async_numel = min_async_numel
        # gloo has no AVG: sum, then scale
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._pending = []       # (work, grad)
        self._done = set()
        self._handles = []
        self._partial = {}
        self._partial_tables = []
        self._peer = None
        self._peer_names = {}
        if self.world > 1 and overlap:
            # (Scheduling note: the main table's all-reduce can only hide under compute that runs after the final level's
            # backward, so a data-parallel caller wants the proposal levels' backward on the main stream —
            # `fused.set_overlap_prop_bwd(False)`, which bench.py calls for world > 1.  This class does not touch it.)
            for p in self.params:
                if p.numel() >= min_async_numel:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
            # `partial_tables`: (hash-table parameter, level groups) pairs whose gradient is reduced level group by level
            # group from inside the backward (fused.register_partial_grad_sink).  Needs `p.grad is None` before backward,
            # so that autograd adopts the kernel's gradient buffer instead of accumulating into an older one.
            # `peer=True`: those pieces travel on the copy engines between IPC-mapped buffers (peer_exchange.py) instead of
            # NCCL all-reduces, and the table's gradient lives in the exchange's own buffer.
            if peer and partial_tables:
                from .peer_exchange import PeerExchange
                specs = {}
                for i, (p, groups) in enumerate(partial_tables):
                    T = p.shape[0] // max(g[1] for g in groups)       # rows per level
                    specs[f"t{i}"] = (p.shape[0], p.shape[1], [(l0 * T, l1 * T) for l0, l1 in groups])
                    self._peer_names[id(p)] = f"t{i}"
                self._peer = PeerExchange(specs, self.params[0].device, group)
            for p, groups in partial_tables:
                from . import fused
                alloc = None
                if self._peer is not None:
                    def alloc(name=self._peer_names[id(p)]):
                        buf = self._peer.grad_buffer(name)
                        buf.zero_()
                        return buf
                fused.register_partial_grad_sink(p, lambda dtable, lo, hi, p=p: self._on_partial(p, dtable, lo, hi), groups, alloc)
                self._partial_tables.append(p)

    def _on_partial(self, p: nn.Parameter, dtable: torch.Tensor, row_lo: int, row_hi: int) -> None:
        # (runs with the scatter's stream current: the exchange is ordered behind the kernel that produced these rows)
        if self._peer is not None:
            name = self._peer_names[id(p)]
            assert dtable.data_ptr() == self._peer.grad_ptr(name), "peer exchange: gradient not in the exchange's buffer"
            k = [g[0] for g in self._peer.specs[name][2]].index(row_lo)
            self._peer.exchange(name, k)
            self._partial[id(p)] = dtable.data_ptr()
            return
        # (a fresh tensor over the same storage, not a view: a view would keep a reference to `dtable` and autograd would
        # then copy the gradient instead of adopting the buffer)
        F = dtable.shape[1]
        piece = torch.empty(0, device=dtable.device, dtype=dtable.dtype).set_(
            dtable.untyped_storage(), dtable.storage_offset() + row_lo * F, (row_hi - row_lo, F), (F, 1))
        self._pending.append((dist.all_reduce(piece, op=self._op, group=self.group, async_op=True), piece))
        self._partial[id(p)] = dtable.data_ptr()

    def _on_grad_ready(self, p: nn.Parameter) -> None:
        if p.grad is None:
            return
        if id(p) in self._partial:
            # already travelling piece by piece; the pieces were reduced in place in the buffer autograd must have adopted
            if p.grad.data_ptr() != self._partial.pop(id(p)):
                raise RuntimeError("partial gradient exchange: set the hash table's .grad to None before backward "
                                   "(its gradient was accumulated into an older buffer, the reduced pieces are lost)")
            self._done.add(id(p))
            return
        self._pending.append((dist.all_reduce(p.grad, op=self._op, group=self.group, async_op=True), p.grad))
        self._done.add(id(p))

    def finish(self) -> None:
        """Reduce everything not already in flight (as one flat buffer per dtype) and wait."""
        if self.world <= 1:
            return
        rest = [p for p in self.params if id(p) not in self._done]
        if self.overlap:
            rest = [p for p in rest if p.grad is not None]        # same set on every rank by contract
        big = [p for p in rest if p.numel() >= self.min_async_numel]
        small = [p for p in rest if p.numel() < self.min_async_numel]
        for p in big:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            self._pending.append((dist.all_reduce(p.grad, op=self._op, group=self.group, async_op=True), p.grad))
        if small:
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in small])
            dist.all_reduce(flat, op=self._op, group=self.group)
            if not self._avg:
                flat.div_(self.world)
            off = 0
            for p in small:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[off:off + n].view_as(p))
                off += n
        if getattr(self, "_peer", None) is not None:
            self._peer.finish()
        for work, grad in self._pending:
            work.wait()
            if not self._avg:
                grad.div_(self.world)
        self._pending.clear()
        self._done.clear()

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
        for p in getattr(self, "_partial_tables", []):
            from . import fused
            fused.unregister_partial_grad_sink(p)
        if getattr(self, "_peer", None) is not None:
            self._peer.close()
            self._peer = None


def level_groups(num_levels: int, cuts=None, world: int = 8):
    """Level groups of the partial exchange of a hash-table gradient: a small first group (the exchange starts early), small
    last groups (little is still in flight when the backward ends).  `cuts`: explicit group boundaries.  Defaults for 16
    levels: (2, 5, 9, 13, 15) — 13.02 ms/step at 8 GPUs against 13.28 with (6, 11, 14), profiles/r2_e_n8_variants.txt; at 2
    ranks (NCCL's LL protocol, few large messages are better) the coarser (6, 11, 14): 11.94 against 12.3 ms."""
    if cuts is None:
        fr = (0.375, 0.6875, 0.875) if world <= 2 else (0.125, 0.3125, 0.5625, 0.8125, 0.9375)
        cuts = sorted({round(num_levels * f) for f in fr} - {0, num_levels})
    edges = [0, *[c for c in cuts if 0 < c < num_levels], num_levels]
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `n_units` independent units (rays, tiles) for `rank` — no communication."""
    per = (n_units + world - 1) // world
    lo = min(rank * per, n_units)
    return lo, min(lo + per, n_units)

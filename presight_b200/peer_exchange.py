"""Gradient exchange over peer memory, on the copy engines (one node; csrc/peer_exchange.cu).

The reference averages gradients with DDP's bucketed NCCL all-reduce (pipelines/PreSight/my_pipeline.py:121-124).  For the
512 MiB main-table gradient that costs SMs exactly when the backward has none to spare (parallel.py).  Here every rank's
gradient buffer G, a staging buffer and a flag array are plain device allocations exported to the other ranks of the node
through CUDA IPC, and a row range [lo, hi) of G is averaged with copies only:

  reduce-scatter   rank q copies, for every peer r, its rows of r's shard into r's staging slot q, then the step number
                   into r's flag (q, phase 1);
  reduce           rank r waits for the phase-1 flags of the range (one polling warp), then
                   G[shard r] = (G[shard r] + sum of the staged pieces) / world   (ps_peer_reduce);
  all-gather       rank r copies its reduced shard into every peer's G, then the step number into the phase-2 flag.

The three phases run either as copy-engine transfers (mode "dma": no SM is touched, but every transfer and every flag is its
own DMA operation; one stream per destination keeps them running side by side — the default) or as two kernels of a few small
CTAs that write to the peers with P2P stores (mode "kernel", ps_peer_exchange_range: measured slower — 32 CTAs do not fill
NVLink, more CTAs take the SMs the backward needs).

`finish()` waits for all phase-2 flags.  Everything is stream-ordered on one communication stream; nothing synchronises
the hosts.  Buffers are reused every step: a rank starts pushing step t+1 only after its own finish() of step t, which has
seen every peer's phase-2 flag — and a peer raises that flag after it has consumed the staged pieces of step t.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import call


class _DevArray:
    """Zero-copy torch view of a raw device allocation (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _alloc(nbytes: int) -> int:
    p = C.c_void_p()
    call("ps_peer_alloc", nbytes, C.byref(p))
    return p.value


def _export(ptr: int) -> bytes:
    h = (C.c_char * 64)()
    call("ps_peer_export", ptr, h)
    return bytes(h)


def _open(handle: bytes) -> int:
    p = C.c_void_p()
    call("ps_peer_open", (C.c_char * 64).from_buffer_copy(handle), C.byref(p))
    return p.value


class PeerExchange:
    """Averages fp32 gradient buffers across the ranks of one node.  `specs`: {name: (rows, cols, [(row_lo, row_hi), ...])} —
    the row ranges are exchanged one by one (`exchange(name, k)`), each split evenly over the ranks."""

    def __init__(self, specs: Dict[str, Tuple[int, int, Sequence[Tuple[int, int]]]], device: torch.device,
                 group=None, mode: str = None, ctas: int = None) -> None:
        import os
        self.mode = mode or os.environ.get("PS_PEER_MODE", "dma")
        self.ctas = ctas or int(os.environ.get("PS_PEER_CTAS", "32"))
        assert self.mode in ("kernel", "dma")
        assert dist.is_initialized()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device(device)
        self.specs = {k: (int(r), int(c), [tuple(map(int, g)) for g in gs]) for k, (r, c, gs) in specs.items()}
        for name, (rows, cols, groups) in self.specs.items():
            for lo, hi in groups:
                if (hi - lo) % self.world or ((hi - lo) // self.world * cols) % 4:
                    raise ValueError(f"peer exchange: rows [{lo}, {hi}) x {cols} of {name} do not split into {self.world} "
                                     "16-byte-aligned shards")
        self.names = list(self.specs)
        self.max_groups = max(len(g) for _, _, g in self.specs.values())
        # one allocation per rank: [G and staging of every buffer | flags]
        self.offsets, off = {}, 0
        for name in self.names:
            rows, cols, _ = self.specs[name]
            nbytes = (rows * cols * 4 + 255) // 256 * 256
            self.offsets[name] = (off, off + nbytes)            # (G, staging)
            off += 2 * nbytes
        self.flags_off = off
        self.n_flags = 2 * len(self.names) * self.max_groups * self.world
        total = off + self.n_flags * 4
        self.base = _alloc(total)
        self._holder = _DevArray(self.base, total)
        self._bytes = torch.as_tensor(self._holder, device=self.dev)
        self._bytes.zero_()
        torch.cuda.synchronize(self.dev)
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, _export(self.base), group=group)
        self.peer_base = [self.base if r == self.rank else _open(handles[r]) for r in range(self.world)]
        dist.barrier(group)
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)
        self.push = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(self.world - 1)]
        self.cur = torch.zeros(1, dtype=torch.int32, device=self.dev)      # the step number, source of every flag copy
        self.counters = torch.zeros(2 * len(self.names) * self.max_groups, dtype=torch.int32, device=self.dev)
        self._bases = (C.c_void_p * self.world)(*self.peer_base)
        self.step = 0
        self._open_step = False

    # ---- addressing ----------------------------------------------------------------------------------------------
    def _flag(self, base: int, phase: int, buf: int, g: int, src_rank: int) -> int:
        idx = ((phase * len(self.names) + buf) * self.max_groups + g) * self.world + src_rank
        return base + self.flags_off + idx * 4

    def grad_buffer(self, name: str) -> torch.Tensor:
        """This rank's gradient buffer [rows, cols] (NOT zeroed; a fresh tensor object every call, so that autograd can adopt
        it as the parameter's .grad without a copy)."""
        rows, cols, _ = self.specs[name]
        g_off, _ = self.offsets[name]
        return self._bytes[g_off:g_off + rows * cols * 4].view(torch.float32).view(rows, cols)

    def grad_ptr(self, name: str) -> int:
        return self.base + self.offsets[name][0]

    # ---- protocol ------------------------------------------------------------------------------------------------
    def _begin(self) -> None:
        if not self._open_step:
            self.step += 1
            # (the previous step's flag copies read `cur`: finish() made the caller's stream wait for them, so order the
            # update behind the caller's stream — a flag that carried the next step's number would never match)
            self.comm.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.comm):
                self.cur.fill_(self.step)
            filled = torch.cuda.Event()
            filled.record(self.comm)
            for ps in self.push:
                ps.wait_event(filled)
            self._open_step = True

    def exchange(self, name: str, k: int) -> None:
        """Average row range k of buffer `name`.  Call with the stream that produced those rows current."""
        self._begin()
        rows, cols, groups = self.specs[name]
        lo, hi = groups[k]
        buf = self.names.index(name)
        g_off, s_off = self.offsets[name]
        shard_rows = (hi - lo) // self.world
        shard_bytes = shard_rows * cols * 4
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))
        self.comm.wait_event(ready)
        st = self.comm.cuda_stream
        me, cur = self.rank, self.cur.data_ptr()
        self._used = getattr(self, "_used", set())
        self._used.add((buf, k))
        if self.mode == "kernel":
            f1 = self._flag(0, 0, buf, k, 0)
            f2 = self._flag(0, 1, buf, k, 0)
            call("ps_peer_exchange_range", self._bases, self.world, me, g_off, s_off, lo * cols * 4, shard_bytes, f1, f2, self.step,
                 1.0 / self.world, self.counters.data_ptr() + 8 * (buf * self.max_groups + k), self.ctas, st)
            return

        def g_addr(base, r):          # rows of rank r's shard inside the range, in the gradient buffer at `base`
            return base + g_off + (lo + r * shard_rows) * cols * 4

        def s_addr(base, q):          # staging slot of source rank q for this range
            return base + s_off + lo * cols * 4 + q * shard_bytes

        # One stream per destination: the transfers to different peers (and their flags) run side by side on the copy engines
        # instead of paying their start-up latencies one after the other.
        # reduce-scatter: my rows of every peer's shard -> the peer's staging slot `me`, then the flag
        for d in range(1, self.world):
            r = (me + d) % self.world
            ps = self.push[d - 1]
            ps.wait_event(ready)
            call("ps_peer_copy", s_addr(self.peer_base[r], me), g_addr(self.base, r), shard_bytes, ps.cuda_stream)
            call("ps_peer_copy", self._flag(self.peer_base[r], 0, buf, k, me), cur, 4, ps.cuda_stream)
        call("ps_peer_copy", self._flag(self.base, 0, buf, k, me), cur, 4, st)
        # reduce my shard
        call("ps_peer_wait_flags", self._flag(self.base, 0, buf, k, 0), self.world, 1, self.step, st)
        srcs = (C.c_void_p * (self.world - 1))(*[s_addr(self.base, q) for q in range(self.world) if q != me])
        call("ps_peer_reduce", g_addr(self.base, me), srcs, self.world - 1, shard_rows * cols, 1.0 / self.world, st)
        reduced = torch.cuda.Event()
        reduced.record(self.comm)
        # all-gather: my reduced shard -> every peer's gradient buffer, then the flag
        for d in range(1, self.world):
            r = (me + d) % self.world
            ps = self.push[d - 1]
            ps.wait_event(reduced)
            call("ps_peer_copy", g_addr(self.peer_base[r], me), g_addr(self.base, me), shard_bytes, ps.cuda_stream)
            call("ps_peer_copy", self._flag(self.peer_base[r], 1, buf, k, me), cur, 4, ps.cuda_stream)
        call("ps_peer_copy", self._flag(self.base, 1, buf, k, me), cur, 4, st)

    def finish(self) -> None:
        """Every range exchanged this step is complete in this rank's buffers when the current stream passes this point."""
        if not self._open_step:
            return
        st = self.comm.cuda_stream
        for (buf, k) in sorted(self._used):
            call("ps_peer_wait_flags", self._flag(self.base, 1, buf, k, 0), self.world, 1, self.step, st)
        self._used.clear()
        main = torch.cuda.current_stream(self.dev)
        main.wait_stream(self.comm)
        for ps in self.push:
            main.wait_stream(ps)          # my own transfers have left this rank's buffers (they are rewritten next step)
        self._open_step = False

    def close(self) -> None:
        torch.cuda.synchronize(self.dev)
        dist.barrier(self.group)
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                call("ps_peer_close", p)
        dist.barrier(self.group)
        self._bytes = None
        call("ps_peer_free", self.base)

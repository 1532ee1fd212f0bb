"""Batch assembly with the pixel chunk resident in HBM.

Reference (data/PreSight/): `MyDataset.load_chunk` builds an `ImageChunk` of pixels on the HOST (my_dataset.py:27-73, 165-330);
`MyDataManager._get_train_batch_loader` wraps it in `DistributedSampler(chunk, world_size, local_rank)` +
`DataLoader(batch_size = train_num_rays_per_batch // world_size, drop_last=True)` (my_datamanager.py:203-219) and
`next_train_image` copies every batch to the GPU and turns its ray indices into a RayBundle (:257-285).

Here the chunk is uploaded once (a nuScenes chunk is a few GB; a B200 has 180) and a step's batch is ONE gather kernel
(`ps_assemble_batch`) plus the ray-generation kernel: no worker processes, no per-step host work, no per-step H2D copy.
The order of the batches is the reference's, bit for bit: the sampler's permutation is reproduced with the same generator
(`torch.randperm(n, generator=manual_seed(seed + epoch))`, padded and strided by rank as DistributedSampler does) and
uploaded once per chunk.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterator, Optional, Tuple

import torch
from torch import Tensor

from .._lib import call, ptr, stream

IMAGE_INDEX, PIXEL_INDEX, RGB, DEPTH, FEATURES, RAY_INDEX = "image_index", "pixel_index", "rgb", "depth", "features", "ray_index"
WIDTH, VIDEO_ID, SEG, SKY = "width", "video_id", "seg", "sky"


@dataclass
class ImageChunk:
    """The reference's ImageChunk fields (my_dataset.py:28-47), any device."""
    rgbs: Tensor                 # [n,3] float32
    segs: Optional[Tensor]       # [n] uint8
    skies: Tensor                # [n] float32 (1 = sky)
    depths: Tensor               # [n] float32
    features: Optional[Tensor]   # [n,C] float32
    pixel_indices: Tensor        # [n] int64
    image_indices: Tensor
    video_ids: Tensor
    widths: Tensor

    def __len__(self) -> int:
        return self.rgbs.shape[0]

    def to(self, device) -> "ImageChunk":
        def mv(t, dt):
            return None if t is None else t.to(device=device, dtype=dt).contiguous()
        return ImageChunk(mv(self.rgbs, torch.float32), mv(self.segs, torch.uint8), mv(self.skies, torch.float32),
                          mv(self.depths, torch.float32), mv(self.features, torch.float32), mv(self.pixel_indices, torch.int64),
                          mv(self.image_indices, torch.int64), mv(self.video_ids, torch.int64), mv(self.widths, torch.int64))


def sampler_indices(n: int, rank: int, world: int, seed: int = 0, epoch: int = 0) -> Tensor:
    """The index sequence torch.utils.data.DistributedSampler(dataset of length n, world, rank, shuffle=True, drop_last=False)
    yields (the reference constructs it with exactly these defaults and never calls set_epoch)."""
    g = torch.Generator()
    g.manual_seed(seed + epoch)
    indices = torch.randperm(n, generator=g)
    num_samples = math.ceil(n / world)
    total = num_samples * world
    pad = total - n
    if pad > 0:
        reps = math.ceil(pad / n)
        indices = torch.cat([indices, indices.repeat(reps)[:pad]]) if pad > n else torch.cat([indices, indices[:pad]])
    return indices[rank:total:world].contiguous()


class DeviceBatchLoader:
    """Iterates over the batches the reference's DataLoader would produce from `chunk` on this rank, assembled on the device.

        loader = DeviceBatchLoader(chunk, batch_size, rank, world, device, ray_generator, pose_scale_factor)
        for ray_bundle, batch in loader: ...          # one pass over the chunk (drop_last=True), then StopIteration
    """

    def __init__(self, chunk: ImageChunk, batch_size: int, rank: int, world: int, device, ray_generator=None,
                 pose_scale_factor: float = 1.0, seed: int = 0, epoch: int = 0) -> None:
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DeviceBatchLoader needs a CUDA device (no CPU fallback exists)")
        self.chunk = chunk.to(self.dev)
        self.batch_size = int(batch_size)
        self.indices = sampler_indices(len(chunk), rank, world, seed, epoch).to(self.dev)
        self.n_batches = self.indices.numel() // self.batch_size          # drop_last=True
        self.ray_generator = ray_generator
        self.pose_scale_factor = float(pose_scale_factor)
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._b = 0

    def __len__(self) -> int:
        return self.n_batches

    def __iter__(self) -> Iterator:
        self._b = 0
        return self

    def assemble(self, idx: Tensor) -> Dict[str, Tensor]:
        c, B = self.chunk, idx.numel()
        dev = self.dev
        out = {RGB: torch.empty(B, 3, device=dev), SKY: torch.empty(B, device=dev), DEPTH: torch.empty(B, device=dev),
               IMAGE_INDEX: torch.empty(B, dtype=torch.int64, device=dev), VIDEO_ID: torch.empty(B, dtype=torch.int64, device=dev),
               RAY_INDEX: torch.empty(B, 3, dtype=torch.int64, device=dev)}
        if c.segs is not None:
            out[SEG] = torch.empty(B, dtype=torch.uint8, device=dev)
        C = 0
        if c.features is not None:
            C = c.features.shape[1]
            out[FEATURES] = torch.empty(B, C, device=dev)
        call("ps_assemble_batch", ptr(c.rgbs), ptr(c.segs), ptr(c.skies), ptr(c.depths), ptr(c.features), C,
             ptr(c.pixel_indices), ptr(c.image_indices), ptr(c.video_ids), ptr(c.widths), len(c), ptr(idx), B, ptr(out[RGB]),
             ptr(out.get(SEG)), ptr(out[SKY]), ptr(out[DEPTH]), ptr(out.get(FEATURES)), ptr(out[IMAGE_INDEX]), ptr(out[VIDEO_ID]),
             ptr(out[RAY_INDEX]), ptr(self._bad), stream())
        return out

    def __next__(self) -> Tuple[object, Dict[str, Tensor]]:
        if self._b >= self.n_batches:
            raise StopIteration
        idx = self.indices[self._b * self.batch_size:(self._b + 1) * self.batch_size]
        self._b += 1
        batch = self.assemble(idx)
        if self.ray_generator is None:
            return None, batch
        rb = self.ray_generator(batch[RAY_INDEX])
        rb.metadata[VIDEO_ID] = batch[VIDEO_ID].view(-1, 1)
        rb.metadata["pose_scale_factor"] = torch.full((idx.numel(), 1), self.pose_scale_factor, device=self.dev)
        return rb, batch

    def check(self) -> None:
        """Raises if any index so far fell outside the chunk (one host read; the reference's indexing raises at once)."""
        if int(self._bad.item()):
            raise IndexError("DeviceBatchLoader: sampler index outside the chunk")

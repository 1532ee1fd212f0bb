#!/usr/bin/env python
"""Time of the device-side batch assembly (ps_assemble_batch + ps_generate_rays) for one training batch, against the bytes
it has to move: chunk of 16 M pixels with 64-d features resident in HBM, 65 536 rays per batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
from presight_b200.cameras.ray_generator import RayGenerator
from presight_b200.data import DeviceBatchLoader, ImageChunk

n, B, C = 16 * 1024 * 1024, 65536, 64
g = torch.Generator().manual_seed(0)
chunk = ImageChunk(rgbs=torch.rand(n, 3, generator=g), segs=torch.randint(0, 19, (n,), generator=g, dtype=torch.uint8),
                   skies=torch.zeros(n), depths=torch.rand(n, generator=g), features=torch.randn(n, C, generator=g),
                   pixel_indices=torch.randint(0, 1600 * 900, (n,), generator=g), image_indices=torch.randint(0, 6, (n,), generator=g),
                   video_ids=torch.randint(0, 7, (n,), generator=g), widths=torch.full((n,), 1600, dtype=torch.int64))
c2w = torch.eye(4)[:3].repeat(6, 1, 1)
gen = RayGenerator(c2w, torch.full((6,), 1200.0), torch.full((6,), 1200.0), torch.full((6,), 800.0), torch.full((6,), 450.0)).cuda()
loader = DeviceBatchLoader(chunk, B, 0, 1, "cuda", ray_generator=gen)
it = iter(loader)
for _ in range(5):
    next(it)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k = 50
a.record()
for _ in range(k):
    next(it)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / k
row_bytes = 12 + 1 + 4 + 4 + C * 4 + 4 * 8          # read per gathered row
out_bytes = 12 + 1 + 4 + 4 + C * 4 + 8 + 8 + 24 + 12 + 12 + 4 + 4   # batch + ray bundle written
print(json.dumps({"batch_assembly_ms": ms, "rays": B, "chunk_pixels": n, "feature_channels": C,
                  "bytes_per_ray": row_bytes + out_bytes, "GBps": (row_bytes + out_bytes) * B / ms / 1e6,
                  "note": "gather of random rows: every 4..256-byte field of a row is its own 32-byte sector(s)"}))

"""Batch assembly: the device loader against the reference's DataLoader / DistributedSampler pipeline."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.batching_oracle import ImageChunkRef, reference_batches  # noqa: E402


def make_chunk(n, C=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    widths = torch.full((n,), 1600, dtype=torch.int64)
    widths[n // 2:] = 800                        # two camera resolutions
    return dict(rgbs=torch.rand(n, 3, generator=g), segs=torch.randint(0, 19, (n,), generator=g, dtype=torch.uint8),
                skies=(torch.rand(n, generator=g) < 0.2).float(), depths=torch.rand(n, generator=g) * 80,
                features=torch.randn(n, C, generator=g), pixel_indices=torch.randint(0, 800 * 450, (n,), generator=g),
                image_indices=torch.randint(0, 240, (n,), generator=g), video_ids=torch.randint(0, 7, (n,), generator=g),
                widths=widths)


@pytest.mark.parametrize("n,world", [(1000, 1), (1001, 2), (4097, 8), (7, 4), (3, 8)])
def test_sampler_indices_match_distributed_sampler(n, world):
    from torch.utils.data import DistributedSampler
    from presight_b200.data.batch_loader import sampler_indices
    ds = list(range(n))
    for rank in range(world):
        ref = list(DistributedSampler(ds, world, rank))
        assert sampler_indices(n, rank, world).tolist() == ref


@pytest.mark.gpu
@pytest.mark.parametrize("with_features", [True, False])
def test_device_batches_match_reference_loader(with_features):
    from presight_b200.data import DeviceBatchLoader, ImageChunk
    n, B, world = 5003, 512, 2
    f = make_chunk(n)
    if not with_features:
        f["features"] = None
    ref_chunk = ImageChunkRef(**f)
    for rank in range(world):
        loader = DeviceBatchLoader(ImageChunk(**f), B, rank, world, "cuda")
        ref = list(reference_batches(ref_chunk, B, rank, world))
        got = list(loader)
        assert len(got) == len(ref) == len(loader) and len(ref) >= 4
        for (_, gb), rb in zip(got, ref):
            for k, v in rb.items():
                assert torch.equal(gb[k].cpu(), v), k          # pure data movement + integer arithmetic: bit-exact
        loader.check()


@pytest.mark.gpu
def test_device_loader_reports_bad_index_and_feeds_ray_generator():
    from presight_b200.cameras.ray_generator import RayGenerator
    from presight_b200.data import DeviceBatchLoader, ImageChunk
    f = make_chunk(2048)
    f["image_indices"] = torch.randint(0, 6, (2048,))
    c2w = torch.eye(4)[:3].repeat(6, 1, 1)
    gen = RayGenerator(c2w, torch.full((6,), 1200.0), torch.full((6,), 1200.0), torch.full((6,), 800.0), torch.full((6,), 450.0)).cuda()
    loader = DeviceBatchLoader(ImageChunk(**f), 256, 0, 1, "cuda", ray_generator=gen, pose_scale_factor=0.01)
    rb, batch = next(iter(loader))
    assert rb.origins.shape == (256, 3) and rb.metadata["video_id"].shape == (256, 1)
    assert float(rb.metadata["pose_scale_factor"][0]) == pytest.approx(0.01)
    assert torch.equal(rb.camera_indices.view(-1), batch["ray_index"][:, 0])
    loader.check()
    loader.indices[3] = 10 ** 9            # an index outside the chunk: flagged, and the row still yields a valid camera index
    loader._b = 0
    rb, batch = next(loader)
    assert int(batch["ray_index"][3].abs().sum()) == 0
    with pytest.raises(IndexError):
        loader.check()


def test_device_loader_refuses_cpu():
    from presight_b200.data import DeviceBatchLoader, ImageChunk
    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceBatchLoader(ImageChunk(**make_chunk(16)), 4, 0, 1, "cpu")

"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's batch assembly (never imported by the product).

Restates, with torch's own DataLoader / DistributedSampler exactly as the reference uses them:
  * ImageChunk.__getitem__                       data/PreSight/my_dataset.py:52-73
  * MyDataManager._get_train_batch_loader        data/PreSight/my_datamanager.py:203-219 (DistributedSampler(chunk, world, rank),
                                                 DataLoader(batch_size, sampler, drop_last=True); num_workers irrelevant to order)
Pinned: the reference's code on this path IS torch.utils.data; the restatement calls the same classes with the same arguments.
"""
import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler


class ImageChunkRef(Dataset):
    def __init__(self, rgbs, segs, skies, depths, features, pixel_indices, image_indices, video_ids, widths):
        self.rgbs, self.segs, self.skies, self.depths, self.features = rgbs, segs, skies, depths, features
        self.pixel_indices, self.image_indices, self.video_ids, self.widths = pixel_indices, image_indices, video_ids, widths

    def __len__(self):
        return len(self.rgbs)

    def __getitem__(self, idx):              # my_dataset.py:52-73
        item = {"rgb": self.rgbs[idx], "seg": self.segs[idx], "depth": self.depths[idx], "sky": self.skies[idx],
                "image_index": self.image_indices[idx], "video_id": self.video_ids[idx]}
        pixel_index, width = self.pixel_indices[idx], self.widths[idx]
        item["ray_index"] = torch.LongTensor([item["image_index"], pixel_index // width, pixel_index % width])
        if self.features is not None:
            item["features"] = self.features[idx]
        return item


def reference_batches(chunk: ImageChunkRef, batch_size: int, rank: int, world: int):
    """my_datamanager.py:203-219 (the branch taken for world_size > 0)."""
    sampler = DistributedSampler(chunk, world, rank)
    return DataLoader(chunk, batch_size=batch_size, sampler=sampler, num_workers=0, pin_memory=False, drop_last=True)

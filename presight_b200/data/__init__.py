from .batch_loader import DeviceBatchLoader, ImageChunk  # noqa: F401

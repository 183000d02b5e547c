"""Data-parallel building blocks: bucketed zero-copy gradient sync (optionally
fused with the optimizer), DistributedDataParallel, CrossBarrier."""
from .bucket import BucketedGradSync  # noqa: F401

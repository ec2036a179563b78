from oracle.upstream import ShapeSpec, Conv2d, get_norm, cat, nonzero_tuple, NaiveSyncBatchNorm, FrozenBatchNorm2d  # noqa
from . import batch_norm  # noqa

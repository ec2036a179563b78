from oracle.upstream import FrozenBatchNorm2d, NaiveSyncBatchNorm  # noqa

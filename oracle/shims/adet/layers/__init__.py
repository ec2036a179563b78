from oracle.upstream import DFConv2d, NaiveGroupNorm, ml_nms  # noqa

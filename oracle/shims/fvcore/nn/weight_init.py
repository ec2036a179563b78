from oracle.upstream import c2_msra_fill, c2_xavier_fill  # noqa

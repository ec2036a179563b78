from oracle.upstream import compute_locations, reduce_sum  # noqa

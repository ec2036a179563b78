from oracle.upstream import get_world_size  # noqa

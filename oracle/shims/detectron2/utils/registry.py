from oracle.upstream import Registry  # noqa

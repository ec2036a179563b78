from oracle.upstream import detector_postprocess  # noqa

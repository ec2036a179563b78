from oracle.upstream import Boxes, Instances, ImageList  # noqa

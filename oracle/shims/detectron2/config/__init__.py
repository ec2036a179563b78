import functools
from sylph_few_shot_detection_b200.config import CfgNode  # container only


def configurable(init_func=None, *, from_config=None):
    """detectron2 `configurable`: `Cls(cfg, ...)` -> `Cls(**Cls.from_config(cfg, ...))`."""
    assert init_func is not None and from_config is None, "only the __init__ decorator form is used by the reference"

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        cfg_like = args[0] if args else kwargs.get("cfg")
        if isinstance(cfg_like, CfgNode) and hasattr(type(self), "from_config"):
            explicit = type(self).from_config(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)
    return wrapped

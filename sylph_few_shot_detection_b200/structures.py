"""`Boxes` / `Instances` / `ShapeSpec` / `Registry` used at the plugin boundary.

When detectron2 is installed its own classes are used, so objects produced here are the exact types the reference's
runners and evaluators expect (`detectron2.structures`, `detectron2.utils.registry`).  detectron2 is not installed
in the build image, so small stand-ins with the same attribute surface are provided for the fields the hot path
touches: `gt_boxes.tensor`, `pred_boxes`, `scores`, `pred_classes`, `locations`, `fpn_levels`, `image_size`.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch

try:  # pragma: no cover - exercised only where detectron2 exists
    import detectron2 as _d2  # type: ignore
    if getattr(_d2, "__sylph_shim__", False):  # the oracle's stand-in package is test infrastructure, not detectron2
        raise ImportError("oracle shim on sys.path")
    from detectron2.layers import ShapeSpec  # type: ignore
    from detectron2.structures import Boxes, Instances  # type: ignore
    from detectron2.utils.registry import Registry  # type: ignore
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class ShapeSpec:  # type: ignore[no-redef]
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

        def __repr__(self):
            return f"ShapeSpec(channels={self.channels}, stride={self.stride})"

    class Registry:  # type: ignore[no-redef]
        def __init__(self, name: str):
            self._name = name
            self._map: Dict[str, Any] = {}

        def register(self, obj: Any = None):
            def add(o):
                if o.__name__ in self._map:
                    raise KeyError(f"'{o.__name__}' already registered in '{self._name}'")
                self._map[o.__name__] = o
                return o
            return add if obj is None else add(obj)

        def get(self, name: str):
            if name not in self._map:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return self._map[name]

        def __contains__(self, name: str) -> bool:
            return name in self._map

    class Boxes:  # type: ignore[no-redef]
        def __init__(self, tensor: torch.Tensor):
            t = torch.as_tensor(tensor, dtype=torch.float32)
            self.tensor = t.reshape(-1, 4)

        def __len__(self) -> int:
            return self.tensor.shape[0]

        def to(self, *a, **k) -> "Boxes":
            return Boxes(self.tensor.to(*a, **k))

        def __getitem__(self, item) -> "Boxes":
            return Boxes(self.tensor[item].reshape(-1, 4))

    class Instances:  # type: ignore[no-redef]
        def __init__(self, image_size: Tuple[int, int], **fields: Any):
            object.__setattr__(self, "_image_size", tuple(image_size))
            object.__setattr__(self, "_fields", {})
            for k, v in fields.items():
                self.set(k, v)

        @property
        def image_size(self) -> Tuple[int, int]:
            return self._image_size

        def set(self, name: str, value: Any) -> None:
            if self._fields:
                assert len(value) == len(self), f"field '{name}' has length {len(value)}, expected {len(self)}"
            self._fields[name] = value

        def __setattr__(self, name: str, value: Any) -> None:
            if name.startswith("_"):
                object.__setattr__(self, name, value)
            else:
                self.set(name, value)

        def __getattr__(self, name: str) -> Any:
            fields = object.__getattribute__(self, "_fields")
            if name not in fields:
                raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
            return fields[name]

        def has(self, name: str) -> bool:
            return name in self._fields

        def get_fields(self) -> Dict[str, Any]:
            return self._fields

        def __len__(self) -> int:
            for v in self._fields.values():
                return len(v)
            return 0

        def to(self, *a, **k) -> "Instances":
            out = Instances(self._image_size)
            for name, v in self._fields.items():
                out.set(name, v.to(*a, **k) if hasattr(v, "to") else v)
            return out

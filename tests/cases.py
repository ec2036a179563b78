"""Configs of the golden cases, restated as overrides on top of the defaults (the GPU box has no /root/reference,
so the reference YAML files cannot be read there).  tests/test_config.py checks, in the build container, that these
overrides reproduce the hot-path keys of the reference YAMLs."""
import os

import torch


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

from sylph_few_shot_detection_b200.presets import preset_cfg


def cfg_for(config_name: str, opts=None):
    return preset_cfg(config_name, ["MODEL.DEVICE", "cpu"] + list(opts or []))


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| relative to max |b| (the tensor scale)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0 if a.numel() == 0 else float("inf")
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / ||b||_2."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0 if a.numel() == 0 else float("inf")
    return float((a - b).norm() / (b.norm() + 1e-30))


def fresh_rendezvous() -> str:
    """Path of a not-yet-existing file for a torch.distributed FileStore rendezvous.  The multi-process CPU tests use it
    instead of a TCP port picked with bind(0): such a port can be taken (or still be in TIME_WAIT) by the time the
    workers listen on it, which showed up as a rare failure of the gloo tests."""
    import os
    import tempfile
    fd, path = tempfile.mkstemp(prefix="sylph_rdzv_")
    os.close(fd)
    os.remove(path)
    return path


def init_gloo(rank: int, world: int, rendezvous: str) -> None:
    import os

    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", init_method="file://" + rendezvous, rank=rank, world_size=world)

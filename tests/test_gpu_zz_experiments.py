"""Parity of opt-in engine experiments that have NOT yet been run on hardware (built after this round's GPU budget was
spent).  They are off by default in the product, so these tests are opt-in as well: SYLPH_RUN_UNVERIFIED=1 runs them
(tools/gpu_round2.sh does); once an experiment has been measured it either becomes the default and its test moves into
tests/test_gpu_cases.py, or it is removed."""
import os

import pytest
import torch

from tests.test_gpu_cases import _images, _setup

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SYLPH_RUN_UNVERIFIED") != "1", reason="unverified experiment: set SYLPH_RUN_UNVERIFIED=1")]


@pytest.mark.parametrize("chunks", ["1", "2,3,1,0", "1,1,2,4"])
def test_chunked_trunk_pass_is_bit_identical(chunks, monkeypatch):
    """SYLPH_TRUNK_CHUNK: the blocks of a ResNet stage over a few images at a time (L2-resident activations) instead of
    the whole batch per layer.  Same kernels over the same tiles: the pyramids must be bit-identical."""
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    ims = [im.cuda() for im in _images(5, 160, 224, 41)]
    monkeypatch.delenv("SYLPH_TRUNK_CHUNK", raising=False)
    _, _, model, _ = _setup(seed=8)
    model.engine.extract_features(SLOT_SUPPORT, ims)
    want = [model.engine.export_features(SLOT_SUPPORT, l).clone() for l in range(5)]
    monkeypatch.setenv("SYLPH_TRUNK_CHUNK", chunks)          # read by sylph_create
    _, _, chunked, _ = _setup(seed=8)
    chunked.engine.extract_features(SLOT_SUPPORT, ims)
    for l in range(5):
        assert torch.equal(chunked.engine.export_features(SLOT_SUPPORT, l), want[l]), f"p{l + 3}"
    assert chunked.engine.launch_count() > model.engine.launch_count()

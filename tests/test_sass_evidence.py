"""Static proof, from the SASS of the built library, that the dense layers run on the Blackwell paths the design claims
(mnemonics per /opt/skills/guides/B200_PROFILING.md): every convolution kernel issues tcgen05.mma (UTCHMMA), reads its
accumulators with tcgen05.ld (LDTM) and stages operands with TMA (UTMALDG); nothing in the library uses the legacy
mma.sync path (HMMA); every kernel takes part in programmatic dependent launch.  No GPU needed (cuobjdump)."""
import re
import shutil
import subprocess

import pytest

from sylph_few_shot_detection_b200 import _lib


@pytest.fixture(scope="module")
def sass_by_kernel():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    _lib.build()
    text = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            kernels[cur].append(m.group(1))
    assert len(kernels) >= 40
    return kernels


def _count(ops, prefix):
    return sum(1 for o in ops if o.startswith(prefix))


def test_convolution_kernels_are_tcgen05_tma_kernels(sass_by_kernel):
    conv = {k: v for k, v in sass_by_kernel.items() if re.search(r"conv_gemm_f16_kernel|conv3x3_pair_kernel|conv1x1_pair_staged_kernel", k)}
    assert len(conv) >= 15                                   # the instantiations the engine dispatches to
    for name, ops in conv.items():
        assert _count(ops, "UTCHMMA") >= 4, name             # tcgen05.mma kind::f16
        assert _count(ops, "LDTM") >= 1, name                # tcgen05.ld (TMEM -> registers in the epilogue)
        assert _count(ops, "UTMALDG") >= 2, name             # TMA tensor loads of the A and B operands
        assert _count(ops, "UTCBAR") >= 1, name              # tcgen05.commit -> mbarrier
        assert _count(ops, "SYNCS") >= 10, name              # mbarrier pipeline
    staged = [k for k in conv if _count(conv[k], "UTMASTG") >= 1]
    assert len(staged) >= 6                                  # TMA-out epilogues of the HBM-bound 1x1 convolutions


def test_no_legacy_tensor_core_path_anywhere(sass_by_kernel):
    for name, ops in sass_by_kernel.items():
        assert _count(ops, "HMMA") == 0 and _count(ops, "HGMMA") == 0 and _count(ops, "IMMA") == 0, name


def test_every_kernel_uses_programmatic_dependent_launch(sass_by_kernel):
    for name, ops in sass_by_kernel.items():
        assert _count(ops, "ACQBULK") >= 1, name             # griddepcontrol.wait
        assert _count(ops, "PREEXIT") >= 1, name             # griddepcontrol.launch_dependents


def test_exchange_kernels_use_system_scope_release_acquire(sass_by_kernel):
    prod = next(v for k, v in sass_by_kernel.items() if "normalize_scatter_codes_kernel" in k)
    cons = next(v for k, v in sass_by_kernel.items() if "collect_codes_kernel" in k)
    assert any(o.startswith("MEMBAR") and ".SYS" in o for o in prod)                       # fence before the signal
    assert any(o.startswith("ATOMG") and ".SYS" in o for o in prod) or any(o.startswith("RED") and ".SYS" in o for o in prod)
    assert any(o.startswith("LD") and "STRONG.SYS" in o for o in cons)                    # ld.acquire.sys on the counter
    assert _count(cons, "NANOSLEEP") >= 1

"""The C-ABI library loads and exports every symbol include/gprot_b200.h declares (CPU only: no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gprot_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gprot_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(built):
    from gprotation_b200 import _lib

    names = _declared()
    assert len(names) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes stub binds exactly the declared surface
    assert _lib.load().gprot_abi_version() == 2


def test_no_torch_or_cxx_types_in_header():
    src = open(os.path.join(ROOT, "include", "gprot_b200.h")).read()
    assert "torch" not in src.lower().replace("no c++ or torch types", "")
    assert 'extern "C"' in src and "std::" not in src


def test_library_is_built_for_sm_100a_with_tma_and_dmma(built):
    import shutil
    import subprocess

    from gprotation_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass          # fp64 tensor pipe
    assert "UBLKCP" in sass              # TMA bulk copies


def test_create_without_device_fails_loudly(built):
    import torch

    from gprotation_b200 import GPEngine, GProtError

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(GProtError) as ei:
        GPEngine(0)
    assert "no CPU fallback" in str(ei.value)


def test_lnprior_batch_argument_errors(built):
    from gprotation_b200 import _lib

    lib = _lib.load()
    out = np.empty(1)
    assert lib.gprot_lnprior_batch(1, None, None, None, None, 0.0, 0.0, None, 0, out.ctypes.data) != 0
    assert lib.gprot_lnprior_batch(0, None, None, None, None, 0.0, 0.0, None, 0, None) == 0

"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/b200romp.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from romp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200romp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200romp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(raw, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) == set(syms)


def test_version_and_error_string(lib):
    assert lib.b200romp_version() == 200
    assert isinstance(lib.b200romp_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from romp_b200 import ROMP, romp_settings
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ROMP(romp_settings([]), state_dict={}, smpl_pack={})


def test_library_is_sm100a():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_settings_match_reference_defaults():
    from romp_b200 import romp_settings
    s = romp_settings([])
    assert s.center_thresh == 0.25 and s.calc_smpl is True and s.root_align is False and s.GPU == 0
    assert s.mode == "image" and s.onnx is False and s.temporal_optimize is False

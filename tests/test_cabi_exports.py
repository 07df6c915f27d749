"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol that
include/fse_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fse_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fse_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for s in ("fse_denoiser_create", "fse_denoise_step", "fse_posterior_step", "fse_sample", "fse_sample_host",
              "fse_vocoder_forward", "fse_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_built):
    handle = ctypes.CDLL(lib_built)
    for s in declared_symbols():
        assert hasattr(handle, s), f"libfse_b200.so does not export {s}"


def test_ctypes_binding_covers_header(lib_built):
    from speech_editing_toolkit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().fse_version() >= 100


def test_sass_is_blackwell_native(lib_built):
    """tcgen05.mma / TMA / TMEM loads must be present in the shipped SASS (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speech_editing_toolkit_b200 import FseError
    from speech_editing_toolkit_b200.engine import Denoiser
    with pytest.raises(FseError):
        Denoiser()

"""Make the UNMODIFIED reference (/root/reference) importable in the build container.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (fixture generation), by
the `-m "not gpu"` tests that pin the numpy oracle against the real reference, and by bench.py's
reference arm.  /root/reference does not exist on the GPU box: there the root is oracle/_ref, the
verbatim git-ignored copy of the hot path's import closure made by oracle/build_ref.py (it travels
with the snapshot like a built .so).  Never imported by the product package.

The reference's math needs none of librosa / pyloudnorm / textgrid / webrtcvad /
skimage, but `modules/speech_editing/spec_denoiser/fs.py:14-15` pulls them in
transitively via `utils/audio/__init__.py:1-5`; they are absent here, so they are
stubbed with MagicMock before import.
"""
import os
import sys
from unittest.mock import MagicMock

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # verbatim copy made by oracle/build_ref.py


def _pick_root() -> str:
    env = os.environ.get("FSE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/modules/speech_editing"):
        return "/root/reference"
    return _VENDORED        # the GPU box: only the git-ignored copy exists


REF_ROOT = _pick_root()

_STUBS = ["librosa", "librosa.filters", "librosa.core", "librosa.feature", "pyloudnorm", "textgrid",
          "webrtcvad", "skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot"]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "modules", "speech_editing"))


def install(config: str = "egs/spec_denoiser.yaml", overrides: str = ""):
    """Put the reference on sys.path, stub absent deps and populate its global hparams."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)   # base_config chains in the yaml are relative to the repo root
    try:
        from utils.commons import hparams as hp_mod
        saved_argv = sys.argv
        sys.argv = [saved_argv[0]]
        try:
            hp = hp_mod.set_hparams(config=config, exp_name="", hparams_str=overrides, print_hparams=False)
        finally:
            sys.argv = saved_argv
    finally:
        os.chdir(cwd)
    return hp

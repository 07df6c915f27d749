"""speech_editing_toolkit_b200 — B200-native FluentSpeech spec_denoiser sampling + HiFi-GAN forward.

Only the hot path of Zain-Jiang/Speech-Editing-Toolkit named by BASELINE.json lives here:
CUDA kernels + C ABI under csrc/, and the host-side mirror of the reference's plugin seams.
"""
from ._lib import FseError, MODES  # noqa: F401

__all__ = ["FseError", "MODES"]

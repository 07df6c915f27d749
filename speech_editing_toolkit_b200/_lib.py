"""ctypes binding of the C ABI declared in include/fse_b200.h.

There is deliberately NO fallback: if libfse_b200.so is missing or a call fails, an exception is
raised.  (The oracle under oracle/ is test infrastructure and is never imported from here.)
"""
import ctypes as C
import os

from .build import LIB_PATH

FSE_MODE_TC_BF16 = 0
FSE_MODE_SIMT_F32 = 1
FSE_MODE_SIMT_BF16 = 2
FSE_MODE_TC_TF32 = 3
MODES = {"tc_bf16": FSE_MODE_TC_BF16, "simt_f32": FSE_MODE_SIMT_F32, "simt_bf16": FSE_MODE_SIMT_BF16, "tc_tf32": FSE_MODE_TC_TF32}


class FseError(RuntimeError):
    pass


class DenoiserConfig(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("hidden", C.c_int32), ("channels", C.c_int32), ("layers", C.c_int32),
                ("dilation_cycle_length", C.c_int32), ("mode", C.c_int32)]


class VocoderConfig(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("num_upsamples", C.c_int32),
                ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
                ("num_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 4),
                ("resblock_dilations", (C.c_int32 * 3) * 4), ("mode", C.c_int32), ("resblock", C.c_int32)]


class MelEncoderConfig(C.Structure):
    _fields_ = [("n_mels", C.c_int32), ("hidden", C.c_int32), ("mode", C.c_int32)]


class CondEncoderConfig(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("vocab", C.c_int32), ("enc_layers", C.c_int32), ("enc_dilations", C.c_int32 * 8),
                ("enc_kernel_size", C.c_int32), ("layers_in_block", C.c_int32), ("enc_post_net_kernel", C.c_int32),
                ("dur_predictor_layers", C.c_int32), ("dur_predictor_kernel", C.c_int32), ("pitch_predictor_layers", C.c_int32),
                ("predictor_kernel", C.c_int32), ("use_pitch_embed", C.c_int32), ("use_uv", C.c_int32),
                ("spk_embed_dim", C.c_int32), ("mode", C.c_int32)]


class CampNetConfig(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("vocab", C.c_int32), ("n_mels", C.c_int32), ("enc_layers", C.c_int32), ("dec_layers", C.c_int32),
                ("heads", C.c_int32), ("ffn_kernel", C.c_int32), ("fine_blocks", C.c_int32), ("fine_kernel", C.c_int32), ("mode", C.c_int32)]


class WgradProblem(C.Structure):
    """fse_wgrad_problem (include/fse_b200.h)"""
    _fields_ = [("P", C.c_void_p), ("ldp", C.c_int64), ("Q", C.c_void_p), ("ldq", C.c_int64), ("M", C.c_int32), ("N", C.c_int32),
                ("offs", C.POINTER(C.c_int32)), ("ntaps", C.c_int32), ("out", C.c_void_p), ("ld_m", C.c_int64), ("ld_n", C.c_int64), ("ld_j", C.c_int64)]


class MelFrontendConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("fft_size", C.c_int32), ("hop_size", C.c_int32), ("win_length", C.c_int32), ("num_mels", C.c_int32),
                ("fmin", C.c_float), ("fmax", C.c_float), ("eps", C.c_float)]


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_int64)]


_P = C.c_void_p
_F = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol of include/fse_b200.h
SIGNATURES = {
    "fse_last_error": (C.c_char_p, []),
    "fse_version": (C.c_int, []),
    "fse_denoiser_create": (C.c_int, [C.POINTER(DenoiserConfig), C.POINTER(_P)]),
    "fse_denoiser_destroy": (None, [_P]),
    "fse_denoiser_load_weights": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32]),
    "fse_denoiser_set_schedule": (C.c_int, [_P, C.c_int32, _F, _F, _F]),
    "fse_denoiser_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32]),
    "fse_denoise_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_posterior_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint64, C.c_uint32, _P, C.c_int32, C.c_int32, _P]),
    "fse_sample": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_sample_host": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, _P, C.c_int32, C.c_int32]),
    "fse_denoiser_last_launches": (C.c_int64, [_P]),
    "fse_vocoder_create": (C.c_int, [C.POINTER(VocoderConfig), C.POINTER(_P)]),
    "fse_vocoder_destroy": (None, [_P]),
    "fse_vocoder_load_weights": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32]),
    "fse_vocoder_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32]),
    "fse_vocoder_forward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_vocoder_forward_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32]),
    "fse_vocoder_last_launches": (C.c_int64, [_P]),
    "fse_mel_encoder_create": (C.c_int, [C.POINTER(MelEncoderConfig), C.POINTER(_P)]),
    "fse_mel_encoder_destroy": (None, [_P]),
    "fse_mel_encoder_load_weights": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32]),
    "fse_mel_encoder_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32]),
    "fse_mel_encoder_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_mel_encoder_last_launches": (C.c_int64, [_P]),
    "fse_cond_encoder_create": (C.c_int, [C.POINTER(CondEncoderConfig), C.POINTER(_P)]),
    "fse_cond_encoder_destroy": (None, [_P]),
    "fse_cond_encoder_load_weights": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32]),
    "fse_cond_encoder_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "fse_cond_encoder_last_launches": (C.c_int64, [_P]),
    "fse_cond_text_encoder": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_cond_style_embed": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "fse_cond_dur_input": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "fse_cond_masked_dur": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fse_cond_duration": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_cond_length_cumsum": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "fse_cond_length_fill": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fse_cond_frames": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                  _P, C.c_int64, _P]),
    "fse_campnet_create": (C.c_int, [C.POINTER(CampNetConfig), C.POINTER(_P)]),
    "fse_campnet_destroy": (None, [_P]),
    "fse_campnet_load_weights": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32]),
    "fse_campnet_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "fse_campnet_last_launches": (C.c_int64, [_P]),
    "fse_campnet_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_edit_prepare": (C.c_int, [_P] * 11 + [C.c_int32] * 4 + [_P]),
    "fse_edit_plan": (C.c_int, [_P] * 11 + [C.c_int32] * 4 + [_P]),
    "fse_edit_assemble": (C.c_int, [_P] * 15 + [C.c_int32] * 5 + [_P]),
    "fse_mel_frontend_create": (C.c_int, [C.POINTER(MelFrontendConfig), C.POINTER(_P)]),
    "fse_mel_frontend_destroy": (None, [_P]),
    "fse_mel_frontend_frames": (C.c_int64, [_P, C.c_int64]),
    "fse_mel_frontend_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int64]),
    "fse_mel_frontend_forward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, C.c_int64, _P]),
    "fse_mel_frontend_last_launches": (C.c_int64, [_P]),
    "fse_train_create": (C.c_int, [C.POINTER(DenoiserConfig), C.POINTER(_P)]),
    "fse_train_destroy": (None, [_P]),
    "fse_train_load_weights_device": (C.c_int, [_P, C.POINTER(Tensor), C.c_int32, _P]),
    "fse_train_workspace_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32]),
    "fse_train_layout": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int32]),
    "fse_train_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_train_backward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_train_last_launches": (C.c_int64, [_P]),
    "fse_wgrad_workspace_bytes": (C.c_int64, [C.c_int32] * 6),
    "fse_wgrad": (C.c_int, [C.c_int32, _P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                            _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P]),
    "fse_wgrad_group_workspace_bytes": (C.c_int64, [C.c_int32, C.POINTER(WgradProblem), C.c_int32, C.c_int32, C.c_int32]),
    "fse_wgrad_group": (C.c_int, [C.c_int32, C.POINTER(WgradProblem), C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_mel_loss_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "fse_mel_loss_forward": (C.c_int, [_P, _P, C.c_float, C.c_float, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_mel_loss_backward": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "fse_denoiser_profile": (C.c_int, [_P, C.c_int32]),
    "fse_denoiser_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fse_vocoder_profile": (C.c_int, [_P, C.c_int32]),
    "fse_vocoder_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fse_debug_conv_gemm": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32]),
}

_lib = None


def lib():
    """Load libfse_b200.so (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().fse_last_error()
        raise FseError(f"fse error {rc}: {msg.decode() if msg else '?'}")

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


def test_cond_encoder_config_validation_happens_before_any_device_call(lib_built):
    """Argument checks of fse_cond_encoder_create (include/fse_b200.h) return FSE_EINVAL with a message and never reach the
    device, so they can be exercised without a GPU; a valid config then fails loudly (FSE_ECUDA) when no sm_100 device exists."""
    import ctypes as C
    import torch
    from speech_editing_toolkit_b200 import _lib
    L = _lib.lib()

    def cfg(**kw):
        c = _lib.CondEncoderConfig()
        c.hidden, c.vocab, c.enc_layers, c.enc_kernel_size, c.layers_in_block, c.enc_post_net_kernel = 192, 80, 4, 5, 2, 3
        for i in range(4):
            c.enc_dilations[i] = 1
        c.dur_predictor_layers, c.dur_predictor_kernel, c.pitch_predictor_layers, c.predictor_kernel = 3, 5, 5, 5
        c.use_pitch_embed, c.use_uv, c.spk_embed_dim, c.mode = 1, 1, 256, 0
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    for bad, word in ((dict(hidden=200), b"hidden"), (dict(hidden=1024), b"hidden"), (dict(vocab=0), b"vocab"), (dict(enc_layers=9), b"enc_layers"),
                      (dict(mode=7), b"mode"), (dict(dur_predictor_layers=0), b"predictor"), (dict(spk_embed_dim=-1), b"spk_embed_dim")):
        h = C.c_void_p()
        assert L.fse_cond_encoder_create(C.byref(cfg(**bad)), C.byref(h)) == -1, bad
        assert word in L.fse_last_error(), (bad, L.fse_last_error())
    c = cfg()
    c.enc_dilations[2] = 0
    assert L.fse_cond_encoder_create(C.byref(c), C.byref(C.c_void_p())) == -1
    assert L.fse_cond_encoder_create(None, None) == -1
    assert L.fse_cond_encoder_workspace_bytes(None, 1, 1, 1) == 0 and L.fse_cond_encoder_last_launches(None) == 0
    assert L.fse_cond_text_encoder(None, None, None, 1, 1, None, 0, None) == -1
    if not torch.cuda.is_available():
        assert L.fse_cond_encoder_create(C.byref(cfg()), C.byref(C.c_void_p())) == -2         # no device: FSE_ECUDA, no fallback


def test_campnet_config_validation_happens_before_any_device_call(lib_built):
    import ctypes as C
    import torch
    from speech_editing_toolkit_b200 import _lib
    L = _lib.lib()

    def cfg(**kw):
        c = _lib.CampNetConfig()
        c.hidden, c.vocab, c.n_mels, c.enc_layers, c.dec_layers, c.heads, c.ffn_kernel, c.fine_blocks, c.fine_kernel, c.mode = 192, 80, 80, 3, 6, 2, 9, 5, 5, 0
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    for bad, word in ((dict(hidden=256), b"heads"), (dict(heads=0), b"heads"), (dict(vocab=0), b"vocab"), (dict(n_mels=81), b"n_mels"),
                      (dict(ffn_kernel=8), b"ffn_kernel"), (dict(fine_kernel=13), b"fine_kernel"), (dict(dec_layers=0), b"layer"), (dict(mode=9), b"mode")):
        assert L.fse_campnet_create(C.byref(cfg(**bad)), C.byref(C.c_void_p())) == -1, bad
        assert word in L.fse_last_error(), (bad, L.fse_last_error())
    assert L.fse_campnet_create(None, None) == -1
    assert L.fse_campnet_workspace_bytes(None, 1, 1, 1) == 0 and L.fse_campnet_last_launches(None) == 0
    assert L.fse_campnet_forward(None, None, None, None, None, None, None, None, 1, 1, 1, None, 0, None) == -1
    if not torch.cuda.is_available():
        assert L.fse_campnet_create(C.byref(cfg()), C.byref(C.c_void_p())) == -2              # no device: FSE_ECUDA, no fallback


def test_header_is_plain_c_and_links_from_a_c_program(lib_built, tmp_path):
    """The boundary is a C ABI: include/fse_b200.h compiles as C99 (no C++-isms, no torch types) and a plain C program links
    against libfse_b200.so and gets an FSE_EINVAL + message back from an argument check (no device needed)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "fse_b200.h")], check=True)
    src = tmp_path / "demo.c"
    src.write_text('#include "fse_b200.h"\n#include <string.h>\n'
                   "int main(void) {\n"
                   "  fse_campnet_config c; fse_campnet* h = 0; memset(&c, 0, sizeof c);\n"
                   "  int rc = fse_campnet_create(&c, &h);\n"
                   "  return (rc == FSE_EINVAL && strlen(fse_last_error()) > 0 && fse_version() >= 100) ? 0 : 1;\n"
                   "}\n")
    exe = tmp_path / "demo"
    libdir = os.path.dirname(lib_built)
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-L", libdir, "-lfse_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_edit_region_argument_checks_without_device(lib_built):
    """fse_edit_* validate pointers and sizes before touching the device; with valid arguments and no GPU they fail loudly."""
    import ctypes as C
    import numpy as np
    import torch
    from speech_editing_toolkit_b200 import _lib
    L = _lib.lib()
    a = np.zeros(8, dtype=np.int64)
    f = np.zeros(8, dtype=np.float32)
    P = lambda x: C.c_void_p(x.ctypes.data)
    assert L.fse_edit_prepare(None, None, None, None, None, None, None, None, None, None, None, 1, 1, 1, 1, None) == -1
    assert L.fse_edit_prepare(P(a), P(a), None, P(a), P(a), None, None, P(a), P(a), P(a), P(f), 1, 0, 1, 1, None) == -1
    assert b"positive" in L.fse_last_error()
    assert L.fse_edit_plan(P(a), P(a), None, P(a), None, P(a), None, None, P(a), P(a), P(a), 1, 4, 4, 4, None) == -1      # edited_mel2ph missing
    assert L.fse_edit_assemble(P(a), None, P(a), P(a), P(a), P(a), P(a), None, None, None, P(a), P(f), P(f), P(f), P(f), 1, 4, 4, 4, 8, None) == -1
    if not torch.cuda.is_available():    # host pointers are never dereferenced: the device check comes first and fails with FSE_ECUDA
        assert L.fse_edit_prepare(P(a), P(a), None, P(a), P(a), None, None, P(a), P(a), P(a), P(f), 1, 4, 4, 4, None) == -2


def test_mel_frontend_config_validation_without_device(lib_built):
    import ctypes as C
    import torch
    from speech_editing_toolkit_b200 import _lib
    L = _lib.lib()

    def cfg(**kw):
        c = _lib.MelFrontendConfig()
        c.sample_rate, c.fft_size, c.hop_size, c.win_length, c.num_mels, c.fmin, c.fmax, c.eps = 22050, 1024, 256, 1024, 80, 55.0, 7600.0, 1e-6
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    for bad, word in ((dict(hop_size=100), b"hop_size"), (dict(fft_size=1000), b"fft_size"), (dict(win_length=800), b"win_length"),
                      (dict(num_mels=81), b"num_mels"), (dict(fmin=8000.0), b"fmin"), (dict(fmax=20000.0), b"fmax"), (dict(eps=0.0), b"eps")):
        assert L.fse_mel_frontend_create(C.byref(cfg(**bad)), C.byref(C.c_void_p())) == -1, bad
        assert word in L.fse_last_error(), (bad, L.fse_last_error())
    assert L.fse_mel_frontend_frames(None, 100) == 0 and L.fse_mel_frontend_workspace_bytes(None, 1, 256) == 0
    if not torch.cuda.is_available():
        assert L.fse_mel_frontend_create(C.byref(cfg()), C.byref(C.c_void_p())) == -2


def test_training_entry_points_validate_arguments_without_device(lib_built):
    """fse_wgrad / fse_wgrad_group / fse_mel_loss_* report argument errors before any device call; with valid arguments and no GPU they
    fail loudly (no CPU fallback)."""
    import ctypes as C
    import numpy as np
    import torch
    from speech_editing_toolkit_b200 import _lib
    L = _lib.lib()
    buf = np.zeros(4096, dtype=np.float32)
    base = (buf.ctypes.data + 15) // 16 * 16
    P = lambda off=0: C.c_void_p(base + off)
    TC_BF16 = _lib.MODES["tc_bf16"]
    offs = (C.c_int32 * 1)(0)
    args = lambda **kw: dict(dict(mode=TC_BF16, P=P(), ldp=64, Q=P(), ldq=64, B=1, T=8, M=64, N=64, offs=offs, ntaps=1, out=P(), ws=P()), **kw)

    def call(a):
        return L.fse_wgrad(a["mode"], a["P"], a["ldp"], a["Q"], a["ldq"], a["B"], a["T"], a["M"], a["N"], a["offs"], a["ntaps"], a["out"], 64, 1, 0, a["ws"], 1 << 20, None)
    for bad, word in ((dict(P=None), b"null"), (dict(mode=_lib.MODES["simt_f32"]), b"tensor-core"), (dict(T=0), b"positive"), (dict(ntaps=17), b"ntaps"),
                      (dict(ldp=32), b"pitch"), (dict(ldq=65), b"aligned"), (dict(P=P(2)), b"aligned")):
        assert call(args(**bad)) == -1, bad
        assert word in L.fse_last_error(), (bad, L.fse_last_error())
    assert L.fse_wgrad_group(TC_BF16, None, 1, 1, 8, P(), 16, None) == -1
    assert L.fse_wgrad_group(TC_BF16, (_lib.WgradProblem * 5)(), 5, 1, 8, P(), 16, None) == -1 and b"per launch" in L.fse_last_error()
    assert L.fse_mel_loss_forward(None, P(), 0.5, 0.5, P(), 1, 1, 8, 80, P(), 1 << 20, None) == -1
    assert L.fse_mel_loss_forward(P(), P(), 0.5, 0.5, P(), 1, 1, 8, 200, P(), 1 << 20, None) == -1 and b"n_mels" in L.fse_last_error()
    assert L.fse_mel_loss_forward(P(), P(), 0.5, 0.5, P(), 1, 1, 8, 80, P(), 16, None) == -1 and b"workspace" in L.fse_last_error()
    assert L.fse_mel_loss_backward(P(), P(), None, 0.5, 0.5, P(), 1, 8, 80, P(), 1 << 20, None) == -1
    assert L.fse_mel_loss_workspace_bytes(0, 8, 80) == 0 and L.fse_mel_loss_workspace_bytes(2, 100, 80) > 3 * 2 * 100 * 80 * 4
    if not torch.cuda.is_available():
        assert call(args()) == -2 and L.fse_mel_loss_forward(P(), P(), 0.5, 0.5, P(), 1, 1, 8, 80, P(), 1 << 20, None) == -2

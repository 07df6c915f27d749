"""First-contact check of the condition-encoder (fse_cond_*) and MelEncoder (fse_mel_encoder_*) entry points on a GPU box,
WITHOUT torch: ctypes on libfse_b200.so + libcudart, numpy, and the numpy oracle as the checker.  It starts in a second
(no `import torch` page-in), never aborts on a mismatch and prints one line per output with the error against the
reference fixture and against the oracle, so that one short GPU call yields a full diagnosis.

    python tests/tools/cond_check.py [simt_f32 tc_bf16 ...] | tee gpurun_out/cond_check.txt

TEST TOOL (imports oracle/): not part of the shipped path."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cond_encoder_oracle as CO          # noqa: E402
from speech_editing_toolkit_b200 import _lib, synth   # noqa: E402

T0 = time.time()
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        RT = C.CDLL(name)
        break
    except OSError:
        RT = None
if RT is None:
    raise SystemExit("libcudart not found")
RT.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
RT.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
RT.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
RT.cudaFree.argtypes = [C.c_void_p]
RT.cudaGetErrorString.restype = C.c_char_p
FAILS = []


def rt(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: cuda error {rc} {RT.cudaGetErrorString(rc).decode()}")


def dmalloc(nbytes):
    p = C.c_void_p()
    rt(RT.cudaMalloc(C.byref(p), max(int(nbytes), 16)), "cudaMalloc")
    rt(RT.cudaMemset(p, 0xFF, max(int(nbytes), 16)), "cudaMemset")      # NaN pattern: unwritten outputs show up
    return p


def h2d(a):
    a = np.ascontiguousarray(a)
    p = dmalloc(a.nbytes)
    rt(RT.cudaMemcpy(p, C.c_void_p(a.ctypes.data), a.nbytes, 1), "cudaMemcpy h2d")
    return p


def d2h(p, shape, dtype):
    a = np.empty(shape, dtype=dtype)
    rt(RT.cudaMemcpy(C.c_void_p(a.ctypes.data), p, a.nbytes, 2), "cudaMemcpy d2h")
    return a


def sync(what):
    rt(RT.cudaDeviceSynchronize(), f"sync after {what}")


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


def report(tag, name, got, want, tol_abs=None, tol_rel=None, exact=False):
    got = np.asarray(got); want = np.asarray(want)
    if got.shape != want.shape:
        FAILS.append(f"{tag} {name}"); print(f"[FAIL] {tag:28s} {name:22s} shape {got.shape} vs {want.shape}", flush=True); return
    if exact:
        nbad = int((got != want).sum())
        ok = nbad == 0
        msg = f"mismatches {nbad}/{got.size}" + ("" if ok else f" first at {tuple(int(i) for i in np.argwhere(got != want)[0])}: {got[got != want][0]} vs {want[got != want][0]}")
    else:
        nan = int((~np.isfinite(got)).sum())
        g = np.nan_to_num(got.astype(np.float64), nan=1e30, posinf=1e30, neginf=-1e30)
        mx, rl = float(np.abs(g - want).max()), rel_l1(g, want)
        ok = nan == 0 and (tol_abs is None or mx < tol_abs) and (tol_rel is None or rl < tol_rel)
        where = tuple(int(i) for i in np.unravel_index(np.abs(g - want).argmax(), got.shape))
        msg = f"max-abs {mx:.3e} at {where}  rel-L1 {rl:.3e}  non-finite {nan}"
    if not ok:
        FAILS.append(f"{tag} {name}")
    print(f"[{'ok' if ok else 'FAIL'}] {tag:28s} {name:22s} {msg}", flush=True)


def tensor_table(sd):
    keep = []
    arr = (_lib.Tensor * len(sd))()
    for i, (k, v) in enumerate(sd.items()):
        a = np.ascontiguousarray(v, dtype=np.float32)
        keep.append(a)
        arr[i].name = k.encode(); arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float)); arr[i].numel = a.size
    return arr, len(sd), keep


class Cond:
    def __init__(self, L, sd, vocab, mode):
        self.L = L
        cfg = _lib.CondEncoderConfig()
        cfg.hidden, cfg.vocab, cfg.enc_layers = 192, vocab, 4
        for i in range(4):
            cfg.enc_dilations[i] = 1
        cfg.enc_kernel_size, cfg.layers_in_block, cfg.enc_post_net_kernel = 5, 2, 3
        cfg.dur_predictor_layers, cfg.dur_predictor_kernel, cfg.pitch_predictor_layers, cfg.predictor_kernel = 3, 5, 5, 5
        cfg.use_pitch_embed, cfg.use_uv, cfg.spk_embed_dim, cfg.mode = 1, 1, 256, _lib.MODES[mode]
        self.h = C.c_void_p()
        _lib.check(L.fse_cond_encoder_create(C.byref(cfg), C.byref(self.h)))
        arr, n, keep = tensor_table(sd)
        _lib.check(L.fse_cond_encoder_load_weights(self.h, arr, n))

    def ws(self, B, Tt, T):
        n = self.L.fse_cond_encoder_workspace_bytes(self.h, B, Tt, T)
        p = dmalloc(n + 1024)
        return C.c_void_p((p.value + 1023) // 1024 * 1024), n

    def forward(self, batch, use_pred_pitch, masked_dur=None):
        """FastSpeech.forward(skip_decoder=True) composed from the entry points exactly as modules.FastSpeechB200 does."""
        L, h = self.L, self.h
        txt, mel2ph = batch["txt_tokens"], batch["mel2ph"]
        B, Tt = txt.shape
        T = mel2ph.shape[1]
        out = {}
        d_txt, d_m2p = h2d(txt), h2d(mel2ph)
        d_mask, d_f0, d_uv, d_spk = h2d(batch["time_mel_masks"]), h2d(batch["f0"]), h2d(batch["uv"]), h2d(batch["spk_embed"])
        ws, nws = self.ws(B, Tt, T)
        d_enc = dmalloc(B * Tt * 192 * 4)
        _lib.check(L.fse_cond_text_encoder(h, d_txt, d_enc, B, Tt, ws, nws, None)); sync("text_encoder")
        out["encoder_out"] = d2h(d_enc, (B, Tt, 192), np.float32)
        d_style = dmalloc(B * 192 * 4)
        _lib.check(L.fse_cond_style_embed(h, d_spk, d_style, B, None)); sync("style_embed")
        out["style_embed"] = d2h(d_style, (B, 1, 192), np.float32)
        d_dinp = dmalloc(B * Tt * 192 * 4)
        _lib.check(L.fse_cond_dur_input(h, d_enc, d_style, d_txt, d_dinp, B, Tt, None)); sync("dur_input")
        out["dur_inp"] = d2h(d_dinp, (B, Tt, 192), np.float32)
        d_md = dmalloc(B * Tt * 8)
        _lib.check(L.fse_cond_masked_dur(h, d_m2p, d_mask, d_txt, d_md, B, T, Tt, None)); sync("masked_dur")
        out["masked_dur_gt"] = d2h(d_md, (B, Tt), np.int64)
        if masked_dur is not None:
            d_md = h2d(masked_dur)
        d_dur = dmalloc(B * Tt * 4)
        _lib.check(L.fse_cond_duration(h, d_dinp, d_md, d_txt, d_dur, B, Tt, ws, nws, None)); sync("duration")
        out["dur"] = d2h(d_dur, (B, Tt), np.float32)
        d_cs, d_tot = dmalloc(B * Tt * 8), dmalloc(B * 8)
        _lib.check(L.fse_cond_length_cumsum(h, d_dur, d_txt, d_cs, d_tot, B, Tt, None)); sync("length_cumsum")
        tot = d2h(d_tot, (B,), np.int64)
        tmax = int(tot.max())
        if 0 < tmax < 100000:
            d_lr = dmalloc(B * tmax * 8)
            _lib.check(L.fse_cond_length_fill(h, d_cs, d_lr, B, Tt, tmax, None)); sync("length_fill")
            out["mel2ph_pred"] = d2h(d_lr, (B, tmax), np.int64)
        else:
            out["mel2ph_pred"] = np.zeros((B, 0), np.int64); print("    length totals", tot)
        d_dec, d_pp, d_fd, d_fdp, d_pitch = dmalloc(B * T * 192 * 4), dmalloc(B * T * 8), dmalloc(B * T * 4), dmalloc(B * T * 4), dmalloc(B * T * 8)
        _lib.check(L.fse_cond_frames(h, d_enc, d_style, d_m2p, d_mask, d_f0, d_uv, int(use_pred_pitch), d_dec, d_pp, d_fd, d_fdp, d_pitch,
                                     B, Tt, T, ws, nws, None)); sync("frames")
        out["decoder_inp"] = d2h(d_dec, (B, T, 192), np.float32)
        out["pitch_pred"] = d2h(d_pp, (B, T, 2), np.float32)
        out["f0_denorm"], out["f0_denorm_pred"] = d2h(d_fd, (B, T), np.float32), d2h(d_fdp, (B, T), np.float32)
        out["pitch"] = d2h(d_pitch, (B, T), np.int64)
        out["launches"] = int(L.fse_cond_encoder_last_launches(h))
        return out


def check_cond(L, mode):
    g = np.load(os.path.join(ROOT, "tests", "golden", "cond_encoder.npz"))
    seed, B, T, vocab = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["vocab"])
    sd = synth.fastspeech_state_dict(seed, vocab)
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(seed, B, T, vocab=vocab), item=1, n_tokens=3)
    f32 = mode == "simt_f32"
    ta, tr = (2e-4, None) if f32 else (None, 2e-2)
    enc = Cond(L, sd, vocab, mode)
    for flag in (False, True):
        sfx = "_predpitch" if flag else ""
        tag = f"{mode} fixture{sfx}"
        out = enc.forward(batch, flag)
        if not flag:
            report(tag, "encoder_out", out["encoder_out"], g["encoder_out"], ta, tr)
            report(tag, "style_embed", out["style_embed"], g["style_embed"], 1e-5)
            report(tag, "masked_dur_gt", out["masked_dur_gt"], CO.masked_dur_gt(batch["mel2ph"], batch["time_mel_masks"], batch["txt_tokens"]), exact=True)
        report(tag, "dur", out["dur"], g["dur" + sfx], ta, tr)
        report(tag, "pitch_pred", out["pitch_pred"], g["pitch_pred" + sfx], ta, tr)
        report(tag, "f0_denorm", out["f0_denorm"], g["f0_denorm" + sfx], 0.05 if (f32 or not flag) else None)
        if f32:
            report(tag, "f0_denorm_pred", out["f0_denorm_pred"], g["f0_denorm_pred" + sfx], 0.05)
        else:   # bf16 operands flip the sign of uv logits that sit near zero: those frames read 0 Hz instead of f0 (or back)
            d = np.abs(out["f0_denorm_pred"] - g["f0_denorm_pred" + sfx])
            print(f"[info] {tag:28s} f0_denorm_pred: {int((d > 1.0).sum())} of {d.size} frames differ by > 1 Hz (voiced/unvoiced flips), "
                  f"median |diff| {float(np.median(d)):.3f} Hz")
        if f32 or not flag:
            report(tag, "pitch bins", out["pitch"], g["pitch" + sfx], exact=True)
        else:
            print(f"[info] {tag:28s} pitch bins: {int((out['pitch'] != g['pitch' + sfx]).sum())} of {out['pitch'].size} differ, "
                  f"max |diff| {int(np.abs(out['pitch'] - g['pitch' + sfx]).max())}")
        report(tag, "decoder_inp", out["decoder_inp"], g["decoder_inp" + sfx], ta, tr)
        print(f"       launches of the frame stage: {out['launches']}")
    out = enc.forward(batch, False, masked_dur=g["masked_dur_in"])
    report(f"{mode} masked_dur call", "dur", out["dur"], g["dur_masked_dur"], ta, tr)
    if f32:
        report(f"{mode} masked_dur call", "mel2ph_pred", out["mel2ph_pred"], g["mel2ph_pred"], exact=True)
    # multi-tile ragged batch against the oracle (both arithmetic contracts)
    vocab2, B2, T2 = 60, 3, 330
    sd2 = synth.fastspeech_state_dict(77, vocab2)
    b2 = synth.synthetic_edit_batch(78, B2, T2, vocab=vocab2, frames_per_phone=4)
    b2 = synth.pad_edit_batch(synth.pad_edit_batch(b2, 1, 20, 4), 2, 5, 4)
    args = (b2["txt_tokens"], b2["time_mel_masks"], b2["mel2ph"], b2["spk_embed"], b2["f0"], b2["uv"])
    ref = CO.fastspeech_forward(sd2, *args, use_pred_pitch=False)
    out = Cond(L, sd2, vocab2, mode).forward(b2, False)
    tag = f"{mode} 3x330 vs oracle"
    for k in ("encoder_out", "dur", "pitch_pred", "decoder_inp"):
        report(tag, k, out[k], ref[k], ta, tr)
    report(tag, "pitch bins", out["pitch"], ref["pitch"], exact=True)
    report(tag, "mel2ph_pred", out["mel2ph_pred"], CO.length_regulator(out["dur"], b2["txt_tokens"] == 0), exact=True)
    if not f32:
        refb = CO.fastspeech_forward(sd2, *args, use_pred_pitch=False, gemm_dtype="bf16")
        for k in ("encoder_out", "dur", "pitch_pred", "decoder_inp"):
            report(f"{mode} 3x330 vs bf16 oracle", k, out[k], refb[k], None, 8e-3)


def check_mel_encoder(L, mode):
    g = np.load(os.path.join(ROOT, "tests", "golden", "mel_encoder.npz"))
    B, T = int(g["B"]), int(g["T"])
    cfg = _lib.MelEncoderConfig()
    cfg.n_mels, cfg.hidden, cfg.mode = 80, 192, _lib.MODES[mode]
    h = C.c_void_p()
    _lib.check(L.fse_mel_encoder_create(C.byref(cfg), C.byref(h)))
    arr, n, keep = tensor_table(synth.mel_encoder_state_dict(int(g["seed"])))
    _lib.check(L.fse_mel_encoder_load_weights(h, arr, n))
    ref, mask = synth.synthetic_ref_and_mask(int(g["seed"]), B, T)
    d_x = h2d((ref * (1 - mask)).astype(np.float32))
    nws = L.fse_mel_encoder_workspace_bytes(h, B, T)
    p = dmalloc(nws + 1024)
    ws = C.c_void_p((p.value + 1023) // 1024 * 1024)
    d_out = dmalloc(B * T * 192 * 4)
    _lib.check(L.fse_mel_encoder_forward(h, d_x, None, None, d_out, B, T, ws, nws, None)); sync("mel_encoder")
    f32 = mode == "simt_f32"
    report(f"{mode} mel_encoder", "out", d2h(d_out, (B, T, 192), np.float32), g["out"], 2e-4 if f32 else None, None if f32 else 1e-2)
    _lib.check(L.fse_mel_encoder_forward(h, d_x, h2d(g["decoder_inp"]), h2d(g["nonpad"]), d_out, B, T, ws, nws, None)); sync("mel_encoder fused")
    report(f"{mode} mel_encoder", "cond (fused)", d2h(d_out, (B, T, 192), np.float32), g["cond"], 2e-4 if f32 else None, None if f32 else 1e-2)


def main():
    modes = sys.argv[1:] or ["simt_f32", "tc_bf16", "simt_bf16"]
    L = _lib.lib()
    print(f"libfse_b200 version {L.fse_version()}  (+{time.time() - T0:.1f}s)", flush=True)
    for mode in modes:
        for fn in (check_cond, check_mel_encoder):
            try:
                fn(L, mode)
            except Exception:
                FAILS.append(f"{mode} {fn.__name__} raised")
                print(f"[FAIL] {mode} {fn.__name__} raised:\n{traceback.format_exc()}", flush=True)
        print(f"--- {mode} done (+{time.time() - T0:.1f}s)", flush=True)
    print("SUMMARY:", "ALL OK" if not FAILS else f"{len(FAILS)} FAILED: {FAILS}")
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())

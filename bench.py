#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: FluentSpeech 100-step spec_denoiser sampling
+ HiFi-GAN V1 forward on synthetic (cond, mel) batches of 32 utterances x 1024 frames per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the path's torch CPU port on host cores)

One "step" = one pass of the hot path over one batch: S=100 reverse-diffusion iterations of the DiffNet
denoiser (4 kernels each: input projection, the fused 20-layer kernel, folded skip GEMM, output projection + posterior)
followed by the vocoder forward (79 kernels).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, HOP = 22050, 256
WORKLOAD = "FluentSpeech 100-step sampling + HiFi-GAN V1, B=32 x T=1024 per GPU (BASELINE configs[1]+vocoder; configs[2] at 8 GPUs)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--mode", default="tc_tf32", choices=["tc_tf32", "tc_bf16", "simt_f32", "simt_bf16"],
                    help="arithmetic of the contractions. Default tc_tf32 = tcgen05 kind::tf32, the reference's own GPU arithmetic "
                         "(cuDNN TF32 convolutions); tc_bf16 = bf16 operands (faster, narrower than the reference: reported as a sub-record)")
    ap.add_argument("--no-alt-mode", action="store_true", help="skip the short second measurement in the other tensor-core mode")
    ap.add_argument("--no-campnet", action="store_true", help="skip the CampNet (BASELINE configs[3]) sub-record")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (BASELINE configs[4]) sub-record")
    ap.add_argument("--no-vocoder", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager-gpu-baseline", action="store_true", help="(default at N=1; kept for compatibility)")
    ap.add_argument("--no-eager-gpu-baseline", action="store_true",
                    help="skip timing the reference's own modules eagerly on this GPU (cuDNN/cuBLAS; rank 0, N=1)")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


DTYPE = {"tc_tf32": "tf32", "tc_bf16": "bf16", "simt_bf16": "bf16", "simt_f32": "f32"}


def tensor_peak(mode, pk):
    """Sustained tensor peak the stream kernel is measured against.  MEASURED_PEAKS.json holds the dense bf16 figure only; the
    tf32 kind runs at half the bf16 rate on this part (B200_PROFILING.md: 1.1 vs 2.25 PFLOP/s dense), so its peak is taken as
    half of the MEASURED sustained bf16 throughput."""
    if mode == "tc_tf32":
        return pk["bf16_sustained"] / 2.0, pk["src"] + ", sustained bf16 / 2 (tf32 kind = half the bf16 rate)"
    return pk["bf16_sustained"], pk["src"] + ", sustained bf16"


def library_gemm_check(dev, seconds=1.5):
    """Informational cross-check of the roofline denominators (not used for `frac`): cuBLAS 8192^3 GEMMs launched back to back for
    ~`seconds` each, bf16 and fp32-with-TF32 — the library's own sustained rate on THIS box under THIS power cap, next to the
    driver-measured bf16 figure in MEASURED_PEAKS.json and the tf32 = bf16 / 2 rule `peak` is derived with."""
    import torch
    out = {}
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.randn(n, n, device=dev, dtype=dt)
            b = torch.randn(n, n, device=dev, dtype=dt)
            c = torch.empty(n, n, device=dev, dtype=dt)
            for _ in range(5):
                torch.matmul(a, b, out=c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            reps = max(10, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
            e0.record()
            for _ in range(reps):
                torch.matmul(a, b, out=c)
            e1.record(); torch.cuda.synchronize()
            out[f"cublas_{name}_sustained_tflops"] = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
            del a, b, c
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out["how"] = f"torch.matmul {n}^3 back to back for ~{seconds} s per dtype on this box, after the timed regions"
    return out


def stream_traffic(mode, B, T):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the streamed residual-layer kernel at the bench shape, read
    from the committed ncu artefact profiles/stream_kernel_traffic.json (written by tools/ncu_traffic.py from an `ncu --set full`
    capture of this kernel); None when no capture of this mode / shape is committed."""
    p = os.path.join(ROOT, "profiles", "stream_kernel_traffic.json")
    try:
        d = json.load(open(p))
        e = d.get(f"{mode}:{B}x{T}")
        return (int(e["dram_bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def roofline_record(mode, prof_d, B, T, pk):
    g_ms, g_n = prof_d["gate_gemm"]
    fused = prof_d["res_gemm"][1] == 0
    if fused:   # one launch = all 20 residual layers: per layer gate GEMM 0.983 MFLOP/frame + residual GEMM 0.131 MFLOP/frame
        flops = 20 * 2.0 * B * T * (512 * 960 + 256 * 256)
        kname = (f"denoiser_stream_kernel<pair{', tf32' if mode == 'tc_tf32' else ''}> (20 x [k=3 conv + cond 1x1 + gate -> residual 1x1] "
                 "as one stream of (layer, unit) items; persistent, cta_group::2)")
    else:
        flops = 2.0 * B * T * 512 * 960                  # algorithmic: 0.983 MFLOP/frame (k=3 conv 512x768 + cond 512x192)
        kname = "conv_gemm_tc_kernel<EpiGate> (dilated k=3 conv + conditioner 1x1 + gate)"
    ach = flops / (g_ms / g_n * 1e-3) / 1e12 if g_n else 0.0
    peak, src = tensor_peak(mode, pk)
    traffic, tsrc = stream_traffic(mode, B, T) if fused and os.environ.get("FSE_FUSED_STREAM", "1") != "0" else (None, None)
    return {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": src, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": tsrc, "avg_launch_us": g_ms / g_n * 1e3 if g_n else None, "launches_timed": g_n, "flops_per_launch": flops}


# ------------------------------------------------------------------ reference arm (the reference's own modules)
_REF_MODELS = {}


def _ref_models(timesteps, device):
    from oracle import ref_runner as R
    key = (timesteps, str(device))
    if key not in _REF_MODELS:
        _REF_MODELS[key] = R.build_models(timesteps, device)
    return _REF_MODELS[key]


def _pick_threads(timesteps):
    """"All the host threads it can use": the box may expose more logical CPUs than the container's quota, and oversubscribed
    OpenMP teams collapse, so probe a few team sizes on one p_sample iteration and keep the fastest."""
    import torch
    from oracle import ref_runner as R
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    model, _ = _ref_models(timesteps, "cpu")
    tb = R.synthetic_inputs(2, 1024, "cpu")
    best = None
    with torch.no_grad():
        cond = R.run_condition(model, tb)
        x = torch.randn(2, 1, 80, 1024)
        t = torch.full((2,), timesteps - 1, dtype=torch.long)
        for n in sorted({avail, max(avail // 2, 1), 32, 16, 8}):
            if n > avail:
                continue
            torch.set_num_threads(n)
            model.p_sample(x, t, cond)
            t0 = time.perf_counter()
            model.p_sample(x, t, cond)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
    return best[1]


def cpu_sample(timesteps, threads=None, B=4, T=1024, iters=3, Bv=1, Tv=256):
    """Bounded CPU sample of the same workload on the reference's OWN modules (oracle/ref_runner.py): condition encoder
    (fs + mel_encoder) and `iters` p_sample iterations at B x T, HifiGanGenerator at Bv x Tv, extrapolated to
    cond + S iterations + vocoder per frame.  Falls back to the validated torch port only when neither
    /root/reference nor oracle/_ref exists (kind "port")."""
    import torch
    from oracle import ref_runner as R
    if R.available():
        if threads is None:
            threads = _pick_threads(timesteps)
        r = R.time_reference(B, T, timesteps, iters=iters, Bv=Bv, Tv=Tv, device="cpu", threads=threads, models=_ref_models(timesteps, "cpu"))
        return {"value": r["frames_per_s"], "unit": "mel-frames/s", "cores": int(threads), "kind": "reference", "sample": r["sample"],
                "per_frame_s": r["per_frame"]}
    from oracle import fluentspeech_oracle as O
    from oracle import torch_port as P
    from speech_editing_toolkit_b200 import schedule, synth
    p = P.to_torch(synth.denoiser_state_dict(1234))
    hp = P.to_torch(synth.hifigan_state_dict(1234))
    sched = {k: torch.from_numpy(v) for k, v in schedule.diffusion_buffers(timesteps).items()}
    cond = torch.from_numpy(synth.synthetic_cond(1, B, T)).transpose(1, 2).contiguous()
    threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    torch.set_num_threads(threads)
    P.sample_loop(p, sched, cond, timesteps, steps=1)                  # warm-up
    t0 = time.perf_counter()
    P.sample_loop(p, sched, cond, timesteps, steps=iters)
    t_fs = (time.perf_counter() - t0) / (B * T * iters)                # seconds per frame-iteration
    mel = torch.randn(Bv, 80, Tv) - 3
    P.hifigan_forward(hp, O.HIFIGAN_V1, mel[:, :, :32])
    t0 = time.perf_counter()
    P.hifigan_forward(hp, O.HIFIGAN_V1, mel)
    t_voc = (time.perf_counter() - t0) / (Bv * Tv)
    sec_per_frame = t_fs * timesteps + t_voc
    return {"value": 1.0 / sec_per_frame, "unit": "mel-frames/s", "cores": threads, "kind": "port",
            "sample": f"torch CPU port (oracle/torch_port.py; no reference tree on this box): denoiser B={B}xT={T} for {iters} of {timesteps} "
                      f"iterations ({t_fs * 1e6:.2f} us/frame-iteration) + HiFi-GAN B={Bv}xT={Tv} ({t_voc * 1e6:.1f} us/frame), extrapolated",
            "per_frame_s": {"cond": 0.0, "iter": t_fs, "vocoder": t_voc}}


def cpu_baseline_record(timesteps):
    """cpu_baseline of the B200 line: all-cores row (the value) + the as-shipped OMP_NUM_THREADS=1 row (tasks/run.py:3)."""
    allc = cpu_sample(timesteps)
    one = cpu_sample(timesteps, threads=1, B=1, T=1024, iters=2, Bv=1, Tv=128)
    rec = {k: allc[k] for k in ("value", "unit", "cores", "kind", "sample")}
    rec["single_thread_as_shipped"] = {"value": one["value"], "unit": "mel-frames/s", "cores": 1, "sample": one["sample"],
                                       "note": "tasks/run.py:3 sets OMP_NUM_THREADS=1 for the reference CLI"}
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    threads = None
    for _ in range(min(args.warmup, 1)):
        threads = cpu_sample(args.timesteps)["cores"]
    t0 = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        vals.append(cpu_sample(args.timesteps, threads=threads))
        threads = vals[-1]["cores"]
    wall = time.perf_counter() - t0
    best = max(vals, key=lambda v: v["value"])
    xs = [x["value"] for x in vals]
    v = float(np.mean(xs))
    frames = args.batch * args.frames
    line = {"impl": "reference", "metric": "mel-frames/sec (100-step sampling + HiFi-GAN)", "value": v, "unit": "mel-frames/s",
            "rtf": (1.0 / v) / (HOP / SR), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": frames / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "frames": args.frames, "timesteps": args.timesteps,
                       "vocoder": "HiFi-GAN V1 (assumed config, SURVEY fact 4)",
                       "note": "CPU arm: the reference's own modules on the host cores; each step is a bounded sample (B=4 x T=1024, 3 of S "
                               "iterations + condition encoder + vocoder B=1 x T=256), extrapolated; ms_per_step is the extrapolated full-batch time"},
            "cpu_baseline": {k: best[k] for k in ("cores", "kind", "sample")} | {"value": v, "unit": "mel-frames/s"},
            "spread": {"min": float(min(xs)), "max": float(max(xs)), "rel": float((max(xs) - min(xs)) / v)},
            "e2e": {"value": v, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line))


def eager_gpu_sample(timesteps, B, T, dev):
    """The reference's own way of running this path on a GPU: its unmodified modules after `.cuda()` — eager PyTorch, one
    cuDNN / ATen launch per op, fp32 with TF32 convolutions (torch default) — timed by oracle/ref_runner.py: condition encoder
    + 3 p_sample iterations at the full batch + HifiGanGenerator on a bounded batch, extrapolated to S iterations + vocoder.
    This is the "reference single-GPU PyTorch" figure north_star compares against.  Reported, never the product path."""
    import torch
    from oracle import ref_runner as R
    if not R.available():
        return {"unavailable": "no reference tree (/root/reference or oracle/_ref) on this box"}
    Bv = min(B, 4)                               # the eager vocoder holds fp32 activations of every stage: bound its batch
    r = R.time_reference(B, T, timesteps, iters=3, Bv=Bv, Tv=T, device=dev)
    _REF_MODELS.clear()
    torch.cuda.empty_cache()
    return {"value": r["frames_per_s"], "unit": "mel-frames/s", "rtf": r["sec_per_frame"] / (HOP / SR),
            "kind": "reference (unmodified modules, eager torch on cuda, TF32 convs as torch defaults)", "sample": r["sample"],
            "allow_tf32": {"cudnn": bool(torch.backends.cudnn.allow_tf32), "matmul": bool(torch.backends.cuda.matmul.allow_tf32)}}


def campnet_record(pk, B=64, T=1024, mode="tc_bf16", iters=5):
    """BASELINE configs[3]: CampNet (egs/campnet.yaml) mask-predict forward, batch 64 x 1024 frames x 128 tokens, inputs resident,
    through CampNetB200.forward (ctypes -> fse_campnet_forward).  Algorithmic FLOPs as in tools/campnet_bench.py."""
    import torch
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.modules import CampNetB200
    net = CampNetB200(80, 100, dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80, b200_mode=mode)).cuda()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.campnet_state_dict(1234, 80).items()}, strict=False)
    b = synth.synthetic_campnet_batch(1, B, T, vocab=80)
    txt, mels, m = (torch.from_numpy(b[k]).cuda() for k in ("txt_tokens", "mels", "time_mel_masks"))
    for _ in range(3):
        out = net(txt, mels=mels, time_mel_masks=m, infer=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = net(txt, mels=mels, time_mel_masks=m, infer=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    Tt, H = txt.shape[1], 192
    enc_tok = 3 * (2 * H * 3 * H + 2 * H * H + 2 * H * 4 * H * 9 + 2 * 4 * H * H)
    dec_frm = 6 * (2 * H * 3 * H + 2 * H * H + 2 * H * H + 2 * H * H + 2 * H * 4 * H * 9 + 2 * 4 * H * H)
    fine_frm = 10 * (2 * H * 2 * H * 5 + 2 * 2 * H * H) + 2 * H * H * 3
    mel_enc = 2 * (2 * 80 * H + 2 * 2 * H * H) + 2 * 2 * H * 80
    gemm = B * Tt * (enc_tok + 6 * 2 * H * 2 * H) + B * T * (dec_frm + fine_frm + mel_enc)
    attn = B * (3 * 4 * Tt * Tt * H + 6 * 4 * T * T * H + 6 * 4 * T * Tt * H)
    ach = (gemm + attn) / (ms / 1e3) / 1e12
    rec = {"workload": f"CampNet (egs/campnet.yaml) mask-predict forward, batch {B} x {T} frames x {Tt} tokens (BASELINE configs[3])",
           "mode": mode, "dtype": DTYPE[mode], "ms_per_forward": ms, "value": B * T / (ms / 1e3), "unit": "mel-frames/s", "launches": int(net.engine().last_launches),
           "algorithmic_tflop": (gemm + attn) / 1e12, "attention_tflop": attn / 1e12,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                        "scope": "whole forward (all kernels), algorithmic FLOPs / CUDA-event time"},
           "finite": bool(torch.isfinite(out["mel_out_fine"]).all().item())}
    del net
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:                                   # make sure the in-tree library matches the sources (no-op when current)
        from speech_editing_toolkit_b200 import build as _b
        try:
            _b.build()
        except Exception as e:                      # no nvcc on this box: use the shipped .so
            print(f"bench: build skipped ({e})", file=sys.stderr)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from speech_editing_toolkit_b200 import schedule, synth
    from speech_editing_toolkit_b200.engine import Denoiser, Vocoder

    B, T, S = args.batch, args.frames, args.timesteps
    den = Denoiser(mode=args.mode)
    den.load_state_dict(synth.denoiser_state_dict(1234))
    buf = schedule.diffusion_buffers(S)
    den.set_schedule(buf["posterior_mean_coef1"], buf["posterior_mean_coef2"], buf["posterior_log_variance_clipped"])
    voc = None
    if not args.no_vocoder:
        voc = Vocoder(mode=args.mode)
        voc.load_state_dict(synth.hifigan_state_dict(1234))

    # synthetic batch of this rank's shard (utterances rank*B .. rank*B+B-1)
    batch = synth.synthetic_edit_batch(1000 + rank, B, T)
    cond_h = torch.from_numpy(synth.synthetic_cond(1000 + rank, B, T)).pin_memory()
    ref_h = torch.from_numpy(batch["ref_mels"]).pin_memory()
    mask_h = torch.from_numpy(batch["time_mel_masks"]).pin_memory()
    cond_d, ref_d, mask_d = cond_h.to(dev), ref_h.to(dev), mask_h.to(dev)
    gathered = torch.empty(world * B, T, 80, device=dev) if world > 1 else None
    mel_h = torch.empty(B, T, 80).pin_memory()
    wav_h = torch.empty(B, T * HOP).pin_memory()

    def step_resident(i):
        mel = den.sample(cond_d, None, seed=i, ref_mel=ref_d, mask=mask_d)
        wav = voc.forward(mel) if voc else None
        if world > 1:
            dist.all_gather_into_tensor(gathered, mel)       # the one collective of the path
        return mel, wav

    def step_e2e(i):
        c = cond_h.to(dev, non_blocking=True); r = ref_h.to(dev, non_blocking=True); m = mask_h.to(dev, non_blocking=True)
        mel = den.sample(c, None, seed=i, ref_mel=r, mask=m)
        mel_h.copy_(mel, non_blocking=True)
        if voc:
            wav_h.copy_(voc.forward(mel), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # the whole model behind the reference's own call (GaussianDiffusion.forward(infer=True), spec_denoiser.py:154-185):
    # text tokens -> condition encoder -> MelEncoder -> S-step sampling (+ mask compositing), then the vocoder
    model, txt_host = None, None
    if not args.no_e2e or not args.no_kernel_timing:
        from speech_editing_toolkit_b200 import plugin
        hp = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=20, residual_channels=256, dilation_cycle_length=1,
                  timesteps=S, timescale=1, diff_loss_type="l1", spec_min=[], spec_max=[], keep_bins=80, schedule_type="vpsde",
                  diff_decoder_type="wavenet_b200", b200_mode=args.mode)
        model = plugin.build_diffusion(hp, phone_encoder=list(range(80)))
        t_ = lambda sd: {k: torch.from_numpy(v) for k, v in sd.items()}
        model.fs.load_state_dict(t_(synth.fastspeech_state_dict(1234, 80)), strict=False)       # decoder.* / mel_out.* unused (skip_decoder)
        model.mel_encoder.load_state_dict(t_(synth.mel_encoder_state_dict(1234)))
        model.denoise_fn.load_state_dict(t_(synth.denoiser_state_dict(1234)))
        model = model.to(dev).eval()
        txt_host = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("txt_tokens", "mel2ph", "time_mel_masks", "spk_embed", "ref_mels", "f0", "uv")}
        txt_dev = {k: v.to(dev) for k, v in txt_host.items()}

    def text_to_mel(t, i):
        ret = model(t["txt_tokens"], t["time_mel_masks"][:, :, None], t["mel2ph"], t["spk_embed"], t["ref_mels"], t["f0"], t["uv"],
                    infer=True, use_pred_pitch=True, seed=i, composite=True)
        return ret["mel_out"]

    def step_e2e_text(i):
        t = {k: v.to(dev, non_blocking=True) for k, v in txt_host.items()}
        mel = text_to_mel(t, i)
        mel_h.copy_(mel, non_blocking=True)
        if voc:
            wav_h.copy_(voc.forward(mel), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_cond_encoder(i):                       # condition encoder + MelEncoder alone, inputs resident (for the breakdown)
        t = txt_dev
        ret = model.fs(t["txt_tokens"], t["time_mel_masks"][:, :, None], t["mel2ph"], t["spk_embed"], t["f0"], t["uv"], None,
                       skip_decoder=True, infer=True, use_pred_pitch=True)
        model.mel_encoder.fused(t["ref_mels"] * (1 - t["time_mel_masks"][:, :, None]), ret["decoder_inp"],
                                (t["mel2ph"] > 0).float()[:, :, None])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, profile=False):
        barrier()
        if profile:
            den.profile(True)
            if voc:
                voc.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    # Timed region 1 (the headline `value`): K steps of the product path exactly as a user runs it - the 100-iteration loop is a
    # replayed CUDA graph (fse_sample captures a call signature on its second sight; the warm-up steps above did that).
    clk = ClockSampler(local)
    clk.start()
    use_prof = not args.no_kernel_timing
    ms_total = timed(step_resident, args.steps, profile=False)
    launches = den.last_launches + (voc.last_launches if voc else 0)
    ms_step = ms_total / args.steps
    frames_total = world * B * T
    value = frames_total / (ms_step / 1e3)
    # Timed region 2 (roofline / breakdown): the same K steps with a CUDA-event pair around every kernel on the launch stream.
    # Events cannot be recorded inside a graph replay, so this pass launches the same kernels eagerly; its step time is reported
    # as ms_per_step_with_kernel_events.  The clock sampler spans both regions.
    ms_prof = timed(step_resident, args.steps, profile=True) / args.steps if use_prof else ms_step
    clocks = clk.stop()
    prof_d = den.profile_read() if use_prof else None
    prof_v = voc.profile_read() if (use_prof and voc) else None
    if use_prof:
        den.profile(False)
        if voc:
            voc.profile(False)

    e2e = None
    cond_ms = None
    if not args.no_e2e:
        d2h = mel_h.numel() * 4 + (wav_h.numel() * 4 if voc else 0)
        for i in range(3):          # warm-up: fse_sample captures a call signature on its second sight (first eager, then capture, then replay)
            step_e2e(i)
        ms_cond = timed(step_e2e, args.steps) / args.steps
        h2d_cond = cond_h.numel() * 4 + ref_h.numel() * 4 + mask_h.numel() * 4
        from_cond = {"value": frames_total / (ms_cond / 1e3), "unit": "mel-frames/s", "h2d_bytes_per_step": h2d_cond,
                     "d2h_bytes_per_step": d2h, "ms_per_step": ms_cond,
                     "api": "Denoiser.sample + Vocoder.forward (ctypes -> fse_sample / fse_vocoder_forward) from a ready cond[B,T,192], pinned host in/out"}
        for i in range(3):
            step_e2e_text(i)
        ms_e2e = timed(step_e2e_text, args.steps) / args.steps
        h2d = sum(v.numel() * v.element_size() for v in txt_host.values())
        e2e = {"value": frames_total / (ms_e2e / 1e3), "unit": "mel-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e,
               "api": "GaussianDiffusionB200.forward(txt_tokens, time_mel_masks, mel2ph, spk_embed, ref_mels, f0, uv, infer=True, "
                      "use_pred_pitch=True) + Vocoder.forward: text tokens -> wav, every sub-module native "
                      "(ctypes -> fse_cond_* / fse_mel_encoder_forward / fse_sample / fse_vocoder_forward), pinned host in/out",
               "from_cond": from_cond}
    if use_prof and model is not None:
        step_cond_encoder(0)
        cond_ms = timed(step_cond_encoder, args.steps) / args.steps

    def train_record(steps=5, mode="tc_bf16"):
        """BASELINE configs[4]: the denoiser branch of the FluentSpeech training step in bf16, data-parallel over the ranks of this
        launch (one bucketed NCCL all-reduce per residual layer, overlapped with the remaining weight-gradient GEMMs), 32 x 1024-frame
        synthetic batches per GPU: q_sample, DiffNet forward + backward (native data path, tcgen05 weight-gradient GEMMs over MN-major operands), masked
        l1 + ssim loss, fused AdamW.  The condition encoder's forward / backward is not part of it (cond is a resident input)."""
        from speech_editing_toolkit_b200 import train
        from speech_editing_toolkit_b200.modules import DiffNetB200
        hp_t = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=20, residual_channels=256, dilation_cycle_length=1, b200_mode=mode)
        net = DiffNetB200(80, hp_t).to(dev).train()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(1234).items()})
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.98), weight_decay=0.0, fused=True)
        sched = {k: torch.from_numpy(v).to(dev) for k, v in schedule.diffusion_buffers(S).items()}
        data = {"ref_mels": ref_d, "time_mel_masks": mask_d, "cond": cond_d}
        red = train.BucketedAllReduce(dict(net.named_parameters())) if world > 1 else None
        losses = None
        for _ in range(2):
            losses = train.train_step(net, sched, data, opt, reducer=red)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            losses = train.train_step(net, sched, data, opt, reducer=red)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            tt_ = torch.tensor([ms], device=dev)
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
            ms = float(tt_.item())
        flops = 3 * 25.13e6 * B * T                      # forward + backward = 3 x the forward's algorithmic FLOPs
        rec = {"workload": f"FluentSpeech training step, denoiser branch (DiffNet fwd + bwd, masked l1 + ssim, AdamW), {B} x {T} frames per GPU, "
                           f"DDP x{world} (BASELINE configs[4])", "mode": mode, "dtype": DTYPE[mode], "ms_per_step": ms, "steps": steps,
               "value": world * B * T / (ms / 1e3), "unit": "mel-frames/s (training)", "n_gpus": world,
               "achieved_tflops_per_gpu": flops / (ms / 1e3) / 1e12, "loss": losses,
               "allreduce": "one async NCCL all-reduce per residual layer's gradients, overlapped with the following layers' weight-gradient GEMMs" if world > 1 else None}
        del net, opt
        torch.cuda.empty_cache()
        return rec

    train_rec = None
    if not args.no_train:
        try:
            train_rec = train_record()
        except Exception as e:                           # a reported extra: never takes the bench line down
            train_rec = {"error": repr(e)}
            if world > 1:
                raise

    def alt_record(mode, steps=3):
        """The same resident step in the other tensor-core mode (3 steps after 3 warm-ups, rank 0's GPU only, no collective):
        reported next to the headline so both precisions are on the driver's record."""
        d2 = Denoiser(mode=mode)
        d2.load_state_dict(synth.denoiser_state_dict(1234))
        d2.set_schedule(buf["posterior_mean_coef1"], buf["posterior_mean_coef2"], buf["posterior_log_variance_clipped"])
        v2 = None
        if voc:
            v2 = Vocoder(mode=mode)
            v2.load_state_dict(synth.hifigan_state_dict(1234))

        def one(i):
            mel = d2.sample(cond_d, None, seed=i, ref_mel=ref_d, mask=mask_d)
            if v2:
                v2.forward(mel)
        for i in range(3):
            one(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        d2.profile(True)                              # second pass with per-kernel events for the roofline / breakdown
        if v2:
            v2.profile(True)
        for i in range(steps):
            one(i)
        torch.cuda.synchronize()
        pd = d2.profile_read()
        pv = v2.profile_read() if v2 else None
        rec = {"mode": mode, "dtype": DTYPE[mode], "value": B * T / (ms / 1e3), "unit": "mel-frames/s (one GPU)", "ms_per_step": ms, "steps": steps,
               "roofline": roofline_record(mode, pd, B, T, peaks()),
               "breakdown": {"denoiser_ms_per_step": {k: v[0] / steps for k, v in pd.items()}}}
        if pv:
            rec["breakdown"]["vocoder_ms_per_step"] = {k: v[0] / steps for k, v in pv.items()}
        del d2, v2
        torch.cuda.empty_cache()
        return rec

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    roofline = roofline_record(args.mode, prof_d, B, T, pk) if prof_d else None
    if roofline is not None and world == 1:
        try:
            roofline["library_check"] = library_gemm_check(dev)
        except Exception as e:      # informational only
            roofline["library_check"] = {"error": f"{type(e).__name__}: {e}"}
    breakdown = {}
    if prof_d:
        breakdown["denoiser_ms_per_step"] = {k: v[0] / args.steps for k, v in prof_d.items()}
    if prof_v:
        breakdown["vocoder_ms_per_step"] = {k: v[0] / args.steps for k, v in prof_v.items()}
    if cond_ms is not None:      # once per batch, in front of the loop: part of e2e, not of the resident `value` step (sampling + vocoder)
        breakdown["cond_encoder_plus_mel_encoder_ms"] = cond_ms
    eager = None
    if not args.no_eager_gpu_baseline and world == 1:
        try:
            eager = eager_gpu_sample(S, B, T, dev)
        except Exception as e:                       # a reported extra: never takes the bench line down
            eager = {"error": repr(e)}
    alt = None
    if not args.no_alt_mode and args.mode in ("tc_tf32", "tc_bf16"):
        alt_mode = "tc_bf16" if args.mode == "tc_tf32" else "tc_tf32"
        try:
            alt = alt_record(alt_mode)
        except Exception as e:                       # a reported extra: never takes the bench line down
            alt = {"mode": alt_mode, "error": repr(e)}
    camp = None
    if not args.no_campnet and world == 1:
        try:
            camp = campnet_record(pk)
        except Exception as e:                       # a reported extra: never takes the bench line down
            camp = {"error": repr(e)}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_record(S)
    audio_s = frames_total * HOP / SR
    line = {
        "metric": "mel-frames/sec (100-step sampling + HiFi-GAN)", "value": value, "unit": "mel-frames/s",
        "rtf": (ms_step / 1e3) / audio_s, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "ms_per_step_with_kernel_events": ms_prof, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE[args.mode], "data": "synthetic",
        "config": {"workload": WORKLOAD if (B, T, S) == (32, 1024, 100) and voc else f"custom B={B} T={T} S={S} vocoder={bool(voc)}",
                   "batch_per_gpu": B, "frames": T, "timesteps": S, "vocoder": "HiFi-GAN V1 (assumed config, SURVEY fact 4)" if voc else None,
                   "mode": args.mode, "noise": "in-kernel Philox4x32-10",
                   "launch": "sampling loop = one CUDA graph replay per step (captured by fse_sample), vocoder = stream launches; "
                             "roofline / breakdown from a second pass of the same steps with per-kernel CUDA events (eager launches)", "weights": "seeded random (synth.py), reference state_dict layout",
                   "l2": "per-step working set (activations + vocoder workspace, >5 GB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"batch-sharded x{world}, one all_gather of mels per step" if world > 1 else "single GPU"},
        "gpu_launches": int(launches * args.steps), "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        "breakdown": breakdown,
    }
    if alt is not None:
        line["alt_mode"] = alt
    if camp is not None:
        line["campnet"] = camp
    if train_rec is not None:
        line["train_step"] = train_rec
    if eager is not None:
        line["eager_gpu_baseline"] = eager
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

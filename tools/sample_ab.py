"""A/B of environment switches of the sampling path (read at handle creation): 100-step sampling at 32 x 1024, graph replay, CUDA events.
    python tools/sample_ab.py tc_tf32 FSE_STREAM_PDL=0 FSE_STREAM_PDL=1
Every setting's mel must be bit-identical to the first one's (these switches change schedules, never arithmetic)."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_editing_toolkit_b200 import schedule, synth
from speech_editing_toolkit_b200.engine import Denoiser

mode = sys.argv[1] if len(sys.argv) > 1 else "tc_tf32"
settings = [a for a in sys.argv[2:] if "=" in a] or ["FSE_STREAM_PDL=0", "FSE_STREAM_PDL=1"]
B, T, S = 32, 1024, 100
sd = synth.denoiser_state_dict(1234)
b = schedule.diffusion_buffers(S)
cond = torch.from_numpy(synth.synthetic_cond(1, B, T)).cuda()
ref = None
for setting in settings:
    kv = dict(x.split("=", 1) for x in setting.split(","))
    os.environ.update(kv)
    d = Denoiser(mode=mode)
    d.load_state_dict(sd)
    d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
    for i in range(3):                       # eager, capture, first replay
        mel = d.sample(cond, None, seed=7)
    torch.cuda.synchronize()
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mel = d.sample(cond, None, seed=7)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ref is None:
        ref = mel.clone()
    crc = zlib.crc32(mel.cpu().numpy().tobytes())
    print(f"{mode} {setting}: {min(ts):8.2f} ms best, {sorted(ts)[len(ts) // 2]:8.2f} ms median per {S}-step sampling of {B} x {T}; crc {crc:08x}, "
          f"bit-identical to the first setting: {bool(torch.equal(mel, ref))}", flush=True)
    for k in kv:
        os.environ.pop(k, None)
    del d

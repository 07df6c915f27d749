"""A/B of the streamed kernel's L2 eviction-priority switches (FSE_STREAM_L2HINT, denoiser_stream.cuh kL2Hint* bits):
    python tools/l2hint_sweep.py tc_tf32 0,1,3,7,19        # 100-step sampling at 32 x 1024, graph replay, CUDA events
    python tools/l2hint_sweep.py tc_tf32 3 --ncu           # short eager run for an ncu pass over the kernel (no timing)
Results must not depend on the mask: every mask's mel is compared bit for bit with mask 0's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_editing_toolkit_b200 import schedule, synth
from speech_editing_toolkit_b200.engine import Denoiser

mode = sys.argv[1] if len(sys.argv) > 1 else "tc_tf32"
masks = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "0,1,3,7,19").split(",")]
ncu = "--ncu" in sys.argv
B, T, S = 32, 1024, (6 if ncu else 100)
sd = synth.denoiser_state_dict(1234)
b = schedule.diffusion_buffers(S)
cond = torch.from_numpy(synth.synthetic_cond(1, B, T)).cuda()
ref = None
for m in masks:
    os.environ["FSE_STREAM_L2HINT"] = str(m)
    d = Denoiser(mode=mode)
    d.load_state_dict(sd)
    d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
    if ncu:
        d.sample(cond, None, seed=7)
        torch.cuda.synchronize()
        continue
    for i in range(3):                       # eager, capture, first replay
        mel = d.sample(cond, None, seed=7)
    torch.cuda.synchronize()
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mel = d.sample(cond, None, seed=7)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ref is None:
        ref = mel.clone()
    same = bool(torch.equal(mel, ref))
    import zlib
    crc = zlib.crc32(mel.cpu().numpy().tobytes())
    print(f"{mode} pdl={os.environ.get('FSE_STREAM_PDL', '0')} crc={crc:08x} l2_hint={m:2d}: {min(ts):8.2f} ms best, {sorted(ts)[len(ts) // 2]:8.2f} ms median per {S}-step sampling of {B} x {T}; bit-identical to mask {masks[0]}: {same}", flush=True)
    del d

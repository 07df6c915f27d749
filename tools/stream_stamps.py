"""Timeline of the streamed residual-layer kernel (clock64 stamps of CTA 0, FSE_DBG_STAMPS=1): python tools/stream_stamps.py tc_tf32"""
import os, sys
os.environ["FSE_DBG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_editing_toolkit_b200 import schedule, synth
from speech_editing_toolkit_b200.engine import Denoiser
mode = sys.argv[1] if len(sys.argv) > 1 else "tc_tf32"
B, T, S = 32, 1024, 3
d = Denoiser(mode=mode)
d.load_state_dict(synth.denoiser_state_dict(1234))
b = schedule.diffusion_buffers(S)
d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
cond = torch.from_numpy(synth.synthetic_cond(1, B, T)).cuda()
print(f"== {mode}", file=sys.stderr)
for i in range(2):
    d.sample(cond, None, seed=i)
torch.cuda.synchronize()

"""One vocoder forward at the bench shape (for `ncu --metrics gpu__time_duration.sum`): python tools/voc_launches.py tc_tf32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_editing_toolkit_b200 import synth
from speech_editing_toolkit_b200.engine import Vocoder
mode = sys.argv[1] if len(sys.argv) > 1 else "tc_tf32"
v = Vocoder(mode=mode)
v.load_state_dict(synth.hifigan_state_dict(1234))
mel = torch.randn(32, 1024, 80, device="cuda") * 1.5 - 3
for _ in range(2):
    v.forward(mel)
torch.cuda.synchronize()

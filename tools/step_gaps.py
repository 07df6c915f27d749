"""Device-side timeline of one bench step (100-step sampling + vocoder, 32 x 1024, CUPTI through torch.profiler): how much of the step is
gaps between kernels, and after which kernels they occur.  usage: python tools/step_gaps.py [tc_tf32|tc_bf16]"""
import collections
import json
import os
import re
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_editing_toolkit_b200 import schedule, synth  # noqa: E402
from speech_editing_toolkit_b200.engine import Denoiser, Vocoder  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tc_tf32"
    B, T, S = 32, 1024, 100
    dev = torch.device("cuda:0")
    den = Denoiser(mode=mode)
    den.load_state_dict(synth.denoiser_state_dict(1234))
    bufs = schedule.diffusion_buffers(S)
    den.set_schedule(*[torch.from_numpy(bufs[k]).to(dev) for k in ("posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped")])
    voc = Vocoder(mode=mode)
    voc.load_state_dict(synth.hifigan_state_dict(1234))
    cond = torch.from_numpy(synth.synthetic_cond(1000, B, T)).to(dev)
    for i in range(3):
        mel = den.sample(cond, None, seed=i)
        voc.forward(mel)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        mel = den.sample(cond, None, seed=7)
        voc.forward(mel)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    busy = sum(e["dur"] for e in ev)
    gaps = collections.defaultdict(lambda: [0, 0.0])
    end = ev[0]["ts"] + ev[0]["dur"]
    for a, b in zip(ev, ev[1:]):
        g = b["ts"] - max(end, a["ts"] + a["dur"])
        end = max(end, b["ts"] + b["dur"])
        key = re.sub(r"<.*", "", a["name"].split("(")[0])[-40:] + " -> " + re.sub(r"<.*", "", b["name"].split("(")[0])[-40:]
        if g > 0:
            gaps[key][0] += 1
            gaps[key][1] += g
    print(f"{mode}: {len(ev)} device activities, span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, gaps {(t1 - t0 - busy) / 1e3:.2f} ms")
    for k, (n, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"  {us / 1e3:7.3f} ms x{n:<4d} avg {us / n:6.1f} us  {k}")


if __name__ == "__main__":
    main()

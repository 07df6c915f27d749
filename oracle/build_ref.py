"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference modules ON THE HOT PATH, so that the UNMODIFIED
reference can run where /root/reference does not exist (the GPU box: oracle/_ref travels with the snapshot like a built .so).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python with no setup.py / pyproject, so the
base contract's `pip install --target baseline/_ref /root/reference` has nothing to build; this script is that install.  It
imports the path's entry modules from /root/reference in a subprocess (GaussianDiffusion + DiffNet + FastSpeech + MelEncoder,
HifiGanGenerator, CampNet), records the import closure (about 25 files) and copies exactly those files plus egs/*.yaml —
no audio, no checkpoints — writing nothing outside oracle/_ref/.  oracle/_ref/ is listed in .gitignore: reference sources never
enter this repository's history.  Users: bench.py's reference arm (`--impl reference`, `cpu_baseline`, `eager_gpu_baseline`)
and the tests that pin the oracles against the live reference.

  python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FSE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

_CLOSURE = r"""
import os, sys, json
sys.path.insert(0, {root!r})
os.environ["FSE_REFERENCE_ROOT"] = {src!r}
from oracle import refshim
refshim.install("egs/spec_denoiser.yaml", overrides="timesteps=10")
import modules.speech_editing.spec_denoiser.spec_denoiser, modules.speech_editing.spec_denoiser.diffnet
import modules.speech_editing.spec_denoiser.fs, modules.speech_editing.commons.mel_encoder
import modules.vocoder.hifigan.hifigan, modules.speech_editing.campnet.campnet
src = os.path.realpath({src!r}) + os.sep
out = sorted(os.path.relpath(os.path.realpath(m.__file__), src) for m in list(sys.modules.values())
             if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(src))
print("CLOSURE=" + json.dumps(out))
"""


def build(verbose: bool = False) -> str:
    """Copy the hot path's import closure to oracle/_ref/.  Returns the destination, or '' when neither the source tree nor a
    previous copy exists (GPU box: the prebuilt copy is used as-is)."""
    if not os.path.isdir(os.path.join(SRC, "modules", "speech_editing")):
        return DST if os.path.isdir(os.path.join(DST, "modules")) else ""
    code = _CLOSURE.format(root=os.path.dirname(HERE), src=SRC)
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    line = [l for l in proc.stdout.splitlines() if l.startswith("CLOSURE=")]
    if proc.returncode != 0 or not line:
        raise RuntimeError("oracle/build_ref.py: could not import the reference path:\n" + proc.stdout + proc.stderr)
    files = json.loads(line[0][len("CLOSURE="):])
    files += [os.path.join("egs", f) for f in sorted(os.listdir(os.path.join(SRC, "egs"))) if f.endswith(".yaml")]
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in files:
        s, t = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(t), exist_ok=True)
        shutil.copyfile(s, t)
        manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()[:16]
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=0, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied verbatim from {SRC}")
    return DST


if __name__ == "__main__":
    out = build(verbose=True)
    sys.exit(0 if out else 1)

"""Build the CUDA extension (C-ABI shared library) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
libfse_b200.so is git-ignored but travels to the GPU box with the repo snapshot.  Each source is compiled
to its own object (in parallel, only when it or a header changed) and the objects are linked into the library.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libfse_b200.so")
SOURCES = ["denoiser.cu", "hifigan.cu", "mel_encoder.cu", "cond_encoder.cu", "campnet.cu", "edit_region.cu", "mel_frontend.cu", "mel_loss.cu", "wgrad.cu", "debug.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA toolkit is required to build libfse_b200.so")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "fse_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def needs_build() -> bool:
    return _stale(LIB_PATH, [os.path.join(CSRC, s) for s in _sources()] + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = _headers()
    jobs = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj + ".tmp"]
        if verbose:
            print(" ".join(cmd))
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            return f"nvcc failed on {src}:\n{proc.stdout}{proc.stderr}"
        os.replace(obj + ".tmp", obj)
        return None

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as pool:
        errors = [e for e in pool.map(compile_one, jobs) if e]
    if errors:
        raise RuntimeError("\n".join(errors))
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in _sources()]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_PATH + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

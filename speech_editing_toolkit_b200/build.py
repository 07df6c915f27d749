"""Build the CUDA extension (C-ABI shared library) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
libfse_b200.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfse_b200.so")
SOURCES = ["denoiser.cu", "hifigan.cu", "mel_encoder.cu", "cond_encoder.cu", "campnet.cu", "edit_region.cu", "mel_frontend.cu", "debug.cu"]
HEADERS = ["conv_gemm.cuh", "epilogues.cuh", "fse_common.cuh", "ptx_sm100.cuh", "denoiser_fused.cuh", "denoiser_stream.cuh", "rowwise.cuh", "attention_tc.cuh", "edit_region_core.h", "mel_frontend_weights.h",
           "../../include/fse_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA toolkit is required to build libfse_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH + ".tmp"]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

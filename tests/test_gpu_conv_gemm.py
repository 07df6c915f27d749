"""The conv-as-shifted-GEMM primitive in isolation (C ABI test hook fse_debug_conv_gemm): tcgen05 path vs
the CUDA-core path vs a float64 host computation over the same bf16 operands."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def run(mode, A, W, B, T, C0, offs, N, BN, KB, shared_a=0):
    from speech_editing_toolkit_b200 import _lib
    out = torch.full((B * T, N), float("nan"), dtype=torch.float32, device="cuda")
    arr = (C.c_int32 * len(offs))(*offs)
    _lib.check(_lib.lib().fse_debug_conv_gemm(_lib.MODES[mode], C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                              C.c_void_p(out.data_ptr()), B, T, C0, len(offs), arr, N, BN, KB,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream), None, shared_a))
    torch.cuda.synchronize()
    return out.cpu().numpy()


CASES = [
    # B,  T,   C0, offsets,                       N,    BN,  KB
    (1, 128, 64, [0], 32, 32, 64),                                   # smallest: one k-block, one tile
    (1, 128, 256, [0], 256, 256, 64),                                # 4 k-blocks, N=256
    (2, 200, 256, [-1, 0, 1], 512, 256, 64),                         # denoiser dilated conv (2 n-tiles, ragged T)
    (1, 130, 80, [0], 256, 256, 64),                                 # input projection: C0=80 padded to 128 by OOB fill
    (1, 96, 256, [0], 80, 80, 64),                                   # output projection: N = 80 (CH=16 epilogue)
    (1, 300, 32, [-15, -12, -9, -6, -3, 0, 3, 6, 9, 12, 15], 32, 32, 32),   # HiFi-GAN stage 4, k=11 d=3, KB=32
    (2, 65, 512, [0, -1], 2048, 256, 64),                            # transposed conv as 2-tap GEMM, 8 n-tiles
    (1, 257, 64, [-3, -2, -1, 0, 1, 2, 3], 64, 64, 64),              # k=7, three tiles
    (3, 1, 128, [-1, 0, 1], 128, 128, 64),                           # T = 1
    (2, 300, 192, [0], 192, 192, 64),                                # hidden-192 projections (CampNet / condition encoder): one 192-wide tile
    (1, 260, 192, [-4, -3, -2, -1, 0, 1, 2, 3, 4], 384, 192, 64),    # CampNet FFN conv k=9, N = 384 as two 192-wide tiles
    (1, 200, 384, [0], 192, 192, 64),                                # FFN 1x1 back to 192
    (2, 150, 192, [-8, -7, -6, -5, -4, -3, -2, -1, 0], 384, 192, 64),  # CampNet decoder FFN: causal 'LEFT' taps
]


def _operands(B, T, C0, offs, N, KB):
    rs = np.random.RandomState(B * 7 + T + C0 + N)
    nkb = (C0 + KB - 1) // KB
    Kp = len(offs) * nkb * KB
    A = torch.from_numpy(rs.standard_normal((B, T, C0)).astype(np.float32)).cuda().to(torch.bfloat16).contiguous()
    Wfull = np.zeros((N, Kp), dtype=np.float32)
    for j in range(len(offs)):
        Wfull[:, j * nkb * KB:j * nkb * KB + C0] = rs.standard_normal((N, C0)).astype(np.float32) / np.sqrt(C0 * len(offs))
    W = torch.from_numpy(Wfull).cuda().to(torch.bfloat16).contiguous()
    # float64 host reference over the bf16-rounded operands
    Af = A.float().cpu().numpy().astype(np.float64)
    Wf = W.float().cpu().numpy().astype(np.float64)
    ref = np.zeros((B, T, N))
    for j, off in enumerate(offs):
        sh = np.zeros_like(Af)
        lo, hi = max(0, -off), min(T, T - off)
        if hi > lo:
            sh[:, lo:hi] = Af[:, lo + off:hi + off]
        ref += sh @ Wf[:, j * nkb * KB:j * nkb * KB + C0].T
    return A, W, ref.reshape(B * T, N)


@pytest.mark.parametrize("B,T,C0,offs,N,BN,KB", CASES)
def test_conv_gemm_tc_vs_simt_vs_host(lib_built, B, T, C0, offs, N, BN, KB):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    A, W, ref = _operands(B, T, C0, offs, N, KB)
    simt = run("simt_bf16", A, W, B, T, C0, offs, N, BN, KB)
    assert np.abs(simt - ref).max() < 1e-3, "CUDA-core path disagrees with the host reference"
    tc = run("tc_bf16", A, W, B, T, C0, offs, N, BN, KB)
    assert np.isfinite(tc).all(), f"tensor-core path left {np.isnan(tc).sum()} outputs unwritten"
    err = np.abs(tc - ref)
    assert err.max() < 1e-3, f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


SHARED_CASES = [c for c in CASES if len(c[3]) >= 2 and max(c[3]) - min(c[3]) <= 128]


@pytest.mark.parametrize("B,T,C0,offs,N,BN,KB", SHARED_CASES)
def test_conv_gemm_shared_a_schedule(lib_built, B, T, C0, offs, N, BN, KB):
    """One activation load per channel block; every tap is a row-shifted UMMA descriptor of that smem copy
    (start address + rows*row_bytes, base-offset field 0: the swizzle is keyed on absolute smem address bits)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    A, W, ref = _operands(B, T, C0, offs, N, KB)
    tc = run("tc_bf16", A, W, B, T, C0, offs, N, BN, KB, shared_a=1)
    assert np.isfinite(tc).all()
    err = np.abs(tc - ref)
    assert err.max() < 1e-3, f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


def _taps(k, d):
    return [(j - (k - 1) // 2) * d for j in range(k)]


SHARED_MT_CASES = [
    # B,  T,    C0,  offsets,       N,   BN,  KB, MT      (the HiFi-GAN resblock shapes: one n-tile, MT sub-tiles per job)
    (2, 2500, 32, _taps(11, 5), 32, 32, 32, 8),           # stage 4: 1074 job rows arrive as 5 boxes of 216
    (1, 1100, 32, _taps(3, 1), 32, 32, 32, 8),
    (2, 1300, 64, _taps(7, 3), 64, 64, 64, 4),            # stage 3
    (1, 700, 128, _taps(11, 5), 128, 128, 64, 2),         # stage 2: two channel blocks
    (1, 300, 256, _taps(7, 5), 256, 256, 64, 1),          # stage 1: span 30, MT = 1
]


@pytest.mark.parametrize("B,T,C0,offs,N,BN,KB,MT", SHARED_MT_CASES)
def test_conv_gemm_shared_a_multi_tile(lib_built, B, T, C0, offs, N, BN, KB, MT):
    """Shared-A schedule with MT 128-frame sub-tiles per job: the job's rows (128*MT + tap span) are loaded once per
    channel block as several TMA boxes laid end to end, every (tap, sub-tile) is a row-shifted descriptor."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    A, W, ref = _operands(B, T, C0, offs, N, KB)
    tc = run("tc_bf16", A, W, B, T, C0, offs, N, BN, KB, shared_a=1 + 16 * (MT - 1))
    assert np.isfinite(tc).all()
    err = np.abs(tc - ref)
    assert err.max() < 1e-3, f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"

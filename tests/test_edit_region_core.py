"""The per-item region-surgery functions of csrc/edit_region_core.h — the code the CUDA kernels of edit_region.cu loop over —
compiled for the HOST with g++ (tests/tools/edit_region_host.cpp) and checked bit-exactly against the tensors recorded from the
reference's own forward_model code (tests/golden/edit_region.npz) and against the oracle on random utterances."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import golden
from oracle import edit_region_oracle as EO
from speech_editing_toolkit_b200 import synth
from test_edit_region_oracle import case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("er") / "liber_host.so"
    subprocess.run([gxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", os.path.join(ROOT, "tests", "tools", "edit_region_host.cpp"),
                    "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.er_assemble.restype = C.c_longlong
    return lib


def p(a):
    return C.c_void_p(a.ctypes.data)


def run_core(lib, item, edited_mel2ph):
    region = np.array([*item["words_region"][0], *item["edited_words_region"][0]], dtype=np.int64)
    T, Tp, Tpe, Te, M = len(item["mel2ph"]), len(item["ph2word"]), len(item["edited_ph2word"]), len(edited_mel2ph), item["mel"].shape[1]
    stride = Tpe + 3                                             # a padded row: the tail past Tpe must be zeroed
    md = np.full(stride, -7, dtype=np.int64); mm = np.full(T, -7, dtype=np.int64); mo = np.full(T, -7, dtype=np.float32)
    lib.er_prepare(p(item["mel2ph"]), p(item["mel2word"]), T, p(item["ph2word"]), p(item["dur"]), Tp, Tpe, stride, p(region), p(md), p(mm), p(mo))
    emp = np.ascontiguousarray(edited_mel2ph, dtype=np.int64)
    plan = np.zeros(8, dtype=np.int64)
    args = (p(item["mel2ph"]), p(item["mel2word"]), T, p(item["edited_ph2word"]), Tpe, p(region), p(emp), Te, p(item["mel"]), p(item["f0"]), p(item["uv"]), M)
    Tn = int(lib.er_assemble(*args, p(plan), None, None, None, None, None))
    out = dict(mel2ph=np.full(Tn, -7, dtype=np.int64), ref_mels=np.full((Tn, M), np.nan, dtype=np.float32), f0=np.full(Tn, np.nan, dtype=np.float32),
               uv=np.full(Tn, np.nan, dtype=np.float32), time_mel_masks=np.full(Tn, np.nan, dtype=np.float32))
    assert lib.er_assemble(*args, p(plan), p(out["mel2ph"]), p(out["ref_mels"]), p(out["f0"]), p(out["uv"]), p(out["time_mel_masks"])) == Tn
    return md, mm, mo, plan, out


@pytest.mark.parametrize("i", [0, 1, 2])
def test_core_matches_tensors_recorded_from_the_reference_code(host, i):
    item, want = case(golden("edit_region.npz"), i)
    md, mm, mo, plan, out = run_core(host, item, want["edited_mel2ph_pred"])
    Tpe = len(item["edited_ph2word"])
    assert np.array_equal(md[:Tpe], want["masked_dur"]) and (md[Tpe:] == 0).all()
    assert np.array_equal(mm, want["masked_mel2ph"]) and np.array_equal(mo, want["time_mel_masks_orig"].astype(np.float32))
    assert np.array_equal(out["mel2ph"], want["mel2ph"])
    assert np.array_equal(out["time_mel_masks"], want["time_mel_masks"][:, 0])
    for k in ("ref_mels", "f0", "uv"):
        assert np.array_equal(out[k], want[k]), k
    assert plan[0] == len(want["mel2ph"])


def test_core_matches_oracle_on_random_utterances(host):
    rs = np.random.RandomState(5)
    for trial in range(60):
        n_words = int(rs.randint(3, 12))
        w0 = int(rs.randint(1, n_words + 1)); w1 = int(rs.randint(w0, n_words + 1))
        new = tuple(int(x) for x in rs.randint(1, 5, size=int(rs.randint(1, 4))))
        item = synth.synthetic_edit_item(100 + trial, n_words=n_words, n_mels=8, edit_span=(w0, w1), new_span_phones=new)
        # a predicted alignment of the edited text: 1-7 frames per phone, some phones with zero frames
        d = rs.randint(0, 8, size=len(item["edited_ph2word"]))
        c0, c1 = item["edited_words_region"][0]
        if d[(item["edited_ph2word"] >= c0) & (item["edited_ph2word"] <= c1)].sum() == 0:
            d[np.argmax(item["edited_ph2word"] >= c0)] = 2         # the reference needs a non-empty edited span (max() of it)
        emp = np.repeat(np.arange(1, len(d) + 1), d).astype(np.int64)
        md, mm, mo, plan, out = run_core(host, item, emp)
        o_md, o_mm, o_mo = EO.prepare(item["mel2ph"], item["mel2word"], item["ph2word"], item["dur"], len(item["edited_ph2word"]), item["words_region"][0])
        assert np.array_equal(md[:len(o_md)], o_md) and np.array_equal(mm, o_mm) and np.array_equal(mo, o_mo)
        want = EO.assemble(item["mel2ph"], item["mel2word"], item["edited_ph2word"], emp, item["words_region"][0], item["edited_words_region"][0],
                           item["mel"], item["f0"], item["uv"])
        for k in ("mel2ph", "ref_mels", "f0", "uv", "time_mel_masks"):
            assert np.array_equal(out[k], want[k]), (trial, k)
        assert tuple(int(x) for x in plan[:4]) == want["plan"]

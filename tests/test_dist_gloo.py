"""N>1 host logic on CPU: world_size-2 gloo run of the batch sharding + terminal all_gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_editing_toolkit_b200 import dist as fdist


def test_shard_bounds_cover_and_balance():
    for n in (1, 5, 32, 33, 256):
        for w in (1, 2, 3, 4, 8):
            spans = [fdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cond = torch.arange(n * 3 * 2, dtype=torch.float32).reshape(n, 3, 2)
        mask = torch.arange(n, dtype=torch.float32).reshape(n, 1)
        fn = lambda c, m: c * 2 + m[:, :, None]           # stand-in for "sample my shard"
        out = fdist.run_sharded(fn, [cond, mask])
        ok = torch.equal(out, fn(cond, mask))
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 5])
def test_run_sharded_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == (n, 3, 2) for r in res)


def _text_worker(rank, world, port, n, q):
    """The editing batch as the reference collates it (int64 tokens / alignment next to fp32 frames): every tensor is sliced on
    the batch axis with the same bounds and the per-rank mels are gathered once."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from speech_editing_toolkit_b200 import synth
        b = synth.synthetic_edit_batch(3, n, 16, n_mels=4, vocab=20, frames_per_phone=4)
        keys = ("txt_tokens", "mel2ph", "time_mel_masks", "ref_mels", "f0")
        inputs = [torch.from_numpy(b[k]) for k in keys]

        def fake_model(txt, mel2ph, mask, ref, f0):            # stand-in for GaussianDiffusionB200.forward on the rank's shard
            assert txt.dtype == torch.int64 and mel2ph.dtype == torch.int64 and txt.shape[0] == ref.shape[0]
            return ref * (1 - mask[:, :, None]) + (f0 * mel2ph.float())[:, :, None] * mask[:, :, None] + txt.float().sum(1)[:, None, None]

        out = fdist.run_sharded(fake_model, inputs)
        q.put((rank, bool(torch.equal(out, fake_model(*inputs))), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [6, 7])
def test_text_batch_sharding_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_text_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == (n, 16, 4) for r in res)


def _allreduce_worker(rank, world, port, q):
    """train.BucketedAllReduce (the DDP plumbing of the training step, utils/commons/trainer.py:475-479): gradient groups handed over by
    the backward's hook are all-reduced asynchronously, what the hook never saw is reduced in finish(); every .grad ends as the mean."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from speech_editing_toolkit_b200 import train
        torch.manual_seed(0)
        params = {f"p{i}": torch.nn.Parameter(torch.zeros(3 + i, 2)) for i in range(5)}
        local = {k: torch.full_like(p, float(rank + 1)) * (i + 1) for i, (k, p) in enumerate(params.items())}
        for k, p in params.items():
            p.grad = local[k].clone()
        red = train.BucketedAllReduce(params)
        assert red.active()
        red({"p0": params["p0"].grad, "p1": params["p1"].grad})          # two hook calls (two "layers"), p3 / p4 never pass the hook
        red({"p2": params["p2"].grad})
        red.finish()
        mean = (1 + world) / 2.0
        ok = all(torch.allclose(p.grad, torch.full_like(p, mean * (i + 1))) for i, (k, p) in enumerate(params.items()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)

"""CPU, world_size 2 over gloo: batch sharding + final all-gather (the only collective on the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from localdiffusion_hallucination_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for total in (1, 7, 16, 33, 128):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        cond = torch.rand(total, 1, 8, 8, generator=g)
        mask = (torch.rand(total, 1, 8, 8, generator=g) > 0.5).float()
        tape = torch.randn(4, total, 1, 8, 8, generator=g)
        # a stand-in "sampler": per-sample, deterministic, uses every input (the real one needs a GPU)
        fn = lambda c, m, z: c * 2 + m + z.sum(dim=0)
        got = parallel.sample_sharded(fn, cond, mask, tape)
        want = fn(cond, mask, tape)
        ok = bool(torch.equal(got, want))
        # noise=None: the GLOBAL seed-10 tape is drawn and sliced per rank (every rank seeding its own local tape would hand all
        # shards the same rows): equals the single-process run on the same global tape
        got2 = parallel.sample_sharded(fn, cond, mask, None, steps=4)
        want2 = fn(cond, mask, parallel.global_noise_tape(tuple(cond.shape), 4, cond.device))
        ok = ok and bool(torch.equal(got2, want2))
        # stacked pair output [2, B, ...] of a never-fused run (ddpm.py:965-970) is gathered along dim 1
        fnp = lambda c, m, z: torch.stack((c + z[0], m - z[1]))
        got3 = parallel.sample_sharded(fnp, cond, mask, tape, pair=True)
        ok = ok and bool(torch.equal(got3, fnp(cond, mask, tape))) and got3.shape[1] == total
        q.put((rank, ok, tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_sharded_sampling_is_world_size_independent(total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res) and all(shape[0] == total for *_, shape in res)

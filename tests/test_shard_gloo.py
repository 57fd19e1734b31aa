"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: the sharding / overlap all-gather
logic of futuresdr_b200.shard.ShardedFir with the oracle as the per-rank kernel must reproduce
the single-stream reference result exactly (same sums, same order -> bit-exact)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ntaps, decim, S, steps, q):
    sys.path.insert(0, ROOT)
    import oracle as orc
    from futuresdr_b200.shard import ShardedFir
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    total = world * S * steps
    x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)).astype(np.complex64)
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)

    def compute(src, out):
        c, p, st, o = orc.decim_fir(taps, decim, src.numpy(), out.numel())
        out[:p] = torch.from_numpy(o)
        return c, p, st

    sh = ShardedFir(taps, S, np.complex64, decim=decim, device=torch.device("cpu"), compute=compute)
    outs = []
    for t in range(steps):
        lo = (t * world + rank) * S
        sh.chunk.copy_(torch.from_numpy(x[lo:lo + S]))
        out = torch.zeros(S // decim, dtype=torch.complex64)
        c, p, st = sh.step(out)
        outs.append((t * world + rank, out[:p].numpy().copy()))
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)
    if rank == 0:
        pieces = sorted([pc for g in gathered for pc in g], key=lambda a: a[0])
        got = np.concatenate([p for _, p in pieces])
        _, _, _, ref = orc.decim_fir(taps, decim, x, total)
        q.put((got.size == ref.size, bool(np.array_equal(got, ref))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ntaps,decim,S", [(2, 256, 1, 4096), (2, 52, 4, 4096), (3, 33, 3, 3000),
                                                  (2, 1, 1, 64), (4, 64, 2, 2048)])
def test_sharded_fir_equals_single_stream(world, ntaps, decim, S):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + ntaps * 7 + decim) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, ntaps, decim, S, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same_len, same = q.get(timeout=5)
    assert same_len and same


@pytest.mark.parametrize("ntaps,decim,S", [(256, 1, 4096), (52, 4, 4096), (1, 1, 64)])
def test_single_rank_stream_continuity(ntaps, decim, S):
    """world == 1 (no process group): the rank carries the tail of its own previous chunk as history with one
    copy per step; the concatenated outputs equal the single-stream reference bit for bit."""
    import oracle as orc
    from futuresdr_b200.shard import ShardedFir
    rng = np.random.default_rng(5)
    steps = 4
    x = (rng.standard_normal(S * steps) + 1j * rng.standard_normal(S * steps)).astype(np.complex64)
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)

    def compute(src, out):
        c, p, st, o = orc.decim_fir(taps, decim, src.numpy(), out.numel())
        out[:p] = torch.from_numpy(o)
        return c, p, st

    sh = ShardedFir(taps, S, np.complex64, decim=decim, device=torch.device("cpu"), compute=compute)
    assert sh.world == 1
    got = []
    for t in range(steps):
        sh.chunk.copy_(torch.from_numpy(x[t * S:(t + 1) * S]))
        out = torch.zeros(S // decim, dtype=torch.complex64)
        c, p, st = sh.step(out)
        got.append(out[:p].numpy().copy())
    got = np.concatenate(got)
    _, _, _, ref = orc.decim_fir(taps, decim, x, x.size)
    assert got.size == ref.size and np.array_equal(got, ref)

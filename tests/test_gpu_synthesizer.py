"""SURVEY §8f row 2 (dual of the channelizer): PfbSynthesizer on the device vs the oracle restatement of
src/blocks/pfb/synthesizer.rs:80-144, driven with the same sequence of work() calls."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,ntaps", [(4, 8), (8, 8 * 12), (16, 16 * 8 + 3), (6, 60), (64, 64 * 16)])
def test_synthesizer_parity(rng, N, ntaps):
    import torch
    from futuresdr_b200.blocks import PfbSynthesizer, WorkIo
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    nv = 5000
    x = (rng.standard_normal((N, nv)) + 1j * rng.standard_normal((N, nv))).astype(np.complex64)
    for chunks in ([1 << 30], [1, 2, 3, 500, 77, 1 << 30]):
        ref = orc.PfbSynthesizer(N, taps)
        blk = PfbSynthesizer(N, taps)
        pos, ci, want, got = 0, 0, [], []
        for _ in range(1000):
            step = min(chunks[min(ci, len(chunks) - 1)], nv - pos)
            ci += 1
            cap = step * N + 3 * N + 5
            c0, p0, o0 = ref.work(x[:, pos:pos + step], cap)
            blk.set_inputs(x[:, pos:pos + step])
            blk.output.reserve(cap)
            blk.work(WorkIo())
            torch.cuda.synchronize()
            assert (blk.in_pos, blk.output.len) == (c0, p0)
            want.append(o0); got.append(blk.output.get().cpu().numpy().copy())
            pos += c0
            if pos >= nv:
                break
        want, got = np.concatenate(want), np.concatenate(got)
        T = int(np.ceil(ntaps / N))
        assert got.size == want.size == (nv - T + 1) * N
        scale = float(np.max(np.abs(taps))) * T * N * float(np.max(np.abs(x)))
        assert np.max(np.abs(got - want)) <= 1e-5 * scale


def test_synthesizer_capacity_rule(rng):
    # `out.len() - produced > num_channels` (synthesizer.rs:96): strictly more than N free items are needed
    import torch
    from futuresdr_b200.blocks import PfbSynthesizer, WorkIo
    N, taps = 8, rng.uniform(-1, 1, 32).astype(np.float32)
    x = (rng.standard_normal((N, 100)) + 1j * rng.standard_normal((N, 100))).astype(np.complex64)
    for cap in (N, N + 1, 2 * N, 2 * N + 1, 5 * N + 3):
        ref, blk = orc.PfbSynthesizer(N, taps), PfbSynthesizer(N, taps)
        c0, p0, _ = ref.work(x, cap)
        blk.set_inputs(x)
        blk.output.reserve(cap)
        blk.work(WorkIo())
        torch.cuda.synchronize()
        assert (blk.in_pos, blk.output.len) == (c0, p0), cap

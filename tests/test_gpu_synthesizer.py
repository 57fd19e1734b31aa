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


@pytest.mark.parametrize("N,T", [(64, 16), (128, 8), (256, 8), (256, 32), (32, 20), (16, 32), (4, 3), (8, 5)])
def test_synthesizer_fused_shapes(N, T, monkeypatch):
    """Shapes of the fused steady-state kernel (gather + IFFT + FIR bank in one launch, spun vectors in a shared-memory
    ring across tiles) against the oracle, in several calls so that tiles, the ring hand-over between CTAs and the
    history left for the next call all get exercised; and against the three-kernel path (B2S_SYNTH_NO_FUSED=1)."""
    import torch
    from futuresdr_b200.blocks import PfbSynthesizer, WorkIo

    def drive():
        rng = np.random.default_rng(77)
        taps = rng.uniform(-1, 1, N * T - 1).astype(np.float32)
        ob = 256 // max(1, N // 16)
        nv = 7 * ob + 37
        x = (rng.standard_normal((N, nv)) + 1j * rng.standard_normal((N, nv))).astype(np.complex64)
        ref, blk = orc.PfbSynthesizer(N, taps), PfbSynthesizer(N, taps)
        pos, want, got = 0, [], []
        for step in (T + 3, 3 * ob + 5, 2 * ob, ob + 1, 1 << 30):
            step = min(step, nv - pos)
            if step <= 0:
                break
            cap = step * N + 3 * N + 5
            c0, p0, o0 = ref.work(x[:, pos:pos + step], cap)
            blk.set_inputs(x[:, pos:pos + step])
            blk.output.reserve(cap)
            blk.work(WorkIo())
            torch.cuda.synchronize()
            assert (blk.in_pos, blk.output.len) == (c0, p0)
            want.append(o0); got.append(blk.output.get().cpu().numpy().copy())
            pos += c0
        want, got = np.concatenate(want), np.concatenate(got)
        scale = float(np.max(np.abs(taps))) * T * N * float(np.max(np.abs(x)))
        assert got.size == want.size > 0 and np.max(np.abs(got - want)) <= 1e-5 * scale
        return got

    a = drive()
    monkeypatch.setenv("B2S_SYNTH_NO_FUSED", "1")
    b = drive()
    assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1e-4 * np.max(np.abs(b))

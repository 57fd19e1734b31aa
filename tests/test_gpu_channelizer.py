"""PfbChannelizer on the device vs the oracle restatement of
src/blocks/pfb/channelizer.rs (parity unpinned in the reference).  The call pattern matters for the
reference (the call that completes the window fill does not consume), so oracle and device are driven
with the SAME sequence of work() calls."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


def _noise(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def _drive(rng, N, ntaps, oversample, n, chunks):
    import torch
    from futuresdr_b200.blocks import PfbChannelizer, WorkIo
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    x = _noise(rng, n)
    ref = orc.PfbChannelizer(N, taps, oversample)
    blk = PfbChannelizer(N, taps, oversample)
    blk.reserve_outputs(n // blk.decimation_factor + 8)
    xd = torch.from_numpy(x).cuda()
    ref_out, pos, ci = [], 0, 0
    for _ in range(10_000):
        step = chunks[min(ci, len(chunks) - 1)]
        ci += 1
        seg = x[pos:pos + step]
        c0, p0, ca0, o0 = ref.work(seg, 1 << 20)
        blk.input.set(xd[pos:pos + step])
        io = WorkIo()
        before = blk.produced
        blk.work(io)
        assert (blk.input.pos, blk.produced - before, io.call_again) == (c0, p0, ca0)
        ref_out.append(o0)
        pos += c0
        if pos >= n - blk.decimation_factor + 1 and not ca0 and p0 == 0 and c0 == 0:
            break
        if c0 == 0 and p0 == 0 and not ca0 and step >= n:
            break
    torch.cuda.synchronize()
    want = np.concatenate(ref_out, axis=1)
    got = blk.outputs[:, :blk.produced].cpu().numpy()
    assert got.shape == want.shape and want.shape[1] > 0
    T = int(np.ceil(ntaps / N))
    scale = float(np.max(np.abs(taps))) * T * N * float(np.max(np.abs(x)))
    assert np.max(np.abs(got - want)) <= 1e-5 * scale
    return got


@pytest.mark.parametrize("N,ntaps,oversample", [(4, 8, 1.0), (8, 8 * 12, 1.0), (16, 16 * 8 + 5, 2.0), (6, 60, 1.0),
                                                (64, 64 * 16, 1.0), (12, 12 * 7, 3.0),
                                                (16, 16 * 4, 1.0), (8, 8 * 8, 1.0), (32, 32 * 16 - 3, 1.0)])
def test_channelizer_parity(rng, N, ntaps, oversample):
    n = 40_000
    _drive(rng, N, ntaps, oversample, n, [1 << 30])              # everything offered at once
    _drive(rng, N, ntaps, oversample, n, [5, 3, 1000, 77, 9999, 1 << 30])   # ragged calls


@pytest.mark.parametrize("N,T", [(128, 16), (256, 8), (64, 20), (32, 3), (16, 32), (4, 5), (256, 17)])
def test_channelizer_fused_shapes(rng, N, T, monkeypatch):
    """Shapes of the fused steady-state kernel (FIR bank + IFFT + transposed store in one launch: N a power of two
    <= 256, critically sampled, T <= 32 padded to 8/16/32 taps) against the oracle, and against the generic
    three-kernel path (B2S_CHAN_NO_FUSED=1) which must agree to rounding."""
    n = N * 3000 + 17
    a = _drive(np.random.default_rng(123), N, N * T - 1, 1.0, n, [N * 40 + 3, 1 << 30])
    monkeypatch.setenv("B2S_CHAN_NO_FUSED", "1")
    b = _drive(np.random.default_rng(123), N, N * T - 1, 1.0, n, [N * 40 + 3, 1 << 30])
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= 1e-4 * np.max(np.abs(b))


def test_channelizer_tone_lands_in_its_channel(rng):
    import torch
    from futuresdr_b200.blocks import PfbChannelizer, WorkIo
    N = 8
    taps = orc.kaiser_lowpass(0.4 / N, 0.1 / N, 1e-3).astype(np.float32)
    n = N * 2000
    x = np.exp(2j * np.pi * (3 / N) * np.arange(n)).astype(np.complex64)
    blk = PfbChannelizer(N, taps, 1.0)
    blk.reserve_outputs(n // N + 8)
    blk.input.set(torch.from_numpy(x).cuda())
    for _ in range(4):
        io = WorkIo()
        blk.work(io)
        if not io.call_again:
            break
    torch.cuda.synchronize()
    p = np.abs(blk.outputs[:, blk.produced - 1].cpu().numpy())
    assert np.argmax(p) == 3 and p[3] > 0.9 and np.all(np.delete(p, 3) < 1e-2)

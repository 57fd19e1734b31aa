"""the spectrum pipe tail -- Fft(2048, Forward, shift) -> |x|^2 -> MovingAvg<2048>(0.1, 3)
(examples/spectrum/src/bin/cpu.rs:21-28) on the device vs the oracle, plus MovingAvg alone."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


def test_moving_avg_bit_exact_and_chunked(rng):
    import torch
    from futuresdr_b200.blocks import MovingAvg, WorkIo
    W, nch = 2048, 41
    x = rng.random(W * nch + 100).astype(np.float32) * 10
    x[5] = np.inf; x[W + 7] = np.nan                     # non-finite items only decay the average (:88-90)
    ref = orc.MovingAvg(W, 0.1, 3)
    blk = MovingAvg(W, 0.1, 3)
    xd = torch.from_numpy(x).cuda()
    pos, got, want = 0, [], []
    for step, cap in ((W * 5 + 17, W * 10), (W - 1, W), (W * 20, W * 2), (10 ** 9, W * 100)):
        seg = x[pos:pos + step]
        c0, p0, o0 = ref.work(seg, cap)
        blk.input.set(xd[pos:pos + step])
        blk.output.reserve(cap)
        blk.work(WorkIo())
        torch.cuda.synchronize()
        assert (blk.input.pos, blk.output.len) == (c0, p0)
        got.append(blk.output.get().cpu().numpy().copy()); want.append(o0)
        pos += c0
    got, want = np.concatenate(got), np.concatenate(want)
    assert got.size == want.size > 0 and np.array_equal(got, want)     # same IEEE ops in the same order


def test_spectrum_pipe(rng):
    import torch
    from futuresdr_b200.blocks import Apply, ApplyOp, Fft, FftDirection, MovingAvg, WorkIo
    N, frames = 2048, 64
    x = (rng.standard_normal(N * frames) + 1j * rng.standard_normal(N * frames)).astype(np.complex64)
    x += (3 * np.exp(2j * np.pi * 0.2 * np.arange(x.size))).astype(np.complex64)
    # oracle pipe
    _, X = orc.fft_block(x, N, fft_shift=True)
    P = orc.norm_sqr(X)
    _, _, want = orc.MovingAvg(N, 0.1, 3).work(P, P.size)
    # device pipe: everything stays in HBM between blocks
    fft = Fft.with_options(N, FftDirection.Forward, True, None)
    mag = Apply(ApplyOp.NormSqr)
    keep = MovingAvg(N, 0.1, 3)
    xd = torch.from_numpy(x).cuda()
    Xd = torch.empty_like(xd)
    Pd = torch.empty(x.size, dtype=torch.float32, device="cuda")
    assert fft.transform(xd, Xd) == x.size
    assert mag.apply(Xd, Pd) == x.size
    keep.input.set(Pd)
    keep.output.reserve(P.size)
    keep.work(WorkIo())
    torch.cuda.synchronize()
    got = keep.output.get().cpu().numpy()
    assert got.size == want.size == (frames // 3) * N
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(want)
    peak = np.argmax(got[-N:])
    assert abs(peak - (N // 2 + int(0.2 * N))) <= 1               # tone at +0.2 fs after fftshift


@pytest.mark.parametrize("N,decay,hist", [(2048, 0.1, 3), (4096, 0.1, 3), (256, 0.5, 1), (1024, 0.01, 7), (64, 1.0, 2), (8192, 0.25, 4)])
def test_fused_spectrum_pipe_vs_oracle_chain(rng, N, decay, hist):
    """SpectrumPipe (FFT + |x|^2 + MovingAvg in one pass, blocked scan) against the oracle chain evaluated in one
    piece, fed in ragged calls with small output capacities so that group boundaries, the carried state and the
    emission counter all get exercised.  Counts exact; values within 1e-5 of the largest average."""
    import torch
    from futuresdr_b200.blocks import SpectrumPipe
    frames = 397 if N <= 2048 else 131
    n = N * frames
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    x += (3 * np.exp(2j * np.pi * 0.2 * np.arange(n))).astype(np.complex64)
    x[N * 5 + 3] = complex(np.inf, 0.0)                  # |X|^2 of that frame is non-finite: only decays the average
    _, X = orc.fft_block(x, N, fft_shift=True)
    P = orc.norm_sqr(X)
    ref = orc.MovingAvg(N, decay, hist)
    blk = SpectrumPipe(N, decay, hist)
    xd = torch.from_numpy(x).cuda()
    pos, got, want = 0, [], []
    plan = [(N * 7 + 5, N * 1), (N * 50, N * 3), (N - 1, N * 4), (N * 3, 0), (N * 100, N * 1000), (10 ** 9, N * 1000)]
    for step, cap in plan:
        seg_items = min(step, n - pos)
        c0, p0, o0 = ref.work(P[pos:pos + seg_items], cap)
        od = torch.zeros(max(cap, 1), dtype=torch.float32, device="cuda")[:cap]
        c, p = blk.process(xd[pos:pos + seg_items], od)
        torch.cuda.synchronize()
        assert (c, p) == (c0, p0), (step, cap, c, p, c0, p0)
        got.append(od[:p].cpu().numpy().copy()); want.append(o0)
        pos += c
    got, want = np.concatenate(got), np.concatenate(want)
    assert got.size == want.size > 0
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - want[fin])) <= 1e-5 * np.max(want[fin])


def test_fused_spectrum_log10(rng):
    import torch
    from futuresdr_b200.blocks import SpectrumPipe
    N, frames = 2048, 90
    x = (rng.standard_normal(N * frames) + 1j * rng.standard_normal(N * frames)).astype(np.complex64)
    _, X = orc.fft_block(x, N, fft_shift=True)
    _, _, lin = orc.MovingAvg(N, 0.1, 3).work(orc.norm_sqr(X), N * frames)
    blk = SpectrumPipe(N, 0.1, 3, log10_scale=10.0)
    od = torch.zeros(lin.size, dtype=torch.float32, device="cuda")
    c, p = blk.process(torch.from_numpy(x).cuda(), od)
    torch.cuda.synchronize()
    assert p == lin.size
    assert np.max(np.abs(od.cpu().numpy() - 10.0 * np.log10(lin))) <= 1e-3      # dB

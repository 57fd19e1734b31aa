"""GPU parity tests for the remaining §8 rows through the C ABI: polyphase resampler, Fft
block, Apply catalogue, PfbArbResampler, the device buffer ring, and the block-level harness
(Mocker semantics).  Oracle = oracle/ (CPU restatement of the cited reference lines)."""
import ctypes as C

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import futuresdr_b200 as fb
    import futuresdr_b200.blocks  # noqa: F401
    return fb


def _noise(rng, n, cplx=True):
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ---------------------------------------------------------------------------------------------
# PolyphaseResamplingFir (polyphase_resampling_fir.rs)
# ---------------------------------------------------------------------------------------------
def _run_filter(f, x, cap):
    import torch
    xd = _dev(x)
    out = torch.zeros(max(cap, 1), dtype=xd.dtype, device="cuda")[:cap]
    c, p, st = f.filter(xd, out)
    torch.cuda.synchronize()
    return c, p, int(st), out[:p].cpu().numpy()


def test_resampler_known_answers(fb):
    # polyphase_resampling_fir.rs:174-260
    taps = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    f = fb.PolyphaseResamplingFir(3, 2, taps, sample_dtype=np.float32)
    assert f.length() == 6
    x5, x8 = np.array([1, 2, 3, 4, 5], np.float32), np.arange(1, 9, dtype=np.float32)
    assert _run_filter(f, x5, 8)[:3] == (2, 3, 0) and list(_run_filter(f, x5, 8)[3]) == [6.0, 12.0, 16.0]
    assert _run_filter(f, x5, 0)[:3] == (0, 0, 1)
    assert _run_filter(f, x5, 3)[:3] == (2, 3, 2)
    c, p, st, o = _run_filter(f, x8, 3)
    assert (c, p, st) == (2, 3, 1) and list(o) == [6.0, 12.0, 16.0]
    c, p, st, o = _run_filter(f, x8[2:], 3)
    assert (c, p, st) == (2, 3, 1) and list(o) == [16.0, 30.0, 30.0]
    c, p, st, o = _run_filter(f, x8[4:], 3)
    assert (c, p, st) == (2, 3, 2) and list(o) == [26.0, 48.0, 44.0]
    f = fb.PolyphaseResamplingFir(2, 1, [1.0, 2.0], sample_dtype=np.float32)
    c, p, st, o = _run_filter(f, np.array([1, 2, 3, 4], np.float32), 10)
    assert (c, p, st) == (3, 6, 0) and list(o) == [1.0, 2.0, 2.0, 4.0, 3.0, 6.0]
    f = fb.PolyphaseResamplingFir(1, 3, [1.0, 2.0], sample_dtype=np.float32)
    c, p, st, o = _run_filter(f, x8, 8)
    assert (c, p, st) == (6, 2, 0) and list(o) == [4.0, 13.0]


@pytest.mark.parametrize("L,M", [(3, 2), (2, 3), (1, 4), (5, 1), (48, 125), (160, 147), (7, 64), (4, 3), (4, 5), (2, 1), (3, 4)])
@pytest.mark.parametrize("cplx", [True, False])
def test_resampler_parity(fb, rng, L, M, cplx):
    taps = orc.kaiser_multirate(L, M, 12, 1e-4)        # FirBuilder::resampling default design
    n = 20000
    x = _noise(rng, n, cplx)
    f = fb.PolyphaseResamplingFir(L, M, taps, sample_dtype=x.dtype)
    for cap in (n * L // M + 64, 1000):
        c0, p0, s0, ref = orc.resamp_fir(taps, L, M, x, cap)
        c, p, st, o = _run_filter(f, x, cap)
        assert (c, p, st) == (c0, p0, s0)
        if p:
            # per-output tap set is one polyphase arm: scale by the largest arm's L1 norm
            arm_l1 = max(np.sum(np.abs(taps[b::L])) for b in range(L))
            assert np.max(np.abs(o - ref)) <= 1e-5 * arm_l1 * np.max(np.abs(x))


# ---------------------------------------------------------------------------------------------
# Fft block (src/blocks/fft.rs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fft_forward_all_sizes(fb, rng, n):
    import torch
    from futuresdr_b200.blocks import Fft
    nfft = 37 if n <= 4096 else 5
    x = _noise(rng, n * nfft + min(3, n - 1))            # trailing items (< n) must stay unconsumed
    fft = Fft(n)
    xd = _dev(x)
    out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
    m = fft.transform(xd, out)
    torch.cuda.synchronize()
    m0, ref = orc.fft_block(x, n)
    assert m == m0 == n * nfft
    got = out[:m].cpu().numpy().reshape(nfft, n)
    ref = ref.reshape(nfft, n)
    # per-transform: ||X - X_ref||_inf <= 1e-5 * max|X|   (SURVEY §8c FFT tolerance)
    assert np.all(np.max(np.abs(got - ref), axis=1) <= 1e-5 * np.max(np.abs(ref), axis=1))
    # cross-check the oracle itself against numpy's pocketfft
    assert np.allclose(ref, np.fft.fft(x[:m].reshape(nfft, n).astype(np.complex128), axis=1), rtol=0, atol=1e-3 * np.sqrt(n))


@pytest.mark.parametrize("inverse,shift,norm", [(False, True, None), (True, False, None), (True, True, None),
                                                (False, False, 1.0 / 4096), (True, True, 0.25)])
def test_fft_options(fb, rng, inverse, shift, norm):
    import torch
    from futuresdr_b200.blocks import Fft, FftDirection
    n, nfft = 4096, 9
    x = _noise(rng, n * nfft)
    fft = Fft.with_options(n, FftDirection.Inverse if inverse else FftDirection.Forward, shift, norm)
    out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
    m = fft.transform(_dev(x), out)
    torch.cuda.synchronize()
    m0, ref = orc.fft_block(x, n, inverse=inverse, fft_shift=shift, normalize=norm)
    assert m == m0
    got, ref = out.cpu().numpy().reshape(nfft, n), ref.reshape(nfft, n)
    assert np.all(np.max(np.abs(got - ref), axis=1) <= 1e-5 * np.max(np.abs(ref), axis=1))


def test_fft_roundtrip_and_parseval_full_size(fb):
    """Size-independent properties on a 64 Mi-sample chunk (16384 transforms of 4096)."""
    import torch
    from futuresdr_b200.blocks import Fft, FftDirection
    n, total = 4096, 64 * 1024 * 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.view_as_complex(torch.randn(total, 2, generator=g, device="cuda"))
    X = torch.empty_like(x)
    y = torch.empty_like(x)
    assert Fft(n).transform(x, X) == total
    assert Fft.with_options(n, FftDirection.Inverse, False, 1.0 / n).transform(X, y) == total
    assert float((y - x).abs().max()) <= 1e-5 * float(x.abs().max()) * 4
    ex = (x.abs() ** 2).view(-1, n).sum(1)
    eX = (X.abs() ** 2).view(-1, n).sum(1) / n
    assert float(((ex - eX).abs() / ex).max()) <= 1e-5
    # against torch.fft (cuFFT) on a slice, as an independent implementation
    ref = torch.fft.fft(x[: 64 * n].view(64, n), dim=1)
    assert float((X[: 64 * n].view(64, n) - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_fft_unsupported_sizes_fail_loudly(fb):
    from futuresdr_b200.blocks import Fft
    with pytest.raises(fb.B200SdrError):
        Fft((1 << 24) + 1)              # non-power-of-two beyond the four-step Bluestein scratch limit
    with pytest.raises(fb.B200SdrError):
        Fft(1 << 27)
    with pytest.raises(fb.B200SdrError):
        Fft(1)


def test_fft_size_handler_replans(fb, rng):
    """The `fft_size` message handler (fft.rs:124-151): a new length takes effect for the next work() call, keeping
    direction / shift / normalisation; Null answers the length; other values are refused."""
    import torch
    from futuresdr_b200.blocks import Fft, FftDirection
    fft = Fft.with_options(1024, FftDirection.Forward, True, 0.5)
    assert fft.fft_size() == 1024
    assert fft.fft_size("x") == "InvalidValue" and fft.fft_size() == 1024
    assert fft.fft_size(3000) == "Ok" and fft.fft_size() == 3000
    x = _noise(rng, 3000 * 4)
    out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
    assert fft.transform(_dev(x), out) == x.size
    torch.cuda.synchronize()
    m0, ref = _fft_block_numpy(x, 3000, False, True, 0.5)
    got, r = out.cpu().numpy().reshape(4, 3000), ref.reshape(4, 3000)
    assert np.all(np.max(np.abs(got - r), axis=1) <= 1e-5 * np.max(np.abs(r), axis=1))
    with pytest.raises(fb.B200SdrError):
        fft.set_fft_size(1)
    assert fft.fft_size() == 3000                     # a refused length leaves the block as it was


def _fft_block_numpy(x, n, inverse, shift, norm):
    """Fft::work (fft.rs:160-221) with numpy's pocketfft in double precision as the transform: the reference for
    lengths the oracle's direct O(n^2) DFT cannot reach in test time."""
    m = (x.size // n) * n
    fr = x[:m].astype(np.complex128).reshape(-1, n)
    if inverse and shift:
        fr = np.roll(fr, -(n // 2), axis=1)                    # buff[k] = i[(k + n/2) % n]   (:179-185)
    X = np.fft.ifft(fr, axis=1) * n if inverse else np.fft.fft(fr, axis=1)
    if not inverse and shift:
        X = np.roll(X, -(n // 2), axis=1)                      # o[k] = X[(k + n/2) % n]      (:196-204)
    if norm is not None:
        X = X * np.float32(norm)
    return m, X.reshape(-1)


@pytest.mark.parametrize("n", [32768, 65536, 1 << 20, 8193, 10007, 12000, 100003])
def test_fft_large_sizes_four_step(fb, rng, n):
    """rustfft plans ANY length (fft.rs:98-103).  Beyond one shared-memory transform (16384 points, 8192 for other
    lengths) the four-step algorithm runs through HBM on top of two shared-memory plans, with Bluestein on top of
    that for lengths that are not powers of two (10007 and 100003 are prime)."""
    import torch
    from futuresdr_b200.blocks import Fft, FftDirection
    nfft = 2
    x = _noise(rng, n * nfft + 3)
    for inverse, shift, norm in ((False, False, None), (False, True, None), (True, True, 1.0 / n), (True, False, None)):
        fft = Fft.with_options(n, FftDirection.Inverse if inverse else FftDirection.Forward, shift, norm)
        out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
        m = fft.transform(_dev(x), out)
        torch.cuda.synchronize()
        m0, ref = _fft_block_numpy(x, n, inverse, shift, norm)
        assert m == m0 == n * nfft
        got, r = out[:m].cpu().numpy().reshape(nfft, n), ref.reshape(nfft, n)
        assert np.all(np.max(np.abs(got - r), axis=1) <= 1e-5 * np.max(np.abs(r), axis=1)), (n, inverse, shift)
    if n in (32768, 8193):                                     # and the oracle itself on the sizes it can afford
        m0, ref = orc.fft_block(x, n, fft_shift=True)
        fft = Fft.with_options(n, FftDirection.Forward, True, None)
        out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
        fft.transform(_dev(x), out)
        torch.cuda.synchronize()
        got, r = out[:m0].cpu().numpy().reshape(nfft, n), ref.reshape(nfft, n)
        assert np.all(np.max(np.abs(got - r), axis=1) <= 1e-5 * np.max(np.abs(r), axis=1))


@pytest.mark.parametrize("n", [3, 5, 6, 7, 12, 100, 1000, 1536, 4095, 5000, 8191])
def test_fft_arbitrary_sizes_bluestein(fb, rng, n):
    """rustfft plans any length; non-powers-of-two run the fused Bluestein kernel."""
    import torch
    from futuresdr_b200.blocks import Fft, FftDirection
    nfft = 11 if n <= 1536 else 3
    x = _noise(rng, n * nfft + min(2, n - 1))
    for inverse, shift, norm in ((False, False, None), (False, True, None), (True, True, 1.0 / n), (True, False, None)):
        fft = Fft.with_options(n, FftDirection.Inverse if inverse else FftDirection.Forward, shift, norm)
        out = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
        m = fft.transform(_dev(x), out)
        torch.cuda.synchronize()
        m0, ref = orc.fft_block(x, n, inverse=inverse, fft_shift=shift, normalize=norm)
        assert m == m0 == n * nfft
        got, r = out[:m].cpu().numpy().reshape(nfft, n), ref.reshape(nfft, n)
        assert np.all(np.max(np.abs(got - r), axis=1) <= 1e-5 * np.max(np.abs(r), axis=1)), (n, inverse, shift)


# ---------------------------------------------------------------------------------------------
# Apply catalogue (src/blocks/apply.rs + the closures of the reference graphs)
# ---------------------------------------------------------------------------------------------
def test_apply_scale_like_vulkan_test(fb, rng):
    # tests/vulkan.rs:56-76: 10 000 random f32 through `x * 12`, |orig*12 - v| < f32::EPSILON, length kept
    import torch
    from futuresdr_b200.blocks import Apply, ApplyOp, Mocker
    orig = rng.random(10_000).astype(np.float32)
    blk = Apply(ApplyOp.ScaleF32, 12.0)
    m = Mocker(blk)
    m.input(orig)
    m.init_output(10_000)
    io = m.run_until_finished()
    v = m.output().cpu().numpy()
    assert io.finished and v.size == orig.size
    assert np.all(np.abs(orig * np.float32(12.0) - v) < np.finfo(np.float32).eps)
    assert np.array_equal(v, orc.scale_f32(orig, 12.0))


def test_apply_quad_demod_stateful_chunks(fb, rng):
    import torch
    from futuresdr_b200.blocks import Apply, ApplyOp
    x = _noise(rng, 100_000)
    ref, _ = orc.quad_demod(x)
    blk = Apply(ApplyOp.QuadDemod)
    xd = _dev(x)
    out = torch.zeros(x.size, dtype=torch.float32, device="cuda")
    pos = 0
    for step in (1, 7, 4096, 50_000, 10 ** 9):        # carry crosses call boundaries
        n = min(step, x.size - pos)
        assert blk.apply(xd[pos:pos + n], out[pos:pos + n]) == n
        pos += n
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    # atan2f differs by a few ulp between libm and CUDA; phases live in [-pi, pi]
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.pi
    blk.reset()
    blk.apply(xd[:10], out[:10])
    torch.cuda.synchronize()
    assert np.max(np.abs(out[:10].cpu().numpy() - ref[:10])) <= 1e-5 * np.pi
    # packed variant used in front of PfbArbResampler
    blk2 = Apply(ApplyOp.QuadDemodC32)
    outc = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
    blk2.apply(xd, outc)
    torch.cuda.synchronize()
    oc = outc.cpu().numpy()
    assert np.max(np.abs(oc.real - ref)) <= 1e-5 * np.pi and np.all(oc.imag == 0)


def test_apply_misc_ops(fb, rng):
    import torch
    from futuresdr_b200.blocks import Apply, ApplyOp
    x = _noise(rng, 33_333)
    xd = _dev(x)
    o = torch.zeros(x.size, dtype=torch.float32, device="cuda")
    Apply(ApplyOp.NormSqr).apply(xd, o)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), orc.norm_sqr(x))
    Apply(ApplyOp.MagC32).apply(xd, o)
    torch.cuda.synchronize()
    assert np.allclose(o.cpu().numpy(), np.abs(x), rtol=1e-6)
    xr = rng.random(1000).astype(np.float32) + 0.1
    orr = torch.zeros(1000, device="cuda")
    Apply(ApplyOp.ExpF32).apply(_dev(xr), orr)
    torch.cuda.synchronize()
    assert np.allclose(orr.cpu().numpy(), np.exp(xr), rtol=1e-6)
    Apply(ApplyOp.Log10F32, 10.0).apply(_dev(xr), orr)
    torch.cuda.synchronize()
    assert np.allclose(orr.cpu().numpy(), 10 * np.log10(xr), rtol=1e-5, atol=1e-6)
    oc = torch.zeros(x.size, dtype=torch.complex64, device="cuda")
    Apply(ApplyOp.ScaleC32, 0.5).apply(xd, oc)
    torch.cuda.synchronize()
    assert np.array_equal(oc.cpu().numpy(), x * np.float32(0.5))
    # m = min(len(i), len(o))  (apply.rs:109)
    c = Apply(ApplyOp.ScaleF32, 2.0).apply(_dev(xr), orr[:10])
    assert c == 10


# ---------------------------------------------------------------------------------------------
# PfbArbResampler (src/blocks/pfb/arb_resampler.rs)
# ---------------------------------------------------------------------------------------------
def _pfb_run(fb, rate, taps, nfilt, x, chunks, out_cap):
    import torch
    from futuresdr_b200.blocks import PfbArbResampler, WorkIo
    blk = PfbArbResampler(rate, taps, nfilt)
    xd = _dev(x)
    outs, pos, ci = [], 0, 0
    blk.output.reserve(out_cap)
    for _guard in range(10_000):
        step = chunks[min(ci, len(chunks) - 1)]
        ci += 1
        blk.input.set(xd[pos:pos + step])
        blk.output.len = 0
        io = WorkIo()
        blk.work(io)
        torch.cuda.synchronize()
        c = blk.input.pos
        outs.append(blk.output.get().cpu().numpy().copy())
        pos += c
        if pos >= x.size and not io.call_again:
            break
        assert c > 0 or io.call_again, "no progress"
    return np.concatenate(outs)


@pytest.mark.parametrize("rate,nfilt,ntaps", [(0.768, 32, 32 * 16), (1.5, 4, 8), (2.37, 16, 16 * 12 + 5),
                                              (0.3333, 32, 32 * 8), (1.0, 8, 64), (0.1, 64, 64 * 20)])
def test_pfbarb_parity(fb, rng, rate, nfilt, ntaps):
    if ntaps == 8:
        taps = np.array([0.0, 0.25, 0.5, 0.25, 0.0, 0.0, 0.0, 0.0], np.float32)   # doc example :66-69
    else:
        taps = (orc.kaiser_lowpass(0.4 / nfilt / max(1.0, 1.0 / rate), 0.1 / nfilt, 1e-3).astype(np.float64))
        taps = np.resize(taps, ntaps).astype(np.float32) * nfilt
    x = _noise(rng, 50_000)
    ref = orc.PfbArb(rate, taps, nfilt).run(x, out_cap_per_call=1 << 20)
    cap = int(60_000 * max(rate, 1.0)) + 1024
    # same call pattern as the oracle run: everything at once
    got = _pfb_run(fb, rate, taps, nfilt, x, [1 << 30], cap)
    assert got.size == ref.size                                   # output COUNT is exact
    T = int(np.ceil(taps.size / nfilt))
    arm_l1 = max(np.sum(np.abs(taps[b::nfilt])) for b in range(nfilt))
    assert np.max(np.abs(got - ref)) <= 1e-5 * arm_l1 * np.max(np.abs(x))
    # streaming in ragged chunks gives the same stream (state carried across calls)
    got2 = _pfb_run(fb, rate, taps, nfilt, x, [3, 1, T, 4097, 10_001, 1 << 30], cap)
    ref2 = orc.PfbArb(rate, taps, nfilt)
    assert got2.size == ref.size
    assert np.max(np.abs(got2 - ref)) <= 1e-5 * arm_l1 * np.max(np.abs(x))


@pytest.mark.parametrize("rate", [0.768, 2.3])
def test_pfbarb_long_stream_wraps_the_periodic_schedule(fb, rng, rate, monkeypatch):
    """The timing recurrence is periodic (2 730 668 input samples at rate 0.768, 3 647 221 at 2.3): the device indexes
    a table of ONE period built at plan time.  A stream of > 2 periods, cut at awkward places, must give the oracle's
    exact output count and values; so must the per-call host replay it replaces (B2S_PFBARB_NO_PERIODIC=1)."""
    nfilt, T = 32, 8
    taps = np.resize(orc.kaiser_lowpass(0.4 / nfilt, 0.1 / nfilt, 1e-3), nfilt * T).astype(np.float32) * nfilt
    n = 7_600_000
    x = _noise(rng, n)
    ref = orc.PfbArb(rate, taps, nfilt).run(x, out_cap_per_call=1 << 24)
    cap = int(3_100_000 * max(rate, 1.0)) + 4096
    chunks = [1_000_003, 1_730_665, 8, 2_730_668, 1 << 30]
    got = _pfb_run(fb, rate, taps, nfilt, x, chunks, cap)
    assert got.size == ref.size
    arm_l1 = max(np.sum(np.abs(taps[b::nfilt])) for b in range(nfilt))
    tol = 1e-5 * arm_l1 * np.max(np.abs(x))
    assert np.max(np.abs(got - ref)) <= tol
    monkeypatch.setenv("B2S_PFBARB_NO_PERIODIC", "1")
    got2 = _pfb_run(fb, rate, taps, nfilt, x[:3_000_000], chunks, cap)
    monkeypatch.delenv("B2S_PFBARB_NO_PERIODIC")
    ref2 = orc.PfbArb(rate, taps, nfilt).run(x[:3_000_000], out_cap_per_call=1 << 24)
    assert got2.size == ref2.size and np.max(np.abs(got2 - ref2)) <= tol


def test_pfbarb_bad_arguments(fb):
    from futuresdr_b200.blocks import PfbArbResampler
    with pytest.raises(AssertionError):
        PfbArbResampler(0.0, np.ones(8, np.float32), 4)
    with pytest.raises(AssertionError):
        PfbArbResampler(1.0, np.ones(2, np.float32), 4)


# ---------------------------------------------------------------------------------------------
# block-level harness: Fir block through the Mocker (tests/fir.rs, perf/fir/fir.rs:94-98)
# ---------------------------------------------------------------------------------------------
def test_fir_block_mocker(fb, rng):
    from futuresdr_b200.blocks import FirBuilder, Mocker
    blk = FirBuilder.fir([1.0, 1.0, 1.0], sample_dtype=np.float32)
    m = Mocker(blk)
    m.input(np.arange(1, 7, dtype=np.float32))
    m.init_output(6)
    io = m.run_until_finished()
    v = m.output().cpu().numpy()
    assert io.finished and v.size == 4
    assert np.all(np.abs(v - [6, 9, 12, 15]) < np.finfo(np.float32).eps)
    # chain of 3 x 64-tap stages: n - stages*63 items come out (perf/fir/fir.rs:94-98)
    x = _noise(rng, 100_000, cplx=False)
    ref = x
    cur = x
    for s in range(3):
        taps = rng.random(64).astype(np.float32)
        blk = FirBuilder.fir(taps, sample_dtype=np.float32)
        m = Mocker(blk)
        m.input(cur)
        m.init_output(cur.size)
        m.run_until_finished()
        cur = m.output().cpu().numpy()
        _, _, _, ref = orc.fir(taps, ref, ref.size)
    assert cur.size == x.size - 3 * 63 == ref.size
    assert np.max(np.abs(cur - ref)) <= 1e-4 * np.max(np.abs(ref))
    # FirBuilder defaults use the reference's designs
    d = FirBuilder.decimating(4)
    assert d.n_taps() == 52
    r = FirBuilder.resampling(6, 4)                   # reduced by gcd to 3/2 (fir.rs:197-199)
    assert r.filter.interp == 3 and r.filter.decim == 2 and r.n_taps() == 72


# ---------------------------------------------------------------------------------------------
# device buffer ring (buffer/vulkan/{h2d,d2h}.rs semantics; tests/vulkan.rs data path)
# ---------------------------------------------------------------------------------------------
def test_ring_circuit_h2d_kernel_d2h(fb, rng):
    from futuresdr_b200 import _lib
    from futuresdr_b200._lib import lib, check
    import torch
    ctx = fb.default_context()
    n_items, chunk, halo = 100_000, 8192, 63
    taps = rng.uniform(-1, 1, 64).astype(np.float32)
    x = _noise(rng, n_items)
    fir = fb.FirFilter(taps)
    r_in, r_out = C.c_void_p(), C.c_void_p()
    check(lib.b2s_ring_create(ctx.handle, 8, chunk, halo, 3, 1, C.byref(r_in)), ctx.handle)
    check(lib.b2s_ring_create(ctx.handle, 8, chunk, 0, 3, 1, C.byref(r_out)), ctx.handle)
    assert lib.b2s_ring_free_slots(r_in) == 3 and lib.b2s_ring_full_slots(r_in) == 0
    got, pos, prev, prev_valid = [], 0, None, 0
    while pos < n_items:
        s = C.c_void_p()
        assert lib.b2s_ring_acquire_empty(r_in, C.byref(s)) == 0
        n = min(chunk, n_items - pos)
        host = np.ctypeslib.as_array(C.cast(lib.b2s_slot_host_ptr(s), C.POINTER(C.c_float)), shape=(2 * chunk,))
        host[: 2 * n] = x[pos:pos + n].view(np.float32)
        check(lib.b2s_ring_submit_full(r_in, s, n, 1), ctx.handle)                   # H2D edge
        full, valid = C.c_void_p(), C.c_size_t(0)
        assert lib.b2s_ring_acquire_full(r_in, C.byref(full), C.byref(valid)) == 0 and valid.value == n
        if prev is not None:
            check(lib.b2s_ring_carry_halo(r_in, prev, prev_valid, halo, full), ctx.handle)   # history in HBM
            check(lib.b2s_ring_release(r_in, prev), ctx.handle)
        h = lib.b2s_slot_halo_valid(full)
        so = C.c_void_p()
        assert lib.b2s_ring_acquire_empty(r_out, C.byref(so)) == 0
        c, p, st = C.c_size_t(0), C.c_size_t(0), C.c_int32(0)
        d_in = lib.b2s_slot_device_ptr(full) - h * 8
        check(lib.b2s_fir_exec(fir._h, C.c_void_p(d_in), n + h, C.c_void_p(lib.b2s_slot_device_ptr(so)), chunk,
                               C.byref(c), C.byref(p), C.byref(st)), ctx.handle)
        assert c.value == n + h - 63 if n + h >= 63 else True
        check(lib.b2s_ring_submit_full(r_out, so, p.value, 0), ctx.handle)
        fo, vo = C.c_void_p(), C.c_size_t(0)
        assert lib.b2s_ring_acquire_full(r_out, C.byref(fo), C.byref(vo)) == 0
        check(lib.b2s_slot_fetch_to_host(fo, vo.value), ctx.handle)                 # D2H edge
        check(lib.b2s_slot_wait(fo), ctx.handle)
        ho = np.ctypeslib.as_array(C.cast(lib.b2s_slot_host_ptr(fo), C.POINTER(C.c_float)), shape=(2 * chunk,))
        got.append(ho[: 2 * vo.value].copy().view(np.complex64))
        check(lib.b2s_ring_release(r_out, fo), ctx.handle)
        prev, prev_valid = full, n
        pos += n
    check(lib.b2s_ring_release(r_in, prev), ctx.handle)
    assert lib.b2s_ring_free_slots(r_in) == 3 and lib.b2s_ring_free_slots(r_out) == 3
    # misuse is refused, exhaustion is EAGAIN
    s1, s2, s3, s4 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert [lib.b2s_ring_acquire_empty(r_out, C.byref(v)) for v in (s1, s2, s3)] == [0, 0, 0]
    assert lib.b2s_ring_acquire_empty(r_out, C.byref(s4)) == _lib.EAGAIN
    assert lib.b2s_ring_release(r_out, s1) == 0 and lib.b2s_ring_release(r_out, s1) == _lib.ESTATE
    assert lib.b2s_ring_release(r_out, s2) == 0 and lib.b2s_ring_release(r_out, s3) == 0
    got = np.concatenate(got)
    _, _, _, ref = orc.fir(taps, x, n_items)
    assert got.size == ref.size == n_items - 63                                   # length: in - (ntaps-1), once
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.sum(np.abs(taps)) * np.max(np.abs(x))
    lib.b2s_ring_destroy(r_in)
    lib.b2s_ring_destroy(r_out)

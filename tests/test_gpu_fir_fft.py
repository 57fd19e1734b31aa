"""Parity of the overlap-save FFT FIR (long filters, Complex<f32> samples) against the oracle."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import futuresdr_b200 as fb
    return fb


def _noise(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def _check(fb, rng, taps, n, cap=None):
    import torch
    x = _noise(rng, n)
    f = fb.FirFilter(taps, algo=fb.ALGO_FFT)
    assert f.algo == fb.ALGO_FFT
    cap = n if cap is None else cap
    yd = torch.full((max(cap, 1),), 3.0, dtype=torch.complex64, device="cuda")[:cap]
    c, p, st = f.filter(torch.from_numpy(x).cuda(), yd)
    torch.cuda.synchronize()
    c0, p0, s0, ref = orc.fir(taps, x, cap)
    assert (c, p, int(st)) == (c0, p0, s0)
    if p:
        tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
        assert np.max(np.abs(yd[:p].cpu().numpy() - ref)) <= tol


@pytest.mark.parametrize("ntaps", [64, 255, 258, 777, 1024, 2048, 2049])
@pytest.mark.parametrize("ctaps", [False, True])
def test_fft_fir_parity(fb, rng, ntaps, ctaps):
    taps = _noise(rng, ntaps) if ctaps else rng.uniform(-1, 1, ntaps).astype(np.float32)
    _check(fb, rng, taps, 40000 + ntaps)


def test_fft_fir_ragged(fb, rng):
    taps = rng.uniform(-1, 1, 1024).astype(np.float32)
    for n in (1024, 1025, 3073 + 1023, 3074 + 1023, 4096, 9999, 100_003):
        for cap in (n, 1, 3073, 3074):
            _check(fb, rng, taps, n, cap=cap)


def test_fft_fir_auto_selection_and_limits(fb):
    assert fb.FirFilter(np.ones(1024, np.float32)).algo == fb.ALGO_FFT          # config-5 filter
    assert fb.FirFilter(np.ones(300, np.complex64)).algo == fb.ALGO_FFT         # long complex taps
    assert fb.FirFilter(np.ones(4000, np.float32)).algo == fb.ALGO_DIRECT       # beyond NF/2+1
    with pytest.raises(fb.B200SdrError):
        fb.FirFilter(np.ones(4000, np.float32), algo=fb.ALGO_FFT)
    with pytest.raises(fb.B200SdrError):
        fb.DecimatingFirFilter(2, np.ones(1024, np.float32), algo=fb.ALGO_FFT)
    with pytest.raises(fb.B200SdrError):
        fb.FirFilter(np.ones(1024, np.float32), sample_dtype=np.float32, algo=fb.ALGO_FFT)


def test_fft_fir_full_chunk_vs_direct(fb):
    import torch
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 16 * 1024 * 1024
    x = torch.view_as_complex(torch.randn(n + 1023, 2, generator=g, device="cuda"))
    taps = np.random.default_rng(7).uniform(-1, 1, 1024).astype(np.float32)
    yf = torch.empty(n, dtype=torch.complex64, device="cuda")
    yd = torch.empty(n, dtype=torch.complex64, device="cuda")
    assert fb.FirFilter(taps, algo=fb.ALGO_FFT).filter(x, yf)[:2] == (n, n)
    assert fb.FirFilter(taps, algo=fb.ALGO_DIRECT).filter(x, yd)[:2] == (n, n)
    scale = float(np.sum(np.abs(taps))) * float(x.abs().max())
    assert float((yf - yd).abs().max()) <= 1e-5 * scale

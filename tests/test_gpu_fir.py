"""Parity of the CUDA FIR / decimating FIR (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): within 1e-5 relative of the futuredsp CPU result.  The
reference sums in strict tap order in f32; any parallel evaluation reorders the sum, so the
tolerance is scaled the way SURVEY.md §7 "Parity metric" states it:
    |y - y_ref| <= 1e-5 * ||taps||_1 * max|x|
(element-wise relative error is meaningless for white-noise outputs that pass through 0).
Counts (consumed, produced, status) must match the reference EXACTLY.
"""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def fb():
    import torch
    import futuresdr_b200 as fb
    assert torch.cuda.is_available()
    return fb


def _noise(rng, n, cplx=True):
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def _run_dev(filt, x, cap):
    import torch
    xi = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = torch.full((max(cap, 1),), 7.0, dtype=xi.dtype, device="cuda")[:cap]
    c, p, st = filt.filter(xi, out)
    torch.cuda.synchronize()
    return c, p, int(st), out[:p].cpu().numpy()


def _tol(taps, x):
    return RTOL * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x))) + 1e-30


# ---- the reference's own known-answer vectors, through the GPU -------------------------------
def test_known_answer_fir(fb):
    f = fb.FirFilter([1.0, 2.0, 3.0], sample_dtype=np.float32)
    assert f.length() == 3
    c, p, st, o = _run_dev(f, np.array([1, 2, 3], np.float32), 3)
    assert (c, p, st) == (1, 1, 0) and o[0] == 10.0                      # fir.rs:283-294
    assert _run_dev(f, np.array([1, 2, 3], np.float32), 0)[:3] == (0, 0, 1)
    c, p, st, o = _run_dev(f, np.array([1, 2, 3, 4, 5], np.float32), 2)
    assert (c, p, st) == (2, 2, 1) and list(o) == [10.0, 16.0]           # fir.rs:309-318
    f = fb.FirFilter([1.0, 2.0], sample_dtype=np.float32)                # fir.rs:321-343
    assert _run_dev(f, np.array([1, 2, 3, 4, 5], np.float32), 3)[:3] == (3, 3, 1)
    assert _run_dev(f, np.array([1, 2, 3, 4], np.float32), 3)[:3] == (3, 3, 2)
    f = fb.FirFilter([1.0, 1.0, 1.0], sample_dtype=np.float32)           # tests/fir.rs:7-31
    c, p, st, o = _run_dev(f, np.arange(1, 7, dtype=np.float32), 6)
    assert p == 4 and np.all(np.abs(o - [6, 9, 12, 15]) < np.finfo(np.float32).eps)


def test_known_answer_decimating(fb):
    f = fb.DecimatingFirFilter(2, [1.0, 2.0, 3.0], sample_dtype=np.float32)   # decimating_fir.rs:341-394
    for x, cap, want, vals in [
        ([0, 1, 2, 3], 3, (2, 1, 0), [10.0]), ([0, 1, 2, 3, 4], 3, (2, 1, 0), [10.0]),
        ([0, 1, 2, 3, 4], 1, (2, 1, 2), [10.0]), ([0, 1, 2, 3, 4, 5], 1, (2, 1, 1), [10.0]),
        ([0, 1, 2, 3, 4, 5], 3, (4, 2, 0), [10.0, 22.0]), ([0, 1, 2, 3, 4, 5], 0, (0, 0, 1), []),
    ]:
        c, p, st, o = _run_dev(f, np.array(x, np.float32), cap)
        assert (c, p, st) == want and list(o) == vals
    f = fb.DecimatingFirFilter(3, [1.0, 2.0, 1.0], sample_dtype=np.float32)   # :396-441
    for x, cap, want, vals in [
        ([0, 1, 2, 3], 3, (0, 0, 0), []), ([0, 1, 2, 3, 4, 5], 3, (3, 1, 0), [12.0]),
        ([0, 1, 2, 3, 4, 5], 1, (3, 1, 2), [12.0]), ([0, 1, 2, 3, 4, 5, 6], 3, (3, 1, 0), [12.0]),
        ([0, 1, 2, 3, 4, 5, 6, 7], 3, (6, 2, 0), [12.0, 24.0]),
    ]:
        c, p, st, o = _run_dev(f, np.array(x, np.float32), cap)
        assert (c, p, st) == want and list(o) == vals


# ---- seeded parity vs the oracle -------------------------------------------------------------
@pytest.mark.parametrize("ntaps", [1, 2, 3, 7, 8, 9, 63, 64, 65, 256, 257, 1024])
@pytest.mark.parametrize("kind", ["f32", "c32", "c32c"])
def test_fir_parity(fb, rng, ntaps, kind):
    n = 5000 + ntaps
    x = _noise(rng, n, cplx=kind != "f32")
    taps = _noise(rng, ntaps, cplx=True) if kind == "c32c" else rng.uniform(-1, 1, ntaps).astype(np.float32)
    f = fb.FirFilter(taps, sample_dtype=x.dtype, algo=fb.ALGO_DIRECT)
    c0, p0, s0, ref = orc.fir(taps, x, n)
    c, p, st, o = _run_dev(f, x, n)
    assert (c, p, st) == (c0, p0, s0)
    assert np.max(np.abs(o - ref)) <= _tol(taps, x)


@pytest.mark.parametrize("decim", [2, 3, 4, 5, 8, 16, 25, 64])
@pytest.mark.parametrize("ntaps", [5, 52, 129])
def test_decimating_parity(fb, rng, decim, ntaps):
    n = 9000
    x = _noise(rng, n)
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    f = fb.DecimatingFirFilter(decim, taps, sample_dtype=np.complex64)
    for cap in (n, 37):
        c0, p0, s0, ref = orc.decim_fir(taps, decim, x, cap)
        c, p, st, o = _run_dev(f, x, cap)
        assert (c, p, st) == (c0, p0, s0)
        assert np.max(np.abs(o - ref)) <= _tol(taps, x)


def test_decimating_default_taps_c32c(fb, rng):
    # FirBuilder::decimating(4) taps (52, kaiser) and a complex-tap decimator (XlatingFir core)
    taps = fb.firdes.kaiser.lowpass(0.25, 0.1, 1e-4)
    assert np.array_equal(taps, orc.kaiser_lowpass(0.25, 0.1, 1e-4))
    x = _noise(rng, 40000)
    f = fb.DecimatingFirFilter(4, taps)
    c0, p0, s0, ref = orc.decim_fir(taps, 4, x, 40000)
    c, p, st, o = _run_dev(f, x, 40000)
    assert (c, p, st) == (c0, p0, s0) and np.max(np.abs(o - ref)) <= _tol(taps, x)
    ct = _noise(rng, 31)
    f = fb.DecimatingFirFilter(5, ct)
    c0, p0, s0, ref = orc.decim_fir(ct, 5, x, 40000)
    c, p, st, o = _run_dev(f, x, 40000)
    assert (c, p, st) == (c0, p0, s0) and np.max(np.abs(o - ref)) <= _tol(ct, x)


def test_ragged_and_edge_sizes(fb, rng):
    taps = rng.uniform(-1, 1, 33).astype(np.float32)
    f = fb.FirFilter(taps)
    for n in (0, 1, 32, 33, 34, 1023, 1024, 1025, 1056, 1057, 2048 + 32, 3000):
        x = _noise(rng, n)
        for cap in (0, 1, n, n + 100):
            c0, p0, s0, ref = orc.fir(taps, x, cap)
            c, p, st, o = _run_dev(f, x, cap)
            assert (c, p, st) == (c0, p0, s0), (n, cap)
            if p:
                assert np.max(np.abs(o - ref)) <= _tol(taps, x)


def test_unaligned_device_slices(fb, rng):
    import torch
    taps = rng.uniform(-1, 1, 64).astype(np.float32)
    f = fb.FirFilter(taps)
    x = _noise(rng, 10000)
    xd = torch.from_numpy(x).cuda()
    out = torch.zeros(10000, dtype=torch.complex64, device="cuda")
    # slice offsets of 1 item = 8 bytes: breaks 16-byte alignment on both sides
    c, p, st = f.filter(xd[1:], out[3:])
    torch.cuda.synchronize()
    c0, p0, s0, ref = orc.fir(taps, x[1:], 10000 - 3)
    assert (c, p, int(st)) == (c0, p0, s0)
    assert np.max(np.abs(out[3:3 + p].cpu().numpy() - ref)) <= _tol(taps, x)


def test_streaming_chunks_equal_one_shot(fb, rng):
    # Fir::work semantics (src/blocks/fir.rs:75-94): consume `consumed`, keep the tail as history
    taps = rng.uniform(-1, 1, 100).astype(np.float32)
    x = _noise(rng, 20000)
    f = fb.DecimatingFirFilter(3, taps)
    _, p_all, _, ref = orc.decim_fir(taps, 3, x, 20000)
    outs, pos, steps = [], 0, [4096, 1000, 7777, 5000]
    while True:
        step = steps.pop(0) if steps else 99999
        end = min(pos + step, x.size)
        c, p, st, o = _run_dev(f, x[pos:end], 4000)
        outs.append(o)
        pos += c
        if end == x.size and st != 1:      # input.finished() && status != InsufficientOutput
            break
    got = np.concatenate(outs)
    # a tail shorter than one full window may remain unconsumed, exactly like the reference
    assert got.size <= p_all and p_all - got.size <= (x.size - pos) // 3 + 1
    assert np.max(np.abs(got - ref[:got.size])) <= _tol(taps, x)


def test_host_slices_drop_in(fb, rng):
    # Filter::filter(&[In], &mut [Out]) with host memory end to end through the C ABI
    taps = rng.uniform(-1, 1, 256).astype(np.float32)
    f = fb.FirFilter(taps, algo=fb.ALGO_DIRECT)
    x = _noise(rng, 300000)
    out = np.zeros(x.size, np.complex64)
    c, p, st = f.filter(x, out)
    c0, p0, s0, ref = orc.fir(taps, x, x.size)
    assert (c, p, int(st)) == (c0, p0, s0)
    assert np.max(np.abs(out[:p] - ref)) <= _tol(taps, x)


def test_linearity_and_shift_invariance_full_size(fb):
    """Size-independent properties at the BASELINE chunk size (64 Mi samples, 256 taps):
    filter(a*x) == a*filter(x) and filter(shift(x)) == shift(filter(x)) on device."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(0x5EED)
    n = 64 * 1024 * 1024
    x = torch.view_as_complex(torch.randn(n + 255 + 5, 2, generator=g, device="cuda"))
    taps = np.random.default_rng(7).uniform(-1, 1, 256).astype(np.float32)
    f = fb.FirFilter(taps)
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    y2 = torch.empty(n, dtype=torch.complex64, device="cuda")
    c, p, st = f.filter(x[:n + 255], y)
    assert (c, p) == (n, n)
    f.filter(x[5:n + 255 + 5], y2)
    scale = float(np.sum(np.abs(taps))) * float(x.abs().max())
    assert float((y[5:] - y2[:-5]).abs().max()) <= 1e-5 * scale          # shift invariance
    f.filter((x[:n + 255] * 0.5).contiguous(), y2)
    assert float((y * 0.5 - y2).abs().max()) <= 1e-5 * scale              # homogeneity (0.5 exact)
    # spot-check 3 windows against the oracle at the full size
    xh = x[:n + 255]
    for k0 in (0, n // 2 + 12345, n - 4096):
        seg = xh[k0:k0 + 4096 + 255].cpu().numpy()
        _, _, _, ref = orc.fir(taps, seg, 4096)
        got = y[k0:k0 + 4096].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-5 * scale


def test_bad_arguments_fail_loudly(fb):
    with pytest.raises(fb.B200SdrError):
        fb.FirFilter(np.zeros(0, np.float32))
    with pytest.raises(fb.B200SdrError):
        fb.DecimatingFirFilter(0, [1.0, 2.0])
    with pytest.raises(TypeError):
        fb.FirFilter(np.ones(3, np.complex64), sample_dtype=np.float32)


def test_f64_fir_bit_exact_and_known_answer(rng):
    """The f64 x f64 Filter impl (fir.rs:217-226): the reference's f64 known-answer test (fir.rs:343-365: taps [1, 2],
    4 in / 3 out -> (3, 3, BothSufficient)) and random data bit for bit (un-fused multiply/add in tap order), device
    and host slices, with decimation."""
    import torch
    import futuresdr_b200 as fb
    f = fb.FirFilter([1.0, 2.0], sample_dtype=np.float64)
    o = torch.zeros(3, dtype=torch.float64, device="cuda")
    c, p, st = f.filter(torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64, device="cuda"), o)
    torch.cuda.synchronize()
    assert (c, p, int(st)) == (3, 3, 2) and o.cpu().tolist() == [4.0, 7.0, 10.0]
    for ntaps, decim in ((64, 1), (257, 1), (33, 3), (5000, 2)):
        taps = rng.standard_normal(ntaps)
        x = rng.standard_normal(200_000)
        flt = fb.DecimatingFirFilter(decim, taps, sample_dtype=np.float64)
        od = torch.zeros(x.size, dtype=torch.float64, device="cuda")
        c, p, st = flt.filter(torch.from_numpy(x).cuda(), od)
        torch.cuda.synchronize()
        # oracle: orc_fir_f64_f64 (decim 1) / the same loop at stride D
        g = taps[::-1]
        n_ref = (x.size + 1 - ntaps) // decim
        assert (c, p) == (n_ref * decim, n_ref)
        if decim == 1:
            _, _, _, ref = orc.fir_f64(taps, x, x.size)
            assert np.array_equal(od[:p].cpu().numpy(), ref)
        else:
            k = rng.integers(0, n_ref, 200)
            for kk in k:
                s = 0.0
                seg = x[decim - 1 + kk * decim: decim - 1 + kk * decim + ntaps]
                for t in range(ntaps):
                    s = s + seg[t] * g[t]
                assert od[kk].item() == s
        ho = np.zeros(p, np.float64)
        c2, p2, _ = flt.filter(x, ho)
        assert (c2, p2) == (c, p) and np.array_equal(ho, od[:p].cpu().numpy())

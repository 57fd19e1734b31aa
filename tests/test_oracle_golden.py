"""Pin the CPU oracle against every known-answer vector the reference's own tests hold for
the hot path.  Each test names the reference test it replays."""
import numpy as np
import pytest

import oracle as orc

II, IO, BS = orc.INSUFFICIENT_INPUT, orc.INSUFFICIENT_OUTPUT, orc.BOTH_SUFFICIENT


# ---- crates/futuredsp/src/fir.rs:283-319 direct_fir_kernel --------------------------------
def test_fir_direct_kernel():
    taps = [1.0, 2.0, 3.0]
    c, p, st, o = orc.fir(taps, [1.0, 2.0, 3.0], 3)
    assert (c, p, st) == (1, 1, II) and o[0] == 10.0
    c, p, st, o = orc.fir(taps, [1.0, 2.0, 3.0], 0)
    assert (c, p, st) == (0, 0, IO)
    c, p, st, o = orc.fir(taps, [1.0, 2.0, 3.0, 4.0, 5.0], 2)
    assert (c, p, st) == (2, 2, IO) and list(o) == [10.0, 16.0]


# ---- fir.rs:321-365 terminating_condition (+ f64) -----------------------------------------
@pytest.mark.parametrize("fn", [orc.fir, orc.fir_f64])
def test_fir_terminating_condition(fn):
    taps = [1.0, 2.0]
    c, p, st, _ = fn(taps, [1.0, 2.0, 3.0, 4.0, 5.0], 3)
    assert (c, p, st) == (3, 3, IO)
    c, p, st, _ = fn(taps, [1.0, 2.0, 3.0, 4.0], 3)
    assert (c, p, st) == (3, 3, BS)


# ---- decimating_fir.rs:313-441 -------------------------------------------------------------
def test_decim_one():
    taps = [1.0, 2.0, 3.0]
    c, p, st, o = orc.decim_fir(taps, 1, [1.0, 2.0, 3.0], 3)
    assert (c, p, st) == (1, 1, II) and o[0] == 10.0
    assert orc.decim_fir(taps, 1, [1.0, 2.0, 3.0], 0)[:3] == (0, 0, IO)
    c, p, st, o = orc.decim_fir(taps, 1, [1.0, 2.0, 3.0, 4.0, 5.0], 2)
    assert (c, p, st) == (2, 2, IO) and list(o) == [10.0, 16.0]


def test_decim_two():
    taps = [1.0, 2.0, 3.0]
    cases = [
        ([0, 1, 2, 3], 3, (2, 1, II), [10.0]),
        ([0, 1, 2, 3, 4], 3, (2, 1, II), [10.0]),
        ([0, 1, 2, 3, 4], 1, (2, 1, BS), [10.0]),
        ([0, 1, 2, 3, 4, 5], 1, (2, 1, IO), [10.0]),
        ([0, 1, 2, 3, 4, 5], 3, (4, 2, II), [10.0, 22.0]),
        ([0, 1, 2, 3, 4, 5], 0, (0, 0, IO), []),
    ]
    for x, cap, want, vals in cases:
        c, p, st, o = orc.decim_fir(taps, 2, np.array(x, np.float32), cap)
        assert (c, p, st) == want and list(o) == vals


def test_decim_three():
    taps = [1.0, 2.0, 1.0]
    cases = [
        ([0, 1, 2, 3], 3, (0, 0, II), []),
        ([0, 1, 2, 3, 4, 5], 3, (3, 1, II), [12.0]),
        ([0, 1, 2, 3, 4, 5], 1, (3, 1, BS), [12.0]),
        ([0, 1, 2, 3, 4, 5, 6], 3, (3, 1, II), [12.0]),
        ([0, 1, 2, 3, 4, 5, 6, 7], 3, (6, 2, II), [12.0, 24.0]),
        ([0, 1, 2, 3, 4, 5, 6, 7], 0, (0, 0, IO), []),
    ]
    for x, cap, want, vals in cases:
        c, p, st, o = orc.decim_fir(taps, 3, np.array(x, np.float32), cap)
        assert (c, p, st) == want and list(o) == vals


def test_decim_terminating_condition():
    taps = [1.0, 2.0]
    assert orc.decim_fir(taps, 1, [1.0, 2, 3, 4, 5], 3)[:3] == (3, 3, IO)
    assert orc.decim_fir(taps, 1, [1.0, 2, 3, 4], 3)[:3] == (3, 3, BS)


# ---- polyphase_resampling_fir.rs:174-260 --------------------------------------------------
def test_polyphase_resampling():
    taps = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    x5 = np.array([1, 2, 3, 4, 5], np.float32)
    c, p, st, o = orc.resamp_fir(taps, 3, 2, x5, 8)
    assert (c, p, st) == (2, 3, II) and list(o) == [6.0, 12.0, 16.0]
    assert orc.resamp_fir(taps, 3, 2, x5, 0)[:3] == (0, 0, IO)
    c, p, st, o = orc.resamp_fir(taps, 3, 2, x5, 3)
    assert (c, p, st) == (2, 3, BS) and list(o) == [6.0, 12.0, 16.0]
    x8 = np.array([1, 2, 3, 4, 5, 6, 7, 8], np.float32)
    c, p, st, o = orc.resamp_fir(taps, 3, 2, x8, 3)
    assert (c, p, st) == (2, 3, IO) and list(o) == [6.0, 12.0, 16.0]
    c, p, st, o = orc.resamp_fir(taps, 3, 2, x8[2:], 3)        # streaming continuation
    assert (c, p, st) == (2, 3, IO) and list(o) == [16.0, 30.0, 30.0]
    c, p, st, o = orc.resamp_fir(taps, 3, 2, x8[4:], 3)
    assert (c, p, st) == (2, 3, BS) and list(o) == [26.0, 48.0, 44.0]
    # interp 2 / decim 1
    c, p, st, o = orc.resamp_fir([1.0, 2.0], 2, 1, np.array([1, 2, 3, 4], np.float32), 10)
    assert (c, p, st) == (3, 6, II) and list(o) == [1.0, 2.0, 2.0, 4.0, 3.0, 6.0]
    # interp 1 / decim 3
    c, p, st, o = orc.resamp_fir([1.0, 2.0], 1, 3, x8, 8)
    assert (c, p, st) == (6, 2, II) and list(o) == [4.0, 13.0]


# ---- tests/fir.rs:7-31 block-level: [1..6] * [1,1,1] -> [6,9,12,15] ------------------------
def test_fir_block_level_vector():
    c, p, st, o = orc.fir([1.0, 1.0, 1.0], np.arange(1, 7, dtype=np.float32), 6)
    assert (c, p) == (4, 4) and np.all(np.abs(o - [6, 9, 12, 15]) < np.finfo(np.float32).eps)


# ---- perf/fir/fir.rs:94-98: outputs = samples - stages*63 for 64 taps ---------------------
def test_fir_chain_length():
    rng = np.random.default_rng(2)
    x = rng.standard_normal(4096).astype(np.float32)
    n, stages = x.size, 3
    for _ in range(stages):
        taps = rng.random(64).astype(np.float32)
        _, _, _, x = orc.fir(taps, x, x.size)
    assert x.size == n - stages * 63


# ---- numpy cross-check: FirFilter == np.convolve(x, taps, 'valid') ------------------------
def test_fir_matches_numpy_convolve(rng):
    x = (rng.standard_normal(2000) + 1j * rng.standard_normal(2000)).astype(np.complex64)
    taps = rng.standard_normal(37).astype(np.float32)
    _, p, _, o = orc.fir(taps, x, 4000)
    ref = np.convolve(x.astype(np.complex128), taps.astype(np.float64), "valid")
    assert p == ref.size
    assert np.max(np.abs(o - ref)) <= 1e-5 * np.sum(np.abs(taps)) * np.max(np.abs(x))
    ctaps = (rng.standard_normal(9) + 1j * rng.standard_normal(9)).astype(np.complex64)
    _, p, _, o = orc.fir(ctaps, x, 4000)
    ref = np.convolve(x.astype(np.complex128), ctaps.astype(np.complex128), "valid")
    assert np.max(np.abs(o - ref)) <= 1e-5 * np.sum(np.abs(ctaps)) * np.max(np.abs(x))
    # decimating: every D-th output of the full filter, starting at phase D-1
    for D in (2, 3, 4, 5):
        _, p, _, od = orc.decim_fir(taps, D, x, 4000)
        full = np.convolve(x.astype(np.complex128), taps.astype(np.float64), "valid")
        assert np.allclose(od, full[D - 1::D][:p], atol=1e-4)


# ---- math/special_funs.rs:53-123 besseli0 --------------------------------------------------
def test_besseli0_accuracy():
    xs = [-3.75, -3.0, -2.0, -1.5, -1.0, -0.3, -0.2, -0.1, -0.01, -0.001, 0.0, 0.001, 0.01, 0.1,
          0.2, 0.3, 1.0, 1.5, 2.0, 3.0, 3.75]
    ys = [9.118945860844564, 4.880792585865025, 2.279585302336067, 1.646723189772891,
          1.266065877752008, 1.022626879351597, 1.010025027795146, 1.002501562934095,
          1.000025000156250, 1.000000250000016, 1.0, 1.000000250000016, 1.000025000156250,
          1.002501562934095, 1.010025027795146, 1.022626879351597, 1.266065877752008,
          1.646723189772891, 2.279585302336067, 4.880792585865025, 9.118945860844564]
    for x, y in zip(xs, ys):
        assert abs(orc.besseli0(x) - y) < 1.6e-7 + 1.0e-7 * abs(y)
    xs = [3.8, 4.0, 4.5, 5.0, 5.5, 6.0, 7.0, 8.0, 9.0, 10.0, 20.0]
    ys = [9.516888026098954, 11.301921952136331, 17.481171855609279, 27.239871823604449,
          42.694645151847787, 67.234406976477985, 1.685939085102897e2, 4.275641157218048e2,
          1.093588354511375e3, 2.815716628466255e3, 4.355828255955355e7]
    for x, y in zip(xs, ys):
        assert abs(orc.besseli0(x) - y) < 1.9e-7 + 1.0e-7 * abs(y)


# ---- firdes/basic.rs:465-535 kaiser::lowpass_accuracy (MATLAB fir1, tol 1e-2) -------------
def test_kaiser_lowpass_accuracy():
    want = [0.000801064154378, -0.002365829920883, -0.002317066829825, 0.002912423701086,
            0.004722494338058, -0.002581790957417, -0.007902817296928, 0.000761425035067,
            0.011472606580612, 0.003169041375600, -0.014740633607712, -0.009778385805180,
            0.016687423513410, 0.019601855418468, -0.015887002008125, -0.033375572621574,
            0.010135834366629, 0.052954908730137, 0.005241422655623, -0.085435542746372,
            -0.047877021123625, 0.179797936334912, 0.413161963225821]
    want = want + want[::-1]
    got = orc.kaiser_lowpass(0.2, 0.05, 0.01, dtype=np.float64)
    assert got.size == len(want) == 46
    assert np.max(np.abs(got - want)) < 1e-2


# ---- firdes/basic.rs:697-760 kaiser::multirate_accuracy (designMultirateFIR, tol 1e-5) ----
def test_kaiser_multirate_accuracy():
    half = [0.0, -0.000456080632562, -0.001109227477145, 0.0, 0.004072775613512,
            0.006844614119589, 0.0, -0.016512756288837, -0.024225080517374, 0.0,
            0.048278505847958, 0.066425028523671, 0.0, -0.123967404911009, -0.172122083496355,
            0.0, 0.395134052036115, 0.817675050290108]
    want = half + [1.0] + half[::-1][:-1]
    got = orc.kaiser_multirate(3, 2, 6, 0.0001, dtype=np.float64)
    assert got.size == len(want) == 36
    assert np.max(np.abs(got - want)) < 1e-5


def test_default_decimator_taps_count():
    # FirBuilder::decimating (src/blocks/fir.rs:154): kaiser::lowpass(1/decim, 0.1, 1e-4);
    # 52 taps for decim=4
    assert orc.kaiser_lowpass(0.25, 0.1, 0.0001).size == 52
    # FirBuilder::resampling default: multirate(interp, decim, 12, 1e-4) -> 24*band taps
    assert orc.kaiser_multirate(3, 2, 12, 0.0001).size == 72
    assert orc.kaiser_multirate(1, 1, 12, 0.0001).tolist() == [1.0]

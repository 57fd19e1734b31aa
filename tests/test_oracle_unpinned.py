"""The three rows whose parity is UNPINNED by the reference's own tests (Fft,
PfbArbResampler, Apply/demod): the C oracle is cross-checked here against independent
restatements so that a transcription slip in oracle.c does not go unnoticed."""
import numpy as np
import pytest

import oracle as orc


def test_fft_oracle_vs_numpy(rng):
    for n in (2, 8, 64, 100, 2048, 4096):       # 100: the direct O(N^2) branch
        x = (rng.standard_normal(3 * n + 1) + 1j * rng.standard_normal(3 * n + 1)).astype(np.complex64)
        m, y = orc.fft_block(x, n)
        assert m == 3 * n
        X = x[:m].reshape(3, n).astype(np.complex128)
        assert np.allclose(y.reshape(3, n), np.fft.fft(X, axis=1), atol=2e-6 * np.sqrt(n) * 4, rtol=1e-6)
        _, y = orc.fft_block(x, n, fft_shift=True)                       # fwd: shift AFTER (fft.rs:196-204)
        assert np.allclose(y.reshape(3, n), np.fft.fftshift(np.fft.fft(X, axis=1), axes=1), atol=1e-4, rtol=1e-6)
        _, y = orc.fft_block(x, n, inverse=True)                          # inverse is un-normalised
        assert np.allclose(y.reshape(3, n), np.fft.ifft(X, axis=1) * n, atol=1e-4, rtol=1e-6)
        _, y = orc.fft_block(x, n, inverse=True, fft_shift=True, normalize=0.5)   # inv: shift BEFORE (:179-185)
        ref = np.fft.ifft(np.fft.ifftshift(X, axes=1), axis=1) * n * 0.5
        assert np.allclose(y.reshape(3, n), ref, atol=1e-4, rtol=1e-6)
    # capacity limits m (fft.rs:169-170)
    m, y = orc.fft_block(np.ones(4096 * 3, np.complex64), 4096, out_cap=4096 * 2 + 5)
    assert m == 8192


def test_quad_demod_oracle(rng):
    x = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(np.complex64)
    y, carry = orc.quad_demod(x)
    prev = np.concatenate([[0], x[:-1]])
    assert np.allclose(y, np.angle(x.astype(np.complex128) * np.conj(prev)), atol=2e-6)
    assert carry == (float(x[-1].real), float(x[-1].imag))
    y2a, c = orc.quad_demod(x[:300])
    y2b, _ = orc.quad_demod(x[300:], c)
    assert np.array_equal(np.concatenate([y2a, y2b]), y)                  # carry == closure state


class PyPfbArb:
    """Line-by-line Python transcription of arb_resampler.rs + window_buffer.rs + utilities.rs
    (independent of oracle.c; numpy float32 scalars force f32 rounding after every operation)."""

    def __init__(self, rate, taps, nf):
        f32 = np.float32
        self.nf = nf
        T = int(np.ceil(f32(len(taps)) / f32(nf)))
        self.T = T
        self.arms = []
        for i in range(nf):
            a = list(taps[i::nf])
            a += [0.0] * (T - len(a))
            self.arms.append(np.array(a, np.float32))
        self.circ = np.zeros(2 * T, np.complex64)
        self.start, self.missing = 0, T
        self.rate, self.delay = f32(rate), f32(1.0) / f32(rate)
        self.tau = f32(0); self.bf = f32(0); self.base = 0; self.mu = f32(0)
        self.boundary = False
        self.buff = [np.complex64(0), np.complex64(0)]

    def push(self, s):
        idx = (self.start - self.missing) % self.T
        self.circ[idx] = s; self.circ[idx + self.T] = s
        self.missing = max(self.missing - 1, 0)
        self.start = (self.start + 1) % self.T

    def filt(self, arm):
        win = self.circ[self.start:self.start + self.T]
        re = np.float32(0); im = np.float32(0)
        a = self.arms[arm]
        for t in range(self.T):
            tap = a[self.T - 1 - t]
            re = np.float32(re + np.float32(win[t].real * tap)); im = np.float32(im + np.float32(win[t].imag * tap))
        return re, im

    def upd(self):
        f32 = np.float32
        self.tau = f32(self.tau + self.delay)
        self.bf = f32(self.tau * f32(self.nf))
        self.base = int(np.floor(self.bf))
        self.mu = f32(self.bf - f32(self.base))

    def blend(self):
        f32 = np.float32
        a = f32(f32(1.0) - self.mu)
        (r0, i0), (r1, i1) = self.buff
        return complex(f32(f32(a * r0) + f32(self.mu * r1)), f32(f32(a * i0) + f32(self.mu * i1)))

    def consume_single(self, s, out):
        self.push(s)
        while self.base < self.nf:
            if self.boundary:
                self.buff[1] = self.filt(0)
                out.append(self.blend()); self.upd(); self.boundary = False
            else:
                self.buff[0] = self.filt(self.base)
                if self.base == self.nf - 1:
                    self.boundary = True; self.base = self.nf
                else:
                    self.buff[1] = self.filt(self.base + 1)
                    out.append(self.blend()); self.upd()
        self.tau = np.float32(self.tau - np.float32(1.0))
        self.bf = np.float32(self.bf - np.float32(self.nf))
        self.base -= self.nf

    def run(self, x):
        out, pos = [], 0
        while self.missing and pos < len(x):
            self.push(x[pos]); pos += 1
        for s in x[pos:]:
            self.consume_single(s, out)
        return np.array(out, np.complex64)


@pytest.mark.parametrize("rate,nf,ntaps", [(1.5, 4, 8), (0.768, 8, 37), (2.3, 5, 23), (0.25, 4, 16)])
def test_pfbarb_oracle_vs_python_transcription(rng, rate, nf, ntaps):
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    x = (rng.standard_normal(700) + 1j * rng.standard_normal(700)).astype(np.complex64)
    a = orc.PfbArb(rate, taps, nf).run(x)
    b = PyPfbArb(rate, taps, nf).run(x)
    assert a.size == b.size and a.size > 0
    assert np.array_equal(a, b)                     # same f32 operation order -> bit-exact
    # call pattern independence (state machine carried across work() calls)
    o = orc.PfbArb(rate, taps, nf)
    parts, pos = [], 0
    for step in (1, 2, 5, 100, 3, 10 ** 6):
        while True:
            c, p, ca, out = o.work(x[pos:pos + step], 1 << 16)
            parts.append(out); pos += c
            if not ca or c == 0:
                break
    assert np.array_equal(np.concatenate(parts), a)


def test_pfbarb_window_fill_quirk_is_reproduced():
    # window_buffer.rs:24-32 while filling writes sample j at slot (start_idx - missing) mod L,
    # i.e. 0, 2, 4, ... (not an append).  With an impulse as 2nd input sample and T = 4 the
    # impulse lands in slot 2 of the first window, and the first output shows tap arm[...][T-1-2].
    taps = np.arange(1, 17, dtype=np.float32)       # 4 arms x 4 taps
    x = np.zeros(12, np.complex64); x[1] = 1.0
    y = orc.PfbArb(1.0, taps, 4).run(x)
    ref = PyPfbArb(1.0, taps, 4).run(x)
    assert np.array_equal(y, ref) and y.size > 0


def test_rotator_oracle_short_run_matches_closed_form(rng):
    # over a short run the f32 recurrence is still within ~n*eps of exp(i*n*incr)
    x = (rng.standard_normal(200) + 1j * rng.standard_normal(200)).astype(np.complex64)
    r = orc.Rotator(0.3)
    y = r.rotate(x)
    ref = x.astype(np.complex128) * np.exp(1j * 0.3 * np.arange(1, 201))
    assert np.max(np.abs(y - ref)) < 1e-4
    # state continues across calls
    r2 = orc.Rotator(0.3)
    assert np.array_equal(np.concatenate([r2.rotate(x[:77]), r2.rotate(x[77:])]), y)


def test_xlating_taps_oracle():
    taps = np.array([1.0, 2.0, 3.0], np.float32)
    bpf = orc.xlating_taps(taps, 1000.0, 8000.0)
    want = taps * np.exp(1j * 2 * np.pi * 1000.0 / 8000.0 * np.arange(3))
    assert np.allclose(bpf, want, atol=1e-6)


def test_channelizer_oracle_tone_and_call_pattern(rng):
    N = 8
    taps = orc.kaiser_lowpass(0.4 / N, 0.1 / N, 1e-3).astype(np.float32)
    n = N * 600
    for k in (0, 2, 5):
        x = np.exp(2j * np.pi * (k / N) * np.arange(n)).astype(np.complex64)
        y = orc.PfbChannelizer(N, taps, 1.0).run(x)
        p = np.abs(y[:, -1])
        assert y.shape == (N, n // N) and np.argmax(p) == k and p[k] > 0.9 and np.all(np.delete(p, k) < 1e-2)
    # the call that completes the window fill consumes nothing (channelizer.rs:170-180)
    ch = orc.PfbChannelizer(N, taps, 1.0)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    c, p, ca, _ = ch.work(x, 1 << 16)
    assert (c, p, ca) == (0, 0, True)
    c, p, ca, o = ch.work(x, 1 << 16)
    assert (c, p) == (n, n // N) and o.shape == (N, n // N)
    # oversampled by 2: decimation N/2, twice the output rate
    y2 = orc.PfbChannelizer(N, taps, 2.0).run(x)
    assert y2.shape == (N, n // (N // 2))


def test_moving_avg_oracle_vs_python_transcription(rng):
    """MovingAvg (moving_avg.rs:72-115) has no value-pinning reference test: cross-check the C oracle bit for bit
    against an independent numpy-f32 transcription, including non-finite inputs, the emission counter carried
    across calls and the output-capacity stop."""
    width, decay, history = 8, np.float32(0.3), 3
    o = orc.MovingAvg(width, float(decay), history)
    avg, cnt = np.zeros(width, np.float32), 0
    one = np.float32(1.0)
    for call, (nchunks, cap_chunks) in enumerate([(7, 10), (2, 10), (9, 1), (4, 10)]):
        x = rng.standard_normal(nchunks * width + (call % 3)).astype(np.float32)
        x[rng.integers(0, x.size, 3)] = [np.inf, -np.inf, np.nan]
        want, consumed, produced = [], 0, 0
        while (consumed + 1) * width <= x.size and (produced + 1) * width <= cap_chunks * width:
            t = x[consumed * width:(consumed + 1) * width]
            fin = np.isfinite(t)
            upd = ((one - decay) * avg).astype(np.float32) + (decay * np.where(fin, t, 0)).astype(np.float32)
            avg = np.where(fin, upd.astype(np.float32), (avg * (one - decay)).astype(np.float32)).astype(np.float32)
            cnt += 1
            if cnt == history:
                want.append(avg.copy()); cnt = 0; produced += 1
            consumed += 1
        c, p, out = o.work(x, cap_chunks * width)
        assert (c, p) == (consumed * width, produced * width)
        if want:
            assert np.array_equal(out, np.concatenate(want))


class _PyWindow:
    """window_buffer.rs:4-44, transcribed independently of the C oracle."""

    def __init__(self, n):
        self.n, self.buf, self.start, self.missing = n, np.zeros(2 * n, np.complex64), 0, n

    def push(self, s):
        idx = (self.start - self.missing) % self.n
        self.buf[idx] = s
        self.buf[idx + self.n] = s
        self.missing = max(self.missing - 1, 0)
        self.start = (self.start + 1) % self.n

    def filled(self):
        return self.missing == 0

    def window(self):
        return self.buf[self.start:self.start + self.n]


def test_synthesizer_oracle_vs_python_transcription(rng):
    """PfbSynthesizer (synthesizer.rs:80-144) has no value-pinning reference test: compare the C oracle with an
    independent Python transcription (numpy inverse FFT instead of rustfft, hence a tolerance, not bit equality),
    over several work() calls with different input lengths and output capacities."""
    N, ntaps = 6, 6 * 5 + 2
    taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
    T = int(np.ceil(np.float32(ntaps) / np.float32(N)))
    arms = [np.concatenate([taps[i::N], np.zeros(T - taps[i::N].size, np.float32)]) for i in range(N)]   # utilities.rs:9-19
    wins = [_PyWindow(T) for _ in range(N)]
    all_filled = False
    o = orc.PfbSynthesizer(N, taps)
    for n_in, cap in [(3, 100), (4, 100), (10, 17), (10, 6), (25, 1000)]:
        x = (rng.standard_normal((N, n_in)) + 1j * rng.standard_normal((N, n_in))).astype(np.complex64)
        want, consumed, produced = [], 0, 0
        while n_in - consumed > 0 and (cap - produced > N or not all_filled):              # synthesizer.rs:92-94
            spun = np.fft.ifft(x[:, consumed].astype(np.complex128)) * N                  # un-normalised inverse FFT
            consumed += 1
            for w, arm, s in zip(wins, arms, spun.astype(np.complex64)):
                w.push(s)
                if w.filled():
                    win = w.window()
                    want.append(np.sum(win.astype(np.complex128) * arm[::-1]))            # FirFilter: taps applied reversed
                    produced += 1
            if not all_filled:
                all_filled = all(w.filled() for w in wins)
        c, p, out = o.work(x, cap)
        assert (c, p) == (consumed, produced)
        if want:
            assert np.max(np.abs(out - np.asarray(want))) <= 1e-4 * (np.max(np.abs(taps)) * T * N * np.max(np.abs(x)))


def test_pfbarb_timing_recurrence_is_periodic_and_the_period_matches_the_oracle():
    """The device PfbArbResampler indexes a table of ONE period of the f32 timing recurrence instead of replaying it
    per call.  The period finder is host code of the product library (b2s_pfbarb_period, no GPU needed): over exactly
    one period the oracle state machine must produce exactly `outputs_per_period` outputs, wherever the period starts."""
    import ctypes as C
    import oracle as orc
    from futuresdr_b200._lib import lib
    rng = np.random.default_rng(3)
    for rate, nfilt in ((0.768, 32), (2.3, 32), (1.5, 4), (0.9, 16), (1.0, 8)):
        lam, outs = C.c_uint64(0), C.c_uint64(0)
        assert lib.b2s_pfbarb_period(C.c_float(rate), nfilt, C.byref(lam), C.byref(outs)) == 0
        lam, outs = lam.value, outs.value
        assert lam > 0 and outs > 0
        # outputs per input over a period = the rate, to the precision of the f32 reciprocal
        assert abs(outs / lam - rate) / rate < 1e-6
        T = 4
        taps = np.ones(nfilt * T, np.float32)
        n0 = T + 1000                                  # window fill + an arbitrary offset into the cycle
        reps = max(1, 3_000_000 // lam)                # short periods: count over many of them
        x = (rng.standard_normal(n0 + reps * lam) + 0j).astype(np.complex64)
        a = orc.PfbArb(rate, taps, nfilt).run(x[:n0], out_cap_per_call=1 << 24).size
        b = orc.PfbArb(rate, taps, nfilt).run(x, out_cap_per_call=1 << 24).size
        assert b - a == reps * outs, (rate, lam, outs, b - a)

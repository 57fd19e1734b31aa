"""BASELINE config 3 as a parity case: FM-receiver chain
    FirBuilder::decimating(4) -> Apply(quadrature demod) -> PfbArbResampler(0.768, 32 arms)
streamed chunk by chunk through device-resident buffers (blocks keep their own history / state),
against the same chain evaluated by the oracle in one piece.  As SURVEY.md §7 notes, the chain
does not type-check in the reference as written (demod yields f32, PfbArbResampler takes
Complex32); like the survey we pack the phase as Complex{re: phi, im: 0}."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


def test_fm_chain_streaming_vs_oracle(rng):
    import torch
    import futuresdr_b200 as fb
    from futuresdr_b200.blocks import Apply, ApplyOp, FirBuilder, PfbArbResampler, WorkIo

    n = 1 << 20
    t = np.arange(n)
    msg = np.sin(2 * np.pi * 0.0007 * t)
    phase = 2 * np.pi * 0.05 * np.cumsum(msg)
    x = (np.exp(1j * phase) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)

    # ---- oracle, one piece
    dtaps = orc.kaiser_lowpass(0.25, 0.1, 1e-4)
    _, _, _, d_ref = orc.decim_fir(dtaps, 4, x, n)
    ph_ref, _ = orc.quad_demod(d_ref)
    ptaps = (orc.kaiser_lowpass(0.4 / 32, 0.1 / 32, 1e-3) * 32).astype(np.float32)[: 32 * 16]
    y_ref = orc.PfbArb(0.768, ptaps, 32).run(ph_ref.astype(np.complex64), out_cap_per_call=1 << 22)

    # ---- device chain, ragged chunks
    dec = FirBuilder.decimating(4)
    dem = Apply(ApplyOp.QuadDemodC32)
    pfb = PfbArbResampler(0.768, ptaps, 32)
    xd = torch.from_numpy(x).cuda()
    left = torch.zeros(0, dtype=torch.complex64, device="cuda")      # decimator history (unconsumed tail)
    outs, pos = [], 0
    chunks = [65536, 1000, 131072, 7, 300000, 1 << 30]
    ci = 0
    while pos < n:
        step = min(chunks[min(ci, len(chunks) - 1)], n - pos)
        ci += 1
        cur = torch.cat([left, xd[pos:pos + step]])
        pos += step
        d_out = torch.empty(cur.numel() // 4 + 1, dtype=torch.complex64, device="cuda")
        c, p, st = dec.filter.filter(cur, d_out)
        left = cur[c:].clone()
        if p == 0:
            continue
        ph = torch.empty(p, dtype=torch.complex64, device="cuda")
        assert dem.apply(d_out[:p], ph) == p
        # PfbArbResampler: call work() until this chunk is consumed (window fill sets call_again)
        off = 0
        while off < p:
            pfb.input.set(ph[off:])
            pfb.output.reserve(int((p - off) * 0.768) + 64)
            io = WorkIo()
            pfb.work(io)
            off += pfb.input.pos
            outs.append(pfb.output.get().clone())
            assert pfb.input.pos > 0 or io.call_again
    torch.cuda.synchronize()
    y = torch.cat(outs).cpu().numpy()
    assert y.size == y_ref.size, (y.size, y_ref.size)                  # exact output count
    # phases are O(pi); resampler arms have unit-ish gain
    assert np.max(np.abs(y - y_ref)) <= 1e-4
    assert np.max(np.abs(y.imag)) <= 1e-6


def test_fm_chain_at_bench_chunk_size():
    """The same chain at the bench's chunk geometry: 64 Mi input samples in four 16 Mi-sample chunks through
    device-resident buffers (decimator history = the 52 samples in front of each chunk), against the oracle chain in
    one piece: exact output count at every stage, values within 1e-4 (phases are O(pi))."""
    import torch
    from futuresdr_b200.blocks import Apply, ApplyOp, FirBuilder, PfbArbResampler, WorkIo
    total, S, H = 64 * 1024 * 1024, 16 * 1024 * 1024, 52
    t = np.arange(total, dtype=np.float64)
    phase = -(0.05 / 0.0007) * np.cos(2 * np.pi * 0.0007 * t)
    rng = np.random.default_rng(99)
    x = (np.exp(1j * phase) + 0.05 * (rng.standard_normal(total) + 1j * rng.standard_normal(total))).astype(np.complex64)
    del t, phase
    dtaps = orc.kaiser_lowpass(0.25, 0.1, 1e-4)
    _, _, _, d_ref = orc.decim_fir(dtaps, 4, x, total)
    ph_ref, _ = orc.quad_demod(d_ref)
    ptaps = (orc.kaiser_lowpass(0.4 / 32, 0.1 / 32, 1e-3) * 32).astype(np.float32)[: 32 * 16]
    y_ref = orc.PfbArb(0.768, ptaps, 32).run(ph_ref.astype(np.complex64), out_cap_per_call=1 << 24)
    assert d_ref.size == (total - 51) // 4

    dec = FirBuilder.decimating(4)
    dem = Apply(ApplyOp.QuadDemodC32)
    pfb = PfbArbResampler(0.768, ptaps, 32)
    xd = torch.from_numpy(x).cuda()
    d1 = torch.empty(S // 4 + 16, dtype=torch.complex64, device="cuda")
    d2 = torch.empty(S // 4 + 16, dtype=torch.complex64, device="cuda")
    out = torch.empty(y_ref.size + 4096, dtype=torch.complex64, device="cuda")
    n_dec = n_out = 0
    for c in range(total // S):
        src = xd[:S] if c == 0 else xd[c * S - H:(c + 1) * S]
        cc, p, st = dec.filter.filter(src, d1)
        assert dem.apply(d1[:p], d2) == p
        n_dec += p
        off = 0
        while off < p:
            pfb.input.set(d2[off:p])
            pfb.output.data, pfb.output.len = out[n_out:], 0
            io = WorkIo()
            pfb.work(io)
            off += pfb.input.pos
            n_out += pfb.output.len
            assert pfb.input.pos > 0 or io.call_again
    torch.cuda.synchronize()
    assert n_dec == d_ref.size
    assert n_out == y_ref.size                                         # exact output count of the whole chain
    y = out[:n_out].cpu().numpy()
    assert np.max(np.abs(y - y_ref)) <= 1e-4

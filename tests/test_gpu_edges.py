"""Host edges: VectorSource/Sink, FileSource/Sink (raw cf32 / f32 wire format) and the
pinned H2D / D2H ring, driven end to end through a linear chain of hot-path blocks."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import futuresdr_b200 as fb
    return fb


def test_vector_source_scale_sink_like_vulkan_test(fb, rng):
    """tests/vulkan.rs:56-76: random f32 -> H2D -> x12 on the device -> D2H equals orig*12 within f32 epsilon and
    the LENGTH is preserved (the partial last buffer is flushed on finish, h2d.rs:123-139)."""
    from futuresdr_b200.edges import VectorSource, VectorSink, run_chain
    from futuresdr_b200.blocks import Apply, ApplyOp
    n = 3 * (1 << 20) + 12345
    orig = rng.uniform(-1, 1, n).astype(np.float32)
    src, snk = VectorSource(orig, chunk_items=1 << 19), VectorSink(np.float32, chunk_items=1 << 19)
    run_chain([src, Apply(ApplyOp.ScaleF32, 12.0), snk], buffer_items=1 << 21)
    got = snk.items()
    assert got.size == n
    assert np.max(np.abs(got - orig * 12.0)) <= np.finfo(np.float32).eps * 12
    assert src.h2d_bytes == 4 * n and snk.d2h_bytes == 4 * n


def test_file_source_fir_file_sink_cf32(fb, rng, tmp_path):
    """cf32 capture on disk -> FileSource -> Fir (64 taps) -> FileSink; the file holds interleaved f32 re/im."""
    from futuresdr_b200.edges import FileSource, FileSink, run_chain
    from futuresdr_b200.blocks import FirBuilder
    n = (1 << 20) + 777
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    taps = rng.uniform(-1, 1, 64).astype(np.float32)
    fin, fout = tmp_path / "in.cf32", tmp_path / "out.cf32"
    x.view(np.float32).tofile(fin)                                   # interleaved re, im
    snk = FileSink(fout, np.complex64, chunk_items=1 << 18)
    run_chain([FileSource(fin, np.complex64, chunk_items=1 << 18), FirBuilder.fir(taps, np.complex64), snk],
              buffer_items=1 << 20)
    got = np.fromfile(fout, dtype=np.float32).view(np.complex64)
    _, _, _, ref = orc.fir(taps, x, n)
    assert got.size == ref.size == n - 63 and snk.n_written == n - 63     # tests/fir.rs: length in - (ntaps-1)
    assert np.max(np.abs(got - ref)) <= 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))


def test_file_source_repeat_and_small_buffers(fb, rng, tmp_path):
    """repeat=True re-opens the file at EOF (file_source.rs:66-70); a downstream Head-like cap is emulated by a
    sink that stops the chain after a fixed number of items."""
    from futuresdr_b200.edges import FileSource, VectorSink, StreamBuffer, _ReaderPort, _WriterPort
    from futuresdr_b200.blocks import WorkIo
    x = rng.uniform(-1, 1, 1000).astype(np.float32)
    f = tmp_path / "loop.f32"
    x.tofile(f)
    src = FileSource(f, np.float32, repeat=True, chunk_items=4096)
    buf = StreamBuffer(np.float32, 1 << 14)
    src.output = _WriterPort(buf)
    io = WorkIo()
    src.work(io)
    assert not io.finished and buf.wr == 4096
    got = buf.read_slice().cpu().numpy()
    assert np.array_equal(got, np.tile(x, 5)[:4096])


def test_fm_chain_from_file(fb, rng, tmp_path):
    """BASELINE configs[2] front half through the edges: cf32 file -> decimating FIR (x4, 52-tap kaiser) ->
    quadrature demodulator -> f32 VectorSink, against the oracle applied to the whole file at once."""
    from futuresdr_b200.edges import FileSource, VectorSink, run_chain
    from futuresdr_b200.blocks import FirBuilder, Apply, ApplyOp
    n = 6 * (1 << 18) + 4321
    t = np.arange(n, dtype=np.float64)
    x = (np.exp(1j * (0.05 * t + 3.0 * np.sin(2e-4 * t))) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    f = tmp_path / "fm.cf32"
    x.view(np.float32).tofile(f)
    dec = FirBuilder.decimating(4, np.complex64)
    snk = VectorSink(np.float32, chunk_items=1 << 16)
    run_chain([FileSource(f, np.complex64, chunk_items=1 << 17), dec, Apply(ApplyOp.QuadDemod), snk],
              buffer_items=1 << 19)
    taps = orc.kaiser_lowpass(0.25, 0.1, 1e-4)                      # FirBuilder.decimating default (fir.rs:154)
    _, p, _, y = orc.decim_fir(taps, 4, x, n)
    ref, _ = orc.quad_demod(y[:p])
    got = snk.items()
    assert got.size == ref.size
    assert np.max(np.abs(got - ref)) <= 2e-4          # atan2 of ~unit-modulus products; FIR error 1e-5 relative

"""CPU tests of the stream-buffer / chain-driver logic of futuresdr_b200.edges (no GPU: the buffers live in
host memory and the stages are small Python blocks with the reference's work()/finished rules)."""
import numpy as np
import pytest
import torch

from futuresdr_b200.blocks import WorkIo
from futuresdr_b200.edges import StreamBuffer, _ReaderPort, _WriterPort, run_chain


class _Src:
    """VectorSource-like: emits ``items`` in bursts of at most ``burst``."""
    in_dtype, out_dtype = None, np.float32

    def __init__(self, items, burst):
        self.items, self.pos, self.burst, self.output = torch.from_numpy(items), 0, burst, None

    def work(self, io):
        o = self.output.slice()
        n = min(o.numel(), self.burst, self.items.numel() - self.pos)
        o[:n] = self.items[self.pos:self.pos + n]
        self.pos += n
        self.output.produce(n)
        if self.pos == self.items.numel():
            io.finished = True


class _MovSum:
    """3-tap moving sum with the Fir block's consume/produce/finish rules (fir.rs:75-94)."""
    in_dtype = out_dtype = np.float32

    def __init__(self):
        self.input = self.output = None

    def work(self, io):
        i, o = self.input.slice(), self.output.slice()
        n = min(max(i.numel() - 2, 0), o.numel())
        if n:
            o[:n] = i[0:n] + i[1:n + 1] + i[2:n + 2]
        self.input.consume(n)
        self.output.produce(n)
        insufficient_output = max(i.numel() - 2, 0) > o.numel()
        if self.input.finished() and not insufficient_output:
            io.finished = True


class _Decim2:
    in_dtype = out_dtype = np.float32

    def __init__(self):
        self.input = self.output = None

    def work(self, io):
        i, o = self.input.slice(), self.output.slice()
        n = min(i.numel() // 2, o.numel())
        if n:
            o[:n] = i[0:2 * n:2]
        self.input.consume(2 * n)
        self.output.produce(n)
        if self.input.finished() and i.numel() // 2 <= o.numel():
            io.finished = True


class _Sink:
    in_dtype, out_dtype = np.float32, None

    def __init__(self, burst):
        self.input, self.got, self.burst = None, [], burst

    def work(self, io):
        i = self.input.slice()
        n = min(i.numel(), self.burst)
        self.got.append(i[:n].clone())
        self.input.consume(n)
        if self.input.finished() and n == i.numel():
            io.finished = True


def test_stream_buffer_compacts_and_keeps_order():
    b = StreamBuffer(np.float32, 16, device="cpu")
    w, r = _WriterPort(b), _ReaderPort(b)
    out, nxt = [], 0
    for step in range(200):
        s = w.slice()
        n = min(s.numel(), 1 + step % 7)
        s[:n] = torch.arange(nxt, nxt + n, dtype=torch.float32)
        nxt += n
        w.produce(n)
        have = r.slice()
        k = min(have.numel(), 1 + (step * 3) % 5)
        out.append(have[:k].clone())
        r.consume(k)
    got = torch.cat(out).numpy()
    assert np.array_equal(got, np.arange(got.size, dtype=np.float32))      # nothing lost, duplicated or reordered
    assert b.wr <= 16 and b.rd <= b.wr


@pytest.mark.parametrize("n,burst_in,burst_out,cap", [(1000, 64, 50, 128), (37, 1000, 1000, 64), (5000, 7, 3, 32),
                                                       (2, 10, 10, 16)])
def test_run_chain_matches_whole_vector(n, burst_in, burst_out, cap):
    x = np.arange(n, dtype=np.float32) * 0.5 + 1.0
    src, snk = _Src(x, burst_in), _Sink(burst_out)
    run_chain([src, _MovSum(), _Decim2(), snk], buffer_items=cap, device="cpu")
    got = torch.cat(snk.got).numpy() if snk.got else np.zeros(0, np.float32)
    ms = x[:-2] + x[1:-1] + x[2:] if n > 2 else np.zeros(0, np.float32)
    want = ms[0:2 * (ms.size // 2):2]
    assert got.size == want.size and np.array_equal(got, want)


def test_run_chain_detects_a_stuck_chain():
    class _Never(_MovSum):
        def work(self, io):
            pass                                     # never consumes, never finishes
    with pytest.raises(RuntimeError):
        run_chain([_Src(np.ones(100, np.float32), 10), _Never(), _Sink(10)], buffer_items=64, device="cpu")

"""b2s_fir_exec_hist: Filter::filter on the logical slice hist ++ in with the history handed over as a separate
pointer (another ring slot, or a peer GPU's memory), the misaligned-slice path of the tensor kernel (a ring slot's
[halo | chunk] starts 8 bytes off a 16-byte boundary when the history is 255 items), the device flags of the
cross-GPU handshake, and the single-process end of futuresdr_b200.shard's peer exchange.  Oracle: oracle/ (CPU)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle as orc  # noqa: E402


def _noise(n, dtype, seed):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype) == np.complex64:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def _ref(taps, decim, x, cap):
    c, p, st, o = orc.decim_fir(taps, decim, x, cap)
    return c, p, int(st), o


def _exec_hist(fir, hist_t, in_t, out_t, wait=None, done=None, publish=None):
    from futuresdr_b200._lib import lib, check, Handshake
    c, p, st = C.c_size_t(0), C.c_size_t(0), C.c_int32(0)
    hs = Handshake()
    if wait:
        hs.wait_flag, hs.wait_value = wait
    if done:
        hs.done_flag, hs.done_value = done
    if publish:
        hs.publish_flag, hs.publish_value = publish
    check(lib.b2s_fir_exec_hist(fir._h, C.c_void_p(hist_t.data_ptr() if hist_t is not None and hist_t.numel() else 0),
                                hist_t.numel() if hist_t is not None else 0, C.c_void_p(in_t.data_ptr()), in_t.numel(),
                                C.c_void_p(out_t.data_ptr()), out_t.numel(), C.byref(hs) if (wait or done or publish) else None,
                                C.byref(c), C.byref(p), C.byref(st)), fir.ctx.handle)
    return c.value, p.value, st.value


@pytest.mark.parametrize("dtype", [np.complex64, np.float32])
@pytest.mark.parametrize("ntaps,decim,algo", [(256, 1, "tensor"), (255, 1, "tensor"), (129, 1, "tensor"), (52, 4, "tensor"),
                                              (24, 2, "tensor"), (257, 1, "tensor"), (64, 1, "direct"), (33, 3, "direct"),
                                              (1024, 1, "auto")])
def test_exec_hist_matches_contiguous_filter(dtype, ntaps, decim, algo):
    import futuresdr_b200 as fb
    if ntaps == 1024 and np.dtype(dtype) != np.complex64:
        pytest.skip("overlap-save path is Complex<f32> only")
    n = 40000 + 4 * decim
    n -= n % (4 * decim)
    H = ((ntaps + decim - 2) // decim) * decim
    taps = np.random.default_rng(3).uniform(-1, 1, ntaps).astype(np.float32)
    x = _noise(H + n, dtype, 17)
    fir = fb.DecimatingFirFilter(decim, taps, dtype, algo={"tensor": fb.ALGO_TENSOR, "direct": fb.ALGO_DIRECT, "auto": fb.ALGO_AUTO}[algo])
    # history in its own allocation; the slice has H items of scratch in front (ring-slot contract)
    hist = torch.from_numpy(x[:H]).cuda()
    pad = ((H + 255) // 256) * 256
    buf = torch.zeros(pad + n, dtype=hist.dtype, device="cuda")
    buf[pad:] = torch.from_numpy(x[H:]).cuda()
    out = torch.zeros(n // decim + 8, dtype=hist.dtype, device="cuda")
    c, p, st = _exec_hist(fir, hist, buf[pad:], out)
    torch.cuda.synchronize()
    rc, rp, rst, ro = _ref(taps, decim, x, out.numel())
    assert (c, p, st) == (rc, rp, rst)
    tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
    err = float(np.max(np.abs(out[:p].cpu().numpy() - ro)))
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("dtype,shift", [(np.complex64, 1), (np.float32, 1), (np.float32, 2), (np.float32, 3)])
@pytest.mark.parametrize("ntaps,decim", [(256, 1), (100, 1), (52, 4)])
def test_tensor_path_on_item_aligned_slices(dtype, shift, ntaps, decim):
    """A slice that starts `shift` items past a 16-byte boundary (what a ring slot's [halo | chunk] looks like) stays
    on the tensor kernel: it starts `shift` items early and shifts the Toeplitz operand by as many zero taps.  The
    items in front of the slice are poisoned with NaN -- they must not leak into the result."""
    import futuresdr_b200 as fb
    n = 50000
    taps = np.random.default_rng(4).uniform(-1, 1, ntaps).astype(np.float32)
    x = _noise(n, dtype, 23)
    fir = fb.DecimatingFirFilter(decim, taps, dtype, algo=fb.ALGO_TENSOR)
    buf = torch.full((n + 8,), float("nan"), dtype=torch.complex64 if np.dtype(dtype) == np.complex64 else torch.float32, device="cuda")
    buf[shift:shift + n] = torch.from_numpy(x).cuda()
    out = torch.zeros(n, dtype=buf.dtype, device="cuda")
    c, p, st = fir.filter(buf[shift:shift + n], out)
    torch.cuda.synchronize()
    rc, rp, rst, ro = _ref(taps, decim, x, n)
    assert (c, p, int(st)) == (rc, rp, rst)
    got = out[:p].cpu().numpy()
    assert np.all(np.isfinite(got.view(np.float32)))
    tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
    assert float(np.max(np.abs(got - ro))) <= tol


def test_exec_hist_flags_and_timeout():
    """The handshake: a satisfied wait lets the kernel through and the done flag is stored; an unsatisfied wait gives
    up after its time-out, flags the context, and b2s_ctx_sync reports B2S_ETIMEOUT instead of hanging the GPU."""
    import futuresdr_b200 as fb
    from futuresdr_b200 import _lib
    from futuresdr_b200._lib import lib, check
    ntaps, n = 256, 1 << 16
    taps = np.random.default_rng(5).uniform(-1, 1, ntaps).astype(np.float32)
    x = _noise(ntaps - 1 + n, np.complex64, 29)
    fir = fb.FirFilter(taps, algo=fb.ALGO_TENSOR)
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    ready, consumed = flags.data_ptr(), flags.data_ptr() + 4
    hist = torch.from_numpy(x[:ntaps - 1]).cuda()
    buf = torch.zeros(256 + n, dtype=torch.complex64, device="cuda")
    buf[256:] = torch.from_numpy(x[ntaps - 1:]).cuda()
    out = torch.zeros(n, dtype=torch.complex64, device="cuda")
    check(lib.b2s_flag_set(fir.ctx.handle, C.c_void_p(ready), 7), fir.ctx.handle)
    published = flags.data_ptr() + 8
    c, p, st = _exec_hist(fir, hist, buf[256:], out, wait=(ready, 7), done=(consumed, 7), publish=(published, 41))
    fir.ctx.sync()
    v = C.c_uint32(0)
    check(lib.b2s_flag_read(fir.ctx.handle, C.c_void_p(consumed), C.byref(v)), fir.ctx.handle)
    assert v.value == 7 and p == n
    check(lib.b2s_flag_read(fir.ctx.handle, C.c_void_p(published), C.byref(v)), fir.ctx.handle)
    assert v.value == 41
    _, _, _, ro = _ref(taps, 1, x, n)
    tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
    assert float(np.max(np.abs(out.cpu().numpy() - ro))) <= tol
    # never published: time-out, reported at the next sync
    _exec_hist(fir, hist, buf[256:], out, wait=(ready, 8), done=(consumed, 8))
    with pytest.raises(fb.B200SdrError) as ei:
        fir.ctx.sync()
    assert ei.value.code == _lib.ETIMEOUT
    fir.ctx.sync()      # status word was cleared: the context stays usable


@pytest.mark.parametrize("dtype", [np.complex64, np.float32])
@pytest.mark.parametrize("ntaps,decim,S", [(256, 1, 1 << 16), (52, 4, 1 << 14), (1024, 1, 1 << 15), (64, 1, 8192), (255, 1, 1 << 15)])
def test_peer_exchange_single_rank_stream(ntaps, decim, S, dtype):
    """world == 1 through the ring: chunk t's history is the tail of the previous SLOT (no copy on the tensor path);
    the concatenated outputs equal the single-stream oracle result, first chunk shorter by ntaps-1."""
    from futuresdr_b200.shard import ShardedFir
    steps = 5
    taps = np.random.default_rng(6).uniform(-1, 1, ntaps).astype(np.float32)
    x = _noise(S * steps, dtype, 31)
    sh = ShardedFir(taps, S, dtype, decim=decim, exchange="peer")
    got = []
    for t in range(steps):
        sh.chunk.copy_(torch.from_numpy(x[t * S:(t + 1) * S]).cuda())
        out = torch.zeros(S // decim, dtype=torch.complex64 if np.dtype(dtype) == np.complex64 else torch.float32, device="cuda")
        c, p, st = sh.step(out)
        got.append((out, p))
    sh.ctx.sync()
    got = np.concatenate([o[:p].cpu().numpy() for o, p in got])
    sh.close()
    _, _, _, ref = _ref(taps, decim, x, x.size)
    assert got.size == ref.size
    tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
    assert float(np.max(np.abs(got - ref))) <= tol


def test_sharded_parity_on_all_gpus(tmp_path):
    """The torchrun parity script (peer and NCCL exchange, 7 filter shapes) at the largest world size this box has."""
    import json
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rep = tmp_path / "shard.json"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "scripts", "shard_parity.py"),
                        "--json", str(rep)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rows = json.load(open(rep))
    assert rows and all(row["ok"] for row in rows)

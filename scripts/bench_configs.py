"""Secondary measurements for the non-headline BASELINE configs (parity-tested elsewhere; these
are NOT bench.py lines): per-kernel device-resident throughput with CUDA events, reported as
Msamples/s (input samples) and algorithmic GB/s vs the HBM peak.  Run under gpurun.

    python scripts/bench_configs.py [--quick]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import futuresdr_b200 as fb  # noqa: E402
from futuresdr_b200 import blocks as B  # noqa: E402

PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(name, n_in, bytes_alg, sec, extra=""):
    gbs = bytes_alg / sec / 1e9
    print(json.dumps({"kernel": name, "Msamples_s": round(n_in / sec / 1e6, 1), "ms": round(sec * 1e3, 4),
                      "alg_GBs": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 4), "note": extra}), flush=True)


def want(section):
    """--only a,b,c runs just the named sections (fir, f32, chain, fft, resamp, next, scale)."""
    for i, a in enumerate(sys.argv):
        if a == "--only" and i + 1 < len(sys.argv):
            return section in sys.argv[i + 1].split(",")
    return True


def main():
    quick = "--quick" in sys.argv
    n = (16 if quick else 64) * 1024 * 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.view_as_complex(torch.randn(n + 4096, 2, generator=g, device="cuda"))
    y = torch.empty(n + 4096, dtype=torch.complex64, device="cuda")
    rng = np.random.default_rng(7)
    import oracle as orc
    xr = torch.view_as_real(x).reshape(-1)[: 2 * n]
    yr = torch.view_as_real(y).reshape(-1)[: 2 * n]

    if want("fir"):
        # FIR tap sweep, direct vs tensor
        for ntaps in (8, 16, 32, 64, 128, 256):
            taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
            for algo, nm in ((fb.ALGO_DIRECT, "direct"), (fb.ALGO_TENSOR, "tensor")):
                if algo == fb.ALGO_TENSOR and ntaps < 16:
                    continue
                f = fb.FirFilter(taps, algo=algo)
                sec = timeit(lambda: f.filter(x[: n + ntaps - 1], y[:n]))
                report(f"fir_c32_{ntaps}taps_{nm}", n, 16 * n, sec)
        # 1024-tap (config 5 per-GPU kernel)
        taps = rng.uniform(-1, 1, 1024).astype(np.float32)
        f = fb.FirFilter(taps)
        nn = n // 4
        sec = timeit(lambda: f.filter(x[: nn + 1023], y[:nn]), iters=3, warm=1)
        report("fir_c32_1024taps_auto", nn, 16 * nn, sec, extra=f"algo={f.algo}")
    if want("fir1024") and not want("fir"):
        taps = rng.uniform(-1, 1, 1024).astype(np.float32)
        f = fb.FirFilter(taps)
        nn = n // 4
        sec = timeit(lambda: f.filter(x[: nn + 1023], y[:nn]), iters=3, warm=1)
        report("fir_c32_1024taps_auto", nn, 16 * nn, sec, extra=f"algo={f.algo}")
    if want("fft4096") and not want("fft"):
        fft = B.Fft(4096)
        sec = timeit(lambda: fft.transform(x[:n], y[:n]))
        report("fft_4096_fwd", n, 16 * n, sec)
    if want("demod") and not want("chain"):
        n4 = n // 4
        dem = B.Apply(B.ApplyOp.QuadDemodC32)
        z = torch.empty(n4, dtype=torch.complex64, device="cuda")
        sec = timeit(lambda: dem.apply(y[:n4], z))
        report("quad_demod_c32", n4, 16 * n4, sec)
    if want("fused"):
        # fused rows of round 2: spectrum pipe in one pass, channelizer in one launch
        for N in (2048, 4096):
            sp = B.SpectrumPipe(N, 0.1, 3)
            po = torch.empty(n // 3 + 2 * N, dtype=torch.float32, device="cuda")
            sec = timeit(lambda: sp.process(x[:n], po), iters=5, warm=2)
            report(f"spectrum_pipe_fused_{N}", n, 8 * n + 4 * (n // 3), sec, extra="FFT + |x|^2 + MovingAvg(0.1, 3): spectrum_kernel + scan + fixup")
        for N, T in ((64, 16), (16, 16), (256, 8)):
            ctaps = (orc.kaiser_lowpass(0.4 / N, 0.1 / N, 1e-3)).astype(np.float32)
            ctaps = np.resize(ctaps, N * T)
            ch = B.PfbChannelizer(N, ctaps, 1.0)
            ch.reserve_outputs(n // N + 8)
            ch.input.set(x[:n])
            ch.work(B.WorkIo())                               # window fill

            def run_ch():
                ch.input.pos, ch.produced = 0, 0
                ch.work(B.WorkIo())
            sec = timeit(run_ch, iters=5, warm=1)
            report(f"pfb_channelizer_fused_{N}ch_{T}taps", n, 16 * n, sec, extra="FIR bank + IFFT + transposed store in one launch")
    if want("synth"):
        for N, T in ((64, 16), (16, 16)):
            staps = np.resize((orc.kaiser_lowpass(0.4 / N, 0.1 / N, 1e-3)).astype(np.float32), N * T)
            syn = B.PfbSynthesizer(N, staps)
            nv = n // N
            xin = x[:N * nv].view(N, nv)
            syn.set_inputs(xin[:, :4 * T])                    # window fill
            syn.output.reserve(8 * T * N)
            syn.work(B.WorkIo())
            syn.set_inputs(xin)
            syn.output.reserve(n + 2 * N)

            def run_syn():
                syn.in_pos, syn.output.len = 0, 0
                syn.work(B.WorkIo())
            sec = timeit(run_syn, iters=5, warm=1)
            report(f"pfb_synthesizer_{N}ch_{T}taps", n, 16 * n, sec, extra="gather + IFFT + FIR bank")
    if want("f32"):
        # f32 x f32 64 taps (perf/fir config-1 kernel)
        xr = torch.view_as_real(x).reshape(-1)[: 2 * n]
        yr = torch.view_as_real(y).reshape(-1)[: 2 * n]
        t64 = rng.random(64).astype(np.float32)
        for algo, nm in ((fb.ALGO_DIRECT, "direct"), (fb.ALGO_TENSOR, "tensor")):
            f = fb.FirFilter(t64, sample_dtype=np.float32, algo=algo)
            sec = timeit(lambda: f.filter(xr, yr[: 2 * n - 63]))
            report(f"fir_f32_64taps_{nm}", 2 * n, 8 * 2 * n, sec)

    if want("chain"):
        # config 3 pieces: decimator /4 (52 taps) -> quad demod -> PfbArb 0.768
        dec = B.FirBuilder.decimating(4)
        sec = timeit(lambda: dec.filter.filter(x[:n], y[: n // 4]))
        report("decim4_52taps_c32", n, 8 * n + 8 * (n // 4), sec, extra=f"algo={dec.filter.algo}")
        dtaps = fb.firdes.kaiser.lowpass(0.25, 0.1, 1e-4)
        decd = fb.DecimatingFirFilter(4, dtaps, algo=fb.ALGO_DIRECT)
        sec = timeit(lambda: decd.filter(x[:n], y[: n // 4]))
        report("decim4_52taps_c32_direct", n, 8 * n + 8 * (n // 4), sec)
        n4 = n // 4
        dem = B.Apply(B.ApplyOp.QuadDemodC32)
        z = torch.empty(n4, dtype=torch.complex64, device="cuda")
        sec = timeit(lambda: dem.apply(y[:n4], z))
        report("quad_demod_c32", n4, 16 * n4, sec)
        demf = B.Apply(B.ApplyOp.QuadDemod)
        zf = torch.empty(n4, dtype=torch.float32, device="cuda")
        sec = timeit(lambda: demf.apply(y[:n4], zf))
        report("quad_demod_f32", n4, 12 * n4, sec)
        import oracle as orc
        ptaps = (orc.kaiser_lowpass(0.4 / 32, 0.1 / 32, 1e-3) * 32).astype(np.float32)[: 32 * 16]
        pfb = B.PfbArbResampler(0.768, ptaps, 32)
        w = torch.empty(int(n4 * 0.8) + 1024, dtype=torch.complex64, device="cuda")

        def run_pfb():
            pfb.input.set(z)
            pfb.output.data, pfb.output.len = w, 0
            io = B.WorkIo()
            pfb.work(io)
            if io.call_again:
                pfb.input.data = z
                pfb.input.pos = pfb.input.pos
                pfb.work(B.WorkIo())
        t0 = time.perf_counter()
        sec = timeit(run_pfb, iters=3, warm=1)
        report("pfbarb_0.768_32x16", n4, 8 * n4 + 8 * int(n4 * 0.768), sec, extra="includes the host timing-recurrence replay")

    if want("fft"):
        # config 4: FFT 4096
        for nfft_size in (64, 1024, 2048, 4096, 8192, 16384):
            fft = B.Fft(nfft_size)
            sec = timeit(lambda: fft.transform(x[:n], y[:n]))
            report(f"fft_{nfft_size}_fwd", n, 16 * n, sec)
        fft = B.Fft.with_options(4096, B.FftDirection.Forward, True, 1.0 / 4096)
        sec = timeit(lambda: fft.transform(x[:n], y[:n]))
        report("fft_4096_fwd_shift_norm", n, 16 * n, sec)
    if want("resamp"):
        # rational resampler 3/2 (72 taps) and 48/125
        for L, M in ((3, 2), (48, 125)):
            r = B.FirBuilder.resampling(L, M)
            cap = n // 4 * L // M + L
            sec = timeit(lambda: r.filter.filter(x[: n // 4], y[:cap]), iters=5)
            report(f"resamp_{L}_{M}_c32", n // 4, 8 * (n // 4) + 8 * ((n // 4) * L // M), sec)
    if want("next"):
        # XlatingFir, PfbChannelizer, spectrum pipe
        xl = B.XlatingFir(4, 1000.0, 48000.0)
        nx = n // 4

        def run_xl():
            c, p, st = xl.filter.filter(x[:nx], y[: nx // 4])
            xl.rotator.rotate_inplace(y[:p])
        sec = timeit(run_xl, iters=3, warm=1)
        report("xlating_fir_d4_52taps", nx, 8 * nx + 8 * (nx // 4), sec, extra="includes the host replay of the rotator recurrence")
        ctaps = (orc.kaiser_lowpass(0.4 / 64, 0.1 / 64, 1e-3)).astype(np.float32)[: 64 * 16]
        ch = B.PfbChannelizer(64, ctaps, 1.0)
        ch.reserve_outputs(n // 64 + 8)
        ch.input.set(x[:n])
        ch.work(B.WorkIo())                                   # window fill

        def run_ch():
            ch.input.pos, ch.produced = 0, 0
            ch.work(B.WorkIo())
        sec = timeit(run_ch, iters=5, warm=1)
        report("pfb_channelizer_64ch_16taps", n, 16 * n, sec, extra="FIR bank + 64-pt IFFT + transpose (3 kernels)")
        fft2 = B.Fft.with_options(2048, B.FftDirection.Forward, True, None)
        mag = B.Apply(B.ApplyOp.NormSqr)
        keep = B.MovingAvg(2048, 0.1, 3)
        pw = torch.empty(n, dtype=torch.float32, device="cuda")
        po = torch.empty(n // 3 + 4096, dtype=torch.float32, device="cuda")

        def run_spec():
            fft2.transform(x[:n], y[:n])
            mag.apply(y[:n], pw)
            keep.input.set(pw)
            keep.output.data, keep.output.len = po, 0
            keep.work(B.WorkIo())
        sec = timeit(run_spec, iters=5, warm=1)
        report("spectrum_pipe_fft2048_normsqr_mavg", n, 8 * n + 4 * (n // 3), sec, extra="unfused: 3 kernels, 32 B/sample of HBM traffic")
    if want("scale"):
        # element-wise scale (the Vulkan/wgpu shader)
        sc = B.Apply(B.ApplyOp.ScaleF32, 12.0)
        sec = timeit(lambda: sc.apply(xr, yr))
        report("scale_f32_x12", 2 * n, 8 * 2 * n, sec)


if __name__ == "__main__":
    main()

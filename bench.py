#!/usr/bin/env python
"""bench.py -- headline benchmark: Msamples/s of the Complex<f32> 256-tap FIR on 64 Mi-sample
chunks (BASELINE.json configs[1]) at N GPUs, with the HBM roofline fraction of the dominant
kernel and the reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the FIR over one 64 Mi-sample chunk per GPU (weak scaling: each rank
owns the next contiguous chunk of one logical stream; the only exchange is the NCCL all-gather
of the (ntaps-1)-sample overlap region, futuresdr_b200/shard.py).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NTAPS = 256
CHUNK = 64 * 1024 * 1024          # samples per GPU per step (BASELINE configs[1])
BYTES_PER_SAMPLE = 16             # 8 B in + 8 B out (taps/halo amortise to 0)
SEED = 0x5EED
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback


def _peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _taps():
    return np.random.default_rng(7).uniform(-1, 1, NTAPS).astype(np.float32)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML polled every ~2 ms from a
    background thread; falls back to one nvidia-smi query if NVML is unavailable)."""

    def __init__(self, index=0):
        self.index, self.thread, self.stop_flag = index, None, False
        self.sm, self.reasons, self.max_sm = [], set(), None

    def _poll(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(1.0)
        if self.sm:
            return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        try:   # fallback: a single nvidia-smi reading right after the region
            q = "clocks.sm,clocks.max.sm"
            o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                str(self.index)], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(o[0]), "sm_max_mhz": float(o[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi after region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port of
# crates/futuredsp/src/fir.rs:52-91; the Rust cannot be compiled here), all host threads.
# ------------------------------------------------------------------------------------------
_CPU_THREADS = None


def _cpu_threads():
    """Thread count for the CPU legs: all the host threads this process may use.  OpenMP's default
    can exceed the container's CPU allowance (oversubscription made the 128-thread run 8x slower
    than the 64-thread one), so a few candidates are timed on a small sample and the best is kept."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import oracle as orc
    cand = {orc.max_threads()}
    try:
        cand.add(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            cand.add(max(1, int(int(q[0]) / int(q[1]))))
    except Exception:
        pass
    m = max(cand)
    cand |= {max(1, m // 2), max(1, m // 4), min(m, 32), min(m, 16)}
    rng = np.random.default_rng(1)
    n = 1 << 22
    x = (rng.standard_normal(n + NTAPS - 1) + 1j * rng.standard_normal(n + NTAPS - 1)).astype(np.complex64)
    out = np.zeros(n, np.complex64)
    taps = _taps()
    best, best_t = None, None
    for th in sorted(cand):
        dt = None
        for _ in range(3):                      # first pass warms the thread pool; keep the best
            t0 = time.perf_counter()
            orc.fir_c32_f32_mt(taps, x, th, fast=True, out=out)
            d1 = time.perf_counter() - t0
            dt = d1 if dt is None else min(dt, d1)
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    _CPU_THREADS = best
    return best


def cpu_reference(sample_items: int, reps: int = 1):
    import oracle as orc
    threads = _cpu_threads()
    rng = np.random.default_rng(SEED)
    x = (rng.standard_normal(sample_items + NTAPS - 1) + 1j * rng.standard_normal(sample_items + NTAPS - 1)
         ).astype(np.complex64)
    taps = _taps()
    out = np.empty(sample_items, np.complex64)
    best = {}
    for fast in (False, True):
        orc.fir_c32_f32_mt(taps, x[: 65536 + NTAPS - 1], threads, fast=fast)     # warm
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.fir_c32_f32_mt(taps, x, threads, fast=fast, out=out)
            ts.append(time.perf_counter() - t0)
        best["nightly_reassoc" if fast else "stable_strict"] = min(ts)
    variant = min(best, key=best.get)
    return {"seconds": best[variant], "variant": variant, "threads": threads, "all": best,
            "msps": sample_items / best[variant] / 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample of the same workload: 256-tap c32 FIR on white noise
    n = 8 * 1024 * 1024
    times, last = [], None
    for i in range(args.warmup + args.steps):
        r = cpu_reference(n, reps=1)
        last = r
        if i >= args.warmup:
            times.append(r["seconds"])
    sec = statistics.mean(times)
    v = n / sec / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic white noise",
        "config": {"workload": "c32 256-tap FIR (BASELINE configs[1]), CPU sample of 8 Mi samples/step",
                   "ntaps": NTAPS, "sample_items": n},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": last["threads"], "kind": "port",
                         "sample": f"{n} samples/step, oracle port of futuredsp fir.rs:52-91 ({last['variant']}), "
                                   f"{last['threads']} OpenMP threads over contiguous shards"},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import futuresdr_b200 as fb
    from futuresdr_b200.shard import ShardedFir

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    taps = _taps()
    algo = {"auto": fb.ALGO_AUTO, "direct": fb.ALGO_DIRECT, "tensor": fb.ALGO_TENSOR}[args.algo]
    sh = ShardedFir(taps, CHUNK, np.complex64, device=dev, algo=algo, exchange=args.exchange)
    ctx = sh._filter.ctx
    # synthetic white noise generated on the device (Philox), per-rank subsequence; every ring slot is filled
    # before the timed region (inputs resident in HBM when it starts)
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    if sh.exchange == "peer":
        for tns in sh.slot_tensors():
            torch.view_as_real(tns).normal_(generator=g)
    else:
        torch.view_as_real(sh.chunk).normal_(generator=g)
    out = torch.empty(CHUNK, dtype=torch.complex64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): inputs already in HBM
    for _ in range(args.warmup):
        sh.step(out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    produced = 0
    for i in range(args.steps):
        # the FIR kernel alone (for the roofline): events on the launching stream around the launch
        if sh.exchange == "peer":
            sh.on_kernel = lambda tag, _i=i: kev[_i][0 if tag == "begin" else 1].record()
            c, p, st = sh.step(out)
            sh.on_kernel = None
        else:
            orig = sh.compute

            def timed(src, o, _i=i, _f=orig):
                kev[_i][0].record()
                r = _f(src, o)
                kev[_i][1].record()
                return r
            sh.compute = timed
            c, p, st = sh.step(out)
            sh.compute = orig
        produced += p
        ev[i + 1].record()
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = (CHUNK * world * args.steps) / (total_ms * 1e-3) / 1e6

    # ---- end to end through the C-ABI host-slice call (`e2e`): pinned host in/out, H2D + D2H timed
    fir = sh._filter
    n_e2e = CHUNK
    h_in = torch.empty(n_e2e + NTAPS - 1, dtype=torch.complex64).pin_memory()
    torch.view_as_real(h_in).normal_(generator=torch.Generator().manual_seed(SEED + 100 + rank))
    h_out = torch.empty(n_e2e, dtype=torch.complex64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        fir.filter(h_in, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        c, p, st = fir.filter(h_in, h_out)          # returns when h_out is filled
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = (n_e2e * world * e2e_steps) / float(te.item()) / 1e6
    # parity spot check of the e2e output against the device path is done in tests/; here we
    # only make sure the result was produced
    assert p == n_e2e

    if rank == 0:
        peak, peak_src = _peak_hbm()
        k_ms = statistics.mean(kern_ms)
        achieved = BYTES_PER_SAMPLE * CHUNK / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_fir.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get({1: "direct", 2: "tensor"}.get(fir.algo, ""), None)
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            r = cpu_reference(4 * 1024 * 1024, reps=3)
            cpu = {"value": r["msps"], "unit": "Msamples/s", "cores": r["threads"], "kind": "port",
                   "sample": f"4 Mi samples x3 (best), same taps/noise family; oracle port of futuredsp "
                             f"fir.rs:52-91, variant {r['variant']} (strict {4*1024*1024/r['all']['stable_strict']/1e6:.1f} / "
                             f"reassoc {4*1024*1024/r['all']['nightly_reassoc']/1e6:.1f} Msamples/s)"}
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic white noise (device Philox, per-rank subsequence)",
            "config": {"workload": "single-B200 Complex<f32> 256-tap FIR on 64 Mi-sample chunks via device-resident ring (BASELINE configs[1])",
                       "ntaps": NTAPS, "chunk_items": CHUNK, "algo": {1: "direct", 2: "tensor"}.get(fir.algo),
                       "l2": "inputs larger than L2 (512 MiB in + 512 MiB out per step)",
                       "sharding": "contiguous time ranges, NCCL all-gather of the 255-sample overlap" if world > 1 else "none"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * CHUNK},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": (n_e2e + NTAPS - 1) * 8,
                    "d2h_bytes_per_step": n_e2e * 8, "steps": e2e_steps,
                    "api": "FirFilter.filter(host_in, host_out) -> b2s_fir_filter_host (pinned host, chunked H2D/kernel/D2H pipeline)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        try:
            # explanatory only: the FLOPs the tensor kernel EXECUTES (3 split-bf16 products over the block-Toeplitz
            # operand, K = ntaps + 127 rounded up to 16) against the dense bf16 tensor peak -- at 256 taps this, not
            # HBM, is what the kernel runs into.  The judged roofline above stays the HBM one
            # SURVEY.md 8(d) states for this metric.
            if fir.algo == 2:
                ksteps = -(-(NTAPS + 127) // 16)
                tiles = -(-CHUNK // 8192)
                mma_flops = float(tiles) * 3 * ksteps * 2 * 128 * 128 * 16
                tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s burst)"
                try:
                    mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                    if "bf16_tflops" in mp:
                        tpeak, tsrc = float(mp["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops)"
                except Exception:
                    pass
                tach = mma_flops / (k_ms * 1e-3) / 1e12
                line["roofline"]["tensor_executed"] = {"achieved": tach, "peak": tpeak, "unit": "TFLOP/s",
                                                       "frac": tach / tpeak, "flops_per_launch": mma_flops,
                                                       "peak_source": tsrc}
        except Exception:
            pass
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "direct", "tensor"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="halo exchange of the sharded stream: in-kernel peer fetch over NVLink (default) or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Bring-up probe for the tcgen05 FIR: runs the tensor path with each convention switch and
prints error statistics against the direct CUDA-core path and the oracle (run under gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futuresdr_b200 as fb  # noqa: E402
import oracle as orc  # noqa: E402

rng = np.random.default_rng(3)
for flags in (0, 1):
    os.environ["B2S_TC_FLAGS"] = str(flags)
    for cplx, ntaps, n in ((True, 256, 20000), (True, 5, 9000), (True, 200, 300000), (False, 256, 40000)):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) if cplx \
            else rng.standard_normal(n).astype(np.float32)
        taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
        try:
            f = fb.FirFilter(taps, sample_dtype=x.dtype, algo=fb.ALGO_TENSOR)
        except Exception as e:
            print("plan failed", e)
            continue
        xd = torch.from_numpy(x).cuda()
        yd = torch.zeros(n, dtype=xd.dtype, device="cuda")
        c, p, st = f.filter(xd, yd)
        torch.cuda.synchronize()
        _, _, _, ref = orc.fir(taps, x, n)
        got = yd[:p].cpu().numpy()
        scale = float(np.sum(np.abs(taps)) * np.max(np.abs(x)))
        err = np.abs(got - ref)
        bad = np.nonzero(err > 1e-5 * scale)[0]
        print(f"flags={flags} cplx={cplx} ntaps={ntaps} n={n} algo={f.algo} produced={p} "
              f"max_err/scale={err.max() / scale:.3e} n_bad={bad.size} first_bad={bad[:8].tolist()}")
        if bad.size and flags == 0:
            k = int(bad[0])
            print("   got", got[k:k + 3], "ref", ref[k:k + 3])

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import futuresdr_b200 as fb
from futuresdr_b200 import blocks as B
import oracle as orc
n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
ptaps = (orc.kaiser_lowpass(0.4 / 32, 0.1 / 32, 1e-3) * 32).astype(np.float32)[: 32 * 16]
print("ntaps", ptaps.size)
pfb = B.PfbArbResampler(0.768, ptaps, 32)
z = torch.randn(n4, dtype=torch.float32, device="cuda").to(torch.complex64)
w = torch.empty(int(n4 * 0.8) + 1024, dtype=torch.complex64, device="cuda")
for it in range(5):
    pfb.input.set(z)
    pfb.output.data, pfb.output.len = w, 0
    io = B.WorkIo()
    pfb.work(io)
    if io.call_again:
        pfb.work(B.WorkIo())
    torch.cuda.synchronize()
    print("iter", it, "consumed", pfb.input.pos, "produced", pfb.output.len, flush=True)

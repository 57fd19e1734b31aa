"""N>1 parity on real GPUs (run under torchrun, one rank per GPU, NCCL): the sharded FIR stream
(futuresdr_b200.shard.ShardedFir: contiguous time ranges + all-gather of the overlap region) must
equal the single-stream oracle result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/shard_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futuresdr_b200 as fb  # noqa: E402
from futuresdr_b200.shard import ShardedFir  # noqa: E402
import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for ntaps, decim, S, steps in ((256, 1, 1 << 16, 3), (1024, 1, 1 << 15, 2), (52, 4, 1 << 14, 3), (64, 1, 8192, 4)):
        rng = np.random.default_rng(123)
        total = world * S * steps
        x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)).astype(np.complex64)
        taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
        sh = ShardedFir(taps, S, np.complex64, decim=decim, device=dev)
        outs = []
        for t in range(steps):
            lo = (t * world + rank) * S
            sh.chunk.copy_(torch.from_numpy(x[lo:lo + S]).to(dev))
            out = torch.zeros(S // decim, dtype=torch.complex64, device=dev)
            c, p, st = sh.step(out)
            torch.cuda.synchronize()
            outs.append((t * world + rank, out[:p].cpu().numpy()))
        gathered = [None] * world
        dist.all_gather_object(gathered, outs)
        if rank == 0:
            pieces = sorted([pc for g in gathered for pc in g], key=lambda a: a[0])
            got = np.concatenate([p for _, p in pieces])
            _, _, _, ref = orc.decim_fir(taps, decim, x, total)
            tol = 1e-5 * float(np.sum(np.abs(taps))) * float(np.max(np.abs(x)))
            err = float(np.max(np.abs(got - ref))) if got.size == ref.size else float("inf")
            good = got.size == ref.size and err <= tol
            ok = ok and good
            print(f"world={world} ntaps={ntaps} decim={decim} S={S} steps={steps} algo={sh._filter.algo} "
                  f"n_out={got.size} (ref {ref.size}) max_err={err:.3e} tol={tol:.3e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

"""N>1 parity on real GPUs (run under torchrun, one rank per GPU): the sharded FIR stream
(futuresdr_b200.shard.ShardedFir: contiguous time ranges; the left neighbour's tail either fetched by the FIR
kernel itself over NVLink -- exchange="peer", CUDA IPC + device flags -- or all-gathered with NCCL --
exchange="nccl") must equal the single-stream oracle result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/shard_parity.py [--json out.json]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futuresdr_b200 as fb  # noqa: E402
from futuresdr_b200.shard import ShardedFir  # noqa: E402
import oracle as orc  # noqa: E402

CASES = (  # ntaps, decim, S, steps, sample dtype
    (256, 1, 1 << 16, 4, np.complex64), (1024, 1, 1 << 15, 3, np.complex64), (52, 4, 1 << 14, 4, np.complex64),
    (64, 1, 8192, 5, np.complex64), (129, 1, 1 << 15, 3, np.complex64), (33, 2, 1 << 13, 3, np.complex64),
    (257, 1, 1 << 15, 3, np.complex64),
    (64, 1, 1 << 14, 4, np.float32), (100, 4, 1 << 14, 3, np.float32), (255, 1, 1 << 15, 3, np.float32),
)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(local % ndev)
    dev = torch.device("cuda", local % ndev)
    dist.init_process_group("nccl", device_id=dev)
    ok, report = True, []
    for exchange in ("peer", "nccl"):
        for ntaps, decim, S, steps, sdt in CASES:
            rng = np.random.default_rng(123)
            total = world * S * steps
            if sdt == np.complex64:
                x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)).astype(np.complex64)
            else:
                x = rng.standard_normal(total).astype(np.float32)
            tdt = torch.complex64 if sdt == np.complex64 else torch.float32
            taps = rng.uniform(-1, 1, ntaps).astype(np.float32)
            sh = ShardedFir(taps, S, sdt, decim=decim, device=dev, exchange=exchange)
            outs = []
            for t in range(steps):
                lo = (t * world + rank) * S
                sh.chunk.copy_(torch.from_numpy(x[lo:lo + S]).to(dev))
                out = torch.zeros(S // decim, dtype=tdt, device=dev)
                c, p, st = sh.step(out)
                # no synchronisation between steps in peer mode: ordering is the device flags' job
                outs.append((t * world + rank, out, p))
            sh._filter.ctx.sync()
            outs = [(i, o[:p].cpu().numpy()) for i, o, p in outs]
            if hasattr(sh, "close"):
                sh.close()
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
                row = {"world": world, "exchange": exchange, "samples": np.dtype(sdt).name, "ntaps": ntaps, "decim": decim, "chunk": S, "steps": steps,
                       "algo": int(sh._filter.algo), "n_out": int(got.size), "n_ref": int(ref.size), "max_err": err,
                       "tol": tol, "ok": bool(good)}
                report.append(row)
                print(json.dumps(row), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        if "--json" in sys.argv:
            json.dump(report, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
        if not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()

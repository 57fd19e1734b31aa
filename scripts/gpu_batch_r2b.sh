# round-2 batch B (2 GPUs): cross-process peer exchange -- parity and scaling
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
echo "--- shard parity world 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/shard_parity.py --json gpurun_out/shard_parity_w2.json 2>&1 | grep -v "^W\|^\*\*\*" | tail -20
echo "--- bench n1"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-330
echo "--- bench n2 peer"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r2b_n2_peer.json; cut -c1-330 gpurun_out/bench_r2b_n2_peer.json
echo "--- bench n2 nccl"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl 2>&1 | tail -1 > gpurun_out/bench_r2b_n2_nccl.json; cut -c1-330 gpurun_out/bench_r2b_n2_nccl.json

# multi-GPU evidence (run under gpurun --gpus N): shard parity + bench at every power of two up to N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "--- single-GPU tests added since the last batch"
timeout 600 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_fir.py -x -q 2>&1 | tail -4
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "--- shard parity world=$n"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/shard_parity.py 2>&1 | grep -E "world=|Error|error|FAIL" | head -12
  fi
done
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    echo "--- bench gpus=$n"
    if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_r1_n$n.json
    else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_r1_n$n.json; fi
    python -c "import json,sys; d=json.load(open('gpurun_out/scale_r1_n$n.json')); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['roofline']['frac'], round(d['e2e']['value']), d['clocks'])"
  fi
done

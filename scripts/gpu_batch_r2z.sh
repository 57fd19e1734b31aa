# round-2 batch Z (2 GPUs): the final tree exactly as the driver launches it at N = 2 (own arm with default flags, reference
# arm), sharded parity at world size 2, the 2-GPU tests of the suite
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/shard_parity.py --json gpurun_out/shard_parity_final_w2.json 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_final_n2.err | tail -1 > gpurun_out/bench_final_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final_n2.json'))
print('n2 value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'per_rank',d.get('per_rank'))
for s in d.get('secondary',[]): print(s['config']['workload'][:60], s.get('value'), s.get('error'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x -k "shard or peer or exec_hist" 2>&1 | tail -2

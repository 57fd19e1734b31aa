# 2-GPU check of the sharded path: parity against the oracle, then the bench line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_parity.py 2>&1 | grep -E "world=|Error|error|FAIL|Traceback" | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_r1_n2.json
python -c "import json; d=json.load(open('gpurun_out/scale_r1_n2.json')); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['roofline']['frac'], round(d['e2e']['value']), d['clocks'])"

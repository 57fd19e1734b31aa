# round-2 batch E (8 GPUs): sharded parity at world 8 through pytest, scaling bench 1/2/4/8 with the driver's flags
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader | head -8
echo "--- pytest sharded parity at the largest world size"
timeout 900 python -m pytest tests/test_gpu_exec_hist.py -q -k "sharded_parity" 2>&1 | tail -5
for N in 1 2 4 8; do
  echo "--- bench N=$N (driver flags: --steps 20 --warmup 5)"
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err
  fi
  tail -2 gpurun_out/scale_r2_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_r2_n$N.json').read().strip().splitlines()[-1])
    print('N=$N value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'e2e',round(d['e2e']['value']),'sustained',d['sustained'] and round(d['sustained']['ms_per_step'],4))
    print('  per_rank',d.get('per_rank'))
    print('  ceiling',d['e2e']['host_copy_ceiling'])
    for s in d.get('secondary') or []:
        print('  sec:',json.dumps(s)[:600])
except Exception as e:
    print('parse failed',e)
PY
done
echo "--- nccl-mode comparison at N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29790 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-secondary --no-sustained 2>/dev/null | tail -1 > gpurun_out/scale_r2_n8_nccl.json; python -c "import json;d=json.load(open('gpurun_out/scale_r2_n8_nccl.json'));print('nccl N=8',d['ms_per_step'],d['value'])"

# round-2 batch X (1 GPU): bench line with the parallel-pipes CUDA graph of configs[0]
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; tail -3 gpurun_out/bench_r2_final_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
s=d['secondary'][0]
print({k:v for k,v in s.items() if k in ('value','ms_per_pass','cuda_graph','cuda_graph_parallel_pipes','roofline')})
PY

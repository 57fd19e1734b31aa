# round-2 batch K (1 GPU): suite + bench after the small-slice switch
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2k_n1.json 2> gpurun_out/bench_r2k_n1.err; tail -3 gpurun_out/bench_r2k_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2k_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'sustained',d['sustained']['ms_per_step'],d['sustained']['roofline_frac_step'])
for s in d['secondary']:
    print(s['config']['workload'][:60], '| value', s.get('value'), '| frac', (s.get('roofline') or {}).get('frac'), '| e2e', (s.get('e2e') or {}).get('value'), '|', (s.get('fused_spectrum_pipe') or {}).get('value'), s.get('cuda_graph'), s.get('error'))
PY
echo "--- ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400

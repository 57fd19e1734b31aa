# round-2 closing check (1 GPU): the committed tree -- full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r2_closing_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_closing_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
for s in d.get('secondary',[]): print(' ', s['config']['workload'][:50], s.get('value'), s.get('error'))
PY

# round-2 batch I (1 GPU): suite after the last changes, channelizer row-stride A/B, bench line
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "--- configs"; timeout 900 python scripts/bench_configs.py --only fused > gpurun_out/bench_configs_r2_fused.jsonl 2>&1; tail -6 gpurun_out/bench_configs_r2_fused.jsonl | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chan_fused -s 1 -c 1 -o gpurun_out/prof_r2_chan python scripts/bench_configs.py --only fused > /dev/null 2>&1
echo "--- bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2i_n1.json 2> gpurun_out/bench_r2i_n1.err; tail -3 gpurun_out/bench_r2i_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'sustained',d['sustained']['ms_per_step'],d['sustained']['roofline_frac_step'])
for s in d['secondary']:
    print(s['config']['workload'][:60], '| value', s.get('value'), '| frac', (s.get('roofline') or {}).get('frac'), '| e2e', (s.get('e2e') or {}).get('value'), '|', (s.get('fused_spectrum_pipe') or {}).get('value'), s.get('error'))
PY
echo "--- small sweep"; timeout 600 python scripts/small_sweep.py 2>&1 | tail -10

# round-2 batch F (1 GPU): re-verify after the scan rewrite / FFT occupancy fix; sanitizer passes
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "--- configs"; timeout 900 python scripts/bench_configs.py --only fused,fft,chain > gpurun_out/bench_configs_r2.jsonl 2>&1; tail -22 gpurun_out/bench_configs_r2.jsonl | cut -c1-220
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2_fused.csv python scripts/bench_configs.py --only fused > /dev/null 2>&1; grep -E "spectrum" gpurun_out/launches_r2_fused.csv | awk -F'","' '{print $5, $NF}' | cut -c1-80 | head -6
echo "--- bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err; tail -3 gpurun_out/bench_r2f_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for s in d['secondary']:
    print(s['config']['workload'][:60], '| value', s.get('value'), '| frac', (s.get('roofline') or {}).get('frac'), '| cpu', (s.get('cpu_baseline') or {}).get('value'), '| e2e', (s.get('e2e') or {}).get('value'), '|', {k: v for k, v in s.items() if k in ('parity_spot_check', 'cuda_graph', 'error')}, '|', (s.get('fused_spectrum_pipe') or {}).get('value'))
PY
echo "--- sanitizer memcheck (selected tests)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_exec_hist.py tests/test_gpu_spectrum.py tests/test_gpu_channelizer.py -q -x -k "not sharded and not timeout and not fused_shapes" 2>&1 | tail -4
echo "--- sanitizer racecheck (selected tests)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_channelizer.py tests/test_gpu_blocks.py -q -x -k "fused_spectrum_pipe_vs_oracle_chain or channelizer_parity or pfbarb_parity" 2>&1 | tail -4

# round-2 batch H (1 GPU): full suite with the four-step FFT tests, final N=1 bench line, profile of pfb_kernel
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "--- bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err; tail -3 gpurun_out/bench_r2h_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'sustained',d['sustained']['ms_per_step'],d['sustained']['roofline_frac_step'])
for s in d['secondary']:
    print(s['config']['workload'][:60], '| value', s.get('value'), '| frac', (s.get('roofline') or {}).get('frac'), '| cpu', (s.get('cpu_baseline') or {}).get('value'), '| e2e', (s.get('e2e') or {}).get('value'), '|', (s.get('fused_spectrum_pipe') or {}).get('value'), s.get('error'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfb_kernel -s 1 -c 1 -o gpurun_out/prof_r2_pfbarb python scripts/bench_configs.py --only chain > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chan_fused -s 1 -c 1 -o gpurun_out/prof_r2_chan python scripts/bench_configs.py --only fused > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 1 -c 1 -o gpurun_out/prof_r2_spectrum python scripts/bench_configs.py --only fused > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_fft_kernel -s 1 -c 1 -o gpurun_out/prof_r2_firfft python scripts/bench_configs.py --only fir1024 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 1 -c 1 -o gpurun_out/prof_r2_fft4096 python scripts/bench_configs.py --only fft4096 > /dev/null 2>&1
echo "--- fft sizes"; python scripts/bench_configs.py --only fft 2>&1 | tail -8 | cut -c1-150
ls -la gpurun_out/*.ncu-rep | tail -8

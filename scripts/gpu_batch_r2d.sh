# round-2 batch D (1 GPU): whole GPU suite, full bench line with the secondary configs, per-kernel numbers, ncu captures
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "--- bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2d_n1.json 2> gpurun_out/bench_r2d_n1.err; tail -3 gpurun_out/bench_r2d_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'ceil',d['e2e']['host_copy_ceiling'])
print('sustained',d['sustained'])
for s in d['secondary']:
    print(json.dumps(s)[:900])
PY
echo "--- ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700
echo "--- configs"; timeout 900 python scripts/bench_configs.py --only fused,chain,next,fir1024 > gpurun_out/bench_configs_r2.jsonl 2>&1; tail -30 gpurun_out/bench_configs_r2.jsonl | cut -c1-220
NCU="timeout 600 ncu --set full --clock-control none --import-source on"
$NCU -k regex:spectrum_kernel -s 1 -c 1 -o gpurun_out/prof_r2_spectrum python scripts/bench_configs.py --only fused > /dev/null 2>&1
$NCU -k regex:chan_fused -s 1 -c 1 -o gpurun_out/prof_r2_chan python scripts/bench_configs.py --only fused > /dev/null 2>&1
$NCU -k regex:pfb_kernel -s 1 -c 1 -o gpurun_out/prof_r2_pfbarb python scripts/bench_configs.py --only chain > /dev/null 2>&1
$NCU -k regex:fir_fft_kernel -s 1 -c 1 -o gpurun_out/prof_r2_firfft python scripts/bench_configs.py --only fir1024 > /dev/null 2>&1
$NCU -k regex:fft_kernel -s 1 -c 1 -o gpurun_out/prof_r2_fft4096 python scripts/bench_configs.py --only fft4096 > /dev/null 2>&1
$NCU -k regex:apply_kernel -s 1 -c 1 -o gpurun_out/prof_r2_demod python scripts/bench_configs.py --only demod > /dev/null 2>&1
$NCU -k regex:fir_tc_kernel -s 3 -c 1 -o gpurun_out/prof_r2_tc python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --no-sustained > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary --no-sustained > /dev/null 2>&1
ls -la gpurun_out | tail -14
echo "--- overlap-save occupancy A/B"; for mb in 1 2 3; do B2S_FFTFIR_MINB=$mb python scripts/bench_configs.py --only fir1024 2>&1 | tail -1 | cut -c1-150; done
echo "--- fft sizes"; python scripts/bench_configs.py --only fft 2>&1 | tail -8 | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_fused.csv python scripts/bench_configs.py --only fused > /dev/null 2>&1; grep -E "spectrum|chan_fused" gpurun_out/launches_r2_fused.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12

# round-1 evidence batch (run under gpurun on one B200)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "--- bench"; timeout 400 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json; cut -c1-1800 gpurun_out/bench_r1_n1.json
echo "--- bench long (clocks under sustained load)"; timeout 400 python bench.py --steps 4000 --warmup 20 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_n1_long.json; python -c "import json; d=json.load(open('gpurun_out/bench_r1_n1_long.json')); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
echo "--- ref arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-600
echo "--- configs"; timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs_r1.jsonl 2>&1; tail -45 gpurun_out/bench_configs_r1.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc -s 2 -c 1 -o gpurun_out/prof_tc_r1_final python bench.py --steps 3 --warmup 3 --algo tensor --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_direct -s 2 -c 1 -o gpurun_out/prof_decim_r1_final python scripts/bench_configs.py --only chain > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resamp_slide -s 2 -c 1 -o gpurun_out/prof_resamp_r1_final python scripts/bench_configs.py --only resamp > /dev/null 2>&1
ls -la gpurun_out | tail -10

# round-1 evidence batch (run under gpurun on one B200)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "--- bench"; timeout 400 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json; cut -c1-1800 gpurun_out/bench_r1_n1.json
echo "--- ref arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-600
echo "--- configs"; timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs_r1.jsonl 2>&1; tail -40 gpurun_out/bench_configs_r1.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_direct -s 2 -c 1 -o gpurun_out/prof_direct_r1 python bench.py --steps 3 --warmup 3 --algo direct --no-cpu > /dev/null 2>&1

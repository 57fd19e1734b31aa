./tests/cpp/test_host 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_cpp_host.py tests/test_gpu_chain.py tests/test_gpu_edges.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json; cut -c1-400 gpurun_out/bench_r1_n1.json
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs_r1.jsonl 2>&1; grep -E "decim|resamp|demod|pfbarb|chan|spectrum|256taps_tensor" gpurun_out/bench_configs_r1.jsonl | cut -c1-190
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc -s 2 -c 1 -o gpurun_out/prof_decim_tc_r1_final python scripts/bench_configs.py --only chain > /dev/null 2>&1

timeout 1200 python -m pytest tests/test_gpu_channelizer.py tests/test_gpu_spectrum.py tests/test_gpu_synthesizer.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py --only next 2>&1 | tail -4 | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_next.csv python scripts/bench_configs.py --only next --quick > /dev/null 2>&1

timeout 600 python -m pytest tests/test_gpu_channelizer.py tests/test_gpu_synthesizer.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py --only next 2>&1 | tail -3 | cut -c1-200

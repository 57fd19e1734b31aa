# round-2 batch O (1 GPU): fused synthesizer
echo "--- tests"; timeout 900 python -m pytest tests/test_gpu_synthesizer.py tests/test_golden_fixtures.py tests/test_gpu_cpp_host.py -q -x 2>&1 | tail -5
echo "--- memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_synthesizer.py -q -x 2>&1 | tail -3
echo "--- perf fused"; python scripts/bench_configs.py --only synth 2>&1 | tail -3 | cut -c1-160
echo "--- perf generic"; B2S_SYNTH_NO_FUSED=1 python scripts/bench_configs.py --only synth 2>&1 | tail -3 | cut -c1-160

# round-2 batch Q (1 GPU): packed-FP32 instruction rates; FFT-family kernels with FADD2 butterflies
mkdir -p gpurun_out
./scripts/microbench/f32x2_rate > gpurun_out/f32x2_rate.jsonl 2>&1; cat gpurun_out/f32x2_rate.jsonl
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_spectrum.py tests/test_gpu_channelizer.py tests/test_gpu_synthesizer.py tests/test_gpu_fir_fft.py tests/test_gpu_cpp_host.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py --only fir1024,fft,fused,synth > gpurun_out/bench_configs_r2q.jsonl 2>&1; cut -c1-150 gpurun_out/bench_configs_r2q.jsonl

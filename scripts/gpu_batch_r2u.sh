# round-2 batch U (1 GPU): twiddles fetched one pass ahead (overlap-save FIR, fused spectrum / channelizer / synthesizer)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_spectrum.py tests/test_gpu_channelizer.py tests/test_gpu_synthesizer.py tests/test_gpu_fir_fft.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py --only fir1024,fused,synth,resamp > gpurun_out/bench_configs_r2u.jsonl 2>&1; cut -c1-150 gpurun_out/bench_configs_r2u.jsonl

# round-2 batch Y (1 GPU): persistent overlap-save FIR (H in smem, next block's input by cp.async) against one block per CTA
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fir_fft.py tests/test_gpu_blocks.py -m gpu -x -q -k "fir_fft or resampler_parity" 2>&1 | tail -2
for pz in 1 0; do echo "== persist=$pz"; B2S_FFTFIR_PERSIST=$pz timeout 300 python scripts/bench_configs.py --only fir1024 2>&1 | cut -c1-130; done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_fir_fft.py -q -x -k "not full and not large" 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_fir_fft.py -q -x -k "not full and not large" 2>&1 | tail -3

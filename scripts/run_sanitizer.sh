# memcheck of the kernels touched this round on small shapes (sanitizer slows kernels ~50x)
export B2S_SANITIZE=1
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_fir_tensor.py -m gpu -q -x -k "parity and (256 or 129 or 17) or unaligned" 2>&1 | tail -8
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fir.py -m gpu -q -x -k "resampler or decim" 2>&1 | tail -8
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_channelizer.py -m gpu -q -x -k "not tone" 2>&1 | tail -8

# round-2 batch M (1 GPU): FFT 8192 / 16384 thread-count A/B
for v in "256 512" "256 1024" "512 512" "512 1024"; do set -- $v; echo "--- 8K threads $1, 16K threads $2"; B2S_FFT8K_THREADS=$1 B2S_FFT16K_THREADS=$2 python scripts/bench_configs.py --only fft 2>&1 | grep -E "fft_8192|fft_16384" | cut -c1-120; done
echo "--- tests"; python -m pytest tests/test_gpu_blocks.py -q -x -k "fft" 2>&1 | tail -2

# round-2 batch V (1 GPU): PfbArbResampler with the planar tap table (bank-conflict-free tap loads)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_chain.py -m gpu -x -q -k "pfb or Pfb or chain or arb" 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py --only chain 2>&1 | cut -c1-170
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfb_kernel -s 1 -c 1 -o gpurun_out/prof_r2_pfbarb2 python scripts/bench_configs.py --only chain > /dev/null 2>&1
ls -la gpurun_out/prof_r2_pfbarb2.ncu-rep

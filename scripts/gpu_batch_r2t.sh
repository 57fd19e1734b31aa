# round-2 batch T (1 GPU): rational resampler, all L banks per segment load (A/B against one bank at a time) + ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_chain.py -m gpu -x -q -k "resamp or Resamp or polyphase" 2>&1 | tail -3
echo "== banks"; timeout 300 python scripts/bench_configs.py --only resamp 2>&1 | cut -c1-130
echo "== one bank at a time"; B2S_RESAMP_NO_BANKS=1 timeout 300 python scripts/bench_configs.py --only resamp 2>&1 | cut -c1-130
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resamp_slide -s 1 -c 1 -o gpurun_out/prof_r2_resamp python scripts/bench_configs.py --only resamp > /dev/null 2>&1
ls -la gpurun_out/prof_r2_resamp.ncu-rep

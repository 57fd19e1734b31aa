# round-2 batch R (1 GPU): packed complex x real MACs (FFMA2) in the sliding-window FIR / decimator / resampler and the
# fused PFB kernels; resampler output staging in final order
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fir.py tests/test_gpu_blocks.py tests/test_gpu_channelizer.py tests/test_gpu_synthesizer.py tests/test_gpu_chain.py tests/test_gpu_xlating.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py --only fir,chain,resamp,fused,synth > gpurun_out/bench_configs_r2r.jsonl 2>&1; cut -c1-150 gpurun_out/bench_configs_r2r.jsonl

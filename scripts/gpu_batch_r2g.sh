# round-2 batch G (1 GPU): spectrum scan/fixup rework, pfbarb fix
mkdir -p gpurun_out
echo "--- tests"; timeout 1200 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_blocks.py tests/test_gpu_chain.py tests/test_gpu_cpp_host.py -q -x 2>&1 | tail -4
echo "--- sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/dbg_pfb.py 2097152 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_spectrum.py -q -x 2>&1 | tail -3
echo "--- configs"; timeout 900 python scripts/bench_configs.py --only fused,chain,next > gpurun_out/bench_configs_r2.jsonl 2>&1; tail -16 gpurun_out/bench_configs_r2.jsonl | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2_fused.csv python scripts/bench_configs.py --only fused > /dev/null 2>&1; grep -E "spectrum" gpurun_out/launches_r2_fused.csv | awk -F'","' '{n=split($5,a,"("); print a[1], $NF}' | head -6

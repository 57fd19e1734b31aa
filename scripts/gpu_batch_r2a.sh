# round-2 batch A: exec_hist / misaligned tensor path / peer ring at world 1, whole GPU suite, N=1 bench
mkdir -p gpurun_out
echo "--- new tests"; timeout 600 python -m pytest tests/test_gpu_exec_hist.py -x -q 2>&1 | tail -15
echo "--- suite"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_exec_hist.py 2>&1 | tail -6
echo "--- bench n1"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r2a_n1.json; cut -c1-1500 gpurun_out/bench_r2a_n1.json
echo "--- bench n1 nccl-mode (old path)"; timeout 400 python bench.py --steps 20 --warmup 5 --exchange nccl --no-cpu 2>&1 | tail -1 | cut -c1-400

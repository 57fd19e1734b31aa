python -m pytest tests/test_gpu_fir_tensor.py tests/test_gpu_fir.py -m gpu -x -q 2>&1 | tail -3
bash scripts/ab_libs.sh nopf 2>&1 | tail -4
FLAGS="0" bash scripts/ab_timing.sh 2>&1 | grep TCT
FLAGS="12 14" bash scripts/tc_knockout.sh

# secondary kernels: parity + timing
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py --only fir,chain,resamp 2>&1 | grep -v tensor | tail -14 | cut -c1-200
for mb in 32 8 4; do echo -n "e2e chunk ${mb} MiB: "; B2S_HOST_CHUNK_MB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value']), d['roofline']['kernel_ms'])"; done

# per-role cycle accounting of the tensor FIR (variant built with -DB2S_TC_TIMING) + store-flavour A/B
V=futuresdr_b200/variants
for fl in ${FLAGS:-0 2}; do
  echo "=== timing variant, flags=$fl"
  B2S_TC_FLAGS=$fl B2S_LIB=$V/libb200sdr_timing.so B2S_TC_TIMING_DUMP=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | grep -E "^TCT|metric" | cut -c1-120
done
for v in "" $V/libb200sdr_stcs.so; do echo -n "=== lib=${v:-default}  "; B2S_LIB=$v timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4))"; done

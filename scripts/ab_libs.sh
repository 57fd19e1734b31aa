# ab_libs.sh NAME...: bench.py kernel time of the default library and the named variants, on the same box
V=futuresdr_b200/variants
for rep in 1 2; do
for v in "" "$@"; do lib=${v:+$V/libb200sdr_$v.so}; echo -n "lib=${v:-default}  "; B2S_LIB=$lib timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4))"; done
done

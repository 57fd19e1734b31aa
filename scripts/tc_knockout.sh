# times the tensor FIR with parts of the pipeline knocked out (B2S_TC_FLAGS tuning switches):
# bit1 (2): no epilogue stores, bit2 (4): no MMAs, bit3 (8): no conversion/STS
for fl in ${FLAGS:-0 2 4 8 14}; do echo -n "flags=$fl  "; B2S_TC_FLAGS=$fl timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4))"; done

# round-2 batch W (1 GPU): final verification of the tree -- full GPU suite, sanitizer passes, bench line, per-kernel table
mkdir -p gpurun_out
echo "--- suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "--- smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "--- bench"; timeout 1200 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; tail -2 gpurun_out/bench_r2_final_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'sustained',d.get('sustained'))
for s in d.get('secondary',[]):
    print(s['config']['workload'][:70], '| value', s.get('value'), s.get('unit'), '| frac', (s.get('roofline') or {}).get('frac'), '| cpu', (s.get('cpu_baseline') or {}).get('value'))
PY
echo "--- reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
echo "--- configs"; timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs_r2_final.jsonl 2>&1; cut -c1-140 gpurun_out/bench_configs_r2_final.jsonl
echo "--- memcheck"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x \
  -k "not full_chunk and not full_size and not bench_chunk_size and not long_stream and not wraps and not large_sizes and not sharded_parity and not timeout and not many_tiles" \
  > gpurun_out/memcheck_r2_final.log 2>&1
echo "exit $?"; tail -4 gpurun_out/memcheck_r2_final.log | cut -c1-200
echo "--- racecheck (kernels changed last: resampler, pfbarb, fused PFB)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_channelizer.py tests/test_gpu_synthesizer.py -q -x \
  -k "(resamp or pfbarb or fused) and not long_stream and not large" > gpurun_out/racecheck_r2_final.log 2>&1
echo "exit $?"; tail -4 gpurun_out/racecheck_r2_final.log | cut -c1-200

# round-2 batch L (1 GPU): compute-sanitizer memcheck over the GPU suite (minus the multi-GiB / multi-minute cases)
mkdir -p gpurun_out
timeout 3000 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x \
  -k "not full_chunk and not full_size and not bench_chunk_size and not long_stream and not wraps and not large_sizes and not sharded_parity and not timeout and not many_tiles" \
  > gpurun_out/memcheck_r2.log 2>&1
echo "exit $?"; tail -8 gpurun_out/memcheck_r2.log | cut -c1-200; grep -c "Invalid" gpurun_out/memcheck_r2.log

# round-2 batch P (1 GPU): complete per-kernel table + ncu of the fused synthesizer
mkdir -p gpurun_out
timeout 1500 python scripts/bench_configs.py > gpurun_out/bench_configs_r2_all.jsonl 2>&1; tail -60 gpurun_out/bench_configs_r2_all.jsonl | cut -c1-170
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_fused -s 1 -c 1 -o gpurun_out/prof_r2_synth python scripts/bench_configs.py --only synth > /dev/null 2>&1
ls -la gpurun_out/prof_r2_synth.ncu-rep

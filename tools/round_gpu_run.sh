#!/bin/bash
# One-GPU evidence run of a round: tests, smoke, bench lines of every BASELINE config, ncu launch lists and one
# `--set full` capture per dominant kernel.  Usage (from the repo root, on a B200 box):  bash tools/round_gpu_run.sh <outdir>
out=${1:-gpurun_out/final}; mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -12) > $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
b() { name=$1; shift; timeout 400 python bench.py "$@" 2> $out/$name.err | tail -1 > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['config']['engine'][:16], (d.get('cpu_baseline') or {}).get('value'))" $out/$name.json; }
b bench_n168_q80
b bench_q06_q80 --workload qwen3-0.6b-q80 --steps 2
b bench_q06_q4k --workload qwen3-0.6b-q4k --steps 2 --no-cpu-baseline
b bench_n168_f32 --workload nano-168m-f32 --no-cpu-baseline
b bench_n168_q80_exact --exact --steps 2 --no-cpu-baseline
b bench_n168_q80_multikernel --no-stream --no-mega --steps 3 --no-cpu-baseline
b bench_n168_q80_megakernel --no-stream --steps 3 --no-cpu-baseline
b bench_q17_q80 --workload qwen3-1.7b-q80 --steps 1 --no-cpu-baseline
b bench_q4b_q80 --workload qwen3-4b-q80 --steps 1 --no-cpu-baseline
# launch lists of the default command (cluster path) and of the multi-kernel path; never a bench value
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_default_n168.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/under_ncu_default.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches_multikernel_q06.csv python bench.py --workload qwen3-0.6b-q80 --steps 1 --warmup 3 --no-cpu-baseline > $out/under_ncu_q06.log 2>&1
# full captures: the cluster kernel (first launch = the 15 prompt positions) and the multi-kernel path's kernels of one layer
ncu --set full --clock-control none --import-source on -k regex:k_decode_cluster -c 1 -o $out/prof_cluster -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_cluster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_matvec|k_attention' --launch-skip 700 -c 10 -o $out/prof_multikernel_q06 -f python bench.py --workload qwen3-0.6b-q80 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_multi.log 2>&1
ls -la $out | tail -30

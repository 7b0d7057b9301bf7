#!/bin/bash
out=gpurun_out/r2h; mkdir -p $out
b() { name=$1; shift; (timeout 600 python bench.py "$@" --no-cpu-baseline 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['config']['engine'][:16], round(d['token_roofline']['frac_of_peak'],3))" $out/$name.json || tail -c 400 $out/$name.err; }
b bench_q06_q80 --workload qwen3-0.6b-q80 --steps 2
(timeout 200 python tools/gpu_trace.py qwen3-0.6b 2048 2>&1 | tail -8) > $out/trace_q06.log; cat $out/trace_q06.log
b bench_q4b_q80 --workload qwen3-4b-q80 --steps 1
(timeout 300 python tools/gpu_trace.py qwen3-4b 4096 2>&1 | tail -8) > $out/trace_q4b.log; cat $out/trace_q4b.log
b bench_q06_q4k --workload qwen3-0.6b-q4k --steps 2
b bench_n168_f32 --workload nano-168m-f32 --steps 2

#!/bin/bash
out=gpurun_out/r2l; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -5) > $out/pytest_engine.log; tail -2 $out/pytest_engine.log
b() { name=$1; shift; (timeout 900 python bench.py "$@" --no-cpu-baseline 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['config']['engine'][:16], round(d['token_roofline']['frac_of_peak'],3))" $out/$name.json || tail -c 400 $out/$name.err; }
b n168_stream --steps 2
b q06_stream --workload qwen3-0.6b-q80 --steps 2
b q4b_stream --workload qwen3-4b-q80 --steps 1
b q06_q4k_stream --workload qwen3-0.6b-q4k --steps 2
b n168_multi --steps 2 --no-stream
b q06_multi --workload qwen3-0.6b-q80 --steps 2 --no-stream
b q06_q4k_multi --workload qwen3-0.6b-q4k --steps 2 --no-stream
b n168_f32_stream --workload nano-168m-f32 --steps 2
b n168_f32_multi --workload nano-168m-f32 --steps 2 --no-stream

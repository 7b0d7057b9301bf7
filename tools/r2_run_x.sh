#!/bin/bash
out=gpurun_out/r2x; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_long.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -6) > $out/pytest.log; tail -3 $out/pytest.log
b() { name=$1; shift; (timeout 900 python bench.py --no-cpu-baseline --no-extra "$@" 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['run']['engine'][:16], d['parity']['agree'], d['parity']['oracle_margin_at_divergence'])" $out/$name.json || tail -c 600 $out/$name.err; }
b n168 --steps 2
NB200_STREAM=1 b q06_stream --workload qwen3-0.6b-q80 --steps 2
NB200_STREAM=1 b q06q4k_stream --workload qwen3-0.6b-q4k --steps 2
NB200_STREAM=1 b q17_stream --workload qwen3-1.7b-q80 --steps 1 --warmup 1
NB200_STREAM=1 b q4b_stream --workload qwen3-4b-q80 --steps 1 --warmup 1
export NB200_ENGINE_SO=$PWD/nano_b200/lib/libnano_b200_trace.so
NB200_STREAM=1 timeout 400 python tools/gpu_trace.py qwen3-4b 4096 2>&1 | grep -E "token cycles|L/2|inside attention" | cut -c1-450
NB200_STREAM=1 timeout 400 python tools/gpu_trace.py qwen3-0.6b 2048 2>&1 | grep -E "token cycles|L/2|inside attention" | cut -c1-450

#!/bin/bash
out=gpurun_out/r2z; mkdir -p $out
(timeout 1200 python -m pytest tests/test_gpu_engine.py tests/test_gpu_long.py tests/test_gpu_baseline_shapes.py tests/test_gpu_ops.py -x -q 2>&1 | tail -8) > $out/pytest.log; tail -3 $out/pytest.log
b() { name=$1; shift; (timeout 900 python bench.py --no-cpu-baseline --no-extra "$@" 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['run']['engine'][:16], d['run'].get('path_calibration'), d['parity']['agree'])" $out/$name.json || tail -c 600 $out/$name.err; }
NB200_STREAM=0 b q06_multi --workload qwen3-0.6b-q80 --steps 2
NB200_STREAM=0 b q17_multi --workload qwen3-1.7b-q80 --steps 1 --warmup 1
NB200_STREAM=0 b q4b_multi --workload qwen3-4b-q80 --steps 1 --warmup 1
NB200_STREAM=0 b n168_multi --steps 2

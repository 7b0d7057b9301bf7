#!/bin/bash
out=gpurun_out/r2g; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -25) > $out/pytest_engine.log; tail -3 $out/pytest_engine.log
(timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | tail -12) > $out/trace_n168.log; cat $out/trace_n168.log
(timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> $out/bench_n168.err | tail -1) > $out/bench_n168.json; tail -c 600 $out/bench_n168.err
python -c "import json;d=json.load(open('$out/bench_n168.json'));print('n168',d['value'],d['e2e']['value'],d['config']['engine'])"

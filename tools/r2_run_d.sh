#!/bin/bash
out=gpurun_out/r2d; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -25) > $out/pytest_engine.log; tail -3 $out/pytest_engine.log
(timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | tail -14) > $out/trace_n168.log; cat $out/trace_n168.log
(timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> $out/bench_n168.err | tail -1) > $out/bench_n168.json; tail -c 600 $out/bench_n168.err
python -c "import json;d=json.load(open('$out/bench_n168.json'));print('n168',d['value'],d['e2e']['value'],d['config']['engine'])"
(timeout 400 python bench.py --workload qwen3-0.6b-q80 --steps 2 --warmup 3 --no-cpu-baseline 2> $out/bench_q06.err | tail -1) > $out/bench_q06.json; tail -c 600 $out/bench_q06.err
python -c "import json;d=json.load(open('$out/bench_q06.json'));print('q06',d['value'],d['e2e']['value'],d['config']['engine'])"
(timeout 200 python tools/gpu_trace.py qwen3-0.6b 2048 2>&1 | tail -14) > $out/trace_q06.log; cat $out/trace_q06.log

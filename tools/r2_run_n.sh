#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
(timeout 600 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -15) > $out/pytest_engine.log; tail -3 $out/pytest_engine.log
(timeout 1500 python -m pytest tests/test_gpu_long.py -x -q -s 2>&1 | tail -25) > $out/pytest_long.log; tail -8 $out/pytest_long.log
(timeout 1500 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -s 2>&1 | tail -25) > $out/pytest_shapes.log; tail -8 $out/pytest_shapes.log

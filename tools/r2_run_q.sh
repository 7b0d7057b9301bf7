#!/bin/bash
out=gpurun_out/r2q; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_shim.py -x -q 2>&1 | tail -15) > $out/pytest_shim.log; tail -5 $out/pytest_shim.log
(timeout 600 python -m pytest tests/test_gpu_lora.py tests/test_gpu_ops.py -x -q 2>&1 | tail -4) > $out/pytest_misc.log; tail -3 $out/pytest_misc.log

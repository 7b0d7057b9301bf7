#!/bin/bash
out=gpurun_out/r2c; mkdir -p $out
(timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | tail -14) > $out/trace_n168.log; cat $out/trace_n168.log

#!/bin/bash
out=gpurun_out/r2e; mkdir -p $out
(timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | tail -12) > $out/trace_n168.log; cat $out/trace_n168.log

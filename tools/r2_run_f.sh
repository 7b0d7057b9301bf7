#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_decode_stream -c 1 -o $out/prof_stream -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
tail -5 $out/ncu.log; ls -la $out

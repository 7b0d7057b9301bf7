#!/bin/bash
# Builds the microbenchmarks behind the latency / bandwidth floors quoted in DESIGN.md section 5.4 (binaries are git-ignored).
cd "$(dirname "$0")"
for f in *_bench.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o "${f%.cu}" "$f" || exit 1
done

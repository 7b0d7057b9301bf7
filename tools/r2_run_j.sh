#!/bin/bash
out=gpurun_out/r2j; mkdir -p $out
for ch in 32 64 168 512; do
  echo "== NB200_ATTN_CHUNK=$ch"; NB200_ATTN_CHUNK=$ch timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | grep -E "token cycles|layer L/2" | cut -c1-330
done
for ch in 16 32 64 128; do
  echo "== q06 NB200_ATTN_CHUNK=$ch"; NB200_ATTN_CHUNK=$ch timeout 200 python tools/gpu_trace.py qwen3-0.6b 2048 2>&1 | grep -E "token cycles|layer L/2" | cut -c1-330
done

#!/bin/bash
timeout 200 python tools/gpu_trace.py nano-168m 512 2>&1 | grep -E "token cycles|inside attention" | cut -c1-400
timeout 200 python tools/gpu_trace.py qwen3-0.6b 2048 2>&1 | grep -E "token cycles|inside attention" | cut -c1-400

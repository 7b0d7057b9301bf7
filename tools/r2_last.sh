#!/bin/bash
out=gpurun_out/r2last; mkdir -p $out
(timeout 130 python -m pytest tests -m gpu -q 2>&1 | tail -12) > $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
(timeout 200 python bench.py 2> $out/bench_n1_default.err | tail -1) > $out/bench_n1_default.json; head -c 300 $out/bench_n1_default.json; echo
cat > /tmp/cap.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from nano_b200 import engine as E, modelfile as mf
spec = mf.PRESETS["nano-168m"]; seq = 512
eng = E.Engine(mf.cached_model(spec, mf.QUANT_Q80, 128), seq)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)]
eng.decode_greedy(ids, 16, 48)
print(eng.path)
PY
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream -s 1 -c 1 -o $out/prof_stream_n168 -f python /tmp/cap.py > $out/ncu_stream.log 2>&1; tail -2 $out/ncu_stream.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_decode_stream -c 6 --csv --log-file $out/launches_default_n168.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/under_ncu_default.log 2>&1; tail -3 $out/launches_default_n168.csv | cut -c1-200
for m in exchange_bench ring_bench consume_bench; do (timeout 25 tools/micro/$m > $out/micro_$m.log 2>&1; tail -2 $out/micro_$m.log); done
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
(timeout 100 python bench.py --impl reference --steps 2 --warmup 1 2> $out/bench_reference_arm.err | tail -1) > $out/bench_reference_arm.json; head -c 200 $out/bench_reference_arm.json
ls -la $out

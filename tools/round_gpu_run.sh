#!/bin/bash
# One-GPU evidence run of a round: tests, smoke, bench lines, microbenchmarks, ncu launch lists and `--set full` captures.
# Usage (from the repo root, on a B200 box):  bash tools/round_gpu_run.sh <outdir>
out=${1:-gpurun_out/final}; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12) > $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
# the two arms exactly as the driver runs them
(timeout 900 python bench.py --impl reference 2> $out/bench_reference_arm.err | tail -1) > $out/bench_reference_arm.json
(timeout 1500 python bench.py 2> $out/bench_n1_default.err | tail -1) > $out/bench_n1_default.json
python - $out <<'PY'
import json, sys
o = sys.argv[1]
for n in ("bench_reference_arm", "bench_n1_default"):
    try:
        d = json.load(open(f"{o}/{n}.json"))
        print(n, round(d["value"], 1), d.get("e2e", {}).get("value"), (d.get("run") or {}).get("engine", "")[:20],
              {k: round(v["value"], 1) for k, v in (d.get("configs") or {}).items()}, (d.get("exact_mode") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(n, "unreadable:", ex)
PY
b() { name=$1; shift; (timeout 900 python bench.py --no-cpu-baseline --no-extra "$@" 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['run']['engine'][:16], d['run'].get('path_calibration'))" $out/$name.json || tail -c 400 $out/$name.err; }
b bench_n168_f32 --workload nano-168m-f32 --steps 2
b bench_q17_q80 --workload qwen3-1.7b-q80 --steps 1 --warmup 1
b bench_q4b_q80 --workload qwen3-4b-q80 --steps 1 --warmup 1
NB200_STREAM=0 b bench_n168_q80_multikernel --steps 2
NB200_STREAM=1 b bench_q06_q80_stream --workload qwen3-0.6b-q80 --steps 2
NB200_STREAM=0 b bench_q06_q4k_multikernel --workload qwen3-0.6b-q4k --steps 2
# microbenchmarks behind the floors in DESIGN.md 5.4
for m in lat_bench bar_bench exchange_bench ring_bench consume_bench; do
  [ -x tools/micro/$m ] && (timeout 120 tools/micro/$m > $out/micro_$m.log 2>&1; tail -3 $out/micro_$m.log)
done
# launch list of the default command restricted to the step's kernel (never a bench value).  Keep -k / -c: an unrestricted list of
# this command profiles every kernel of the per-kernel orientation pass one by one and takes half an hour.
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_decode_stream -c 6 --csv --log-file $out/launches_default_n168.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/under_ncu_default.log 2>&1
cat > /tmp/cap.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from nano_b200 import engine as E, modelfile as mf
name, quant, seq, n_total = sys.argv[1], {"q80": mf.QUANT_Q80, "q4k": mf.QUANT_Q4K}[sys.argv[2]], int(sys.argv[3]), int(sys.argv[4])
spec = mf.PRESETS[name]
eng = E.Engine(mf.cached_model(spec, quant, 128), seq)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)] if spec.arch == 0 else [1000 + i for i in range(16)]
eng.decode_greedy(ids, 16, n_total)
print(eng.path)
PY
# streaming kernel: launch 0 = 15 prompt positions, launch 1 = 32 decode positions (captured; ~1 minute)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream -s 1 -c 1 -o $out/prof_stream_n168 -f python /tmp/cap.py nano-168m q80 512 48 > $out/ncu_stream.log 2>&1
# multi-kernel path (Qwen3-0.6B Q80): the kernels of one layer early in the decode segment (141 matching launches per token; a large
# --launch-skip is slow under ncu, and a report of more than ~6 kernels exceeds what gpurun copies back together with the one above)
NB200_STREAM=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_matvec|k_attention' -s 2400 -c 5 -o $out/prof_multikernel_q06 -f python /tmp/cap.py qwen3-0.6b q80 2048 40 > $out/ncu_multi.log 2>&1
ls -la $out | tail -40

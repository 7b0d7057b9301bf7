#!/bin/bash
out=gpurun_out/r2p; mkdir -p $out
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err); tail -c 600 $out/bench_n2.err
python - <<'PY'
import json
for line in open('gpurun_out/r2p/bench_n2.json'):
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('value',round(d['value'],1),'n',d['n_gpus'],'e2e',round(d['e2e']['value'],1)); print('tp',d['tp'])
PY
(timeout 600 python -m pytest tests/test_gpu_tp.py -x -q 2>&1 | tail -4) > $out/pytest_tp.log; cat $out/pytest_tp.log

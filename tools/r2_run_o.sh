#!/bin/bash
out=gpurun_out/r2o; mkdir -p $out
(time timeout 1200 python bench.py --steps 5 --warmup 3 > $out/bench_default.json 2> $out/bench_default.err); tail -c 300 $out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o/bench_default.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['run']['engine'][:20],'frac',round(d['roofline']['frac'],4))
print('parity',d['parity']); print('exact',d['exact_mode'])
for k,v in (d['configs'] or {}).items(): print(k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','engine','error')}, 'e2e', v.get('e2e',{}).get('value'), 'parity', v.get('parity'))
print('cpu',d['cpu_baseline'])
PY
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2> $out/bench_ref.err); cat $out/bench_ref.json | cut -c1-900

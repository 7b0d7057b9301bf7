#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
b() { name=$1; shift; (timeout 300 python bench.py --steps 2 --no-cpu-baseline "$@" 2> $out/$name.err | tail -1) > $out/$name.json
      python -c "import json,sys; d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1))" $out/$name.json || tail -c 300 $out/$name.err; }
b base
NB200_STAGE_KB=16 b stage16
NB200_OWNED_ROWS=1000000 NB200_OWNED_KB=1000000 b noowned
NB200_STAGE_KB=16 NB200_OWNED_ROWS=1000000 NB200_OWNED_KB=1000000 b stage16_noowned
NB200_STAGE_KB=16 NB200_OWNED_ROWS=1000000 NB200_OWNED_KB=1000000 NB200_ATTN_CHUNK=64 b stage16_noowned_chunk64
NB200_STAGE_KB=24 NB200_OWNED_ROWS=1000000 NB200_OWNED_KB=1000000 b stage24_noowned

"""%globaltimer stamps inside k_attention_fast (layer L/2, kv head 0) for one token of the multi-kernel path.
   usage: python tools/gpu_attn_trace.py [preset] [seq]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NB200_ATTN_DBG"] = "1"
from nano_b200 import engine as E, modelfile as mf
name = sys.argv[1] if len(sys.argv) > 1 else "qwen3-0.6b"
seq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
spec = mf.PRESETS[name]
eng = E.Engine(mf.cached_model(spec, mf.QUANT_Q80, 128), seq, flags=E.FLAG_NO_STREAM)
print(eng.path)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)] if spec.arch == 0 else [1000 + i for i in range(16)]
eng.decode_greedy(ids, 16, seq)
names = {0: "entry", 1: "pos loaded", 2: "item start", 3: "partial done", 4: "ws written+sync", 5: "ticket done", 16: "last: fence", 17: "last: staged+weights", 18: "last: xba written"}
for pos in (64, seq // 2, seq - 2):
    for rep in range(3):
        eng.forward_nolog(int(ids[pos]), pos)
    t = eng.attn_trace().astype(np.int64)
    base = t[0]
    print(f"pos {pos}: " + "  ".join(f"{names[k]}={t[k] - base}" for k in sorted(names) if t[k]))

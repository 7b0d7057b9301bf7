import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nano_b200 import engine as E, modelfile as mf

def t(label, fn):
    t0 = time.time(); r = fn(); print(f"  {label}: {time.time()-t0:.3f}s", flush=True); return r

spec = mf.PRESETS["toy-qwen3"]
path = mf.cached_model(spec, mf.QUANT_Q80, 64)
for flags, nm in [(0, "default"), (E.FLAG_NO_PDL, "no_pdl"), (E.FLAG_NO_GRAPH, "no_graph"), (E.FLAG_NO_GRAPH | E.FLAG_NO_PDL, "no_graph_no_pdl")]:
    print(nm, flush=True)
    eng = t("create", lambda: E.Engine(path, 64, flags=flags))
    for i in range(3):
        t(f"forward {i}", lambda: eng.forward_nolog(5, i))
    t("logits", lambda: eng.logits())
    ids = np.zeros(65, np.uint32); ids[:4] = [1, 2, 3, 4]
    r = t("decode_greedy 60", lambda: eng.decode_greedy(ids, 4, 64))
    print("  device ms (prefill, decode):", r, flush=True)
    t("close", lambda: eng.close())

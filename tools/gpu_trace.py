"""Per-phase cycle trace of one token in the persistent kernel (CTA 0's clock64 after each grid barrier)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nano_b200 import engine as E, modelfile as mf
name, quant, gs, seq = sys.argv[1] if len(sys.argv) > 1 else "nano-168m", mf.QUANT_Q80, 128, 512
if len(sys.argv) > 2: seq = int(sys.argv[2])
spec = mf.PRESETS[name]
eng = E.Engine(mf.cached_model(spec, quant, gs), seq)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)] if spec.arch == 0 else [1000 + i for i in range(16)]
for _ in range(3): eng.decode_greedy(ids, 16, seq)
for pos in (seq // 2, seq - 2):
    st = eng.trace_token(int(ids[pos]), pos).astype(np.int64)
    L = spec.n_layer
    intra = st[1024:1024 + 48].reshape(3, 16)[:, :7]
    st = st[: 5 * L + 4]
    d = np.diff(st)
    per = d[: 5 * L].reshape(L, 5)
    print(f"pos {pos}: total cycles {st[-1] - st[0]}  per-layer phase medians [qkv, attn, o, w13, w2] = {np.median(per, axis=0).astype(int).tolist()}  "
          f"layer sum median {int(np.median(per.sum(axis=1)))}  tail (cls, finalize+embed) = {d[5 * L:].tolist()}")
    print("   layer0", per[0].tolist(), " layer", L // 2, per[L // 2].tolist())
    for name, row in zip(("qkv", "o", "w13"), intra):
        print(f"   intra {name}: [tile-load issue, (q80 entry), stage, rms, quant, rows, epilogue] =", np.diff(row).tolist())

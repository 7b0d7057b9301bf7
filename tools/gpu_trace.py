"""Per-phase cycle trace of one token in the streaming kernel (CTA 0's clock64 after each grid barrier, plus the
stamps inside layer L/2: after every prologue / consume / barrier)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nano_b200 import engine as E, modelfile as mf
name = sys.argv[1] if len(sys.argv) > 1 else "nano-168m"
seq = int(sys.argv[2]) if len(sys.argv) > 2 else 512
quant = {"q80": mf.QUANT_Q80, "q4k": mf.QUANT_Q4K, "f32": mf.QUANT_F32}[sys.argv[3] if len(sys.argv) > 3 else "q80"]
spec = mf.PRESETS[name]
eng = E.Engine(mf.cached_model(spec, quant, 128, fast=(quant == mf.QUANT_Q80 and spec.n_embd >= 2048)), seq)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)] if spec.arch == 0 else [1000 + i for i in range(16)]
for _ in range(2): eng.decode_greedy(ids, 16, seq)
L = spec.n_layer
for pos in (seq // 2, seq - 2):
    st = eng.trace_token(int(ids[pos]), pos).astype(np.int64)
    st_all = st
    intra = st[1024:1024 + 10]
    st = st[: 5 * L + 2]
    d = np.diff(st)
    per = d[: 5 * L].reshape(L, 5)
    print(f"{name} pos {pos}: token cycles {st[-1] - st[0]}  per-layer phase medians [qkv, attn, o, w13, w2] = {np.median(per, axis=0).astype(int).tolist()}  "
          f"layer sum median {int(np.median(per.sum(axis=1)))}  cls = {int(d[5 * L])}")
    print("   layer0", per[0].tolist(), " layer", L // 2, per[L // 2].tolist())
    di = np.diff(intra).tolist()
    print("   layer L/2 [prep, consume]: qkv", di[0:2], " attn", di[2:3], " o", di[3:5], " w13", di[5:7], " w2", di[7:9])
    dbg = st_all[1100:1100 + 64].reshape(4, 16)
    for nm, row in zip(("qkv", "o", "w13", "w2"), dbg):
        r = row - row[0]
        print(f"   inside {nm}: prep [loaded {r[1]}, inverse {r[2]}, quantised {r[3]}, cbar {r[4]}]  consume [enter {r[5]}, tile ready {r[6]}, (row: start {r[10]}, dp4a done {r[11]}, term {r[12]}, gathered {r[13]}) first row {r[7]}, tile done {r[8]}, out {r[9]}]"
              )
    at = st_all[1100 + 64:1100 + 64 + 16]; at = at - at[0]
    print(f"   inside attention (CTA 0 = kv head 0, split 0): q ready {at[1]}, tiles resident {at[2]}, k/v injected {at[3]}, scores {at[4]}, softmax {at[5]}, P.V {at[6]}, released {at[7]}, partial {at[8]}, (merge: start {at[9]}, partials polled {at[10]}, published {at[11]}); softmax detail: max loop {at[12]}, warp_max {at[13]}, expf(old) {at[14]}, exp loop {at[15]}")

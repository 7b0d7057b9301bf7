"""Cycle stamps of rank 0 in the cluster-resident kernel for one token (layer pattern: qkv[prep,tiles,sync] attn[partial,sync,merge] o[prep,tiles,sync] w13[...] w2[...])."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nano_b200 import engine as E, modelfile as mf
name = sys.argv[1] if len(sys.argv) > 1 else "nano-168m"
seq = int(sys.argv[2]) if len(sys.argv) > 2 else 512
spec = mf.PRESETS[name]
eng = E.Engine(mf.cached_model(spec, mf.QUANT_Q80, 128), seq)
ids = np.zeros(seq + 1, np.uint32); ids[:16] = [17 + i % 10 for i in range(16)] if spec.arch == 0 else [1000 + i for i in range(16)]
for _ in range(2): eng.decode_greedy(ids, 16, seq)
for pos in (seq // 2, seq - 2):
    raw = eng.trace_token(int(ids[pos]), pos).astype(np.int64)
    tiles = raw[1000:1032].reshape(8, 4)
    print("   w13 tiles of layer L/2 (rank 0, warp 0): [refill, wait full, consume+arrive] and gap to next tile:")
    for j in range(8):
        if tiles[j, 0]:
            nxt = tiles[j + 1, 0] - tiles[j, 3] if j < 7 and tiles[j + 1, 0] else -1
            print("     tile", j, np.diff(tiles[j]).tolist(), "epilogue+loop", int(nxt))
    at = raw[1040:1046]
    if at[0]: print("   attention partial (rank 0, thread 0): [q prep, main loop, warp merge, sync, cross-warp merge] =", np.diff(at).tolist())
    st = raw[:1000]
    st = st[st != 0]
    d = np.diff(st)
    per_layer = 3 + 3 + 3 + 3 + 3            # stamps per layer
    L = spec.n_layer
    body = d[: per_layer * L].reshape(L, per_layer)
    names = ["qkv.prep", "qkv.tiles", "qkv.sync", "att.part", "att.sync", "att.merge", "o.prep", "o.tiles", "o.sync", "w13.prep", "w13.tiles", "w13.sync", "w2.prep", "w2.tiles", "w2.sync"]
    med = np.median(body, axis=0).astype(int)
    print(f"pos {pos}: token cycles {st[-1] - st[0]}; per-layer median {int(np.median(body.sum(axis=1)))}")
    print("   " + "  ".join(f"{n}={v}" for n, v in zip(names, med)))
    print("   tail:", d[per_layer * L:].tolist())

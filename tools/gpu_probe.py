"""First-contact diagnostics on the GPU box: prints parity numbers and quick timings (not a test)."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nano_b200 import engine as E, modelfile as mf      # noqa: E402
from oracle import bindings as ob                      # noqa: E402


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[ok] {name}: {r}  ({time.time() - t:.2f}s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}\n{traceback.format_exc()}", flush=True)


def parity(name, quant, gs, S=24, flags=0):
    spec = mf.PRESETS[name]
    path = mf.cached_model(spec, quant, gs)
    eng = E.Engine(path, S, flags=flags); o = ob.NanoOracle(path, S)
    toks = mf.teacher_tokens(S, spec.vocab)
    worst = 0.0; bitsame = True; agree = 0
    for pos in range(S):
        a = eng.forward(toks[pos], pos); b = o.forward(toks[pos], pos)
        worst = max(worst, float(np.abs(a - b).max()))
        bitsame &= bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
        agree += int(np.argmax(a) == np.argmax(b))
    eng.close(); o.close()
    return dict(max_dlogit=worst, bit_identical=bitsame, argmax_agree=f"{agree}/{S}")


def timing(name, quant, gs, S, P=16, reps=3):
    spec = mf.PRESETS[name]
    path = mf.cached_model(spec, quant, gs)
    eng = E.Engine(path, S)
    out = []
    for _ in range(reps):
        ids = np.zeros(S + 1, np.uint32); ids[:P] = [(17 + i % 10) if spec.arch == 0 else 1000 + i for i in range(P)]
        pre, dec = eng.decode_greedy(ids, P, S)
        out.append((S - P) / (dec * 1e-3))
    t0 = time.time(); n = 0
    ids = np.zeros(S + 1, np.uint32); ids[:P] = [(17 + i % 10) if spec.arch == 0 else 1000 + i for i in range(P)]
    for pos in range(min(S - 1, 200)):
        ids[pos + 1] = eng.next_greedy(ids, pos, 1 if pos < P - 1 else 0); n += 1
    api = n / (time.time() - t0)
    r = dict(device_tok_s=[round(v) for v in out], api_tok_s=round(api), launches_per_token=eng.launches_per_token,
             weight_MB=round(eng.weight_bytes / 1e6, 1))
    eng.close()
    return r


if __name__ == "__main__":
    print("devices", E.device_count(), flush=True)
    for cfg in [("toy-nano", mf.QUANT_F32, 128), ("toy-nano", mf.QUANT_Q80, 64), ("toy-nano", mf.QUANT_Q4K, 128),
                ("toy-qwen3", mf.QUANT_F32, 128), ("toy-qwen3", mf.QUANT_Q80, 64), ("toy-qwen3", mf.QUANT_Q4K, 128)]:
        step(f"parity fast {cfg}", lambda c=cfg: parity(*c))
        step(f"parity exact {cfg}", lambda c=cfg: parity(*c, flags=E.FLAG_EXACT))
    step("timing nano-168m q80 seq512", lambda: timing("nano-168m", mf.QUANT_Q80, 128, 512))
    if "--big" in sys.argv:
        step("timing qwen3-0.6b q80 seq2048", lambda: timing("qwen3-0.6b", mf.QUANT_Q80, 128, 2048, reps=2))

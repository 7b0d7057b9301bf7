"""Regenerates the committed golden fixtures from the UNMODIFIED reference (strict build).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

  sort6_model.bin      the reference's only in-tree fixture: the FP32 sort model embedded in
                       infer/main_sort.c:6-3098 (49,452 bytes), extracted through oracle/ref_harness.c
  sort6_kat.json       seq2seq answers of the reference on it (README.md:379 "114515 -> 111455" and four more)
  toy_logits.npz       teacher-forced logits of the strict reference on seeded synthetic toy files
                       (nano_b200/modelfile.py presets, seed 39) at a few positions, F32 / Q80 / Q4K
  lora_logits.npz      the same with a seeded synthetic LoRA plug-in (rank 8, alpha 16) attached to toy-nano
  q4k_kat.npz          Q4K op-level vectors following the recipe of infer/tools/export_q4k.c:394-450
                       (seed 39 xorshift, d=8, n=768): reference quantize_tensor_q4k bytes + matmul_q4k outputs
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nano_b200 import modelfile as mf            # noqa: E402
from oracle import bindings as ob               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOY_CONFIGS = [("toy-nano", mf.QUANT_F32, 128), ("toy-nano", mf.QUANT_Q80, 64), ("toy-nano", mf.QUANT_Q4K, 128),
               ("toy-qwen3", mf.QUANT_F32, 128), ("toy-qwen3", mf.QUANT_Q80, 64), ("toy-qwen3", mf.QUANT_Q4K, 128)]
TOY_SEQ = 24
TOY_POSITIONS = [0, 1, 7, 23]


def xorshift_f32(n, seed=39):
    """utils.c:959-970 random_f32 stream."""
    M = 0xFFFFFFFFFFFFFFFF
    st = seed
    out = np.zeros(n, np.float32)
    for i in range(n):
        st ^= st >> 12; st ^= (st << 25) & M; st ^= st >> 27
        out[i] = np.float32((((st * 0x2545F4914F6CDD1D) & M) >> 32 >> 8) / 16777216.0)
    return out


def main():
    ob.build()
    sm = ob.sort_model_bytes()
    open(os.path.join(HERE, "sort6_model.bin"), "wb").write(sm)
    ref = ob.RefEngine(sm, 6, penalty=0.0, temperature=0.0, top_p=0.0, top_k=1)
    kat = {s: ref.seq2seq(s, 6) for s in ["251212", "114515", "654321", "000000", "909090", "123321", "777111"]}
    json.dump(kat, open(os.path.join(HERE, "sort6_kat.json"), "w"), indent=1)
    print("sort KAT", kat)

    out = {}
    for name, quant, gs in TOY_CONFIGS:
        spec = mf.PRESETS[name]
        path = mf.cached_model(spec, quant, gs)
        r = ob.RefEngine(path, TOY_SEQ, "strict")
        toks = mf.teacher_tokens(TOY_SEQ, spec.vocab)
        rows = []
        for pos in range(TOY_SEQ):
            lg = r.forward(toks[pos], pos)
            if pos in TOY_POSITIONS:
                rows.append(lg)
        key = f"{name}_{quant:02x}_{gs}"
        out[key] = np.stack(rows)
        r.close()
        print(key, out[key].shape, float(np.abs(out[key]).max()))
    np.savez_compressed(os.path.join(HERE, "toy_logits.npz"), **out)

    # LoRA plug-in (infer.c:792-808, 898-903): strict-reference logits with a seeded synthetic plug-in attached
    lora_out = {}
    for quant, gs in [(mf.QUANT_F32, 128), (mf.QUANT_Q80, 64)]:
        spec = mf.PRESETS["toy-nano"]
        path = mf.cached_model(spec, quant, gs)
        r = ob.RefEngine(path, TOY_SEQ, "strict")
        r.load_lora(mf.write_lora(spec, 8, 16, seed=7))
        toks = mf.teacher_tokens(TOY_SEQ, spec.vocab)
        rows = [r.forward(toks[pos], pos) for pos in range(TOY_SEQ)]
        lora_out[f"toy-nano_{quant:02x}_{gs}"] = np.stack([rows[p] for p in TOY_POSITIONS])
        r.close()
    np.savez_compressed(os.path.join(HERE, "lora_logits.npz"), **lora_out)

    # Noise floor of the reference itself: max |logit(fast build) - logit(strict build)| of the SAME source on the
    # SAME file (SURVEY finding 11).  The fast-mode GPU tolerance is max(north-star tolerance, 1.5 x this floor).
    floors = {}
    S = 40
    for name, quant, gs in TOY_CONFIGS + [("mini-qwen3", mf.QUANT_Q80, 128), ("mini-nano", mf.QUANT_Q80, 128), ("mini-nano", mf.QUANT_Q4K, 128)]:
        spec = mf.PRESETS[name]
        path = mf.cached_model(spec, quant, gs)
        a = ob.RefEngine(path, S, "strict")
        toks = mf.teacher_tokens(S, spec.vocab)
        strict = [a.forward(toks[p], p) for p in range(S)]
        worst = 0.0
        for fl in ("fast_v3", "fast_v4"):
            b = ob.RefEngine(path, S, fl)
            for p in range(S):
                worst = max(worst, float(np.abs(b.forward(toks[p], p) - strict[p]).max()))
            b.close()
        a.close()
        floors[f"{name}_{quant:02x}_{gs}"] = worst
    json.dump(floors, open(os.path.join(HERE, "reference_noise_floor.json"), "w"), indent=1)
    print("floors", floors)

    # Q4K KAT
    L = ob.RefEngine.lib("strict")
    d, n = 8, 768
    rnd = xorshift_f32(d * n + n)
    W = (rnd[: d * n] - np.float32(0.5)).astype(np.float32)
    x = (rnd[d * n:] - np.float32(0.5)).astype(np.float32)
    shape_w = (C.c_uint32 * 2)(d, n)
    shape_x = (C.c_uint32 * 1)(n)
    TW = L.quantize_tensor_q4k(W.ctypes.data_as(ob.f32p), 2, shape_w)
    TX = L.quantize_tensor_q4k(x.ctypes.data_as(ob.f32p), 1, shape_x)
    nbw = L.bytes_num_of_q4k_tensor(TW); nbx = L.bytes_num_of_q4k_tensor(TX)
    wb = np.ctypeslib.as_array(C.cast(TW, ob.u8p), shape=(nbw,)).copy()
    xb = np.ctypeslib.as_array(C.cast(TX, ob.u8p), shape=(nbx,)).copy()
    y = np.zeros(d, np.float32)
    L.matmul_q4k(y.ctypes.data_as(ob.f32p), TX, TW, 0)
    np.savez_compressed(os.path.join(HERE, "q4k_kat.npz"), W=W, x=x, w_tensor=wb, x_tensor=xb, y=y)
    print("q4k kat y", y)


if __name__ == "__main__":
    main()

"""Export a Nano checkpoint (the reference's PyTorch `GPT`, model.py:325-417) to a Nano model file: what export.py:228-513 does
(`export_model` F32, `export_quantized` Q80), restated over `modelfile.write_model_from_weights`.  Same header, tokenizer section,
tensor order and Q80 arithmetic, so the output is byte-identical to the reference exporter's for the same model
(tests/test_export_nano.py).  Format tooling, not part of the decode hot path.

    python -m nano_b200.export_nano checkpoint.pt nano.bin --quant q80      # needs the reference's model.py on PYTHONPATH to unpickle
"""
from __future__ import annotations

import argparse
from typing import Optional

import numpy as np

from . import modelfile as mf


def spec_from_nano_config(cfg, tied: bool, name: str = "nano-export") -> mf.ModelSpec:
    n_kv = cfg.n_kv_head if getattr(cfg, "n_kv_head", None) is not None else cfg.n_head
    return mf.ModelSpec(name, mf.ARCH_NANO, int(cfg.block_size), int(cfg.vocab_size), int(cfg.n_layer), int(cfg.n_embd), int(cfg.n_head), int(n_kv),
                        int(cfg.n_hidden), 0, 1 if tied else 0)


def weights_from_nano_state_dict(sd, spec: mf.ModelSpec) -> dict:
    """Reference module names -> writer names (export.py:298-317)."""
    def t(name):
        return sd[name].detach().to("cpu").float().numpy()

    def stack(fmt):
        return np.stack([t(fmt.format(i)) for i in range(spec.n_layer)])

    w = {"emb": t("tok_embeddings.weight"), "final_norm": t("norm.weight"),
         "attn_norm": stack("layers.{}.attention_norm.weight"), "ffn_norm": stack("layers.{}.ffn_norm.weight"),
         "wq": stack("layers.{}.attention.wq.weight"), "wk": stack("layers.{}.attention.wk.weight"),
         "wv": stack("layers.{}.attention.wv.weight"), "wo": stack("layers.{}.attention.wo.weight"),
         "w1": stack("layers.{}.feed_forward.w1.weight"), "w2": stack("layers.{}.feed_forward.w2.weight"),
         "w3": stack("layers.{}.feed_forward.w3.weight")}
    if not spec.tied:
        w["cls"] = t("output.weight")
    return w


def export_nano(model, out_path: str, quant: int = mf.QUANT_F32, gs: int = 128, tokenizer_config: Optional[dict] = None) -> dict:
    """`model`: the reference's GPT (or anything with .config and .state_dict() in its naming)."""
    sd = model.state_dict()
    tied = bool((sd["tok_embeddings.weight"] == sd["output.weight"]).all())             # export.py:261
    spec = spec_from_nano_config(model.config, tied)
    if quant == mf.QUANT_Q80:
        while spec.n_embd % gs:                                                         # export.py:399-401
            gs //= 2
    tok = mf.nano_tokenizer_section_from_config(tokenizer_config) if tokenizer_config else None
    rope = None
    if getattr(model, "freqs_cos", None) is not None:                                   # the checkpoint's own table (export.py:314-315)
        rope = (model.freqs_cos.detach().cpu().float().numpy(), model.freqs_sin.detach().cpu().float().numpy())
    return mf.write_model_from_weights(out_path, spec, quant, gs, weights_from_nano_state_dict(sd, spec), tok, rope=rope)


def export_lora(lora_state_dict, lora_rank: int, lora_alpha: int, base_config, out_path: str) -> dict:
    """A LoRA plug-in (the `lora` entry of a reference LoRA checkpoint: keys `layers.<l>.attention.w{q,k,v,o}.lora_{a,b}.weight`) in the
    layout parse_lora_file reads (infer.c:436-500), as export.py:117-226 writes it: 256-byte header, then per projection the A
    factors of all layers followed by the B factors."""
    import struct
    n_kv = base_config.n_kv_head if getattr(base_config, "n_kv_head", None) is not None else base_config.n_head
    hdr = struct.pack("<II", 0x42443453, 0x55524C4D) + struct.pack("<12i", 2024, 10, 10, 32, int(lora_rank), int(lora_alpha), int(base_config.n_layer),
                                                                  int(base_config.n_embd), int(base_config.n_head), int(n_kv), int(base_config.n_hidden), 0)
    hdr += b"\0" * (256 - len(hdr))
    parts = [hdr]
    for proj in ("wq", "wk", "wv", "wo"):
        for fac in ("lora_a", "lora_b"):
            keys = sorted((k for k in lora_state_dict if f"{proj}.{fac}" in k), key=lambda k: int(k.split(".")[1]))
            assert len(keys) == int(base_config.n_layer), f"{proj}.{fac}: {len(keys)} tensors for {base_config.n_layer} layers"
            for k in keys:
                parts.append(np.ascontiguousarray(lora_state_dict[k].detach().to("cpu").float().numpy(), np.float32).tobytes())
    blob = b"".join(parts)
    with open(out_path, "wb") as f:
        f.write(blob)
    return {"path": out_path, "bytes": len(blob)}


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("checkpoint"); ap.add_argument("out")
    ap.add_argument("--quant", default="f32", choices=["f32", "q80"])
    ap.add_argument("--group-size", type=int, default=128)
    a = ap.parse_args()
    import torch
    ck = torch.load(a.checkpoint, map_location="cpu", weights_only=False)               # export.py:478-512 load_checkpoint
    from model import GPT, ModelConfig                                                  # the reference's model.py
    cfg = ck["model_config"]
    model = GPT(ModelConfig(**cfg) if isinstance(cfg, dict) else cfg)
    model.load_state_dict({k.removeprefix("_orig_mod."): v for k, v in ck["model"].items()}, strict=False)
    model.eval()
    info = export_nano(model, a.out, mf.QUANT_F32 if a.quant == "f32" else mf.QUANT_Q80, a.group_size, ck.get("tokenizer_config"))
    print(f"wrote {info['path']}: {info['bytes']} bytes, {info['spec']}")


if __name__ == "__main__":
    main()

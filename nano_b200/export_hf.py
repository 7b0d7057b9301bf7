"""Export a Hugging Face Qwen3 checkpoint to a bd4sur/Nano model file (what infer/tools/export_qwen.py:639-750
does with its own model class): the tensors are taken from the HF state dict as they are -- the engine's `rope_qwen3`
(infer.c:692-711) rotates the two halves of a head exactly like HF's `rotate_half`, so no q/k permutation is needed
(export_qwen.py:694-697) -- and written in the reference's section order by `modelfile.write_model_from_weights`.

    python -m nano_b200.export_hf /path/to/Qwen3-0.6B qwen3-0b6.bin --quant q80 --group-size 128 [--max-seq-len 40960]

Format tooling, not part of the decode hot path.  There is no network in the build container, so the test
(tests/test_export_hf.py) exports a small randomly initialised `Qwen3ForCausalLM` and compares the exported file's logits
(oracle and the unmodified reference engine) with the HF model's own forward pass.  Two differences between the reference
engine and HF that an exported REAL checkpoint inherits: the engine's rmsnorm adds 1e-5 (infer.c:601-614) where Qwen3
checkpoints use rms_norm_eps = 1e-6, and the tokenizer section (the reference packs vocab.json / merges into it,
export_qwen.py:362-440) has to be supplied by the caller."""
from __future__ import annotations

import argparse
from typing import Optional

import numpy as np

from . import modelfile as mf


def spec_from_hf_config(cfg, max_seq_len: Optional[int] = None, name: str = "hf-qwen3") -> mf.ModelSpec:
    hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    return mf.ModelSpec(name, mf.ARCH_QWEN3, int(max_seq_len or cfg.max_position_embeddings), int(cfg.vocab_size), int(cfg.num_hidden_layers),
                        int(cfg.hidden_size), int(cfg.num_attention_heads), int(cfg.num_key_value_heads), int(cfg.intermediate_size), int(hd),
                        1 if getattr(cfg, "tie_word_embeddings", True) else 0)


def weights_from_hf_state_dict(sd, spec: mf.ModelSpec) -> dict:
    """HF names -> writer names (export_qwen.py:667-726)."""
    def t(name):
        return sd[name].detach().to("cpu").float().numpy()

    L = spec.n_layer
    def stack(fmt):
        return np.stack([t(fmt.format(i)) for i in range(L)])

    w = {"emb": t("model.embed_tokens.weight"), "final_norm": t("model.norm.weight"),
         "attn_norm": stack("model.layers.{}.input_layernorm.weight"), "ffn_norm": stack("model.layers.{}.post_attention_layernorm.weight"),
         "wq": stack("model.layers.{}.self_attn.q_proj.weight"), "wk": stack("model.layers.{}.self_attn.k_proj.weight"),
         "wv": stack("model.layers.{}.self_attn.v_proj.weight"), "wo": stack("model.layers.{}.self_attn.o_proj.weight"),
         "q_norm": stack("model.layers.{}.self_attn.q_norm.weight"), "k_norm": stack("model.layers.{}.self_attn.k_norm.weight"),
         "w1": stack("model.layers.{}.mlp.gate_proj.weight"), "w2": stack("model.layers.{}.mlp.down_proj.weight"),
         "w3": stack("model.layers.{}.mlp.up_proj.weight")}
    if not spec.tied:
        w["cls"] = t("lm_head.weight")
    return w


def export_qwen3(model, out_path: str, quant: int = mf.QUANT_F32, gs: int = 128, max_seq_len: Optional[int] = None,
                 tokenizer_section: Optional[bytes] = None) -> dict:
    """`model`: a transformers Qwen3ForCausalLM (or anything with .config and .state_dict() in HF naming)."""
    spec = spec_from_hf_config(model.config, max_seq_len)
    if quant == mf.QUANT_Q80:
        while spec.n_embd % gs or spec.q_dim % gs or spec.n_hidden % gs:       # export_qwen.py:570-572 backs the group size off the same way
            gs //= 2
    return mf.write_model_from_weights(out_path, spec, quant, gs, weights_from_hf_state_dict(model.state_dict(), spec), tokenizer_section)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("hf_path"); ap.add_argument("out")
    ap.add_argument("--quant", default="q80", choices=["f32", "q80", "q4k"])
    ap.add_argument("--group-size", type=int, default=128)
    ap.add_argument("--max-seq-len", type=int, default=None)
    a = ap.parse_args()
    from transformers import AutoModelForCausalLM
    model = AutoModelForCausalLM.from_pretrained(a.hf_path)
    info = export_qwen3(model, a.out, {"f32": mf.QUANT_F32, "q80": mf.QUANT_Q80, "q4k": mf.QUANT_Q4K}[a.quant], a.group_size, a.max_seq_len)
    print(f"wrote {info['path']}: {info['bytes']} bytes, {info['spec']}")


if __name__ == "__main__":
    main()

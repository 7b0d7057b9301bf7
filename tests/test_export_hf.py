"""CPU test of the Hugging Face Qwen3 -> Nano model-file export (nano_b200/export_hf.py; SURVEY section 8 row f1, VERDICT r1 missing item 7).
There are no checkpoints in this environment, so a small randomly initialised `Qwen3ForCausalLM` stands in for one: the exported
file is run by the oracle (and by the unmodified reference engine, when it is built here) and its logits are compared with the HF
model's own forward pass on the same tokens.  This pins the conventions an export of real weights relies on: tensor naming and
order, no q/k permutation with `rope_qwen3`, per-head q/k norm, theta = 1e6, the tied and the untied classifier."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")

from nano_b200 import export_hf, modelfile as mf
from oracle import bindings as ob

S = 14


def tiny_qwen3(tied: bool, seed: int = 3):
    from transformers import Qwen3Config, Qwen3ForCausalLM
    torch.manual_seed(seed)
    cfg = Qwen3Config(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=32, max_position_embeddings=64, rms_norm_eps=1e-5,      # the engine's rmsnorm constant (infer.c:601-614)
                      tie_word_embeddings=tied, attention_bias=False)
    rp = dict(getattr(cfg, "rope_parameters", None) or {})
    if rp:
        rp["rope_theta"] = 1000000.0; cfg.rope_parameters = rp              # transformers >= 5
    else:
        cfg.rope_theta = 1000000.0
    model = Qwen3ForCausalLM(cfg).float().eval()
    with torch.no_grad():                                                   # HF initialises every norm gain to 1: make them matter
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            elif p.dim() == 2:
                p.mul_(3.0)                                                 # larger logits than the 0.02-std initialisation gives
    return model


def hf_logits(model, toks):
    with torch.no_grad():
        out = model(input_ids=torch.tensor([list(map(int, toks))], dtype=torch.long)).logits[0]
    return out.float().numpy()


@pytest.mark.parametrize("tied", [True, False], ids=["tied", "untied"])
def test_exported_qwen3_matches_the_hf_forward(tmp_path, tied):
    model = tiny_qwen3(tied)
    toks = mf.teacher_tokens(S, 512)
    want = hf_logits(model, toks)                                           # [S, V]: position p sees tokens 0..p
    scale = float(np.abs(want).max())

    if tied:
        path = str(tmp_path / "tiny_f32.bin")
        info = export_hf.export_qwen3(model, path, mf.QUANT_F32)
        assert info["spec"].tied == 1 and info["spec"].head_dim == 32
        o = ob.NanoOracle(path, S)
        got = np.stack([o.forward(int(toks[p]), p) for p in range(S)])
        o.close()
        err = float(np.abs(got - want).max())
        print(f"F32 export vs HF forward: max|dlogit| {err:.3e} (logit scale {scale:.2f})")
        assert err < 2e-4 * max(1.0, scale), err
        if ob.ref_available("strict"):                                      # the unmodified reference engine parses the same file
            r = ob.RefEngine(path, S)
            ref = np.stack([r.forward(int(toks[p]), p) for p in range(S)])
            r.close()
            assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), "oracle != reference on the exported file"

    # Q80 export (the untied classifier exists only in Q80 files, infer.c:206-216): quantisation noise, not a convention error
    pathq = str(tmp_path / "tiny_q80.bin")
    infoq = export_hf.export_qwen3(model, pathq, mf.QUANT_Q80, 64)
    assert infoq["spec"].tied == (1 if tied else 0) and infoq["gs"] == 64
    o = ob.NanoOracle(pathq, S)
    gotq = np.stack([o.forward(int(toks[p]), p) for p in range(S)])
    o.close()
    errq = float(np.abs(gotq - want).max())
    agree = int((gotq.argmax(1) == want.argmax(1)).sum())
    print(f"Q80 export ({'tied' if tied else 'untied'}) vs HF forward: max|dlogit| {errq:.3e} (logit scale {scale:.2f}), argmax agreement {agree}/{S}")
    assert errq < 0.05 * max(1.0, scale), errq
    if ob.ref_available("strict"):
        r = ob.RefEngine(pathq, S)
        refq = np.stack([r.forward(int(toks[p]), p) for p in range(S)])
        r.close()
        assert np.array_equal(refq.view(np.uint32), gotq.view(np.uint32)), "oracle != reference on the exported Q80 file"

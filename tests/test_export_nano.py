"""CPU test of the Nano checkpoint export (nano_b200/export_nano.py): a small randomly initialised instance of the
reference's own PyTorch model (model.py, imported from /root/reference in this container only) is exported by OUR exporter and by the
reference's export.py -- the two files must be byte-identical (F32 and Q80) -- and the file's logits (oracle) must equal the PyTorch
model's forward pass."""
import importlib.util
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from nano_b200 import export_nano, modelfile as mf
from oracle import bindings as ob

REF = os.environ.get("NANO_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model.py")), reason="the reference tree is not present")
S = 12


@pytest.fixture(scope="module")
def ref_modules():
    sys.path.insert(0, REF)
    try:
        import model as ref_model                    # the reference's model.py
        spec = importlib.util.spec_from_file_location("nano_ref_export", os.path.join(REF, "export.py"))
        ref_export = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_export)
    finally:
        sys.path.remove(REF)
    return ref_model, ref_export


def tiny_gpt(ref_model, seed=5):
    torch.manual_seed(seed)
    cfg = ref_model.ModelConfig(block_size=32, vocab_size=160, n_layer=2, n_embd=128, n_head=4, n_kv_head=2, n_hidden=256)
    m = ref_model.GPT(cfg).float().eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            else:
                p.mul_(3.0)
    return m


def test_export_is_byte_identical_to_the_reference_exporter_and_matches_the_model(tmp_path, ref_modules, capsys):
    ref_model, ref_export = ref_modules
    m = tiny_gpt(ref_model)
    V = m.config.vocab_size
    tok = {"itos": [chr(0x4E00 + i) for i in range(V)], "vocab_size": V, "special_tokens": [chr(0x4E00), chr(0x4E01)]}
    toks = mf.teacher_tokens(S, V)
    with torch.no_grad():
        want = np.stack([m(torch.tensor([list(map(int, toks[: p + 1]))], dtype=torch.long))[0][0, -1].float().numpy() for p in range(S)])

    ours, theirs = str(tmp_path / "ours.bin"), str(tmp_path / "theirs.bin")
    export_nano.export_nano(m, ours, mf.QUANT_F32, tokenizer_config=tok)
    ref_export.export_model(m, tok, theirs)
    a, b = open(ours, "rb").read(), open(theirs, "rb").read()
    assert len(a) == len(b), (len(a), len(b))
    assert a == b, "F32 export differs from the reference exporter at byte %d" % next(i for i in range(len(a)) if a[i] != b[i])

    o = ob.NanoOracle(ours, S)
    got = np.stack([o.forward(int(toks[p]), p) for p in range(S)])
    o.close()
    err, scale = float(np.abs(got - want).max()), float(np.abs(want).max())
    print(f"exported F32 file vs the PyTorch model: max|dlogit| {err:.3e} (logit scale {scale:.2f})")
    assert err < 2e-4 * max(1.0, scale)

    oursq, theirsq = str(tmp_path / "ours_q80.bin"), str(tmp_path / "theirs_q80.bin")
    export_nano.export_nano(m, oursq, mf.QUANT_Q80, 64, tokenizer_config=tok)
    ref_export.export_quantized(m, tok, theirsq, group_size=64)
    a, b = open(oursq, "rb").read(), open(theirsq, "rb").read()
    assert len(a) == len(b), (len(a), len(b))
    assert a == b, "Q80 export differs from the reference exporter at byte %d" % next(i for i in range(len(a)) if a[i] != b[i])


def test_lora_export_is_byte_identical_and_the_reference_engine_reproduces_the_lora_model(tmp_path, ref_modules):
    """export.py:117-226 / infer.c:436-500 and the LoRA branches infer.c:792-808, :898-903: base file + plug-in run by the unmodified
    reference engine against the PyTorch model with the same low-rank branches attached."""
    ref_model, ref_export = ref_modules
    m = tiny_gpt(ref_model, seed=9)
    V = m.config.vocab_size
    tok = {"itos": [chr(0x4E00 + i) for i in range(V)], "vocab_size": V, "special_tokens": []}
    base = str(tmp_path / "base.bin")
    export_nano.export_nano(m, base, mf.QUANT_F32, tokenizer_config=tok)           # the base model, before the branches are attached
    m.to_lora(lora_rank=4, lora_alpha=8)
    torch.manual_seed(11)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_b" in n:
                p.copy_(0.05 * torch.randn_like(p))                                  # B starts at zero: make the branch contribute
    m.eval()
    lora_sd = m.get_lora_state_dict()
    ours, theirs = str(tmp_path / "lora_ours.bin"), str(tmp_path / "lora_theirs.bin")
    export_nano.export_lora(lora_sd, 4, 8, m.config, ours)
    ref_export.export_lora(lora_sd, {"lora_rank": 4, "lora_alpha": 8}, m.config, theirs)
    a, b = open(ours, "rb").read(), open(theirs, "rb").read()
    assert a == b, "LoRA export differs from the reference exporter"
    if not ob.ref_available("strict"):
        pytest.skip("oracle/_ref not built: byte identity checked only")
    toks = mf.teacher_tokens(S, V)
    with torch.no_grad():
        want = np.stack([m(torch.tensor([list(map(int, toks[: p + 1]))], dtype=torch.long))[0][0, -1].float().numpy() for p in range(S)])
    r = ob.RefEngine(base, S)
    r.load_lora(a)
    got = np.stack([r.forward(int(toks[p]), p) for p in range(S)])
    r.close()
    err, scale = float(np.abs(got - want).max()), float(np.abs(want).max())
    print(f"reference engine (base file + exported plug-in) vs the PyTorch LoRA model: max|dlogit| {err:.3e} (logit scale {scale:.2f})")
    assert err < 2e-4 * max(1.0, scale)

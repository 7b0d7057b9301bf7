"""GPU tests through the REFERENCE's host API (libnano_infer_b200.so: llm_context_init, generate_next_token,
llm_session_step, seq2seq ...) -- the calls nano_cli / nano_sort / nano_wss make -- against the unmodified
reference (oracle/_ref) or, when that did not travel, the oracle port."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from nano_b200 import build as nb_build, modelfile as mf
from oracle import bindings as ob

pytestmark = pytest.mark.gpu

HAVE_REFHOST = os.path.exists(nb_build.REFHOST_SO)
needs_refhost = pytest.mark.skipif(not HAVE_REFHOST, reason="libnano_refhost.so (reference tokenizer/utils objects) not built")


class Session(C.Structure):       # infer.h:236-250
    _fields_ = [("prompt", C.c_void_p), ("num_prompt_tokens", C.c_uint32), ("max_seq_len", C.c_uint32), ("output_ids", C.POINTER(C.c_uint32)),
                ("output_count", C.c_uint32), ("output_text", C.c_wchar_p), ("next_token", C.c_uint32), ("pos", C.c_uint32),
                ("is_prefilling", C.c_int32), ("t_0", C.c_uint64), ("t_1", C.c_uint64), ("tps", C.c_float)]


_shim = None


def shim():
    global _shim
    if _shim is None:
        C.CDLL(nb_build.REFHOST_SO, mode=C.RTLD_GLOBAL)        # the reference's own tokenizer.c/utils.c/hal objects
        L = C.CDLL(nb_build.SHIM_SO, mode=C.RTLD_GLOBAL)
        L.llm_context_init.restype = C.c_void_p
        L.llm_context_init.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint64]
        L.llm_context_init_from_buffer.restype = C.c_void_p
        L.llm_context_init_from_buffer.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint64]
        L.llm_context_free.argtypes = [C.c_void_p]
        L.generate_next_token.restype = C.c_uint32
        L.generate_next_token.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_int]
        L.seq2seq.argtypes = [C.c_void_p, C.c_wchar_p, C.c_wchar_p, C.c_uint32]
        L.llm_session_init.restype = C.POINTER(Session)
        L.llm_session_init.argtypes = [C.c_void_p, C.c_wchar_p, C.c_uint32, C.c_int32]
        L.llm_session_step.restype = C.c_int32
        L.llm_session_step.argtypes = [C.c_void_p, C.POINTER(Session)]
        L.llm_session_free.argtypes = [C.POINTER(Session)]
        _shim = L
    return _shim


@needs_refhost
def test_sort_demo_through_reference_api(sort_model):
    """main_sort.c's calls: llm_context_init_from_buffer(model, 6, 0, 0, 0, 1, 39) + seq2seq (README.md:379)."""
    L = shim()
    img = np.frombuffer(sort_model, np.uint8).copy()
    ctx = L.llm_context_init_from_buffer(img.ctypes.data, 6, 0.0, 0.0, 0.0, 1, 39)
    kat = json.load(open(os.path.join(GOLDEN, "sort6_kat.json")))
    for src, want in kat.items():
        out = C.create_unicode_buffer(16)
        L.seq2seq(ctx, src, out, 6)
        assert out.value == want, (src, out.value)


@needs_refhost
@pytest.mark.parametrize("penalty", [1.0, 1.3])
def test_generate_next_token_greedy_matches_reference(penalty, monkeypatch):
    """generate_next_token(ctx, ids, pos, is_prefilling) with temperature 0: ids identical to the oracle (exact mode)."""
    monkeypatch.setenv("NB200_EXACT", "1")
    spec = mf.PRESETS["toy-qwen3"]
    path = mf.cached_model(spec, mf.QUANT_Q80, 64)
    S, P = 32, 5
    L = shim()
    ctx = L.llm_context_init(path.encode(), None, S, penalty, 0.0, 0.8, 20, 39)
    ids = (C.c_uint32 * (S + 1))(); want = np.zeros(S + 1, np.uint32)
    for i, t in enumerate([7, 8, 9, 8, 7]):
        ids[i] = t; want[i] = t
    o = ob.NanoOracle(path, S)
    for pos in range(S - 1):
        pre = 1 if pos < P - 1 else 0
        ids[pos + 1] = L.generate_next_token(ctx, ids, pos, pre)
        want[pos + 1] = o.next_greedy(want, pos, pre, penalty)
    assert list(ids)[:S] == want[:S].tolist()
    L.llm_context_free(ctx)
    o.close()


@needs_refhost
@pytest.mark.skipif(not ob.ref_available("strict"), reason="needs the unmodified reference for the sampling path")
def test_temperature_sampling_matches_reference(monkeypatch):
    """temperature 0.7 / top-p 0.8 (main_cli.c:227): logits go to the host and the reference's sampler is restated
    there with the reference's own xorshift coin -- same ids as the reference when the logits are bit-identical."""
    monkeypatch.setenv("NB200_EXACT", "1")
    spec = mf.PRESETS["toy-nano"]
    path = mf.cached_model(spec, mf.QUANT_Q80, 64)
    S, P = 28, 4
    L = shim()
    ctx = L.llm_context_init(path.encode(), None, S, 1.1, 0.7, 0.8, 20, 39)
    ref = ob.RefEngine(path, S, "strict", penalty=1.1, temperature=0.7, top_p=0.8, top_k=20, seed=39)
    ids = (C.c_uint32 * (S + 1))(); want = np.zeros(S + 1, np.uint32)
    for i, t in enumerate([17, 18, 19, 20]):
        ids[i] = t; want[i] = t
    for pos in range(S - 1):
        pre = 1 if pos < P - 1 else 0
        ids[pos + 1] = L.generate_next_token(ctx, ids, pos, pre)
        want[pos + 1] = ref.next(want, pos, pre)
    assert list(ids)[:S] == want[:S].tolist()
    L.llm_context_free(ctx)


@needs_refhost
def test_session_api_runs_and_stops_at_length():
    spec = mf.PRESETS["toy-nano"]
    path = mf.cached_model(spec, mf.QUANT_Q80, 64)
    L = shim()
    ctx = L.llm_context_init(path.encode(), None, 24, 1.0, 0.0, 0.8, 20, 39)
    s = L.llm_session_init(ctx, "0123", 24, 0)
    assert s.contents.num_prompt_tokens == 4
    statuses = []
    for _ in range(30):
        st = L.llm_session_step(ctx, s)
        statuses.append(st)
        if st <= 0:
            break
    assert statuses[:3] == [11, 11, 11] and 12 in statuses          # LLM_RUNNING_IN_PREFILLING x3, then decoding
    assert statuses[-1] in (-10, -20)                                # EOS (ids 0/3) or length limit (infer.h:52-55)
    L.llm_session_free(s)
    L.llm_context_free(ctx)


@pytest.mark.skipif(not os.path.exists(nb_build.NANO_CLI), reason="nano_cli not linked on this box")
def test_nano_cli_binary_starts_and_reports_missing_model():
    """The unmodified main_cli.c linked against our libraries runs; its model path is hard-coded (main_cli.c:12),
    so without that file it must exit through the reference's own error path."""
    r = subprocess.run([nb_build.NANO_CLI], input=b"", capture_output=True, timeout=60)
    assert r.returncode != 0
    assert b"open" in r.stderr.lower() or b"couldn" in r.stderr.lower()

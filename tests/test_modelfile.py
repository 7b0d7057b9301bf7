import numpy as np

from nano_b200 import modelfile as mf


def test_weight_counts_match_survey():
    # W = N168 168,296,448 / Q06 595,984,384 / Q4B 4,022,272,000
    assert mf.PRESETS["nano-168m"].n_weights() == 168_296_448
    assert mf.PRESETS["qwen3-0.6b"].n_weights() == 595_984_384
    assert mf.PRESETS["qwen3-4b"].n_weights() == 4_022_272_000


def test_bytes_per_token_matches_appendix_c():
    n = mf.PRESETS["nano-168m"]
    avg = n.bytes_per_token(mf.QUANT_Q80, 128, (16 + 512 + 1) / 2 - 1)
    assert abs(avg / 1e6 - 192.7) < 1.0
    q = mf.PRESETS["qwen3-0.6b"]
    assert abs(q.bytes_per_token(mf.QUANT_Q80, 128, (16 + 2048 + 1) / 2 - 1) / 1e6 - 850.1) < 2.0
    assert abs(q.bytes_per_token(mf.QUANT_Q4K, 0, (16 + 2048 + 1) / 2 - 1) / 1e6 - 580.0) < 2.0


def test_teacher_tokens_are_the_xorshift_stream():
    t = mf.teacher_tokens(4, 1 << 31)
    # utils.c:959-965 with state 39
    st = 39; M = (1 << 64) - 1; out = []
    for _ in range(4):
        st ^= st >> 12; st ^= (st << 25) & M; st ^= st >> 27
        out.append(((st * 0x2545F4914F6CDD1D) & M) >> 32)
    assert [int(v) for v in t] == [v % 151669 for v in out]


def test_q80_quantiser_roundtrip():
    rng = np.random.default_rng(0)
    w = rng.standard_normal(1024, dtype=np.float32)
    q, s = mf.quantize_q80(w, 128)
    assert q.dtype == np.int8 and np.abs(q).max() == 127
    assert np.abs(q.reshape(-1, 128) * s[:, None] - w.reshape(-1, 128)).max() <= s.max() * 0.5 + 1e-7

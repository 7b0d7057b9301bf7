"""GPU parity on the BASELINE model shapes themselves: Nano-168M Q80 (24 layers) and Qwen3-0.6B Q80 / Q4K
(28 layers, V = 151 936), 40 teacher-forced positions against the strict oracle: logits within the fast-mode policy, greedy ids
agreeing wherever the oracle's margin is real (margins printed), on the default execution path of each model.
"""
import os

import numpy as np
import pytest

from nano_b200 import engine as E, modelfile as mf
from oracle import bindings as ob

pytestmark = pytest.mark.gpu

# The floor = the reference's own -O3 -ffast-math vs strict deviation on the same file over the same 40 positions, measured with the
# unmodified reference in the build container and committed in tests/golden/reference_noise_floor.json (Qwen3-0.6B Q4K: 0.83, with the
# reference's two builds agreeing on only 27 of 40 greedy ids -- SURVEY finding 11 at full scale).
CASES = [("nano-168m", mf.QUANT_Q80, 128), ("qwen3-0.6b", mf.QUANT_Q80, 128), ("qwen3-0.6b", mf.QUANT_Q4K, 128)]


@pytest.mark.parametrize("name,quant,gs", CASES)
def test_baseline_shape_logits_and_ids(name, quant, gs):
    import json
    from conftest import GOLDEN
    floor = json.load(open(os.path.join(GOLDEN, "reference_noise_floor.json")))[f"{name}_{quant:02x}_{gs}"]
    spec = mf.PRESETS[name]
    path = mf.cached_model(spec, quant, gs)
    S = 40
    eng = E.Engine(path, S)
    if quant == mf.QUANT_Q80 and name == "qwen3-0.6b":
        # a large one-GPU Q80 engine times both fast paths when it is created and keeps the faster one; the 40 positions below
        # then also prove that the calibration steps (run at positions S/2..) left nothing behind
        cal = eng.calibration
        assert cal and cal["streaming_us_per_token"] > 0 and cal["multikernel_us_per_token"] > 0, cal
        faster_is_stream = cal["streaming_us_per_token"] <= cal["multikernel_us_per_token"]
        assert eng.path.startswith("streaming") == faster_is_stream, (eng.path, cal)
        print("calibration:", cal)
    else:
        assert eng.calibration is None
    o = ob.NanoOracle(path, S)
    ob.NanoOracle.lib().nor_set_threads(min(32, os.cpu_count() or 1))
    toks = mf.teacher_tokens(S, spec.vocab)
    limit = max(1e-2, 1.5 * floor)
    worst, agree, real, margins = 0.0, 0, 0, []
    for pos in range(S):
        a = eng.forward(toks[pos], pos); b = o.forward(toks[pos], pos)
        worst = max(worst, float(np.abs(a - b).max()))
        top2 = np.partition(b, -2)[-2:]
        m = float(top2[1] - top2[0]); margins.append(m)
        same = int(np.argmax(a)) == int(np.argmax(b))
        agree += same
        if m > 2 * limit:
            real += 1
            assert same, f"{name} pos {pos}: argmax differs although the oracle margin is {m}"
    print(f"{name} {quant:#x} [{eng.path[:24]}]: max|dlogit| {worst:.3e} (limit {limit:.3e}), argmax agreement {agree}/{S}, "
          f"{real} positions with a decisive margin, median margin {np.median(margins):.3e}")
    assert worst <= limit, f"{name} {quant:#x}: {worst} > {limit}"
    eng.close(); o.close()

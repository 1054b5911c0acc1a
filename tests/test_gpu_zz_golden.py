"""GPU test against the committed golden fixtures (tests/golden/oracle_golden.json) with no oracle in the loop.  Kept in its
own module (collected last) so that the oracle-based parity suites run first."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gpu_matches_committed_golden_fixtures(pkg, make_model):
    """The CUDA path against tests/golden/oracle_golden.json directly (no oracle in the loop): argmax per step and the
    SHA-256 of all logits bytes of 12 teacher-forced steps, for the seeded tiny models of the fixture (Llama, Qwen3, Phi-3; Q8_0 and FP16)."""
    import hashlib
    import json
    import os

    import test_oracle

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")) as f:
        gold = json.load(f)
    for shape, quant, lanes in test_oracle._golden_cases(pkg):
        key = f"{shape}/{pkg.gguf.GGMLType.NAMES[quant]}/lanes{lanes}"
        m = make_model(shape, quant, 24, seed=1234)
        plan = pkg.B200MasterPlan.initialize_plan(m, fp16_lanes=lanes)
        try:
            h, toks = hashlib.sha256(), []
            for pos, tok in enumerate(gold[key]["input"]):
                lg, am = plan.forward_decode(int(tok), pos)
                h.update(np.ascontiguousarray(lg, dtype=np.float32).tobytes())
                toks.append(int(am))
            assert toks == gold[key]["argmax"], key
            assert h.hexdigest() == gold[key]["logits_sha256"], key
        finally:
            plan.free()


def test_gpu_requantiser_matches_committed_golden_fixtures(pkg):
    """The device K-quant -> Q8_0 re-quantiser (csrc/kquant.cuh) against the committed hashes, no oracle in the loop."""
    import hashlib
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")) as f:
        gold = json.load(f)["kquant_to_q8_0"]
    for name in ("Q4_K", "Q5_K", "Q6_K"):
        tt = getattr(pkg.gguf.GGMLType, name)
        raw = pkg.synth.random_kquant(tt, 256 * 64, np.random.Generator(np.random.PCG64(tt)), zero_blocks=1)
        assert hashlib.sha256(raw.tobytes()).hexdigest() == gold[name]["src_sha256"], name  # the seeded input itself
        assert hashlib.sha256(pkg.native.requant_kquant(tt, raw, 256 * 64).tobytes()).hexdigest() == gold[name]["q8_0_sha256"], name

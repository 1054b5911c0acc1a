"""GPU test against the committed golden fixtures (tests/golden/oracle_golden.json) with no oracle in the loop.  Kept in its
own module (collected last) so that the oracle-based parity suites run first."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gpu_matches_committed_golden_fixtures(pkg, make_model):
    """The CUDA path against tests/golden/oracle_golden.json directly (no oracle in the loop): argmax per step and the
    SHA-256 of all logits bytes of 12 teacher-forced steps, for the six seeded tiny models of the fixture."""
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

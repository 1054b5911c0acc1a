"""the device-side temperature / top-p sampler (csrc/sampler.cuh behind b200_forward_decode_sample) against the
oracle's restatement of Sampler.java / CategoricalSampler.java / ToppSampler.java: same logits (bit-exact forward), same uniform
number -> same token id, for greedy, categorical and top-p sampling."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", ["tiny-llama", "small-llama"])
def test_device_sampler_matches_oracle(pkg, orc, make_model, shape):
    m = make_model(shape, pkg.gguf.GGMLType.Q8_0, 48)
    c = m.configuration
    plan = pkg.B200MasterPlan.initialize_plan(m)
    om = orc.OracleModel(m)
    stream = orc.bench_tokens(c.vocab_size, 40)
    rng = orc.JavaLXM(12345)
    cases = [(0.0, 0.95), (1.0, 0.0), (0.7, 1.0), (1.0, 0.95), (0.1, 0.95), (1.3, 0.5), (0.8, 0.9), (2.0, 0.99)]
    try:
        for pos in range(32):
            temp, topp = cases[pos % len(cases)]
            r = rng.next_float1()
            tok = int(stream[pos])
            got, info = plan.forward_decode_sample(tok, pos, temp, topp, r, want_info=True)
            ref_logits = om.forward(tok, pos)
            want = orc.sample(ref_logits, temp, topp, r)
            assert got == want, (pos, temp, topp, r, got, want, info)
            if temp > 0 and 0 < topp < 1:
                assert 0 < info[1] <= info[0] <= c.vocab_size
    finally:
        plan.free()
        om.close()


def test_sampler_object_drives_a_generation(pkg, orc, make_model):
    """Sampler.selectSampler's object over the plan vs the same loop over the oracle with the oracle's RNG restatement."""
    m = make_model("tiny-llama", pkg.gguf.GGMLType.Q8_0, 40)
    plan = pkg.B200MasterPlan.initialize_plan(m)
    om = orc.OracleModel(m)
    smp = pkg.sampler.select_sampler(m.configuration.vocab_size, 0.8, 0.95, 42)
    rng = orc.JavaLXM(42)
    tok_g = tok_o = 3
    for pos in range(24):
        tok_g = smp.sample_token(plan, tok_g, pos)
        tok_o = orc.sample(om.forward(tok_o, pos), 0.8, 0.95, rng.next_float1())
        assert tok_g == tok_o, pos
    plan.free()
    om.close()

"""Regenerates tests/golden/oracle_golden.json from the oracle (python tests/golden/make_golden.py).
The reference holds no golden vectors for this path and cannot run here (no JDK), so these
pin the ORACLE against drift; they are not outputs of the reference itself."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import test_oracle  # noqa: E402

pkg, orc = ge.import_package(), ge.import_oracle()
orc.build()
with tempfile.TemporaryDirectory() as d:
    cache = {}

    def make_model(shape, quant, ctx=64, seed=1234):
        key = (shape, quant, seed)
        if key not in cache:
            p = os.path.join(d, f"{shape}-{quant}-{seed}.gguf")
            pkg.synth.write_model(p, shape, quant, seed=seed)
            cache[key] = p
        return pkg.load_model(cache[key], ctx)

    out = {}
    for shape, quant, lanes in test_oracle._golden_cases(pkg):
        out[f"{shape}/{pkg.gguf.GGMLType.NAMES[quant]}/lanes{lanes}"] = test_oracle.golden_run(orc, pkg, make_model, shape, quant, lanes)
    out["kquant_to_q8_0"] = test_oracle.kquant_golden(orc, pkg)
with open(os.path.join(os.path.dirname(__file__), "oracle_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))

"""GPU diagnostic: parity of one model vs oracle for a few tokens under the current env toggles."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg, orc = ge.import_package(), ge.import_oracle()
shape = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sh = pkg.synth.SHAPES[shape]
m = pkg.loader.model_from_tensors(sh, 8, pkg.synth.build_tensors_fast(sh, 8, seed=3), 32)
plan = pkg.B200MasterPlan.initialize_plan(m); om = orc.OracleModel(m)
toks = pkg.llama_bench.synthetic_tokens(sh.vocab, n)
ok = True
for i in range(n):
    lg, am = plan.forward_decode(toks[i], i); ref = om.forward(toks[i], i)
    same = np.array_equal(lg.view(np.uint32), ref.view(np.uint32))
    ok &= same
    print(f"{shape} budget={os.environ.get('B200_SMV_BUDGET_KB','-')} pos {i}: bitexact={same} maxdiff={np.abs(lg-ref).max():.3e}", flush=True)
print("RESULT", "OK" if ok else "MISMATCH")

"""GPU diagnostic: time plan creation and decode steps under the STREAM/PDL toggles (tiny + small shapes)."""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.import_package()
shape = sys.argv[1] if len(sys.argv) > 1 else "small-llama"
with tempfile.TemporaryDirectory() as d:
    sh = pkg.synth.SHAPES[shape]
    m = pkg.loader.model_from_tensors(sh, 8, pkg.synth.build_tensors_fast(sh, 8, seed=3), 64)
    t0 = time.time(); plan = pkg.B200MasterPlan.initialize_plan(m); t1 = time.time()
    print(f"stream={os.environ.get('B200_STREAM','1')} pdl={os.environ.get('B200_PDL','1')} create {t1-t0:.3f}s", flush=True)
    toks = pkg.llama_bench.synthetic_tokens(m.configuration.vocab_size, 40)
    for i in range(3):
        t0 = time.time(); plan.forward_decode(toks[i], i); print(f"  decode {i}: {(time.time()-t0)*1e3:.2f} ms", flush=True)
    t0 = time.time(); ids, ms = plan.decode_sequence(toks[3:35], 32, 3); t1 = time.time()
    print(f"  decode_sequence 32: device {ms:.3f} ms wall {(t1-t0)*1e3:.2f} ms -> {ms/32*1e3:.1f} us/token", flush=True)
    print('  norm phase cycles [wait, load, seqsum, norm+quant]:', plan._native.profile_norm(), flush=True)
    plan.free()

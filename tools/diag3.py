import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.import_package()
rng = np.random.default_rng(0)
for n in (256, 4096, 8192):
    for k in range(4):
        x = rng.standard_normal(n).astype(np.float32)
        t = (x * x).astype(np.float32)
        v, nseg, fb = pkg.native.test_seqsum(t, True)
        ref = np.add.accumulate(t, dtype=np.float32)[-1]
        print(n, "entries", nseg, "fallback_from", fb, "exact", np.float32(v) == ref)

"""Checks the tcgen05 prefill GEMM against an fp32 matmul of the same fp16 operands and times it.
usage: python tools/gemm_check.py [--big]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
from gpullama3_java_b200 import native  # noqa: E402

rng = np.random.default_rng(7)
out = []
shapes = [(128, 128, 64), (256, 384, 512), (512, 4096, 4096)]
if "--big" in sys.argv:
    shapes += [(512, 6144, 4096), (512, 28672, 4096), (512, 4096, 14336), (4096, 4096, 4096), (8192, 8192, 8192)]
two = int(os.environ.get("B200_GEMM_2CTA", "0"))
for (m, n, k) in shapes:
    bn = 256 if two in (512, 1256) else two  # 512: two pair tiles per CTA pair; 1256: persistent pair kernel
    if two and (m % (512 if two == 512 else 256) or n % bn):
        continue
    a = (rng.standard_normal((m, k)) * 0.5).astype(np.float16)
    b = (rng.standard_normal((n, k)) * 0.5).astype(np.float16)
    c, ms = native.gemm_f16(a, b, iters=20)
    if m * n * k <= 512 * 4096 * 4096:
        ref = a.astype(np.float32) @ b.astype(np.float32).T
        err = float(np.max(np.abs(c - ref)))
        scale = float(np.max(np.abs(ref)))
    else:  # spot check 64 rows
        rows = rng.integers(0, m, 64)
        ref = a[rows].astype(np.float32) @ b.astype(np.float32).T
        err = float(np.max(np.abs(c[rows] - ref)))
        scale = float(np.max(np.abs(ref)))
    rec = {"two_cta": two, "resid": os.environ.get("B200_GEMM_RESID", "0"), "stages": os.environ.get("B200_GEMM_STAGES", "3"), "m": m, "n": n, "k": k, "max_abs_err": err, "ref_max": scale, "ms": ms, "tflops": 2.0 * m * n * k / (ms * 1e-3) / 1e12}
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_check_s%s_2cta%d_r%s.json" % (os.environ.get("B200_GEMM_STAGES", "3"), two, os.environ.get("B200_GEMM_RESID", "0")), "w"), indent=1)

"""GPU probe: FP64 issue-rate micro-benchmarks and a per-stage profile of the config-2 pipeline."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200._capi import PM_FLAG_SIMPLE_KERNELS  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialXtX  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

n_st = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = {}
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
acc = PotentialXtX(pd)
ctx = acc.context
if "--no-micro" not in sys.argv:
    for which, name in ((0, "dfma"), (1, "dmma"), (2, "mixed"), (3, "dgemm8192")):
        out[name + "_tflops"] = ctx.microbench(which, 8192)
    print(json.dumps(out), flush=True)
sts = [cases.fcc_supercell(seed=20240 + s) for s in range(n_st)]
axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
rows = n_st * 775
rng = np.random.default_rng(0)
w = rng.uniform(0.2, 1.0, rows)
y = w * rng.normal(size=rows)
for flags, name in ((0, "dmma"), (PM_FLAG_SIMPLE_KERNELS, "simple")):
    if name == "simple" and "--no-simple" in sys.argv:
        continue
    a = PotentialXtX(pd, flags=flags)
    a.stage(axis, pcs, tys, [True] * n_st, w, y)
    a.add_staged()  # warm-up
    a.context.synchronize()
    a.context.profile(True)
    t0 = time.time()
    a.add_staged()
    a.context.synchronize()
    dt = time.time() - t0
    prof = a.context.profile_get()
    a.context.profile(False)
    t0 = time.time()
    reps = 3
    for _ in range(reps):
        a.add_staged()
    a.context.synchronize()
    dt2 = (time.time() - t0) / reps
    print(name, "profiled pass %.1f ms; unprofiled %.1f ms/pass = %.1f structures/s" % (dt * 1e3, dt2 * 1e3, n_st / dt2))
    for k, (ms, ln) in prof.items():
        print("   %-12s %9.3f ms  (%.1f us/structure)" % (k, ms, ms * 1e3 / n_st))

"""Latency of single-structure evaluations (the MD / relaxation use of PypolymlpCalc.eval): ms per pm_eval call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

pd = make_params_dict(**cases.cfg2_model_kwargs(4))
prop = PotentialPropertiesFast(pd, np.random.default_rng(12).normal(size=2030) * 1e-3)
for rep, label in (((2, 2, 2), "32 atoms"), ((4, 4, 4), "256 atoms"), ((4, 4, 8), "512 atoms"), ((8, 8, 8), "2048 atoms")):
    st = cases.fcc_supercell(rep=rep, sigma=0.03, seed=1)
    prop.eval(*st, True)
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        prop.eval(*st, True)
    dt = (time.perf_counter() - t0) / n
    print("%-10s %.3f ms per eval  (%.2f M atoms/s)" % (label, dt * 1e3, st[1].shape[1] / dt * 1e-6))
    if os.environ.get("PM_DEBUG_STAGES"):
        ctx = prop._ctx
        ctx.profile(True)
        prop.eval(*st, True)
        ctx.synchronize()
        print("   ", {k: round(v[0], 3) for k, v in ctx.profile_get().items() if v[0] > 0})
        ctx.profile(False)

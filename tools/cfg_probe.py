"""Stage profile and throughput of the fused feature + X^T X build on the large BASELINE models:
config 3 (binary, order 4, maxl [12,8,2], F = 9385, 216-atom bcc 6x6x3) and config 4 (ternary, F = 45090,
128-atom bcc 4x4x4).  Usage: python tools/cfg_probe.py 3|4 [n_structures]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialXtX  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_st = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if cfg == 3:
    kw, rep, nt = cases.cfg3_model_kwargs(), (6, 6, 3), 2
elif cfg == 4:
    kw, rep, nt = cases.cfg4_model_kwargs(), (4, 4, 4), 3
else:
    kw, rep, nt = cases.cfg2_model_kwargs(4), None, 1
t0 = time.time()
pd = make_params_dict(**kw)
acc = PotentialXtX(pd)
print("model tables + context: %.1f s, F = %d" % (time.time() - t0, acc.model.info()["n_variables"]), flush=True)
if cfg in (3, 4):
    sts = [cases.bcc_supercell(rep=rep, a=3.2, n_type=nt, seed=100 + s) for s in range(n_st)]
else:
    sts = [cases.fcc_supercell(seed=20240 + s) for s in range(n_st)]
axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
n_atoms = pcs[0].shape[1]
rows = n_st * (1 + 6 + 3 * n_atoms)
rng = np.random.default_rng(0)
w = rng.uniform(0.2, 1.0, rows)
y = w * rng.normal(size=rows)
acc.stage(axis, pcs, tys, [True] * n_st, w, y)
acc.add_staged()  # warm-up (allocations)
acc.context.synchronize()
acc.context.profile(True)
acc.add_staged()
acc.context.synchronize()
prof = acc.context.profile_get()
acc.context.profile(False)
reps = 2
t0 = time.time()
for _ in range(reps):
    acc.add_staged()
acc.context.synchronize()
dt = (time.time() - t0) / reps
out = {"config": cfg, "n_structures": n_st, "atoms": n_atoms, "ms_per_pass": dt * 1e3, "structures_per_s": n_st / dt,
       "stage_ms": {k: ms for k, (ms, ln) in prof.items() if ms > 0}}
print(json.dumps(out))

"""Eval (BASELINE config 5) probe: atoms/s through pm_eval with host buffers, stage profile, A/B switches via env."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200._capi import StructureBatch, check, lib  # noqa: E402
from pypolymlp_b200._capi import pd as pd_  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
sts = [cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777 + k) for k in range(n_ev)]
prop = PotentialPropertiesFast(pd, np.random.default_rng(12).normal(size=2030) * 1e-3)
batch = StructureBatch([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts], [True] * n_ev)
out = (np.zeros(n_ev), np.zeros((n_ev * 512, 3)), np.zeros((n_ev, 6)))


def step():
    check(lib().pm_eval(prop._ctx.handle, ctypes.byref(batch.c), pd_(out[0]), pd_(out[1]), pd_(out[2])))


step()
t0 = time.perf_counter()
for _ in range(3):
    step()
dt = (time.perf_counter() - t0) / 3
print("eval %d x 512 atoms: %.2f ms/call = %.2f M atoms/s" % (n_ev, dt * 1e3, n_ev * 512 / dt * 1e-6))
ctx = prop._ctx
ctx.profile(True)
step()             # (one lane, other chunk sizes than the timed calls: this call grows the chunk buffers)
ctx.synchronize()
ctx.profile(True)  # resets the stage sums
step()
ctx.synchronize()
for k, (ms, ln) in ctx.profile_get().items():
    if ms > 0:
        print("   %-12s %8.3f ms" % (k, ms))
ctx.profile(False)
G2 = np.load(os.path.join(cases.GOLDEN, "ref_vectors_r02.npz"))
print("   parity vs reference golden (structure 0): dE %.2e dF %.2e" % (
    abs(out[0][0] - G2["cfg5_e"][0]) / abs(G2["cfg5_e"][0]), np.abs(out[1][:512] - G2["cfg5_f"]).max() / np.abs(G2["cfg5_f"]).max()))

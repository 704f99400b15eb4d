"""atoms/s of the E/F/S evaluation (BASELINE config 5: F=2030 model, 512-atom fcc supercells, sigma 0.03 A)."""
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

n_st = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
coeffs = np.random.default_rng(12).normal(size=2030) * 1e-3
sts = [cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777 + s) for s in range(n_st)]
axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
prop = PotentialPropertiesFast(pd, coeffs)
prop._ctx.profile(False)
prop.eval_multiple(axis, pcs, tys)   # warm-up at full size (device buffers and pinned staging grow on first use)
t0 = time.perf_counter()
prop.eval_multiple(axis, pcs, tys)
dt = time.perf_counter() - t0
print(f"GPU eval_multiple: {n_st} x 512 atoms in {dt * 1e3:.1f} ms -> {n_st * 512 / dt:.3e} atoms/s (host lists in, host arrays out)")
prop._ctx.profile(True)
prop.eval_multiple(axis, pcs, tys)
for k, (ms, ln) in prop._ctx.profile_get().items():
    if ms > 0:
        print("   %-12s %9.3f ms" % (k, ms))
# the pybind11 drop-in (lists of NumPy arrays in through the buffer protocol, NumPy arrays out)
import importlib.util  # noqa: E402

from pypolymlp_b200.build import pybind_module_path  # noqa: E402

spec = importlib.util.spec_from_file_location("libmlpcpp", pybind_module_path())
libmlpcpp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(libmlpcpp)
pp = libmlpcpp.PotentialPropertiesFast(pd, coeffs)
pp.eval_multiple(axis, pcs, tys)
t0 = time.perf_counter()
pp.eval_multiple(axis, pcs, tys)
ea_pb = np.array(pp.get_e_array())
fa_pb = [np.array(f) for f in pp.get_f_array()]
dt = time.perf_counter() - t0
print(f"pybind drop-in eval_multiple + getters: {dt * 1e3:.1f} ms -> {n_st * 512 / dt:.3e} atoms/s")
from oracle import ref  # noqa: E402

if ref.available():
    ev = ref.RefEval(pd, coeffs)
    t0 = time.perf_counter()
    for k in range(2):
        e, f, s = ev.eval(axis[k], pcs[k], tys[k], use_openmp=True)
    dt = (time.perf_counter() - t0) / 2
    print(f"reference CPU eval (OpenMP over atoms, {ref.num_threads()} threads): {dt * 1e3:.1f} ms/structure -> {512 / dt:.3e} atoms/s")
    ea = np.array(prop.get_e_array())
    print("E parity:", abs(ea[1] - e) / abs(e), "F parity:", np.abs(np.array(prop.get_f_array()[1]) - f).max() / np.abs(f).max())

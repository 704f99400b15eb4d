"""Diagnosis of the device eval path on the reference's ternary legacy potentials (SrTiO3): eval (both kernel
flavours) against the design-matrix path X @ coeffs and the published energies."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import polymlp_oracle as po  # noqa: E402
from pypolymlp_b200._capi import PM_FLAG_SIMPLE_KERNELS  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialModel, PotentialPropertiesFast  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402
from test_legacy_io import LEGACY_EVAL, load_legacy_golden, params_from_golden  # noqa: E402

np.set_printoptions(precision=12, linewidth=200)
L = load_legacy_golden()
for key in ("srtio3_gtinv", "srtio3_gtinv_cond", "srtio3_pair"):
    name, st, e_true = LEGACY_EVAL[key][:3]
    pd, c = params_from_golden(name), L[key + "_coeffs"]
    axis, pos, types = L[st + "_axis"], L[st + "_pos"], L[st + "_types"]
    print("==", key, "published E", e_true, "F", len(c))
    for flags in (0, PM_FLAG_SIMPLE_KERNELS):
        try:
            p = PotentialPropertiesFast(pd, c, flags=flags)
            p.eval(axis, pos, types)
            print(" eval flags", flags, "E", repr(p.get_e()), "max|F|", np.abs(p.get_f()).max(), "S", p.get_s())
        except Exception as ex:  # noqa: BLE001
            print(" eval flags", flags, "EXC", repr(ex))
    try:
        x = PotentialModel(pd, [axis], [pos], [types], [1], [True], [len(types)]).get_x()
        y = x @ c
        print(" X@c      E", repr(float(y[0])), "S", y[1:7], "max|F|", np.abs(y[7:]).max())
    except Exception as ex:  # noqa: BLE001
        print(" X EXC", repr(ex))

# a small ternary gtinv model on a skewed cell against the numpy oracle
for kw in (dict(n_type=3, cutoff=4.5, model_type=3, max_p=2, gtinv_order=3, gtinv_maxl=[2, 2], n_gaussians=3),
           dict(n_type=3, cutoff=4.5, model_type=2, max_p=2, gtinv_order=2, gtinv_maxl=[2], n_gaussians=3),
           dict(n_type=2, cutoff=4.5, model_type=3, max_p=2, gtinv_order=3, gtinv_maxl=[2, 2], n_gaussians=3)):
    pd = make_params_dict(**kw)
    tab = po.Tables(pd)
    ax, pc, ty = cases.skewed_cell(kw["n_type"], n_atom=6, seed=5)
    c = np.random.default_rng(3).normal(size=tab.n_variables)
    e, f, s = po.eval_structure(tab, c, ax, pc, ty)
    print("== synthetic", kw["n_type"], kw["model_type"], kw["gtinv_maxl"], "F", tab.n_variables, "oracle E", repr(e))
    for flags in (0, PM_FLAG_SIMPLE_KERNELS):
        p = PotentialPropertiesFast(pd, c, flags=flags)
        p.eval(ax, pc, ty)
        print(" eval flags", flags, "E", repr(p.get_e()), "dF", np.abs(p.get_f() - f).max() / np.abs(f).max(),
              "dS", np.abs(p.get_s() - s).max() / np.abs(s).max())

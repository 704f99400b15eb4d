"""Model description in the shape of the reference's `PolymlpParams.as_dict()`.

Only the keys that cross the pybind11 boundary are produced (compute/py_params.cpp:14-43):
n_type, print_memory, model.{cutoff, pair_type, feature_type, model_type, max_p, max_l,
pair_conditional, pair_params, pair_params_conditional, gtinv.{lm_seq, l_comb, lm_coeffs}}.
"""

import itertools

import numpy as np

from .libmlpcpp import Readgtinv


def set_gaussian_params(params1=(1.0, 1.0, 1), params2=(0.0, 5.0, 7), n_gaussians=None, cutoff=None):
    """Gaussian radial parameters (reference: src/pypolymlp/core/params_utils.py:75-95)."""
    if n_gaussians is None:
        g1 = np.linspace(float(params1[0]), float(params1[1]), int(params1[2]))
        g2 = np.linspace(float(params2[0]), float(params2[1]), int(params2[2]))
        pair_params = [[float(a), float(b)] for a, b in itertools.product(g1, g2)]
    else:
        if cutoff is None:
            raise RuntimeError("Cutoff required for automatic setting of Gaussians.")
        g2 = np.linspace(0.0, 1.0, n_gaussians - 1) * (cutoff - 1.0)
        width = max(1.0, g2[1] - g2[0])
        pair_params = [[float(width), float(p2)] for p2 in g2]
    pair_params.append([0.0, 0.0])
    return pair_params


def make_params_dict(n_type, cutoff, model_type, max_p, gtinv_order, gtinv_maxl, pair_params=None,
                     n_gaussians=None, gaussian_params1=(1.0, 1.0, 1), gaussian_params2=(0.0, 5.0, 7),
                     pair_params_conditional=None, gtinv_version=1, print_memory=False, feature_type="gtinv"):
    """Build the params dict of a polymlp; gtinv tables come from Readgtinv.  feature_type='pair' gives the
    reference's empty gtinv block (order 0, max_l [] -> model.max_l = 0; params_utils.py:57-60)."""
    if pair_params is None:
        pair_params = set_gaussian_params(gaussian_params1, gaussian_params2, n_gaussians, cutoff)
    if feature_type == "pair":
        gtinv_order, gtinv_maxl = 0, []
    rg = Readgtinv(gtinv_order, list(gtinv_maxl), gtinv_version) if feature_type != "pair" else None
    cond = pair_params_conditional is not None
    if not cond:
        pair_params_conditional = {
            (i, j): list(range(len(pair_params))) for i in range(n_type) for j in range(i, n_type)
        }
    return {
        "n_type": int(n_type),
        "print_memory": bool(print_memory),
        "model": {
            "cutoff": float(cutoff),
            "pair_type": "gaussian",
            "feature_type": feature_type,
            "model_type": int(model_type),
            "max_p": int(max_p),
            "max_l": int(max(gtinv_maxl)) if len(gtinv_maxl) else 0,
            "pair_conditional": cond,
            "pair_params": [list(map(float, p)) for p in pair_params],
            "pair_params_conditional": pair_params_conditional,
            "gtinv": {
                "order": int(gtinv_order),
                "max_l": list(gtinv_maxl),
                "version": int(gtinv_version),
                "lm_seq": rg.get_lm_seq() if rg else [],
                "l_comb": rg.get_l_comb() if rg else [],
                "lm_coeffs": rg.get_lm_coeffs() if rg else [],
            },
        },
    }

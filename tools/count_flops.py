"""Algorithmic FLOPs per structure of the BASELINE configs: evaluates
W = W_syrk + W_xty + W_poly + W_deriv + W_anlm from the host model tables (pm_model_count_flops) and the neighbour
statistics of one synthetic structure of the named shape.  Host only (no GPU): python tools/count_flops.py [2 3 4]

The neighbour statistics (ordered pairs per (centre type, neighbour type)) are counted here by a plain numpy image sum;
they depend only on the lattice, the cutoff and the type assignment."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200.libmlpcpp import _Model  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402


def pair_counts(axis, pos, types, n_type, cutoff):
    """pairs[t, u] = ordered pairs (i, image of j) with 1e-10 < r < cutoff, centre type t, neighbour type u."""
    reach = [int(np.ceil(cutoff / np.linalg.norm(np.linalg.inv(axis)[k]) ** -1)) for k in range(3)]
    counts = np.zeros((n_type, n_type), np.int64)
    for a in range(-reach[0], reach[0] + 1):
        for b in range(-reach[1], reach[1] + 1):
            for c in range(-reach[2], reach[2] + 1):
                shift = axis @ np.array([a, b, c], float)
                d = pos[:, None, :] - pos[:, :, None] + shift[:, None, None]      # [xyz, i, j]
                r2 = np.einsum("kij,kij->ij", d, d)
                hit = (r2 < cutoff * cutoff) & (r2 > 1e-20)
                for t in range(n_type):
                    for u in range(n_type):
                        counts[t, u] += hit[np.ix_(types == t, types == u)].sum()
    return counts


CONFIGS = {
    2: ("config 2: order 3, maxl [4,4], model_type 4, 256-atom fcc", cases.cfg2_model_kwargs(4), 1,
        lambda: cases.fcc_supercell(seed=20240)),
    3: ("config 3: binary order 4, maxl [12,8,2], model_type 3, 216-atom bcc", cases.cfg3_model_kwargs(), 2,
        lambda: cases.bcc_supercell(rep=(6, 6, 3), a=3.2, n_type=2, seed=100)),
    4: ("config 4: ternary order 4, maxl [12,8,2], model_type 3, 128-atom bcc", cases.cfg4_model_kwargs(), 3,
        lambda: cases.bcc_supercell(rep=(4, 4, 4), a=3.2, n_type=3, seed=100)),
}

for cfg in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
    name, kw, n_type, make = CONFIGS[cfg]
    model = _Model(make_params_dict(**kw))
    axis, pos, types = make()
    pairs = pair_counts(axis, pos, types, n_type, kw["cutoff"])
    atoms = np.array([(types == t).sum() for t in range(n_type)], np.int64)
    w = model.count_flops(atoms, pairs, force=True)
    total = sum(w.values())
    print(json.dumps({"config": name, "n_features": model.info()["n_variables"], "atoms": int(atoms.sum()),
                      "ordered_pairs": int(pairs.sum()), "rows": int(1 + 6 + 3 * atoms.sum()),
                      "flops": {k: float("%.4g" % v) for k, v in w.items()}, "flops_per_structure": float("%.4g" % total)}))

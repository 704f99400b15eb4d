"""Small fit + eval invocations that reach the round-2 kernels, for `compute-sanitizer --tool memcheck python tools/sanitize_probe.py`."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PM_EVAL_LA_MIN", "1")
import cases  # noqa: E402
from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast, PotentialXtX  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

pd = make_params_dict(**cases.cfg2_model_kwargs(4))
sts = [cases.fcc_supercell(rep=(2, 2, 2), sigma=0.03, seed=5 + k) for k in range(5)] + [cases.fcc_supercell(rep=(3, 2, 2), sigma=0.03, seed=1)]
prop = PotentialPropertiesFast(pd, np.random.default_rng(12).normal(size=2030) * 1e-3)
for variant in ({}, {"PM_EVAL_LB": "0"}, {"PM_EVAL_LA": "0"}, {"PM_EVAL_K2_NO_DMMA": "1"}, {"PM_CELL_LIST": "1"}):
    os.environ.update(variant)
    prop.eval_multiple([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts])
    print("eval", variant, float(np.sum(prop.get_e_array())))
    for k in variant:
        del os.environ[k]
rows = sum(1 + 6 + 3 * s[1].shape[1] for s in sts)
rng = np.random.default_rng(1)
w = rng.uniform(0.5, 1.0, rows)
y = w * rng.normal(size=rows)
acc = PotentialXtX(pd)
acc.add([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts], [True] * len(sts), w, y)
r = acc.finalize()
print("fit", float(np.trace(r["xtx"])))
if len(sys.argv) > 1 and sys.argv[1] == "big":
    # the large-model kernels (k_lrows_big, k_features_v4r, the head passes of k_anlm_v2, k_xrows_v2 ...) on the 16-atom ternary cell
    from pypolymlp_b200.libmlpcpp import PotentialModel

    pd4 = make_params_dict(**cases.cfg4_model_kwargs())
    ax, pc, ty = cases.cfg4_small_cell()
    x = PotentialModel(pd4, [ax], [pc], [ty], [1], [True], [16]).get_x()
    print("big X", x.shape, float(np.abs(x).sum()))

"""DMMA (mma.sync m8n8k4 f64) issue-interval / dependent-latency probe on one SM: cycles per DMMA of one warp for
w warps per CTA and n independent accumulator chains per warp."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pypolymlp_b200.libmlpcpp import _Context, _Model  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

ctx = _Context(_Model(make_params_dict(**cases.si_model_kwargs())))
print("warps chains cycles/DMMA/warp  DMMA/cycle/SM")
for w in (1, 4, 8, 12, 16):
    for n in (1, 2, 3, 4, 6, 8, 12, 16):
        c = ctx.microbench(5, 100 * w + n)
        print(f"{w:5d} {n:6d} {c:10.2f} {w / c:12.4f}")

import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
os.environ["PM_DEBUG_TABLES"] = "1"
from test_legacy_io import LEGACY_EVAL, load_legacy_golden, params_from_golden
from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast
L = load_legacy_golden()
for key in sorted(LEGACY_EVAL):
    name = LEGACY_EVAL[key][0]
    print("==", key, name, flush=True)
    PotentialPropertiesFast(params_from_golden(name), L[key + "_coeffs"])

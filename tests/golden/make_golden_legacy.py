"""Golden data for the legacy `polymlp.lammps` reader (pypolymlp_b200/io_legacy.py), produced by the UNMODIFIED
reference Python loader (src/pypolymlp/core/io_polymlp_legacy.py) imported from /root/reference.  Run in the build
container:

    python tests/golden/make_golden_legacy.py

Outputs
  legacy_params.json   what the reference loader returns for every legacy file bundled with the reference
                       (tests/test_calc/files/mlps/polymlp.lammps.*, tests/files/polymlp.lammps.*) and for our own
                       fixture polymlp.lammps.synthetic: model parameters + digests of the scaled coefficients
  legacy.npz           scaled coefficients of the potentials the reference publishes known answers for
                       (tests/test_calc/test_properties_legacy_{SrTiO3,Ag,MgO}.py) and the structures of those tests
  polymlp.yaml.flexible.{1,2}.SrTiO3   the reference's hybrid ("flexible") SrTiO3 potential re-written by OUR yaml
                       writer from what our reader got out of the reference's files (coefficients checked against the
                       reference loader here); sub-model 2 is an O-only model with type_full = 0
  polymlp.lammps.synthetic   a small legacy file written by OUR writer (binary conditional gtinv model, seeded
                       coefficients / scales); travels with the repo so the reader is covered without /root/reference
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from pypolymlp_b200 import dropin  # noqa: E402

# the reference's Python package imports its pybind11 extension at module scope; its source tree has none, ours stands in
dropin.install("/root/reference/src")

from pypolymlp.core.interface_vasp import Poscar  # noqa: E402  (reference)
from pypolymlp.core.io_polymlp_legacy import load_mlp_lammps as ref_load  # noqa: E402  (reference)

from pypolymlp_b200.io_legacy import save_mlp_lammps  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402
from pypolymlp_b200.libmlpcpp import FeaturesAttr  # noqa: E402

REF_T = "/root/reference/tests"


def describe(params, coeffs):
    m = params.model
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    return {
        "elements": list(params.elements), "n_type": int(params.n_type), "cutoff": float(m.cutoff),
        "pair_type": m.pair_type, "feature_type": m.feature_type, "model_type": int(m.model_type),
        "max_p": int(m.max_p), "max_l": int(m.max_l), "gtinv_order": int(m.gtinv.order),
        "gtinv_maxl": [int(v) for v in m.gtinv.max_l], "gtinv_version": int(m.gtinv.version),
        "pair_conditional": bool(m.pair_conditional),
        "pair_params": [[float(a), float(b)] for a, b in m.pair_params],
        "pair_params_conditional": [[int(k[0]), int(k[1]), [int(v) for v in ids]]
                                    for k, ids in m.pair_params_conditional.items()],
        "type_full": bool(params.type_full), "type_indices": [int(v) for v in params.type_indices],
        "mass": [float(v) for v in params.mass],
        "n_coeffs": int(c.size), "coeffs_sum": float(c.sum()), "coeffs_abs_sum": float(np.abs(c).sum()),
        "coeffs_head": [float(v) for v in c[:4]], "coeffs_sha256": hashlib.sha256(c.tobytes()).hexdigest(),
    }


# ---- our own fixture, written by our writer, read back by the reference loader ------------------------------------
pd = make_params_dict(**cases.binary_model_kwargs())
nf = FeaturesAttr(pd).get_n_features()
rng = np.random.default_rng(2718)
raw, scales = rng.normal(size=nf), np.exp(rng.normal(size=nf))
synthetic = os.path.join(cases.GOLDEN, "polymlp.lammps.synthetic")
save_mlp_lammps(pd, raw, scales, ["Mg", "O"], mass=[24.305, 15.999], filename=synthetic)

table = {}
files = sorted(glob.glob(REF_T + "/test_calc/files/mlps/polymlp.lammps.*") + glob.glob(REF_T + "/files/polymlp.lammps.*"))
for fn in files + [synthetic]:
    params, coeffs = ref_load(fn)
    table[os.path.basename(fn)] = describe(params, coeffs)
    print(os.path.basename(fn), table[os.path.basename(fn)]["n_coeffs"])
with open(os.path.join(cases.GOLDEN, "legacy_params.json"), "w") as f:
    json.dump(table, f, indent=1, sort_keys=True)

# ---- coefficients + structures of the reference's published legacy known answers ----------------------------------
out = {}
for key, name in [("srtio3_pair", "pair.SrTiO3"), ("srtio3_pair_cond", "pair.cond.SrTiO3"),
                  ("srtio3_gtinv", "gtinv.SrTiO3"), ("srtio3_gtinv_cond", "gtinv.cond.SrTiO3"),
                  ("ag_pair", "pair.Ag"), ("mgo_pair", "pair.MgO")]:
    _, coeffs = ref_load(REF_T + "/test_calc/files/mlps/polymlp.lammps." + name)
    out[key + "_coeffs"] = np.asarray(coeffs, np.float64)
for key, poscar in [("srtio3", "POSCAR.perovskite.SrTiO3"), ("ag", "POSCAR.fcc.Ag")]:
    st = Poscar(REF_T + "/test_calc/files/poscars/" + poscar).structure
    out[key + "_axis"] = np.asarray(st.axis, np.float64)
    out[key + "_pos"] = np.asarray(st.axis @ st.positions, np.float64)
    out[key + "_types"] = np.asarray(st.types, np.int32)
np.savez_compressed(os.path.join(cases.GOLDEN, "legacy.npz"), **out)
print({k: v.shape for k, v in out.items()})

# ---- the reference's hybrid potential of test_properties_legacy_SrTiO3.py:72-85, re-written by our writer -----------
from pypolymlp.core.io_polymlp import load_mlp as ref_load_mlp  # noqa: E402  (reference)

from pypolymlp_b200.io_yaml import load_mlp_yaml, save_mlp_yaml  # noqa: E402

for k in (1, 2):
    src = REF_T + "/test_calc/files/mlps/polymlp.yaml.flexible.%d.SrTiO3" % k
    dst = os.path.join(cases.GOLDEN, "polymlp.yaml.flexible.%d.SrTiO3" % k)
    pd_k, coeffs_k, meta_k = load_mlp_yaml(src)
    params_ref, coeffs_ref = ref_load_mlp(src)
    assert np.array_equal(coeffs_k, coeffs_ref) and list(params_ref.elements) == meta_k["elements"]
    assert bool(params_ref.type_full) == meta_k["type_full"] and list(params_ref.type_indices) == meta_k["type_indices"]
    save_mlp_yaml(pd_k, coeffs_k, np.ones(len(coeffs_k)), meta_k["elements"], filename=dst, mass=meta_k["mass"],
                  type_full=meta_k["type_full"], type_indices=meta_k["type_indices"])
    params_back, coeffs_back = ref_load_mlp(dst)      # the reference loader reads our file back identically
    assert np.allclose(coeffs_back, coeffs_ref, rtol=1e-15, atol=0)
    assert params_back.as_dict()["model"] == params_ref.as_dict()["model"]
    assert bool(params_back.type_full) == bool(params_ref.type_full)
    print(os.path.basename(dst), len(coeffs_k), meta_k["elements"], meta_k["type_full"], meta_k["type_indices"])

# ---- cell-shape invariance set of tests/test_calc/test_check_neighbors.py:39-75 ---------------------------------------
# ideal rocksalt MgO re-expressed in 25 unimodular (mostly very skewed) cells.  The reference builds them with
# phonopy (utils/supercell_utils.py:26-53; phonopy is not installed here); for det H = 1 the supercell is the same
# crystal in the lattice A' = A H with fractional coordinates H^-1 x wrapped into [0, 1) (refine rule of :19-23)
def supercell(st, hnf):
    assert round(np.linalg.det(hnf)) == 1
    frac = np.linalg.solve(hnf.astype(float), np.asarray(st.positions, float))
    frac -= np.floor(frac)
    frac[frac > 1 - 1e-13] -= 1.0

    class _S:
        axis, positions, types = np.asarray(st.axis, float) @ hnf, frac, st.types
    return _S


unitcell = Poscar(REF_T + "/test_calc/files/poscars/POSCAR.RS.idealMgO").structure
expansions = [
    [[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[1, 0, 0], [1, 1, 0], [1, 1, 1]], [[1, 0, 0], [1, 1, 0], [0, 0, 1]],
    [[1, 1, 0], [0, 1, 0], [0, 0, 1]], [[1, 0, 0], [0, 1, 0], [1, 0, 1]], [[1, 0, 0], [0, 1, 0], [0, 1, 1]],
    [[1, 0, 0], [-1, 1, 0], [0, 0, 1]], [[1, 0, 0], [-1, 1, 0], [-1, -1, 1]], [[1, 0, 0], [-3, 1, 0], [3, 3, 1]],
    [[1, 0, 0], [3, 1, 0], [3, 3, 1]], [[1, 0, 0], [-3, 1, 0], [-3, -3, 1]], [[1, 0, 0], [5, 1, 0], [5, 5, 1]],
    [[1, 0, 0], [-5, 1, 0], [-5, -5, 1]], [[1, 0, 0], [10, 1, 0], [10, 10, 1]], [[1, 0, 0], [-10, 1, 0], [-10, -10, 1]],
    [[1, 3, 3], [0, 1, 3], [0, 0, 1]], [[1, 4, 4], [0, 1, 4], [0, 0, 1]], [[1, 5, 5], [0, 1, 5], [0, 0, 1]],
    [[1, -5, -5], [0, 1, -5], [0, 0, 1]], [[1, 10, 10], [0, 1, 10], [0, 0, 1]], [[1, -10, -10], [0, 1, -10], [0, 0, 1]],
    [[1, 0, 0], [5, 1, 0], [0, 0, 1]], [[1, 0, 0], [-5, 1, 0], [0, 0, 1]], [[1, 0, 0], [10, 1, 0], [0, 0, 1]],
    [[1, 0, 0], [-10, 1, 0], [0, 0, 1]],
]
cells = {}
for k, hnf in enumerate(expansions):
    st = supercell(unitcell, np.array(hnf))
    cells["axis_%02d" % k] = np.asarray(st.axis, np.float64)
    cells["pos_%02d" % k] = np.asarray(st.axis @ st.positions, np.float64)
    cells["types_%02d" % k] = np.asarray(st.types, np.int32)
np.savez_compressed(os.path.join(cases.GOLDEN, "mgo_cell_shapes.npz"), **cells)
print("mgo_cell_shapes", len(expansions), cells["pos_08"].shape, cells["axis_13"])

# ---- the skewed BiGd2 cell of tests/test_cxx/test_neighbor_variants.py (tests/files/POSCAR-BiGd2) -------------------
st = Poscar(REF_T + "/files/POSCAR-BiGd2").structure
np.savez_compressed(os.path.join(cases.GOLDEN, "bigd2.npz"), axis=np.asarray(st.axis, np.float64),
                    positions=np.asarray(st.positions, np.float64), types=np.asarray(st.types, np.int32))
print("bigd2", np.asarray(st.positions).shape, st.n_atoms)

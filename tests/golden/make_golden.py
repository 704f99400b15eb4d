"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference C++
(oracle/_ref, built from /root/reference by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

Outputs: si_dataset.npz (the reference's bundled Si training set as arrays) and ref_vectors.npz.
"""
import lzma
import os
import sys

import numpy as np
import yaml

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import ref  # noqa: E402
from pypolymlp_b200.params import make_params_dict  # noqa: E402

out = {}

# ---- bundled Si dataset ---------------------------------------------------------------------------
with lzma.open("/root/reference/tests/files/phonopy_training_dataset.yaml.xz", "rt") as f:
    data = yaml.load(f, Loader=yaml.CSafeLoader)
axis = np.array(data["supercell"]["lattice"]).T
frac = np.array([p["coordinates"] for p in data["supercell"]["points"]]).T
np.savez_compressed(os.path.join(cases.GOLDEN, "si_dataset.npz"), axis=axis, frac=frac,
                    displacements=np.array(data["dataset"]["displacements"]),
                    forces=np.array(data["dataset"]["forces"]),
                    energies=np.array(data["dataset"]["supercell_energies"]))
axis, positions_c, forces, energies = cases.load_si_dataset()

# ---- Si model: one structure in full, column sums over the 180 training structures ---------------
pd = make_params_dict(**cases.si_model_kwargs())
rm = ref.RefModel(pd)
types = np.zeros(64, np.int32)
xe, xf, xs = rm.run(axis, positions_c[3], types, True)
out["si3_xe"], out["si3_xf"], out["si3_xs"] = xe, xf, xs
a, d = rm.atom(axis, positions_c[3], types, 5)
out["si3_atom5_anlm"], out["si3_atom5_d"] = a, d
train_ids, _ = cases.split_ids_train_test(200, 0.9)
X = rm.build_x([axis] * len(train_ids), [positions_c[i] for i in train_ids], [types] * len(train_ids),
               [True] * len(train_ids))
out["si_train_shape"] = np.array(X.shape)
out["si_train_colsum"] = X.sum(axis=0)
out["si_train_colsqsum"] = np.square(X).sum(axis=0)

# ---- binary model on a skewed cell (conditional radial sets): X and eval ----------------------------
pd = make_params_dict(**cases.binary_model_kwargs())
rm = ref.RefModel(pd)
ax, pc, ty = cases.skewed_cell(2)
xe, xf, xs = rm.run(ax, pc, ty, True)
out["bin_xe"], out["bin_xf"], out["bin_xs"] = xe, xf, xs
coeffs = np.random.default_rng(11).normal(size=rm.n_features)
e, f, s = ref.RefEval(pd, coeffs).eval(ax, pc, ty)
out["bin_coeffs"], out["bin_e"], out["bin_f"], out["bin_s"] = coeffs, np.array([e]), f, s
for kind in ("full", "half"):
    o = ref.neighbor(kind, ax, pc, 5.0)
    for name, arr in zip(("off", "nb", "dx", "dy", "dz"), o):
        out[f"bin_nbr_{kind}_{name}"] = arr

# ---- ternary, max_p = 3 -----------------------------------------------------------------------------
pd = make_params_dict(**cases.ternary_p3_model_kwargs())
rm = ref.RefModel(pd)
ax, pc, ty = cases.skewed_cell(3, n_atom=7, seed=3)
xe, xf, xs = rm.run(ax, pc, ty, True)
out["ter_xe"], out["ter_xf"], out["ter_xs"] = xe, xf[::3], xs

# ---- small skewed cell: neighbour list with the integer-abs metric branch ---------------------------
ax, pc, ty = cases.small_skewed_cell()
o = ref.neighbor("full", ax, pc, 4.0)
for name, arr in zip(("off", "nb", "dx", "dy", "dz"), o):
    out[f"small_nbr_{name}"] = arr
tr, ax2, pc2 = ref.neighbor_cell(ax, pc, 4.0)
out["small_trans"], out["small_axis"], out["small_pos"] = tr, ax2, pc2

# ---- config-2 model on one 256-atom fcc structure: summaries (full X would be 12.6 MB) ---------------
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
rm = ref.RefModel(pd)
ax, pc, ty = cases.fcc_supercell()
xe, xf, xs = rm.run(ax, pc, ty, True)
rows = np.array([0, 1, 2, 100, 383, 384, 500, 765, 766, 767])
out["fcc_xe"], out["fcc_xs"] = xe, xs
out["fcc_rows"], out["fcc_xf_rows"] = rows, xf[rows]
out["fcc_xf_colsum"], out["fcc_xf_colsqsum"] = xf.sum(axis=0), np.square(xf).sum(axis=0)
o = ref.neighbor("full", ax, pc, 6.0)
out["fcc_nbr_count"] = np.array([o[0][-1]])
out["fcc_nbr_checksum"] = np.array([o[1].sum(), np.square(o[2]).sum() + np.square(o[3]).sum() + np.square(o[4]).sum()])
coeffs = np.random.default_rng(12).normal(size=rm.n_features) * 1e-3
e, f, s = ref.RefEval(pd, coeffs).eval(ax, pc, ty)
out["fcc_coeffs"], out["fcc_e"], out["fcc_f"], out["fcc_s"] = coeffs, np.array([e]), f, s

# ---- config-3 model (binary, order 4, maxl [12,8,2], F = 9385) on a 16-atom bcc cell: summaries -----------
pd = make_params_dict(**cases.cfg3_model_kwargs())
rm = ref.RefModel(pd)
ax, pc, ty = cases.bcc_supercell(rep=(2, 2, 2), a=3.2, n_type=2, seed=5)
xe, xf, xs = rm.run(ax, pc, ty, True)
out["cfg3_xe"], out["cfg3_xs"] = xe, xs
out["cfg3_xf_rows"] = xf[::6]
out["cfg3_xf_colsum"], out["cfg3_xf_colsqsum"] = xf.sum(axis=0), np.square(xf).sum(axis=0)

# ---- pair features (feature_type = "pair", compute/local_pair.cpp): binary conditional model -----------
pd = make_params_dict(**cases.pair_model_kwargs(2))
rm = ref.RefModel(pd)
ax, pc, ty = cases.skewed_cell(2)
xe, xf, xs = rm.run(ax, pc, ty, True)
out["pair_xe"], out["pair_xf"], out["pair_xs"] = xe, xf, xs
coeffs = np.random.default_rng(13).normal(size=rm.n_features)
e, f, s = ref.RefEval(pd, coeffs).eval(ax, pc, ty)
out["pair_coeffs"], out["pair_e"], out["pair_f"], out["pair_s"] = coeffs, np.array([e]), f, s

np.savez_compressed(os.path.join(cases.GOLDEN, "ref_vectors.npz"), **out)

# ---- the reference's MgO fixtures (tests/test_calc/files): structures and published coefficients --------
# Known answers that go with them are quoted in tests/ (test_calc/test_compute_features.py:45-58,
# test_calc/test_properties_MgO.py:9-97).


def read_poscar(path):
    with open(path) as fh:
        ln = fh.read().split("\n")
    scale = float(ln[1])
    lattice = np.array([[float(v) for v in ln[k].split()] for k in (2, 3, 4)]) * scale
    counts = [int(v) for v in ln[6].split()]
    n = sum(counts)
    frac = np.array([[float(v) for v in ln[8 + k].split()[:3]] for k in range(n)])
    axis_ = lattice.T
    types_ = np.concatenate([np.full(c, k, np.int32) for k, c in enumerate(counts)])
    return axis_, axis_ @ frac.T, types_


fdir = "/root/reference/tests/test_calc/files/"
mgo = {}
for key, name in (("rs", "POSCAR.RS.MgO"), ("st1", "POSCAR-00001.MgO"), ("st2", "POSCAR-00002.MgO")):
    mgo[key + "_axis"], mgo[key + "_pos"], mgo[key + "_types"] = read_poscar(fdir + "poscars/" + name)
for key in ("pair", "gtinv"):
    with open(fdir + "mlps/polymlp.yaml." + key + ".MgO") as fh:
        mgo[key + "_coeffs"] = np.array(yaml.safe_load(fh)["coeffs"], dtype=np.float64)
np.savez_compressed(os.path.join(cases.GOLDEN, "mgo.npz"), **mgo)

for k, v in out.items():
    print(k, v.shape)

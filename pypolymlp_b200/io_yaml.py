"""polymlp.yaml reader / writer for single models, in the reference's format
(src/pypolymlp/core/io_polymlp_yaml.py:17-86 save_mlp_yaml, :89-161 load_mlp_yaml).

The file carries everything the evaluation path needs: the model description that becomes the
`params_dict` of the pybind11 boundary (compute/py_params.cpp:14-43) and the regression coefficients
already divided by the feature scales.  `load_mlp_yaml` returns that params_dict (gtinv tables filled
from Readgtinv), the coefficients and the remaining metadata; `save_mlp_yaml` writes a fitted model so
that the reference's own loader reads it back."""

import io

import numpy as np
import yaml

from .params import make_params_dict

# atomic masses are only echoed into the file (LAMMPS interop); unknown elements get 1.0
_MASS = {"H": 1.008, "Li": 6.94, "B": 10.81, "C": 12.011, "N": 14.007, "O": 15.999, "F": 18.998, "Na": 22.99,
         "Mg": 24.305, "Al": 26.982, "Si": 28.085, "P": 30.974, "S": 32.06, "Cl": 35.45, "K": 39.098, "Ca": 40.078,
         "Ti": 47.867, "V": 50.942, "Cr": 51.996, "Mn": 54.938, "Fe": 55.845, "Co": 58.933, "Ni": 58.693,
         "Cu": 63.546, "Zn": 65.38, "Ga": 69.723, "Ge": 72.63, "As": 74.922, "Sr": 87.62, "Y": 88.906, "Zr": 91.224,
         "Nb": 92.906, "Mo": 95.95, "Ag": 107.87, "Sn": 118.71, "Ba": 137.33, "Gd": 157.25, "W": 183.84,
         "Pt": 195.08, "Au": 196.97, "Pb": 207.2, "Bi": 208.98}


def load_mlp_yaml(filename="polymlp.yaml"):
    """Returns (params_dict, coeffs, meta).  meta: elements, mass, type_full, type_indices, enable_spins."""
    if isinstance(filename, io.IOBase):
        yml = yaml.safe_load(filename)
    else:
        with open(filename) as f:
            yml = yaml.safe_load(f)
    elements = list(yml["elements"])
    n_type = len(elements)
    feature_type = yml["feature_type"]
    if yml.get("pair_type", "gaussian") != "gaussian":
        raise ValueError("pair_type must be 'gaussian'")
    cond = {tuple(int(v) for v in tp["atom_type_pair"]): [int(v) for v in tp["pair_params_indices"]]
            for tp in yml["type_pairs"]}
    for i in range(n_type):
        for j in range(i, n_type):
            cond.setdefault((i, j), [])
    gtinv = feature_type == "gtinv"
    pd = make_params_dict(
        n_type=n_type, cutoff=float(yml["cutoff"]), model_type=int(yml["model_type"]), max_p=int(yml["max_p"]),
        gtinv_order=int(yml["gtinv_order"]) if gtinv else 0, gtinv_maxl=list(yml["gtinv_maxl"]) if gtinv else [],
        pair_params=[[float(a), float(b)] for a, b in yml["pair_params"]], pair_params_conditional=cond,
        gtinv_version=int(yml.get("gtinv_version", 1)) if gtinv else 1, feature_type=feature_type)
    if int(yml["max_l"]) != pd["model"]["max_l"]:
        raise ValueError("max_l does not match gtinv_maxl")
    meta = {"elements": elements, "mass": yml.get("mass"), "type_full": bool(yml.get("type_full", 1)),
            "type_indices": list(yml.get("type_indices", range(n_type))), "enable_spins": yml.get("enable_spins")}
    return pd, np.asarray(yml["coeffs"], dtype=np.float64), meta


def save_mlp_yaml(params_dict, coeffs, scales, elements, filename="polymlp.yaml", mass=None, type_full=True,
                  type_indices=None):
    """Writes coeffs / scales with the key order and number format of the reference writer."""
    model = params_dict["model"]
    coeffs = np.asarray(coeffs, float) / np.asarray(scales, float)
    with open(filename, "w") as f:
        print("elements:     ", "[" + ", ".join(str(e) for e in elements) + "]", file=f)
        print("enable_spins: ", [0 for _ in elements], file=f)
        print("cutoff:       ", float(model["cutoff"]), file=f)
        print("pair_type:    ", model.get("pair_type", "gaussian"), file=f)
        print("feature_type: ", model["feature_type"], file=f)
        print("model_type:   ", int(model["model_type"]), file=f)
        print("max_p:        ", int(model["max_p"]), file=f)
        print("max_l:        ", int(model["max_l"]), file=f)
        print(file=f)
        if model["feature_type"] == "gtinv":
            g = model["gtinv"]
            print("gtinv_order:  ", int(g["order"]), file=f)
            print("gtinv_maxl:   ", [int(v) for v in g["max_l"]], file=f)
            print("gtinv_sym:    ", [0 for _ in g["max_l"]], file=f)
            print("gtinv_version:", int(g.get("version", 1)), file=f)
            print(file=f)
        print("electrostatic:", 0, file=f)
        print("mass:         ", [float(m) for m in mass] if mass is not None else [_MASS.get(e, 1.0) for e in elements],
              file=f)
        print(file=f)
        print("n_pair_params:", len(model["pair_params"]), file=f)
        print("pair_params:", file=f)
        for p in model["pair_params"]:
            print("-", [float(p[0]), float(p[1])], file=f)
        print("", file=f)
        cond = model["pair_params_conditional"]
        print("n_type_pairs:", len(cond), file=f)
        print("type_pairs:", file=f)
        for tp, ids in cond.items():
            print("- atom_type_pair:     ", [int(v) for v in tp], file=f)
            print("  pair_params_indices:", [int(v) for v in ids], file=f)
        print(file=f)
        print("type_full:   ", int(bool(type_full)), file=f)
        print("type_indices:", [int(v) for v in type_indices] if type_indices is not None
              else list(range(int(params_dict["n_type"]))), file=f)
        print(file=f)
        print("n_coeffs:", len(coeffs), file=f)
        print("coeffs:", "[" + ", ".join(f"{c:.15e}" for c in coeffs) + "]", file=f)

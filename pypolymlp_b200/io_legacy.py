"""polymlp.lammps (legacy text format) reader / writer and the format dispatch `load_mlp`.

Reference: src/pypolymlp/core/io_polymlp_legacy.py:23-161 (load_mlp_lammps), src/pypolymlp/core/io_polymlp.py:54-75
(load_mlp), :120-140 (convert_to_yaml), :143-156 (is_legacy).  The legacy file is the other model artefact the
evaluation path starts from (SURVEY.md §8f-4, §5 "checkpoint / resume"): one record per line, `values # label`,
in a fixed order with three optional trailing sections (gtinv_version, n_type_pairs, type_full).  It stores the raw
regression coefficients and the feature scales on two separate lines; the coefficients the evaluation uses are
`coeffs / scales` (io_polymlp_legacy.py:159).

`load_mlp_lammps` returns the same triple as `io_yaml.load_mlp_yaml`: the params_dict of the pybind11 boundary
(compute/py_params.cpp:14-43), the scaled coefficients, and the metadata."""

import io

import numpy as np

from .io_yaml import load_mlp_yaml, save_mlp_yaml
from .params import make_params_dict

_TRUE = ("y", "yes", "t", "true", "on", "1")
_FALSE = ("n", "no", "f", "false", "off", "0")


def _flag(token):
    # src/pypolymlp/core/utils.py strtobool: distutils' truth table, ValueError otherwise
    t = str(token).lower()
    if t in _TRUE:
        return True
    if t in _FALSE:
        return False
    raise ValueError("invalid truth value %r" % (token,))


class _Records:
    """The file as a list of (tokens, label) records with a cursor; labels are only consulted for the optional
    trailing sections, exactly as the reference does (`"n_type_pairs" in line`)."""

    def __init__(self, lines):
        self.rec = []
        for ln in lines:
            body, _, label = ln.partition("#")
            self.rec.append((body.split(), label.strip()))
        self.pos = 0

    def take(self, conv=int, many=False):
        if self.pos >= len(self.rec):
            raise ValueError("polymlp.lammps: unexpected end of file at record %d" % self.pos)
        tok, _ = self.rec[self.pos]
        self.pos += 1
        if many:
            return [conv(v) for v in tok]
        if not tok:
            raise ValueError("polymlp.lammps: empty record %d" % (self.pos - 1))
        return conv(tok[0])

    def next_is(self, label):
        """True when the record under the cursor exists and carries `label` (anywhere on the line)."""
        if self.pos >= len(self.rec):
            return False
        tok, lab = self.rec[self.pos]
        return label in lab or label in " ".join(tok)


def _read_lines(filename):
    if isinstance(filename, io.IOBase):
        return filename.readlines()
    with open(filename) as f:
        return f.readlines()


def is_legacy(filename="polymlp.yaml"):
    """First line of a legacy file is `El1 El2 ... # element(s)` (io_polymlp.py:143-156)."""
    if isinstance(filename, io.IOBase):
        pos = filename.tell()
        line = filename.readline()
        filename.seek(pos)
    else:
        with open(filename) as f:
            line = f.readline()
    return "# ele" in line


def load_mlp_lammps(filename="polymlp.lammps"):
    """Returns (params_dict, coeffs / scales, meta).  meta as in load_mlp_yaml, plus the raw `scales`."""
    r = _Records(_read_lines(filename))
    elements = r.take(str, many=True)
    n_type = len(elements)
    cutoff = r.take(float)
    pair_type = r.take(str)
    feature_type = r.take(str)
    model_type = r.take(int)
    max_p = r.take(int)
    max_l = r.take(int)
    if pair_type != "gaussian":
        raise ValueError("pair_type must be 'gaussian'")
    if feature_type == "gtinv":
        gtinv_order = r.take(int)
        gtinv_maxl = r.take(int, many=True)
        gtinv_sym = r.take(_flag, many=True)
    else:  # pair models carry no gtinv block and max_l is forced to 0 (io_polymlp_legacy.py:66-70)
        gtinv_order, gtinv_maxl, gtinv_sym, max_l = 0, [], [], 0

    n_coeffs = r.take(int)
    coeffs = np.array(r.take(float, many=True))
    scales = np.array(r.take(float, many=True))
    if coeffs.shape != scales.shape:
        raise ValueError("polymlp.lammps: %d coefficients but %d scales" % (coeffs.size, scales.size))
    n_pair_params = r.take(int)
    pair_params = [r.take(float, many=True) for _ in range(n_pair_params)]
    mass = r.take(float, many=True)
    r.take(_flag)  # electrostatic: read and ignored by the reference as well

    gtinv_version = 1
    if feature_type == "gtinv" and r.next_is("gtinv_version"):
        gtinv_version = r.take(int)

    cond, pair_conditional = None, False
    if r.next_is("n_type_pairs"):
        pair_conditional = True
        cond = {}
        for _ in range(r.take(int)):
            pair = tuple(r.take(int, many=True)[:2])  # third number = count of radial functions, redundant
            cond[pair] = r.take(int, many=True)

    type_full, type_indices = True, list(range(n_type))
    if r.next_is("type_full"):
        type_full = r.take(_flag)
        type_indices = r.take(int, many=True)

    pd = make_params_dict(n_type=n_type, cutoff=cutoff, model_type=model_type, max_p=max_p, gtinv_order=gtinv_order,
                          gtinv_maxl=gtinv_maxl, pair_params=pair_params, pair_params_conditional=cond,
                          gtinv_version=gtinv_version, feature_type=feature_type)
    pd["model"]["pair_conditional"] = pair_conditional
    if feature_type == "gtinv" and max_l != pd["model"]["max_l"]:
        raise ValueError("max_l does not match gtinv_max_l")
    meta = {"elements": elements, "mass": mass, "type_full": type_full, "type_indices": type_indices,
            "enable_spins": None, "gtinv_sym": gtinv_sym, "n_coeffs": n_coeffs, "scales": scales}
    return pd, coeffs / scales, meta


def save_mlp_lammps(params_dict, coeffs, scales, elements, mass=None, filename="polymlp.lammps",
                    type_full=True, type_indices=None):
    """Writes the legacy text format (record order of load_mlp_lammps; number formats of the files bundled with
    the reference, e.g. tests/test_calc/files/mlps/polymlp.lammps.gtinv.cond.SrTiO3).  `coeffs` are the raw
    regression coefficients and `scales` the feature scales, stored on separate lines."""
    model = params_dict["model"]
    n_type = int(params_dict["n_type"])
    mass = [1.0] * n_type if mass is None else list(mass)
    type_indices = list(range(n_type)) if type_indices is None else list(type_indices)
    out = []
    out.append(" ".join(str(e) for e in elements) + " # elements")
    out.append("%s # cutoff" % float(model["cutoff"]))
    out.append("%s # pair_type" % model.get("pair_type", "gaussian"))
    out.append("%s # feature_type" % model["feature_type"])
    out.append("%d # model_type" % model["model_type"])
    out.append("%d # max_p" % model["max_p"])
    out.append("%d # max_l" % model["max_l"])
    if model["feature_type"] == "gtinv":
        g = model["gtinv"]
        out.append("%d # gtinv_order" % g["order"])
        out.append(" ".join(str(int(v)) for v in g["max_l"]) + " # gtinv_max_l")
        out.append(" ".join("0" for _ in g["max_l"]) + " # gtinv_sym")
    out.append("%d # n_coeffs" % len(coeffs))
    out.append(" ".join("%.15e" % c for c in np.asarray(coeffs, float)) + " # reg. coeffs")
    out.append(" ".join("%.15e" % c for c in np.asarray(scales, float)) + " # scales")
    out.append("%d # n_params" % len(model["pair_params"]))
    for p in model["pair_params"]:
        out.append("%.15f %.15f # pair func. params" % (p[0], p[1]))
    out.append(" ".join("%.15e" % m for m in mass) + " # atomic mass")
    out.append("False # electrostatic")
    if model["feature_type"] == "gtinv":
        out.append("%d # gtinv_version" % model["gtinv"].get("version", 1))
    cond = model["pair_params_conditional"]
    out.append("%d # n_type_pairs" % len(cond))
    for pair, ids in cond.items():
        out.append("%d %d %d # atom type pair" % (pair[0], pair[1], len(ids)))
        out.append(" ".join(str(int(v)) for v in ids) + " # pair params indices")
    out.append("%d # type_full" % int(bool(type_full)))
    out.append(" ".join(str(int(v)) for v in type_indices) + " # type_indices")
    with open(filename, "w") as f:
        f.write("\n".join(out) + "\n")


def load_mlp(filename="polymlp.yaml"):
    """Format dispatch of the reference (io_polymlp.py:54-75): legacy text or yaml, same return triple."""
    if is_legacy(filename):
        return load_mlp_lammps(filename)
    return load_mlp_yaml(filename)


def convert_to_yaml(txt="polymlp.lammps", yaml="polymlp.yaml"):
    """Legacy -> polymlp.yaml with unit scales (io_polymlp.py:120-140: the stored coefficients are already divided by
    the scales).  One file or a list of files (hybrid models: yaml.1, yaml.2, ... in sorted order); returns False as soon
    as an input is not a legacy file, True otherwise.  The sub-model flags type_full / type_indices are carried over."""

    def one(src, dst):
        pd, coeffs, meta = load_mlp_lammps(src)
        save_mlp_yaml(pd, coeffs, np.ones(len(coeffs)), meta["elements"], filename=dst, mass=meta["mass"],
                      type_full=meta["type_full"], type_indices=meta["type_indices"])

    if isinstance(txt, (str, io.IOBase)):
        if not is_legacy(txt):
            return False
        one(txt, yaml)
        return True
    files = sorted(txt)
    for i, f in enumerate(files):
        if not is_legacy(f):
            return False
        one(f, yaml + "." + str(i + 1) if len(files) > 1 else yaml)
    return True


def load_mlps(file_list_or_file):
    """One file or a list of files (hybrid models) -> lists of (params_dict, coeffs, meta), one entry per sub-model
    (io_polymlp.py:78-104 returns PolymlpParams + a coefficient list; here the boundary dicts take their place)."""
    if isinstance(file_list_or_file, (str, io.IOBase)):
        files = [file_list_or_file]
    elif isinstance(file_list_or_file, (list, tuple, np.ndarray)):
        files = list(file_list_or_file)
    else:
        raise RuntimeError("Input object not appropriate for load_mlps.")
    if not files:
        raise RuntimeError("Input object not appropriate for load_mlps.")
    loaded = [load_mlp(f) for f in files]
    return [x[0] for x in loaded], [x[1] for x in loaded], [x[2] for x in loaded]


def find_mlps(path):
    """polymlp.yaml* files of a directory, else polymlp.lammps* files, sorted; None if neither (io_polymlp.py:107-117)."""
    import glob

    for pattern in ("/polymlp.yaml*", "/polymlp.lammps*"):
        files = glob.glob(path + pattern)
        if files:
            return sorted(files)
    return None


def is_hybrid(filename="polymlp.yaml"):
    """True for a list of more than one potential file (io_polymlp.py:159-175)."""
    if isinstance(filename, (str, io.IOBase)):
        return False
    if isinstance(filename, (list, tuple, np.ndarray)):
        return len(filename) > 1
    raise RuntimeError("filename must be strings or array-type.")

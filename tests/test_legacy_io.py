"""Legacy `polymlp.lammps` reader / writer (pypolymlp_b200/io_legacy.py) against the reference's own loader
(src/pypolymlp/core/io_polymlp_legacy.py:23-161; golden table made by tests/golden/make_golden_legacy.py) and against
the known answers the reference publishes for its bundled legacy potentials
(tests/test_calc/test_properties_legacy_{SrTiO3,Ag,MgO}.py).  CPU only: the evaluation side is the oracle."""

import hashlib
import io
import json
import os

import numpy as np
import pytest

import cases
from oracle import polymlp_oracle as po
from oracle import ref
from pypolymlp_b200.io_legacy import convert_to_yaml, is_legacy, load_mlp, load_mlp_lammps, save_mlp_lammps
from pypolymlp_b200.io_yaml import load_mlp_yaml, save_mlp_yaml
from pypolymlp_b200.params import make_params_dict

with open(os.path.join(cases.GOLDEN, "legacy_params.json")) as _f:
    GOLD = json.load(_f)
SYNTHETIC = os.path.join(cases.GOLDEN, "polymlp.lammps.synthetic")
REF_FILES = {name: os.path.join(d, name)
             for d in ("/root/reference/tests/test_calc/files/mlps", "/root/reference/tests/files")
             for name in GOLD if os.path.isfile(os.path.join(d, name))}
EV_TO_GPA = 160.21766208  # src/pypolymlp/core/units.py:22


def params_from_golden(name):
    """params_dict of a bundled legacy potential from the golden table (so the known-answer tests travel)."""
    g = GOLD[name]
    cond = {(a, b): ids for a, b, ids in g["pair_params_conditional"]} if g["pair_conditional"] else None
    pd = make_params_dict(n_type=g["n_type"], cutoff=g["cutoff"], model_type=g["model_type"], max_p=g["max_p"],
                          gtinv_order=g["gtinv_order"], gtinv_maxl=g["gtinv_maxl"], pair_params=g["pair_params"],
                          pair_params_conditional=cond, gtinv_version=g["gtinv_version"],
                          feature_type=g["feature_type"])
    pd["model"]["pair_conditional"] = g["pair_conditional"]
    return pd


def check_against_golden(name, pd, coeffs, meta):
    g, m = GOLD[name], pd["model"]
    assert meta["elements"] == g["elements"] and pd["n_type"] == g["n_type"]
    assert (m["cutoff"], m["pair_type"], m["feature_type"]) == (g["cutoff"], g["pair_type"], g["feature_type"])
    assert (m["model_type"], m["max_p"], m["max_l"]) == (g["model_type"], g["max_p"], g["max_l"])
    assert (m["gtinv"]["order"], list(m["gtinv"]["max_l"]), m["gtinv"]["version"]) == \
        (g["gtinv_order"], g["gtinv_maxl"], g["gtinv_version"])
    assert m["pair_conditional"] == g["pair_conditional"]
    assert m["pair_params"] == g["pair_params"]  # bit-equal: same float() of the same tokens
    assert [[k[0], k[1], list(v)] for k, v in m["pair_params_conditional"].items()] == g["pair_params_conditional"]
    assert meta["type_full"] == g["type_full"] and meta["type_indices"] == g["type_indices"]
    assert meta["mass"] == g["mass"]
    c = np.ascontiguousarray(coeffs, np.float64)
    assert c.size == g["n_coeffs"] == meta["n_coeffs"]
    assert hashlib.sha256(c.tobytes()).hexdigest() == g["coeffs_sha256"]  # coeffs / scales, bit-exact
    assert list(c[:4]) == g["coeffs_head"]


@pytest.mark.parametrize("name", sorted(n for n in GOLD if n != "polymlp.lammps.synthetic"))
def test_reader_matches_reference_loader_on_bundled_files(name):
    if name not in REF_FILES:
        pytest.skip("reference checkout not present")
    assert is_legacy(REF_FILES[name])
    check_against_golden(name, *load_mlp(REF_FILES[name]))


def test_reader_on_own_fixture_matches_reference_loader():
    """polymlp.lammps.synthetic was written by save_mlp_lammps and parsed by the reference loader (golden table)."""
    pd, coeffs, meta = load_mlp_lammps(SYNTHETIC)
    check_against_golden("polymlp.lammps.synthetic", pd, coeffs, meta)
    want = make_params_dict(**cases.binary_model_kwargs())
    assert pd["model"]["gtinv"]["lm_coeffs"] == want["model"]["gtinv"]["lm_coeffs"]
    rng = np.random.default_rng(2718)
    raw = rng.normal(size=coeffs.size)
    scales = np.exp(rng.normal(size=coeffs.size))
    assert np.allclose(coeffs, raw / scales, rtol=1e-14, atol=0)
    assert np.allclose(meta["scales"], scales, rtol=1e-15, atol=0)
    with open(SYNTHETIC) as f:  # file objects are accepted as by the reference (io_polymlp_legacy.py:37-41)
        pd2, coeffs2, _ = load_mlp(f)
    assert np.array_equal(coeffs2, coeffs) and pd2["model"]["pair_params"] == pd["model"]["pair_params"]


def test_writer_reader_round_trip_and_yaml_conversion(tmp_path):
    for kw, elements in [(cases.pair_model_kwargs(2), ["Mg", "O"]), (cases.si_model_kwargs(), ["Si"]),
                         (cases.ternary_p3_model_kwargs(), ["Sr", "Ti", "O"])]:
        pd = make_params_dict(**kw)
        n = po.Tables(pd).n_variables
        rng = np.random.default_rng(n)
        raw, scales = rng.normal(size=n), np.exp(rng.normal(size=n))
        fn = str(tmp_path / ("polymlp.lammps.%d" % n))
        save_mlp_lammps(pd, raw, scales, elements, mass=[1.0 + i for i in range(len(elements))], filename=fn)
        assert is_legacy(fn)
        pd2, coeffs, meta = load_mlp(fn)
        assert pd2["model"]["pair_params_conditional"] == pd["model"]["pair_params_conditional"]
        for key in ("cutoff", "feature_type", "model_type", "max_p", "max_l", "pair_params"):
            assert pd2["model"][key] == pd["model"][key], key
        assert pd2["model"]["gtinv"] == pd["model"]["gtinv"]
        assert meta["elements"] == elements and meta["mass"] == [1.0 + i for i in range(len(elements))]
        assert np.allclose(coeffs, raw / scales, rtol=1e-14, atol=0)
        # legacy -> yaml (io_polymlp.py:120-130): unit scales, same coefficients, same model
        yml = str(tmp_path / ("polymlp.yaml.%d" % n))
        assert convert_to_yaml(fn, yml) is True
        assert not is_legacy(yml)
        pd3, coeffs3, meta3 = load_mlp(yml)
        assert np.allclose(coeffs3, coeffs, rtol=1e-14, atol=0)
        assert pd3["model"]["gtinv"] == pd["model"]["gtinv"] and pd3["model"]["pair_params"] == pd["model"]["pair_params"]
        assert pd3["model"]["pair_params_conditional"] == pd["model"]["pair_params_conditional"]
        assert meta3["elements"] == elements and list(meta3["mass"]) == meta["mass"]


def test_is_legacy_and_dispatch(tmp_path):
    pd = make_params_dict(**cases.si_model_kwargs())
    n = po.Tables(pd).n_variables
    yml = str(tmp_path / "polymlp.yaml")
    save_mlp_yaml(pd, np.arange(n, dtype=float), np.ones(n), ["Si"], filename=yml)
    assert not is_legacy(yml)
    assert np.array_equal(load_mlp(yml)[1], load_mlp_yaml(yml)[1])
    assert convert_to_yaml(yml, str(tmp_path / "out.yaml")) is False   # not a legacy file: the reference returns False
    buf = io.StringIO(open(SYNTHETIC).read())
    assert is_legacy(buf) and buf.tell() == 0  # peeks, does not consume


def test_reader_optional_sections_and_errors(tmp_path):
    """Old files end after the electrostatic line (no gtinv_version / n_type_pairs / type_full): defaults of
    io_polymlp_legacy.py:95-130 (version 1, every radial function for every pair, type_full, identity indices)."""
    lines = open(SYNTHETIC).read().splitlines()
    cut = next(i for i, ln in enumerate(lines) if "electrostatic" in ln)
    fn = str(tmp_path / "old.lammps")
    with open(fn, "w") as f:
        f.write("\n".join(lines[:cut + 1]) + "\n")
    pd, coeffs, meta = load_mlp_lammps(fn)
    n_fn = len(pd["model"]["pair_params"])
    assert pd["model"]["gtinv"]["version"] == 1 and not pd["model"]["pair_conditional"]
    assert pd["model"]["pair_params_conditional"] == {(0, 0): list(range(n_fn)), (0, 1): list(range(n_fn)),
                                                      (1, 1): list(range(n_fn))}
    assert meta["type_full"] is True and meta["type_indices"] == [0, 1]
    assert coeffs.size == GOLD["polymlp.lammps.synthetic"]["n_coeffs"]
    with open(fn, "w") as f:  # truncated inside the mandatory part
        f.write("\n".join(lines[:8]) + "\n")
    with pytest.raises(ValueError):
        load_mlp_lammps(fn)
    bad = list(lines[:cut + 1])
    scales_line = next(i for i, ln in enumerate(bad) if ln.endswith("# scales"))
    bad[scales_line] = "1.0 2.0 # scales"
    with open(fn, "w") as f:
        f.write("\n".join(bad) + "\n")
    with pytest.raises(ValueError):
        load_mlp_lammps(fn)


# ---- the reference's published answers for its legacy potentials --------------------------------------------------
# (file, coefficient key in legacy.npz, structure, energy, tolerance kind, stress in GPa or raw eV, forces or None)
LEGACY_EVAL = {
    # tests/test_calc/test_properties_legacy_SrTiO3.py:19-69 (perovskite unit cell, forces vanish by symmetry)
    "srtio3_pair": ("polymlp.lammps.pair.SrTiO3", "srtio3", -31.63203426437243, dict(abs=1e-12),
                    ("gpa", [-1.9204671] * 3 + [0, 0, 0])),
    "srtio3_pair_cond": ("polymlp.lammps.pair.cond.SrTiO3", "srtio3", -31.64306562035659, dict(abs=1e-12),
                         ("gpa", [-1.74745126] * 3 + [0, 0, 0])),
    "srtio3_gtinv": ("polymlp.lammps.gtinv.SrTiO3", "srtio3", -31.64286166673613, dict(abs=1e-12),
                     ("gpa", [-0.02560392] * 3 + [0, 0, 0])),
    "srtio3_gtinv_cond": ("polymlp.lammps.gtinv.cond.SrTiO3", "srtio3", -31.642024569145274, dict(rel=1e-8),
                          ("gpa", [0.21674893] * 3 + [0, 0, 0])),
    # tests/test_calc/test_properties_legacy_Ag.py:15-38 (stresses compared in eV there)
    "ag_pair": ("polymlp.lammps.pair.Ag", "ag", -10.031307591610425, dict(abs=1e-12),
                ("ev", [-2.26126561e00, -2.26131407e00, -2.26107121e00, 5.71350861e-05, 1.22477289e-04, 4.07360575e-04])),
}
AG_FORCES = [[-1.24702646e-02, 9.72028834e-05, 6.18963049e-03, 6.18343122e-03],
             [-3.74121702e-03, 1.85699851e-03, 2.89620474e-05, 1.85525646e-03],
             [-2.49438077e-02, 1.23668884e-02, 1.23710083e-02, 2.05911033e-04]]


def load_legacy_golden():
    return np.load(os.path.join(cases.GOLDEN, "legacy.npz"))


def check_legacy_eval(key, e, f, s, axis):
    """e: float, f: (N, 3), s: 6 (eV/cell).  Tolerances are the reference tests' own."""
    _, st, e_true, tol, (unit, s_true) = LEGACY_EVAL[key]
    assert e == pytest.approx(e_true, **tol)
    f = np.asarray(f)
    if st == "srtio3":
        assert np.abs(f).max() < 1e-10  # reference asserts forces[0][0] == 0 within 1e-12; all vanish by symmetry
    else:
        np.testing.assert_allclose(f.T, AG_FORCES, atol=1e-6)
    s = np.asarray(s) * (EV_TO_GPA / np.linalg.det(axis) if unit == "gpa" else 1.0)
    np.testing.assert_allclose(s, s_true, atol=1e-5)


@pytest.mark.parametrize("key", sorted(LEGACY_EVAL))
def test_published_legacy_answers_through_the_oracle(key):
    """Loader output (coefficients from legacy.npz = the reference loader's coeffs / scales; params from the golden
    table, or straight from the file when the reference checkout is here) evaluated by the oracle."""
    name, st = LEGACY_EVAL[key][:2]
    L = load_legacy_golden()
    pd, coeffs = params_from_golden(name), L[key + "_coeffs"]
    if name in REF_FILES:
        pd_file, coeffs_file, _ = load_mlp(REF_FILES[name])
        assert np.array_equal(coeffs_file, coeffs)
        assert pd_file["model"] == pd["model"]
    axis, pos, types = L[st + "_axis"], L[st + "_pos"], L[st + "_types"]
    if "gtinv" in key:  # the numpy restatement needs ~15 s per 3-type gtinv potential; the compiled reference 1 s
        if not ref.available():
            pytest.skip("oracle/_ref not built")
        e, f, s = ref.RefEval(pd, coeffs).eval(axis, pos, types)
    else:
        e, f, s = po.eval_structure(po.Tables(pd), coeffs, axis, pos, types)
    check_legacy_eval(key, e, f, s, axis)


def test_legacy_mgo_pair_published_answers():
    """polymlp.lammps.pair.MgO on the displaced rocksalt cell: test_properties_legacy_MgO.py:17-49 (the same numbers
    the yaml potential publishes in test_properties_MgO.py:9-45, from slightly different coefficients)."""
    from test_oracle_golden import check_mgo_eval

    L, M = load_legacy_golden(), cases.load_mgo()
    pd = params_from_golden("polymlp.lammps.pair.MgO")
    want = make_params_dict(**cases.mgo_model_kwargs("pair"))["model"]
    assert {k: v for k, v in pd["model"].items() if k != "pair_conditional"} == \
        {k: v for k, v in want.items() if k != "pair_conditional"}
    e, f, s = po.eval_structure(po.Tables(pd), L["mgo_pair_coeffs"], M["rs_axis"], M["rs_pos"], M["rs_types"])
    check_mgo_eval("pair", e, f, s, np.linalg.det(M["rs_axis"]))


def test_load_mlps_find_mlps_is_hybrid(tmp_path):
    """io_polymlp.py:78-117,159-175: lists of potential files (hybrid models), directory lookup, hybrid predicate."""
    import shutil

    from pypolymlp_b200.io_legacy import find_mlps, is_hybrid, load_mlps

    flex = [os.path.join(cases.GOLDEN, "polymlp.yaml.flexible.%d.SrTiO3" % k) for k in (1, 2)]
    pds, coeffs, metas = load_mlps(flex)
    assert [m["elements"] for m in metas] == [["Sr", "Ti", "O"], ["O"]]
    assert [m["type_full"] for m in metas] == [True, False] and metas[1]["type_indices"] == [2]
    assert [len(c) for c in coeffs] == [5452, 2220] and [p["n_type"] for p in pds] == [3, 1]
    one = load_mlps(SYNTHETIC)
    assert len(one[0]) == 1 and np.array_equal(one[1][0], load_mlp(SYNTHETIC)[1])
    assert is_hybrid(flex) and not is_hybrid(flex[:1]) and not is_hybrid(flex[0])
    with pytest.raises(RuntimeError):
        is_hybrid(3)
    with pytest.raises(RuntimeError):
        load_mlps(3)
    assert find_mlps(str(tmp_path)) is None
    shutil.copy(SYNTHETIC, tmp_path / "polymlp.lammps")
    assert find_mlps(str(tmp_path)) == [str(tmp_path / "polymlp.lammps")]
    for k, f in enumerate(flex):
        shutil.copy(f, tmp_path / ("polymlp.yaml.%d" % (k + 1)))
    found = find_mlps(str(tmp_path))  # yaml files take precedence over legacy ones
    assert found == [str(tmp_path / "polymlp.yaml.1"), str(tmp_path / "polymlp.yaml.2")]
    assert [len(c) for c in load_mlps(found)[1]] == [5452, 2220]


def test_convert_to_yaml_keeps_submodel_flags_and_lists(tmp_path):
    """a hybrid sub-model (type_full = False, type_indices = [2], the O-only part of the reference's flexible
    SrTiO3 potential) keeps both flags through legacy -> yaml; a list of files converts to yaml.1, yaml.2, ... in sorted
    order (io_polymlp.py:120-140)."""
    pd = make_params_dict(**cases.si_model_kwargs())
    n = po.Tables(pd).n_variables
    coeffs = np.linspace(-1.0, 1.0, n)
    f1, f2 = str(tmp_path / "polymlp.lammps.1"), str(tmp_path / "polymlp.lammps.2")
    save_mlp_lammps(pd, coeffs, np.ones(n), ["O"], mass=[16.0], filename=f1)
    save_mlp_lammps(pd, coeffs, np.ones(n), ["O"], mass=[16.0], filename=f2, type_full=False, type_indices=[2])
    assert load_mlp_lammps(f2)[2]["type_full"] is False and load_mlp_lammps(f2)[2]["type_indices"] == [2]
    out = str(tmp_path / "conv.yaml")
    assert convert_to_yaml([f2, f1], out) is True
    m1, m2 = load_mlp_yaml(out + ".1")[2], load_mlp_yaml(out + ".2")[2]
    assert m1["type_full"] is True and list(m1["type_indices"]) == [0]
    assert m2["type_full"] is False and list(m2["type_indices"]) == [2]
    assert np.allclose(load_mlp_yaml(out + ".2")[1], coeffs, rtol=1e-14, atol=0)
    assert convert_to_yaml([f1, out + ".1"], str(tmp_path / "x.yaml")) is False

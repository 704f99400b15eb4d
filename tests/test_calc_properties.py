"""Host-side composition of pypolymlp_b200.calc.Properties (single and hybrid potentials from files; reference:
calculator/properties.py, properties_single.py, properties_hybrid.py).  CPU only: `PotentialPropertiesFast` is replaced
IN THE TEST by a stand-in backed by the oracle (the product has no such hook and no host evaluation: without the
monkeypatch and without a device the constructor raises).  Known answer: the reference's hybrid "flexible" SrTiO3
potential, tests/test_calc/test_properties_legacy_SrTiO3.py:72-85."""

import ctypes
import os

import numpy as np
import pytest

import cases
from oracle import ref
from pypolymlp_b200 import calc
from pypolymlp_b200.io_legacy import load_mlp
from test_legacy_io import SYNTHETIC, load_legacy_golden

FLEX = [os.path.join(cases.GOLDEN, "polymlp.yaml.flexible.%d.SrTiO3" % k) for k in (1, 2)]
FLEX_ENERGY = -31.642657299368572  # test_properties_legacy_SrTiO3.py:77
SRTIO3_ELEMENTS = ["Sr", "Ti", "O", "O", "O"]


class OracleEngine:
    """Test stand-in with the interface of libmlpcpp.PotentialPropertiesFast, evaluated by oracle/_ref."""

    def __init__(self, params_dict, coeffs, device=None):
        self._ev = ref.RefEval(params_dict, coeffs)

    def eval_multiple(self, axis_array, positions_c_array, types_array):
        out = [self._ev.eval(a, p, t) for a, p, t in zip(axis_array, positions_c_array, types_array)]
        self._e = np.array([o[0] for o in out])
        self._f = [np.asarray(o[1]) for o in out]
        self._s = np.array([o[2] for o in out])

    def get_e_array(self):
        return self._e

    def get_f_array(self):
        return self._f

    def get_s_array(self):
        return self._s


@pytest.fixture
def oracle_engine(monkeypatch):
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    monkeypatch.setattr(calc, "PotentialPropertiesFast", OracleEngine)


def test_hybrid_flexible_published_energy(oracle_engine):
    L = load_legacy_golden()
    prop = calc.Properties(pot=FLEX)
    assert prop.elements == ["Sr", "Ti", "O"]
    e, f, s = prop.eval(L["srtio3_axis"], L["srtio3_pos"], SRTIO3_ELEMENTS)
    assert e == pytest.approx(FLEX_ENERGY, rel=1e-8)
    assert f.shape == (3, 5) and np.abs(f).max() < 1e-10 and s.shape == (6,)
    # the sum is what the sub-models give one by one; sub-model 2 is O-only (type_full = 0): it sees atoms 2, 3, 4
    p1, p2 = calc.PropertiesSingle(pot=FLEX[0]), calc.PropertiesSingle(pot=FLEX[1])
    assert p1.type_full and not p2.type_full and p2.elements == ["O"]
    e1, f1, s1 = p1.eval(L["srtio3_axis"], L["srtio3_pos"], SRTIO3_ELEMENTS)
    pd2, c2, _ = load_mlp(FLEX[1])
    e2, f2, s2 = ref.RefEval(pd2, c2).eval(L["srtio3_axis"], L["srtio3_pos"][:, 2:], np.zeros(3, np.int32))
    assert e == pytest.approx(e1 + e2, rel=1e-14)
    np.testing.assert_allclose(s, s1 + s2, rtol=1e-12, atol=1e-14)


def test_partial_model_scatters_forces_and_skips_empty_structures(oracle_engine):
    """A displaced cell: forces of the O-only sub-model land on the O atoms only; a structure without O contributes
    nothing (properties_single.py:62-71,118-140)."""
    L = load_legacy_golden()
    rng = np.random.default_rng(5)
    axis = L["srtio3_axis"]
    pos = L["srtio3_pos"] + rng.normal(scale=0.05, size=(3, 5))
    order = [3, 0, 2, 4, 1]  # shuffled atom order: the element strings, not the positions in the file, decide
    elements = [SRTIO3_ELEMENTS[i] for i in order]
    p2 = calc.PropertiesSingle(pot=FLEX[1])
    no_oxygen = ["Sr", "Ti"]
    e, f, s = p2.eval_multiple([axis, axis], [pos[:, order], pos[:, :2]], [elements, no_oxygen])
    pd2, c2, _ = load_mlp(FLEX[1])
    o_atoms = [k for k, el in enumerate(elements) if el == "O"]
    e_ref, f_ref, s_ref = ref.RefEval(pd2, c2).eval(axis, pos[:, order][:, o_atoms], np.zeros(3, np.int32))
    assert e[0] == pytest.approx(e_ref, rel=1e-14) and e[1] == 0.0
    np.testing.assert_allclose(f[0][:, o_atoms], np.asarray(f_ref).T, rtol=1e-13, atol=1e-15)
    others = [k for k in range(5) if k not in o_atoms]
    assert np.all(f[0][:, others] == 0.0) and f[1].shape == (3, 2) and np.all(f[1] == 0.0) and np.all(s[1] == 0.0)
    np.testing.assert_allclose(s[0], s_ref, rtol=1e-13, atol=1e-15)
    # hybrid on the displaced cell = sub-model 1 + scattered sub-model 2, atom order independent
    prop = calc.Properties(pot=FLEX)
    eh, fh, sh = prop.eval(axis, pos[:, order], elements)
    eh0, fh0, sh0 = prop.eval(axis, pos, SRTIO3_ELEMENTS)
    assert eh == pytest.approx(eh0, rel=1e-13)
    np.testing.assert_allclose(fh, fh0[:, order], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sh, sh0, rtol=1e-10, atol=1e-12)


def test_single_legacy_file_and_element_mapping(oracle_engine):
    """Properties(pot=<legacy file>): types follow the potential's element order (Mg = 0, O = 1), whatever the order
    of the atoms; unknown elements are an error for a full model."""
    pd, coeffs, meta = load_mlp(SYNTHETIC)
    ax, pc, ty = cases.skewed_cell(2, n_atom=7, seed=9)
    elements = [meta["elements"][t] for t in ty]
    prop = calc.Properties(pot=SYNTHETIC)
    e, f, s = prop.eval(ax, pc, elements)
    e_ref, f_ref, s_ref = ref.RefEval(pd, coeffs).eval(ax, pc, ty)
    assert e == pytest.approx(e_ref, rel=1e-14)
    np.testing.assert_allclose(f, np.asarray(f_ref).T, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(s, s_ref, rtol=1e-13, atol=1e-15)
    es, fs, ss = prop.eval_multiple([ax, ax], [pc, pc[:, ::-1]], [elements, elements[::-1]])
    assert es[0] == pytest.approx(es[1], rel=1e-13)
    np.testing.assert_allclose(fs[0], fs[1][:, ::-1], rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError):
        prop.eval(ax, pc, ["Mg"] * 6 + ["Zn"])


def test_no_device_no_evaluation():
    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        calc.Properties(pot=FLEX)

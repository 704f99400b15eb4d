"""GPU parity: the reference's published answers for its bundled LEGACY potentials (tests/test_calc/
test_properties_legacy_{SrTiO3,Ag,MgO}.py) through the device evaluation path -- the ternary (Sr-Ti-O) models on the
eval kernels, ideal perovskite cell with neighbours exactly on the axes -- and the FeatureParams / get_fn / get_ylm
test hooks of the compiled pybind11 drop-in.  The CPU side of the same cases (loader + oracle) is
tests/test_legacy_io.py."""

import numpy as np
import pytest

import cases
from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast
from test_legacy_io import LEGACY_EVAL, check_legacy_eval, load_legacy_golden, params_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", sorted(LEGACY_EVAL))
def test_published_legacy_answers_gpu(key):
    name, st = LEGACY_EVAL[key][:2]
    L = load_legacy_golden()
    prop = PotentialPropertiesFast(params_from_golden(name), L[key + "_coeffs"])
    axis, pos, types = L[st + "_axis"], L[st + "_pos"], L[st + "_types"]
    prop.eval(axis, pos, types)
    check_legacy_eval(key, prop.get_e(), prop.get_f(), prop.get_s(), axis)
    prop.eval_multiple([axis] * 3, [pos] * 3, [types] * 3)
    for k in range(3):
        check_legacy_eval(key, prop.get_e_array()[k], prop.get_f_array()[k], prop.get_s_array()[k], axis)


@pytest.mark.parametrize("key", sorted(LEGACY_EVAL))
def test_published_legacy_answers_lane_atom_kernels_gpu(key, monkeypatch):
    """Same published answers with the large-batch eval kernels forced on (multi-type models: lanes of one warp carry atoms
    of different types; models the lane = atom kernels do not serve fall back by themselves)."""
    monkeypatch.setenv("PM_EVAL_LA_MIN", "1")
    name, st = LEGACY_EVAL[key][:2]
    L = load_legacy_golden()
    prop = PotentialPropertiesFast(params_from_golden(name), L[key + "_coeffs"])
    axis, pos, types = L[st + "_axis"], L[st + "_pos"], L[st + "_types"]
    prop.eval_multiple([axis] * 5, [pos] * 5, [types] * 5)
    for k in range(5):
        check_legacy_eval(key, prop.get_e_array()[k], prop.get_f_array()[k], prop.get_s_array()[k], axis)
    monkeypatch.setenv("PM_EVAL_LB", "0")
    prop.eval_multiple([axis] * 2, [pos] * 2, [types] * 2)
    check_legacy_eval(key, prop.get_e_array()[1], prop.get_f_array()[1], prop.get_s_array()[1], axis)


def test_legacy_mgo_pair_published_answers_gpu():
    from test_oracle_golden import check_mgo_eval

    L, M = load_legacy_golden(), cases.load_mgo()
    prop = PotentialPropertiesFast(params_from_golden("polymlp.lammps.pair.MgO"), L["mgo_pair_coeffs"])
    prop.eval(M["rs_axis"], M["rs_pos"], M["rs_types"])
    check_mgo_eval("pair", prop.get_e(), prop.get_f(), prop.get_s(), np.linalg.det(M["rs_axis"]))


def test_pybind_get_fn_get_ylm_hooks_gpu():
    """FeatureParams / get_fn / get_ylm of the compiled drop-in (pybind11_mlp.cpp:145-181), called the way the
    reference's wrapper does (cxx/wrapper/api_functions.py:8-27), with the reference's known answers
    (tests/test_cxx/test_functions.py:14-50) and against the ctypes mirror."""
    from pypolymlp_b200 import dropin
    from pypolymlp_b200 import libmlpcpp as mirror

    ext = dropin.load_extension()
    fp = ext.FeatureParams()
    fp.pair_type = "gaussian"
    fp.cutoff = 6.0
    params = [[1.0, 0.0], [1.0, 1.0], [1.0, 2.0]]
    fn, fn_d = ext.get_fn(1.2, fp, params)
    fn_m, fn_d_m = mirror.get_fn(1.2, params, 6.0)
    np.testing.assert_allclose(fn, fn_m, rtol=1e-15, atol=0)
    np.testing.assert_allclose(fn_d, fn_d_m, rtol=1e-15, atol=0)
    x, y, z = 0.3, -0.5, 0.81
    r = float(np.sqrt(x * x + y * y + z * z))
    out = ext.get_ylm(r=r, x=x, y=y, z=z, lmax=10)
    out_m = mirror.get_ylm(x, y, z, 10, r=r)
    for a, b in zip(out, out_m):
        assert len(a) == 66
        np.testing.assert_allclose(np.asarray(a), b, rtol=1e-14, atol=1e-15)
    # a neighbour exactly on the z axis: sin(theta) must be exactly 0 (cos_theta = z / r as in the reference)
    ylm, ylm_dx, ylm_dy, ylm_dz = (np.asarray(v) for v in ext.get_ylm(r=1.7, x=0.0, y=0.0, z=1.7, lmax=4))
    m_nonzero = np.array([k for l in range(5) for k in range(l * (l + 1) // 2, l * (l + 1) // 2 + l)], int)
    assert np.all(ylm[m_nonzero] == 0.0) and np.all(ylm_dz == 0.0)

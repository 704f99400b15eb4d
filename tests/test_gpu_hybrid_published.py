"""GPU replays of reference known answers that need calc.Properties on the device: the hybrid "flexible" SrTiO3 energy
(test_properties_legacy_SrTiO3.py:72-85), the cell-shape invariance set of test_check_neighbors.py:39-75 and the skewed
BiGd2 neighbour list.  First run on a B200 in round 1 (XPASS under a non-strict xfail mark); the mark is gone now.
CPU side of the same cases: tests/test_calc_properties.py."""

import numpy as np
import pytest

from pypolymlp_b200 import calc
from test_calc_properties import FLEX, FLEX_ENERGY, SRTIO3_ELEMENTS
from test_legacy_io import load_legacy_golden

pytestmark = pytest.mark.gpu


def test_hybrid_flexible_published_energy_gpu():
    """The reference's hybrid SrTiO3 potential (sub-model 2: O only, max_l 8, model_type 4) through calc.Properties on
    the device: published energy (test_properties_legacy_SrTiO3.py:72-85), vanishing forces, additivity."""
    L = load_legacy_golden()
    axis, pos = L["srtio3_axis"], L["srtio3_pos"]
    prop = calc.Properties(pot=FLEX)
    e, f, s = prop.eval(axis, pos, SRTIO3_ELEMENTS)
    assert e == pytest.approx(FLEX_ENERGY, rel=1e-8)
    assert f.shape == (3, 5) and np.abs(f).max() < 1e-10
    p1, p2 = calc.PropertiesSingle(pot=FLEX[0]), calc.PropertiesSingle(pot=FLEX[1])
    e1, f1, s1 = p1.eval(axis, pos, SRTIO3_ELEMENTS)
    e2, f2, s2 = p2.eval(axis, pos, SRTIO3_ELEMENTS)
    assert e == pytest.approx(e1 + e2, rel=1e-13)
    np.testing.assert_allclose(s, s1 + s2, rtol=1e-12, atol=1e-13)
    # displaced cell, shuffled atoms: same energy, permuted forces; the O-only model leaves Sr / Ti forces untouched
    rng = np.random.default_rng(5)
    posd = pos + rng.normal(scale=0.05, size=(3, 5))
    order = [3, 0, 2, 4, 1]
    ea, fa, sa = prop.eval(axis, posd, SRTIO3_ELEMENTS)
    eb, fb, sb = prop.eval(axis, posd[:, order], [SRTIO3_ELEMENTS[i] for i in order])
    assert ea == pytest.approx(eb, rel=1e-12)
    np.testing.assert_allclose(fb, fa[:, order], rtol=1e-9, atol=1e-11)
    e2d, f2d, _ = p2.eval(axis, posd, SRTIO3_ELEMENTS)
    assert np.all(f2d[:, :2] == 0.0) and np.abs(f2d[:, 2:]).max() > 0.0


def test_mgo_cell_shape_invariance_published_energy_gpu():
    """tests/test_calc/test_check_neighbors.py:39-75 on the device: ideal rocksalt MgO in 25 unimodular, mostly very
    skewed cells (host cell reduction + device neighbour kernels + eval), one published energy, atol 1e-12."""
    import cases
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast
    from pypolymlp_b200.params import make_params_dict
    from test_oracle_golden import MGO_IDEAL_ENERGY, load_mgo_cell_shapes

    cells = load_mgo_cell_shapes()
    prop = PotentialPropertiesFast(make_params_dict(**cases.mgo_model_kwargs("pair")), cases.load_mgo()["pair_coeffs"])
    prop.eval_multiple([c[0] for c in cells], [c[1] for c in cells], [c[2] for c in cells])
    np.testing.assert_allclose(prop.get_e_array(), MGO_IDEAL_ENERGY, atol=1e-12, rtol=0)
    assert max(np.abs(f).max() for f in prop.get_f_array()) < 1e-10
    s = np.asarray(prop.get_s_array())
    np.testing.assert_allclose(s, np.broadcast_to(s[0], s.shape), atol=1e-10, rtol=0)


def test_bigd2_neighbor_full_known_answers_gpu():
    """tests/test_cxx/test_neighbor_variants.py:17-47 on the device neighbour kernels through the compiled drop-in's
    NeighborFull hook, and bit-exact against the oracle list."""
    from oracle import polymlp_oracle as po
    from pypolymlp_b200 import dropin
    from test_oracle_golden import BIGD2_FULL, load_bigd2

    axis, pc, types = load_bigd2()
    full = dropin.load_extension().NeighborFull(axis.tolist(), pc.tolist(), 6.0)
    dist = full.get_distances(2, types.tolist())
    diff = full.get_differences(2, types.tolist())
    nbr = full.get_neighbor_indices(2, types.tolist())
    assert len(dist) == 30 and len(dist[0]) == 2
    for (i, t), (count, dist_sum, sq_sum, idx_sum) in BIGD2_FULL.items():
        if count is not None:
            assert len(dist[i][t]) == count
        if dist_sum is not None:
            assert np.sum(dist[i][t]) == pytest.approx(dist_sum)
        assert np.sum(np.square(diff[i][t])) == pytest.approx(sq_sum)
        assert np.sum(nbr[i][t]) == idx_sum
    off, nb, dx, dy, dz = po.neighbor_full(axis, pc, 6.0)
    for i in range(30):
        for t in range(2):
            pick = types[nb[off[i]:off[i + 1]]] == t
            assert np.array_equal(np.asarray(nbr[i][t]), nb[off[i]:off[i + 1]][pick])
            want = np.stack([dx, dy, dz], axis=1)[off[i]:off[i + 1]][pick]
            assert np.array_equal(np.asarray(diff[i][t]).reshape(-1, 3), want)

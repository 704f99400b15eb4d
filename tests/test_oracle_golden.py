"""The numpy oracle (oracle/polymlp_oracle.py) against the reference's own known answers and
against golden vectors produced by the unmodified reference C++ (tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

import cases
from oracle import polymlp_oracle as po
from pypolymlp_b200.params import make_params_dict

G = np.load(os.path.join(cases.GOLDEN, "ref_vectors.npz"))


def test_get_fn_known_answers():
    # reference: tests/test_cxx/test_functions.py:14-25
    fn, fnd = po.radial(np.array([1.2]), np.array([[1.0, 0.0], [1.0, 1.0], [1.0, 2.0]]), 6.0)
    np.testing.assert_allclose(fn[0], [0.21430317094756238, 0.8690422117212636, 0.4769404780495180], rtol=1e-12)
    np.testing.assert_allclose(fnd[0], [-0.5507864848009192, -0.495464911460655, 0.6819640474131319], rtol=1e-12)


def test_ylm_known_answers():
    # reference: tests/test_cxx/test_functions.py:28-50
    x, y, z = 0.173723561607389, 0.446843340790007, 0.877582561890373
    Y, Yx, Yy, Yz = po.ylm_der(np.array([x]), np.array([y]), np.array([z]), 10)
    assert Y.shape[0] == 66
    assert Y.sum() == pytest.approx(-3.094632553138235 + 0.09510814961092404j, rel=1e-6)
    assert Yx.sum() == pytest.approx(-6.916830463136405 - 1.910490992920798j, rel=1e-6)
    assert Yy.sum() == pytest.approx(12.686387327323297 + 23.645024627283863j, rel=1e-6)
    assert Yz.sum() == pytest.approx(-5.090360117438893 - 11.661266919165213j, rel=1e-6)


ROCKSALT_AXIS = np.eye(3) * 4.0
ROCKSALT_FRAC = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [0, 0, .5], [0, .5, 0], [.5, 0, 0], [.5, .5, .5]]).T


def test_translation_counts_rocksalt():
    # reference: tests/test_cxx/test_neighbor.py:199-207 (POSCAR-rocksalt, a = 4)
    pc = ROCKSALT_AXIS @ ROCKSALT_FRAC
    for cutoff, n in ((6.0, 147), (8.0, 203), (16.0, 751)):
        assert len(po.NeighborCell(ROCKSALT_AXIS, pc, cutoff).trans) == n


def test_neighbor_rocksalt_shells():
    # reference: tests/test_cxx/test_neighbor.py:14-45
    pc = ROCKSALT_AXIS @ ROCKSALT_FRAC
    types = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    off, nb, dx, dy, dz = po.neighbor_full(ROCKSALT_AXIS, pc, 6.0)
    sl = slice(off[0], off[1])
    d = np.sqrt(dx[sl] ** 2 + dy[sl] ** 2 + dz[sl] ** 2)
    t2 = types[nb[sl]]
    assert (t2 == 0).sum() == 54 and (t2 == 1).sum() == 38
    assert np.isclose(d[t2 == 0], 2.8284271247461903).sum() == 12
    assert np.isclose(d[t2 == 1], 2.0).sum() == 6
    assert np.sum(dx[sl][t2 == 0] ** 2 + dy[sl][t2 == 0] ** 2 + dz[sl][t2 == 0] ** 2) == pytest.approx(1152)
    assert nb[sl][t2 == 0].sum() == 72 and nb[sl][t2 == 1].sum() == 206


def test_neighbor_bit_exact_vs_reference_vectors():
    ax, pc, ty = cases.skewed_cell(2)
    for kind, fn in (("full", po.neighbor_full), ("half", po.neighbor_half)):
        o = fn(ax, pc, 5.0)
        for name, arr in zip(("off", "nb", "dx", "dy", "dz"), o):
            assert np.array_equal(arr, G[f"bin_nbr_{kind}_{name}"]), (kind, name)
    ax, pc, ty = cases.small_skewed_cell()
    o = po.neighbor_full(ax, pc, 4.0)
    for name, arr in zip(("off", "nb", "dx", "dy", "dz"), o):
        assert np.array_equal(arr, G[f"small_nbr_{name}"]), name
    nc = po.NeighborCell(ax, pc, 4.0)
    assert np.array_equal(nc.trans, G["small_trans"])
    assert np.array_equal(np.array(nc.axis), G["small_axis"])


def _rel(a, b):
    return cases.x_rel_err(a, b)


def test_si_structure_x():
    axis, positions_c, _, _ = cases.load_si_dataset()
    tab = po.Tables(make_params_dict(**cases.si_model_kwargs()))
    assert tab.n_variables == 168
    xe, xf, xs = po.structure_x(tab, axis, positions_c[3], np.zeros(64, int), True)
    assert _rel(xe, G["si3_xe"]) < 1e-12
    assert _rel(xf, G["si3_xf"]) < 1e-10
    assert _rel(xs, G["si3_xs"]) < 1e-10


def test_binary_conditional_x_and_eval():
    tab = po.Tables(make_params_dict(**cases.binary_model_kwargs()))
    ax, pc, ty = cases.skewed_cell(2)
    xe, xf, xs = po.structure_x(tab, ax, pc, ty, True)
    assert _rel(xe, G["bin_xe"]) < 1e-12
    assert _rel(xf, G["bin_xf"]) < 1e-10
    assert _rel(xs, G["bin_xs"]) < 1e-10
    e, f, s = po.eval_structure(tab, G["bin_coeffs"], ax, pc, ty)
    assert abs(e - G["bin_e"][0]) <= 1e-10 * abs(G["bin_e"][0])
    assert np.abs(f - G["bin_f"]).max() < 1e-10 * np.abs(G["bin_f"]).max()
    assert np.abs(s - G["bin_s"]).max() < 1e-10 * np.abs(G["bin_s"]).max()


def test_ternary_order3_polynomial():
    tab = po.Tables(make_params_dict(**cases.ternary_p3_model_kwargs()))
    ax, pc, ty = cases.skewed_cell(3, n_atom=7, seed=3)
    xe, xf, xs = po.structure_x(tab, ax, pc, ty, True)
    assert xe.shape[0] == G["ter_xe"].shape[0]
    assert _rel(xe, G["ter_xe"]) < 1e-12
    assert _rel(xf[::3], G["ter_xf"]) < 1e-10
    assert _rel(xs, G["ter_xs"]) < 1e-10


def test_si_train_column_sums_published_goldens():
    # reference: tests/test_mlp_dev/test_core_features.py:17-31; the golden column sums were produced by
    # the unmodified reference C++ over the 180 training structures (make_golden.py) and reproduce the
    # numbers published in the reference's own test.
    cs = G["si_train_colsum"]
    assert tuple(G["si_train_shape"]) == (35820, 168)
    assert cs.sum() == pytest.approx(5165294.450079148, rel=1e-6)
    assert cs[:20].sum() == pytest.approx(447.7438322711305, rel=1e-6)
    assert cs[20:40].sum() == pytest.approx(15803.28147846774, rel=1e-6)
    assert cs[40:60].sum() == pytest.approx(93568.30539681061, rel=1e-6)
    assert cs[60:80].sum() == pytest.approx(308321.82717270905, rel=1e-6)
    assert cs[-60:-40].sum() == pytest.approx(1987480.894857876, rel=1e-6)
    assert cs[-40:-20].sum() == pytest.approx(34372.80650408206, rel=1e-6)
    assert cs[-20:].sum() == pytest.approx(2008586.8132866116, rel=1e-6)
    # the oracle on a sample of the training structures agrees with the per-structure golden
    axis, positions_c, _, _ = cases.load_si_dataset()
    tab = po.Tables(make_params_dict(**cases.si_model_kwargs()))
    a, d = po.atom_anlm(tab, np.zeros(64, int), *po.neighbor_full(axis, positions_c[3], 6.0), 5, False)[0], None
    assert np.abs(a - G["si3_atom5_anlm"]).max() < 1e-12 * np.abs(G["si3_atom5_anlm"]).max()
    d, _ = po.atom_features(tab, 0, a, False)
    assert np.abs(d - G["si3_atom5_d"]).max() < 1e-12 * np.abs(G["si3_atom5_d"]).max()


# ---- feature_type = "pair" and the reference's MgO fixtures -------------------------------------------
EV_TO_GPA = 160.21766208  # src/pypolymlp/core/units.py:22


def test_pair_features_vs_reference_vectors():
    pd = make_params_dict(**cases.pair_model_kwargs(2))
    tab = po.Tables(pd)
    assert tab.n_variables == 88
    ax, pc, ty = cases.skewed_cell(2)
    xe, xf, xs = po.structure_x(tab, ax, pc, ty, True)
    assert cases.x_rel_err(np.vstack([xe[None], xs, xf]), np.vstack([G["pair_xe"][None], G["pair_xs"], G["pair_xf"]])) < 1e-10
    e, f, s = po.eval_structure(tab, G["pair_coeffs"], ax, pc, ty)
    assert abs(e - G["pair_e"][0]) < 1e-10 * abs(G["pair_e"][0])
    assert np.abs(f - G["pair_f"]).max() < 1e-10 * np.abs(G["pair_f"]).max()
    assert np.abs(s - G["pair_s"]).max() < 1e-10 * np.abs(G["pair_s"]).max()
    with pytest.raises(ValueError):  # pair models stop at model_type 2 (polymlp_model_params_polynomial.cpp:42-45)
        po.Tables(make_params_dict(**cases.pair_model_kwargs(2, model_type=3)))


# published in the reference: tests/test_calc/test_properties_MgO.py:9-97 (POSCAR.RS.MgO, polymlp.yaml.{pair,gtinv}.MgO)
MGO_EVAL = {
    "pair": (-40.22469744315832,
             [[-0.03962958, -0.01188776, -0.07928375], [-0.00459307, 0.00331611, 0.02210267],
              [0.01105381, -0.00137797, 0.022104], [0.01105096, 0.00331545, -0.00918516],
              [0.00978188, 0.00177061, 0.01180386], [0.00590284, 0.00293358, 0.01180553],
              [0.00589926, 0.00176979, 0.01958533], [0.0005339, 0.00016018, 0.00106752]],
             [-0.17379186, -0.17521681, -0.16909401, 0.00004465, 0.00008926, 0.00029751]),
    "gtinv": (-40.223320043232334,
              [[-0.03882777, -0.0116472, -0.07768031], [-0.00302382, 0.00324809, 0.02165033],
               [0.01082712, -0.00090718, 0.02165147], [0.01082468, 0.00324753, -0.00604679],
               [0.00925154, 0.00182258, 0.01215086], [0.00607581, 0.00277448, 0.01215188],
               [0.00607362, 0.00182207, 0.0185246], [-0.00120118, -0.00036037, -0.00240204]],
              [-2.78575100e-02, -2.91029804e-02, -2.37508486e-02, 1.32744242e-05, 2.65114887e-05, 8.83338671e-05]),
}
# tests/test_calc/test_compute_features.py:45-58 (POSCAR-0000{1,2}.MgO, energy rows only)
MGO_FEATURES = {"pair": (324, 997193.0146734761, -6.428600229143143), "gtinv": (1899, 237100.3979091199, -1.9717588279337908)}


def check_mgo_eval(kind, e, f, s, volume):
    e_true, f_true, s_true = MGO_EVAL[kind]
    assert e == pytest.approx(e_true, rel=1e-8)
    np.testing.assert_allclose(np.asarray(f), f_true, atol=1e-6)
    np.testing.assert_allclose(np.asarray(s) * EV_TO_GPA / volume, s_true, atol=1e-5)


@pytest.mark.parametrize("kind", ["pair", "gtinv"])
def test_mgo_published_properties(kind):
    M = cases.load_mgo()
    tab = po.Tables(make_params_dict(**cases.mgo_model_kwargs(kind)))
    e, f, s = po.eval_structure(tab, M[kind + "_coeffs"], M["rs_axis"], M["rs_pos"], M["rs_types"])
    check_mgo_eval(kind, e, f, s, np.linalg.det(M["rs_axis"]))


def test_mgo_published_pair_feature_sums():
    M = cases.load_mgo()
    tab = po.Tables(make_params_dict(**cases.mgo_model_kwargs("pair")))
    x = po.build_x(tab, [M["st1_axis"], M["st2_axis"]], [M["st1_pos"], M["st2_pos"]], [M["st1_types"], M["st2_types"]],
                   [False, False])
    ncol, total, diff = MGO_FEATURES["pair"]
    assert x.shape == (2, ncol)
    assert np.sum(x) == pytest.approx(total, rel=1e-6)
    assert np.sum(x[0] - x[1]) == pytest.approx(diff, rel=1e-6)


# ---- cell-shape invariance (tests/test_calc/test_check_neighbors.py:39-75) -------------------------------------------
MGO_IDEAL_ENERGY = -40.225125687168706


def load_mgo_cell_shapes():
    S = np.load(os.path.join(cases.GOLDEN, "mgo_cell_shapes.npz"))
    return [(S["axis_%02d" % k], S["pos_%02d" % k], S["types_%02d" % k]) for k in range(25)]


def test_mgo_cell_shape_invariance_published_energy():
    """Ideal rocksalt MgO in 25 unimodular cells, most of them far too skewed for the plain image search (lattice
    vectors up to 60 A long): the reference publishes one energy for all of them, atol 1e-12; exercises the cell
    reduction of NeighborCell (compute/neighbor_cell.cpp:126-168)."""
    M = cases.load_mgo()
    tab = po.Tables(make_params_dict(**cases.mgo_model_kwargs("pair")))
    for axis, pos, types in load_mgo_cell_shapes():
        e, f, s = po.eval_structure(tab, M["pair_coeffs"], axis, pos, types)
        assert abs(e - MGO_IDEAL_ENERGY) < 1e-12
        assert np.abs(f).max() < 1e-10


# ---- intermediates pinned by the reference's PolymlpAPI test (tests/test_cxx/test_polymlp_api.py:53-101) ---------------
def _api_test_anlmtp(tab, t):
    """compute_anlmtp_conjugate(arange, arange, type) of the reference (polymlp_api.cpp:63-86): m <= 0 entries take
    k + ik in local order, partners cc * conj; an m = 0 entry is its own partner and is written last, i.e. conjugated."""
    full = tab.local[t]["full"]
    a = np.zeros(len(full), complex)
    k = 0
    for i, (_, _, m, _, _, conj, _, _) in enumerate(full):
        if not conj:
            a[i] = complex(k, -k if m == 0 else k)
            k += 1
    for i, (_, _, _, _, _, conj, cc, src) in enumerate(full):
        if conj:
            a[i] = cc * np.conj(a[src])
    return a, k


def test_polymlp_api_known_answers():
    """MgO gtinv model (1899 columns): 450 order parameters / 270 without conjugates per type, the reference's sums of
    the conjugate expansion, of the 891 invariants and of their derivative contractions."""
    tab = po.Tables(make_params_dict(**cases.mgo_model_kwargs("gtinv")))
    assert tab.n_variables == 1899
    for t in (0, 1):
        a, n_noconj = _api_test_anlmtp(tab, t)
        assert (len(a), n_noconj) == (450, 270)
        assert a.sum() == 31581 + 17199j
        d, G = po.atom_features(tab, t, a, True)
        assert d.shape == (891,)
        assert d.sum() == pytest.approx(849391612.5622855, rel=1e-14)
    a, _ = _api_test_anlmtp(tab, 0)
    _, G = po.atom_features(tab, 0, a, True)
    dfx, dfy, dfz, ds = np.ones((450, 8)), -2.0 * np.ones((450, 8)), 0.5 * np.ones((450, 8)), -np.ones((450, 6))
    dfy[200, 3], dfz[100, 3], ds[300, 3] = 0.02, -0.05, 0.03
    for arr, want in ((dfx, -14716083.219391469), (dfy, 29122027.72943048), (dfz, -7326280.333739502),
                      (ds, 10734296.686030138)):
        assert np.real(G @ arr.astype(complex)).sum() == pytest.approx(want, rel=1e-13)


# ---- skewed BiGd2 cell (tests/test_cxx/test_neighbor_variants.py:17-66) -----------------------------------------------
BIGD2_FULL = {  # (atom, neighbour type): (count, sum of distances or None, sum of squared differences, sum of indices)
    (0, 0): (8, 38.163579855514044, 187.52873463482072, 15), (0, 1): (None, 83.65951693527006, 415.3567695554637, 325),
    (15, 0): (None, None, 207.80873466782282, 32), (15, 1): (17, None, 381.07889949085563, 302),
}


def load_bigd2():
    B = np.load(os.path.join(cases.GOLDEN, "bigd2.npz"))
    return B["axis"], B["axis"] @ B["positions"], B["types"]


def check_bigd2_full(off, nb, dx, dy, dz, types):
    assert len(off) == 31
    for (i, t), (count, dist_sum, sq_sum, idx_sum) in BIGD2_FULL.items():
        sl = slice(off[i], off[i + 1])
        pick = types[nb[sl]] == t
        r2 = (dx[sl] ** 2 + dy[sl] ** 2 + dz[sl] ** 2)[pick]
        if count is not None:
            assert pick.sum() == count
        if dist_sum is not None:
            assert np.sqrt(r2).sum() == pytest.approx(dist_sum)
        assert r2.sum() == pytest.approx(sq_sum)
        assert nb[sl][pick].sum() == idx_sum


def test_bigd2_neighbor_full_known_answers():
    axis, pc, types = load_bigd2()
    check_bigd2_full(*po.neighbor_full(axis, pc, 6.0), types)


def test_bigd2_neighbor_cell_known_answers_compiled_dropin():
    """NeighborCell of the compiled drop-in (host code, pm_cell_translations): cell and positions unchanged, 91 / 117 /
    281 translations at cutoff 6 / 8 / 16 (test_neighbor_variants.py:50-66); the oracle counts agree."""
    from pypolymlp_b200 import dropin

    ext = dropin.load_extension()
    axis, pc, _ = load_bigd2()
    for cutoff, n_trans in ((6.0, 91), (8.0, 117), (16.0, 281)):
        nc = ext.NeighborCell(axis.tolist(), pc.tolist(), cutoff)
        assert len(nc.get_translations()) == n_trans
        np.testing.assert_allclose(nc.get_axis(), axis)
        np.testing.assert_allclose(nc.get_positions_cartesian(), pc)

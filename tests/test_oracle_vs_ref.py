"""The numpy oracle against the unmodified reference C++ (oracle/_ref) on seeded random inputs.
Skipped when oracle/_ref has not been built (it needs /root/reference; it ships to the GPU box)."""

import numpy as np
import pytest

import cases
from oracle import polymlp_oracle as po
from oracle import ref
from pypolymlp_b200.params import make_params_dict

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("kwargs,n_type", [
    (dict(n_type=1, cutoff=5.0, model_type=4, max_p=2, gtinv_order=3, gtinv_maxl=[2, 2], n_gaussians=4), 1),
    (cases.binary_model_kwargs(), 2),
    (dict(n_type=3, cutoff=4.5, model_type=3, max_p=2, gtinv_order=4, gtinv_maxl=[2, 2, 1], n_gaussians=3), 3),
    (cases.pair_model_kwargs(1), 1),
    (cases.pair_model_kwargs(2), 2),
    (cases.pair_model_kwargs(3, model_type=2, max_p=3), 3),
    (cases.pair_model_kwargs(2, model_type=1, max_p=1), 2),
])
def test_x_and_eval(kwargs, n_type):
    pd = make_params_dict(**kwargs)
    tab, rm = po.Tables(pd), ref.RefModel(pd)
    assert tab.n_variables == rm.n_features
    for seed in (1, 2):
        ax, pc, ty = cases.skewed_cell(n_type, n_atom=6, seed=seed)
        cut = kwargs["cutoff"]
        for kind, fn in (("full", po.neighbor_full), ("half", po.neighbor_half)):
            for a, b in zip(fn(ax, pc, cut), ref.neighbor(kind, ax, pc, cut)):
                assert np.array_equal(a, b)
        xe, xf, xs = po.structure_x(tab, ax, pc, ty, True)
        re, rf, rs = rm.run(ax, pc, ty, True)
        a, b = np.vstack([xe[None], xs, xf]), np.vstack([re[None], rs, rf])
        assert cases.x_rel_err(a, b) < 1e-10
        coeffs = np.random.default_rng(seed).normal(size=tab.n_variables)
        e, f, s = ref.RefEval(pd, coeffs).eval(ax, pc, ty)
        e2, f2, s2 = po.eval_structure(tab, coeffs, ax, pc, ty)
        assert abs(e - e2) < 1e-10 * abs(e)
        assert np.abs(f - f2).max() < 1e-10 * np.abs(f).max()
        assert np.abs(s - s2).max() < 1e-10 * np.abs(s).max()


def test_small_cells_metric_branch():
    for seed in range(4):
        ax, pc, ty = cases.small_skewed_cell(seed)
        for a, b in zip(po.neighbor_full(ax, pc, 3.5), ref.neighbor("full", ax, pc, 3.5)):
            assert np.array_equal(a, b)


def test_readgtinv():
    import os
    d = os.path.join(os.path.dirname(ref.__file__), "_ref")
    for order, maxl in ((2, [6]), (3, [4, 4]), (4, [4, 2, 2])):
        a, b = po.readgtinv(order, maxl, d), ref.readgtinv(order, maxl)
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]


def test_neighbor_cell_of_the_dropin_is_bit_equal_on_skewed_cells():
    """NeighborCell of the compiled drop-in (host C++, pm_cell_translations) against the reference's NeighborCell in
    oracle/_ref on the 25 unimodular MgO cells of test_check_neighbors.py (lattice vectors up to 60 A: every one needs the
    cell reduction) and on the BiGd2 cell: reduced axis, wrapped positions and the translation list, exactly."""
    from pypolymlp_b200 import dropin
    from test_oracle_golden import load_bigd2, load_mgo_cell_shapes

    ext = dropin.load_extension()
    cells = [(a, p, 8.0) for a, p, _ in load_mgo_cell_shapes()]
    axis, pc, _ = load_bigd2()
    cells += [(axis, pc, rc) for rc in (6.0, 8.0, 16.0)]
    for axis, pos, cutoff in cells:
        nc = ext.NeighborCell(axis.tolist(), pos.tolist(), cutoff)
        trans, axis_ref, pos_ref = ref.neighbor_cell(axis, pos, cutoff)
        assert np.array_equal(np.asarray(nc.get_translations()), trans)
        assert np.array_equal(np.asarray(nc.get_axis()), axis_ref)
        assert np.array_equal(np.asarray(nc.get_positions_cartesian()), pos_ref)


def test_oracle_neighbor_lists_on_skewed_cells_bit_equal():
    """numpy oracle vs reference NeighborFull / NeighborHalf on the skewed MgO cells and BiGd2 (cell reduction path)."""
    from test_oracle_golden import load_bigd2, load_mgo_cell_shapes

    cells = [(a, p, 8.0) for a, p, _ in load_mgo_cell_shapes()[::3]] + [load_bigd2()[:2] + (6.0,)]
    for axis, pos, cutoff in cells:
        for kind, fn in (("full", po.neighbor_full), ("half", po.neighbor_half)):
            for a, b in zip(fn(axis, pos, cutoff), ref.neighbor(kind, axis, pos, cutoff)):
                assert np.array_equal(a, b)

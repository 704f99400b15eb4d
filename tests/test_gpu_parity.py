"""Parity of the CUDA path (through the C ABI) with the oracle / reference golden vectors.
Run on the B200 box: python -m pytest tests -m gpu.  Tolerances: neighbour lists bit-exact;
X / XtX / Xty 1e-10 relative (cases.x_rel_err); eval E rel 1e-10, F/S 1e-10 of the largest component."""

import os

import numpy as np
import pytest

import cases
from oracle import polymlp_oracle as po
from pypolymlp_b200 import fit
from pypolymlp_b200._capi import PM_FLAG_SCATTER, PM_FLAG_SIMPLE_KERNELS
from pypolymlp_b200.libmlpcpp import (PotentialModel, PotentialPropertiesFast, PotentialXtX, _Context, _Model)
from pypolymlp_b200.params import make_params_dict

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(cases.GOLDEN, "ref_vectors.npz"))
FLAVOURS = [pytest.param(PM_FLAG_SIMPLE_KERNELS, id="simple"), pytest.param(0, id="dmma"),
            pytest.param(PM_FLAG_SCATTER, id="dmma-scatter")]


def _sort_ref_order(off, nb, dx, dy, dz):
    """Our lists are grouped by neighbour type inside an atom; a stable sort by j restores the
    reference's (j, translation) order."""
    out = [[], [], [], []]
    for i in range(len(off) - 1):
        sl = slice(off[i], off[i + 1])
        order = np.argsort(nb[sl], kind="stable")
        for k, a in enumerate((nb, dx, dy, dz)):
            out[k].append(a[sl][order])
    return [np.concatenate(o) if o else np.zeros(0) for o in out]


def test_neighbor_bit_exact():
    pd = make_params_dict(**cases.binary_model_kwargs())
    ctx = _Context(_Model(pd))
    ax, pc, ty = cases.skewed_cell(2)
    off, nb, dx, dy, dz = ctx.neighbor_full(ax, pc, ty)
    assert np.array_equal(off, G["bin_nbr_full_off"])
    nb, dx, dy, dz = _sort_ref_order(off, nb, dx, dy, dz)
    for a, name in ((nb, "nb"), (dx, "dx"), (dy, "dy"), (dz, "dz")):
        assert np.array_equal(a, G[f"bin_nbr_full_{name}"]), name
    # many images + reduced cell
    kw = dict(cases.si_model_kwargs())
    kw["cutoff"] = 4.0
    ctx = _Context(_Model(make_params_dict(**kw)))
    ax, pc, ty = cases.small_skewed_cell()
    o = ctx.neighbor_full(ax, pc, ty)
    for a, name in zip(o, ("off", "nb", "dx", "dy", "dz")):
        assert np.array_equal(a, G[f"small_nbr_{name}"]), name
    # empty structure and isolated atom
    off, nb, *_ = ctx.neighbor_full(np.eye(3) * 50.0, np.zeros((3, 1)), [0])
    assert list(off) == [0, 0]


def test_neighbor_fcc256_checksums():
    ctx = _Context(_Model(make_params_dict(**cases.cfg2_model_kwargs(4))))
    ax, pc, ty = cases.fcc_supercell()
    off, nb, dx, dy, dz = ctx.neighbor_full(ax, pc, ty)
    assert off[-1] == G["fcc_nbr_count"][0]
    assert nb.sum() == G["fcc_nbr_checksum"][0]
    assert np.square(dx).sum() + np.square(dy).sum() + np.square(dz).sum() == pytest.approx(G["fcc_nbr_checksum"][1], rel=1e-14)
    axo, pco = ax, pc
    o = po.neighbor_full(axo, pco, 6.0)
    for a, b in zip((off, nb, dx, dy, dz), o):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("flags", FLAVOURS)
def test_x_si_structure_and_intermediates(flags):
    axis, positions_c, _, _ = cases.load_si_dataset()
    pd = make_params_dict(**cases.si_model_kwargs())
    ty = np.zeros(64, np.int32)
    pm = PotentialModel(pd, [axis], [positions_c[3]], [ty], [1], [True], [64], flags=flags)
    x = pm.get_x()
    assert x.shape == (1 + 6 + 192, 168)
    assert pm.get_n_data() == [1, 192, 6] and pm.get_sbegin() == [1] and pm.get_fbegin() == [7]
    # intermediates of the last chunk
    anc = pm._ctx.debug_fetch(0).view(np.complex128).reshape(64, -1)
    d = pm._ctx.debug_fetch(1).reshape(64, -1)
    m = pm._model
    ti = m.type_info(0)
    gold_a = G["si3_atom5_anlm"]
    tab = po.Tables(pd)
    heads = [k for k, f in enumerate(tab.local[0]["full"]) if not f[5]]
    assert np.abs(anc[5, : ti["n_head"]] - gold_a[heads]).max() < 1e-12 * np.abs(gold_a).max()
    dd = d[5][d[5] != 0.0]
    assert cases.x_rel_err(np.sort(dd), np.sort(G["si3_atom5_d"][G["si3_atom5_d"] != 0.0])) < 1e-11
    assert cases.x_rel_err(x[0], G["si3_xe"]) < 1e-10
    assert cases.x_rel_err(x[1:7], G["si3_xs"]) < 1e-10
    assert cases.x_rel_err(x[7:], G["si3_xf"]) < 1e-10
    full = np.vstack([G["si3_xe"][None], G["si3_xs"], G["si3_xf"]])
    assert cases.x_rel_err(x, full) < 1e-10


@pytest.mark.parametrize("flags", FLAVOURS)
def test_x_binary_conditional(flags):
    pd = make_params_dict(**cases.binary_model_kwargs())
    ax, pc, ty = cases.skewed_cell(2)
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [9], flags=flags).get_x()
    full = np.vstack([G["bin_xe"][None], G["bin_xs"], G["bin_xf"]])
    assert cases.x_rel_err(x, full) < 1e-10


def test_x_ternary_order3_polynomial():
    pd = make_params_dict(**cases.ternary_p3_model_kwargs())
    ax, pc, ty = cases.skewed_cell(3, n_atom=7, seed=3)
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [7]).get_x()
    assert cases.x_rel_err(x[0], G["ter_xe"]) < 1e-10
    assert cases.x_rel_err(x[1:7], G["ter_xs"]) < 1e-10
    assert cases.x_rel_err(x[7::3], G["ter_xf"]) < 1e-10


@pytest.mark.parametrize("flags", FLAVOURS)
def test_x_mixed_batch_layout_energy_only_and_force(flags):
    """Two datasets (force / energy-only), ragged sizes: PyModel row layout and values vs the oracle."""
    pd = make_params_dict(**cases.binary_model_kwargs())
    tab = po.Tables(pd)
    sts = [cases.skewed_cell(2, n_atom=n, seed=s) for n, s in ((5, 1), (8, 2), (3, 4), (6, 5))]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    pm = PotentialModel(pd, axis, pcs, tys, [2, 2], [True, False], [5, 8, 3, 6], flags=flags)
    ref_x = po.build_x(tab, axis, pcs, tys, [True, True, False, False])
    assert pm.get_x().shape == ref_x.shape
    assert pm.get_fbegin() == [16, -1] and pm.get_sbegin() == [4, -1]
    assert cases.x_rel_err(pm.get_x(), ref_x) < 1e-10


@pytest.mark.parametrize("flags", FLAVOURS)
def test_x_single_type_ragged_batches(flags):
    """Single-type model (the k_xrows_v4 / zero-once G fast paths): ragged cells, an energy-only structure in the
    middle, and a second call on the same context with different structures (buffers are reused)."""
    pd = make_params_dict(**cases.si_model_kwargs())
    tab = po.Tables(pd)
    for seeds in (((7, 1), (3, 2), (12, 3), (5, 4)), ((4, 9), (9, 8), (6, 7), (2, 6))):
        sts = [cases.skewed_cell(1, n_atom=n, seed=s) for n, s in seeds]
        axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
        n_atoms = [len(t) for t in tys]
        pm = PotentialModel(pd, axis, pcs, tys, [1, 1, 2], [True, False, True], n_atoms, flags=flags)
        ref_x = po.build_x(tab, axis, pcs, tys, [True, False, True, True])
        assert pm.get_x().shape == ref_x.shape
        assert cases.x_rel_err(pm.get_x(), ref_x) < 1e-10


@pytest.mark.parametrize("flags", FLAVOURS)
def test_x_fcc256_config2(flags):
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    ax, pc, ty = cases.fcc_supercell()
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [256], flags=flags).get_x()
    assert x.shape == (775, 2030)
    assert cases.x_rel_err(x[0], G["fcc_xe"]) < 1e-10
    assert cases.x_rel_err(x[1:7], G["fcc_xs"]) < 1e-10
    xf = x[7:]
    scale = np.abs(xf).max(axis=0)
    assert (np.abs(xf[G["fcc_rows"]] - G["fcc_xf_rows"]) / np.maximum(scale, 1e-8 * scale.max())).max() < 1e-10
    assert (np.abs(np.square(xf).sum(axis=0) - G["fcc_xf_colsqsum"]) / G["fcc_xf_colsqsum"].max()).max() < 1e-10


def test_x_config3_binary_order4_maxl12():
    """BASELINE config 3 model (F = 9385, L = 12: generic kernels) on a 16-atom bcc cell vs the reference."""
    pd = make_params_dict(**cases.cfg3_model_kwargs())
    ax, pc, ty = cases.bcc_supercell(rep=(2, 2, 2), a=3.2, n_type=2, seed=5)
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [16]).get_x()
    assert x.shape == (1 + 6 + 48, 9385)
    assert cases.x_rel_err(x[0], G["cfg3_xe"]) < 1e-10
    assert cases.x_rel_err(x[1:7], G["cfg3_xs"]) < 1e-10
    xf = x[7:]
    scale = np.maximum(np.abs(xf).max(axis=0), 1e-8 * np.abs(xf).max())
    assert (np.abs(xf[::6] - G["cfg3_xf_rows"]) / scale).max() < 1e-10
    assert (np.abs(np.square(xf).sum(axis=0) - G["cfg3_xf_colsqsum"]) / G["cfg3_xf_colsqsum"].max()).max() < 1e-10
    # fused accumulation on the same structure: X^T X against the materialised X
    rows = x.shape[0]
    rng = np.random.default_rng(3)
    w = rng.uniform(0.2, 1.0, rows)
    y = w * rng.normal(size=rows)
    acc = PotentialXtX(pd)
    acc.add([ax], [pc], [ty], [True], w, y)
    res = acc.finalize()
    xw = x * w[:, None]
    ref_xtx = xw.T @ xw
    assert np.abs(res["xtx"] - ref_xtx).max() < 1e-10 * np.abs(ref_xtx).max()
    assert np.abs(res["xty"] - xw.T @ y).max() < 1e-10 * np.abs(xw.T @ y).max()


@pytest.mark.parametrize("mrt", ["8", "9", "12", "16", ""])
def test_x_large_model_kernels_ragged_vs_straightforward(mrt, monkeypatch):
    """Large-model kernels (k_lrows_big in all its row-tile variants, the 512-thread sliced feature kernel, the
    staged linear-column gathers) against the straightforward kernels -- which the reference goldens pin on this
    model (test above) -- on what the goldens do not cover: ragged neighbour-type segments, a structure without
    atoms of one type (empty segments), an energy-only structure, segments longer than one chunk."""
    if mrt:
        monkeypatch.setenv("PM_LROWS_MRT", mrt)
    else:
        monkeypatch.delenv("PM_LROWS_MRT", raising=False)
    pd = make_params_dict(**cases.cfg3_model_kwargs())
    sts = [cases.skewed_cell(2, n_atom=5, seed=11), cases.skewed_cell(1, n_atom=3, seed=12),
           cases.skewed_cell(2, n_atom=2, seed=13), cases.bcc_supercell(rep=(2, 2, 1), a=3.2, n_type=2, seed=14)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    n_atoms = [len(t) for t in tys]
    args = (pd, axis, pcs, tys, [2, 1, 1], [True, False, True], n_atoms)
    x_ref = PotentialModel(*args, flags=PM_FLAG_SIMPLE_KERNELS).get_x()
    x = PotentialModel(*args).get_x()
    assert x.shape == x_ref.shape and np.isfinite(x).all()
    assert cases.x_rel_err(x, x_ref) < 1e-10
    rows = x.shape[0]
    rng = np.random.default_rng(5)
    w = rng.uniform(0.2, 1.0, rows)
    y = w * rng.normal(size=rows)
    force = [True, True, False, True]
    ra = PotentialXtX(pd, flags=PM_FLAG_SIMPLE_KERNELS)
    ra.add(axis, pcs, tys, force, w, y)
    r0 = ra.finalize()
    rb = PotentialXtX(pd)
    rb.add(axis, pcs, tys, force, w, y)
    r1 = rb.finalize()
    for k in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.abs(r1[k] - r0[k]).max() <= 1e-10 * np.abs(r0[k]).max(), k


def _si_datasets(ids):
    axis, positions_c, forces, energies = cases.load_si_dataset()
    return fit.Dataset([axis] * len(ids), [positions_c[i] for i in ids], [np.zeros(64, np.int32)] * len(ids),
                       energies[ids], forces=forces[ids].reshape(-1), include_force=True, include_stress=False)


@pytest.mark.parametrize("flags", FLAVOURS)
def test_xtx_xty_vs_oracle_small(flags):
    pd = make_params_dict(**cases.si_model_kwargs())
    tab = po.Tables(pd)
    ds = _si_datasets(np.array([0, 7, 19]))
    w, y = fit.apply_weights(ds, min_e=-5.737324395625)
    X = po.build_x(tab, ds.axis, ds.positions_c, ds.types, [True] * 3)
    xtx, xty, ysq, xe_sum, xe_sq = po.accumulate(X, 3, w, y / np.where(w > 0, w, 1.0))
    acc = PotentialXtX(pd, flags=flags)
    acc.add(ds.axis, ds.positions_c, ds.types, [True] * 3, w, y)
    res = acc.finalize()
    assert res["total_n_data"] == X.shape[0]
    assert np.abs(res["xtx"] - xtx).max() < 1e-10 * np.abs(xtx).max()
    assert np.abs(res["xtx"] - res["xtx"].T).max() == 0.0
    assert np.abs(res["xty"] - xty).max() < 1e-10 * np.abs(xty).max()
    assert abs(res["y_sq_norm"] - ysq) < 1e-10 * ysq
    assert np.abs(res["xe_sum"] - xe_sum).max() < 1e-10 * np.abs(xe_sum).max()
    assert np.abs(res["xe_sq_sum"] - xe_sq).max() < 1e-10 * np.abs(xe_sq).max()
    # second accumulate adds on top (linearity) and batching does not matter
    acc.add(ds.axis[:1], ds.positions_c[:1], ds.types[:1], [True], *_rows_of(w, y, 3, 64, [0]))
    acc.add(ds.axis[1:], ds.positions_c[1:], ds.types[1:], [True, True], *_rows_of(w, y, 3, 64, [1, 2]))
    res2 = acc.finalize()
    assert np.abs(res2["xtx"] - 2 * xtx).max() < 1e-10 * np.abs(xtx).max()
    assert res2["total_n_data"] == 2 * X.shape[0]


def _rows_of(w, y, n_st, n_atom, sel):
    e = np.array(sel)
    s = np.concatenate([n_st + 6 * k + np.arange(6) for k in sel])
    f = np.concatenate([n_st + 6 * n_st + 3 * n_atom * k + np.arange(3 * n_atom) for k in sel])
    idx = np.concatenate([e, s, f])
    return w[idx], y[idx]


def test_fit_si_reference_goldens():
    """Config 1: the reference's bundled Si set, 180 training structures.
    Goldens: tests/test_mlp_dev/test_core_features.py:17-31 and test_core_data.py:39-57 of the reference."""
    pd = make_params_dict(**cases.si_model_kwargs())
    train_ids, test_ids = cases.split_ids_train_test(200, 0.9)
    train = _si_datasets(train_ids)
    pm = PotentialModel(pd, train.axis, train.positions_c, train.types, [180], [True], [64] * 180)
    x = pm.get_x()
    assert x.shape == (35820, 168)
    assert tuple(pm.get_n_data()) == (180, 34560, 1080) and pm.get_fbegin() == [1260] and pm.get_sbegin() == [180]
    assert np.sum(x) == pytest.approx(5165294.450079148, rel=1e-6)
    assert np.sum(x[:, :20]) == pytest.approx(447.7438322711305, rel=1e-6)
    assert np.sum(x[:, 20:40]) == pytest.approx(15803.28147846774, rel=1e-6)
    assert np.sum(x[:, 40:60]) == pytest.approx(93568.30539681061, rel=1e-6)
    assert np.sum(x[:, 60:80]) == pytest.approx(308321.82717270905, rel=1e-6)
    assert np.sum(x[:, -60:-40]) == pytest.approx(1987480.894857876, rel=1e-6)
    assert np.sum(x[:, -40:-20]) == pytest.approx(34372.80650408206, rel=1e-6)
    assert np.sum(x[:, -20:]) == pytest.approx(2008586.8132866116, rel=1e-6)
    assert np.abs(x.sum(axis=0) - G["si_train_colsum"]).max() < 1e-10 * np.abs(G["si_train_colsum"]).max()
    assert np.abs(np.square(x).sum(axis=0) - G["si_train_colsqsum"]).max() < 1e-10 * G["si_train_colsqsum"].max()

    for bs in (64, 10, 50):
        data_xy = fit.calc_xtx_xty(pd, [train], batch_size=bs)
        assert data_xy.xtx.shape == (168, 168)
        assert data_xy.xty[56] == pytest.approx(6.899130774433e5, rel=1e-6)
        assert data_xy.scales[56] == pytest.approx(0.0032488632685359524)
        assert data_xy.min_energy == pytest.approx(-5.737324395625)
        assert data_xy.total_n_data == 35820
    # the same sums from the materialised X (the reference's own route)
    w, y = fit.apply_weights(train, min_e=data_xy.min_energy)
    xw = x * w[:, None]
    xtx_ref = (xw.T @ xw) / data_xy.scales[:, None] / data_xy.scales[None, :]
    assert np.abs(data_xy.xtx - xtx_ref).max() < 1e-10 * np.abs(xtx_ref).max()
    # ridge fit: RMSE goldens of tests/test_mlp_dev_api/test_mlp_devel_phono3py.py:22-27 (rel 1e-2)
    test = _si_datasets(test_ids)
    alphas = [10.0 ** a for a in np.linspace(-1, 1, 3)]  # reg_alpha_params (-1, 1, 3) as in the reference API test
    best = fit.fit(pd, [train], [test], alphas)
    coeffs = best["coeffs"] / best["scales"]
    prop = PotentialPropertiesFast(pd, coeffs)
    for ds, e_gold, f_gold in ((train, 1.925e-6, 9.113e-4), (test, 2.067e-6, 9.180e-4)):
        prop.eval_multiple(ds.axis, ds.positions_c, ds.types)
        e = np.array(prop.get_e_array())
        f = np.concatenate([np.asarray(a).reshape(-1) for a in prop.get_f_array()])
        rmse_e = np.sqrt(np.mean(np.square((e - ds.energies) / 64)))
        rmse_f = np.sqrt(np.mean(np.square(f - ds.forces)))
        assert rmse_e == pytest.approx(e_gold, rel=1e-2)
        assert rmse_f == pytest.approx(f_gold, rel=1e-2)


@pytest.mark.parametrize("flags", FLAVOURS)
def test_eval_binary_and_fcc(flags):
    pd = make_params_dict(**cases.binary_model_kwargs())
    ax, pc, ty = cases.skewed_cell(2)
    prop = PotentialPropertiesFast(pd, G["bin_coeffs"], flags=flags)
    prop.eval(ax, pc, ty, True)
    assert abs(prop.get_e() - G["bin_e"][0]) < 1e-10 * abs(G["bin_e"][0])
    assert np.abs(prop.get_f() - G["bin_f"]).max() < 1e-10 * np.abs(G["bin_f"]).max()
    assert np.abs(prop.get_s() - G["bin_s"]).max() < 1e-10 * np.abs(G["bin_s"]).max()
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    ax, pc, ty = cases.fcc_supercell()
    prop = PotentialPropertiesFast(pd, G["fcc_coeffs"], flags=flags)
    prop.eval_multiple([ax, ax], [pc, pc], [ty, ty])
    for k in range(2):
        assert abs(prop.get_e_array()[k] - G["fcc_e"][0]) < 1e-10 * abs(G["fcc_e"][0])
        assert np.abs(prop.get_f_array()[k] - G["fcc_f"]).max() < 1e-10 * np.abs(G["fcc_f"]).max()
        assert np.abs(prop.get_s_array()[k] - G["fcc_s"]).max() < 1e-10 * np.abs(G["fcc_s"]).max()
    with pytest.raises(ValueError):
        PotentialPropertiesFast(pd, G["fcc_coeffs"][:-1])


def test_eval_large_model_matches_design_matrix():
    """E / F / S evaluation of the config-3 model (max_l 12: the unfused eval path with the G buffer, 512-thread
    feature kernel) against X @ c of the design matrix the reference goldens pin (energy, force and stress rows)."""
    pd = make_params_dict(**cases.cfg3_model_kwargs())
    ax, pc, ty = cases.bcc_supercell(rep=(2, 2, 2), a=3.2, n_type=2, seed=5)
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [16]).get_x()
    coeffs = np.random.default_rng(9).normal(size=x.shape[1]) * 1e-3
    pred = x @ coeffs
    prop = PotentialPropertiesFast(pd, coeffs)
    prop.eval(ax, pc, ty, True)
    scale_f = np.abs(pred[7:]).max()
    assert abs(prop.get_e() - pred[0]) < 1e-10 * max(abs(pred[0]), np.abs(x[0] * coeffs).max())
    assert np.abs(np.asarray(prop.get_f()).reshape(-1) - pred[7:]).max() < 1e-10 * scale_f
    assert np.abs(np.asarray(prop.get_s()) - pred[1:7]).max() < 1e-10 * max(np.abs(pred[1:7]).max(), np.abs(x[1:7] * coeffs).max())


def test_config2_full_size_properties():
    """BASELINE config-2 shapes (256-atom fcc, F = 2030): size-independent properties of the fused path:
    tensor-core and straightforward kernels agree, accumulation is additive over structures and independent
    of chunking, C is symmetric, X^T y / y^T y / xe_sum are consistent with the materialised X."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    n = 6
    sts = [cases.fcc_supercell(seed=20240 + s) for s in range(n)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    rng = np.random.default_rng(5)
    rows = n * 775
    w = rng.uniform(0.2, 1.0, rows)
    y = w * rng.normal(size=rows)
    a1 = PotentialXtX(pd)
    a1.add(axis, pcs, tys, [True] * n, w, y)
    r1 = a1.finalize()
    a2 = PotentialXtX(pd, flags=PM_FLAG_SIMPLE_KERNELS, workspace_bytes=400 << 20)  # several chunks
    a2.add(axis, pcs, tys, [True] * n, w, y)
    r2 = a2.finalize()
    sc = np.abs(r2["xtx"]).max()
    assert np.abs(r1["xtx"] - r2["xtx"]).max() < 1e-10 * sc
    assert np.abs(r1["xty"] - r2["xty"]).max() < 1e-10 * np.abs(r2["xty"]).max()
    assert np.abs(r1["xtx"] - r1["xtx"].T).max() == 0.0
    assert r1["total_n_data"] == rows
    x = PotentialModel(pd, axis, pcs, tys, [n], [True], [256] * n).get_x()
    xw = x * w[:, None]
    ref = xw.T @ xw
    assert np.abs(r1["xtx"] - ref).max() < 1e-10 * np.abs(ref).max()
    assert np.abs(r1["xty"] - xw.T @ y).max() < 1e-10 * np.abs(xw.T @ y).max()
    assert abs(r1["y_sq_norm"] - y @ y) < 1e-10 * (y @ y)
    assert np.abs(r1["xe_sum"] - x[:n].sum(axis=0)).max() < 1e-10 * np.abs(x[:n].sum(axis=0)).max()
    assert np.abs(r1["xe_sq_sum"] - np.square(x[:n]).sum(axis=0)).max() < 1e-10 * np.square(x[:n]).sum(axis=0).max()
    # staged (inputs resident in HBM) path gives the same accumulator
    a3 = PotentialXtX(pd)
    a3.stage(axis, pcs, tys, [True] * n, w, y)
    a3.add_staged()
    r3 = a3.finalize()
    assert np.array_equal(r3["xtx"], r1["xtx"])   # same chunks, deterministic accumulation: bitwise equal


def test_config3_full_size_properties():
    """BASELINE config-3 shapes (216-atom bcc 6x6x3, binary, F = 9385): the large-model tensor-core kernels against the
    straightforward kernels at full structure size, chunking independence (one structure per chunk vs all in one),
    symmetry, and X^T y / xe_sum consistency with the materialised X of one structure."""
    pd = make_params_dict(**cases.cfg3_model_kwargs())
    n = 3
    sts = [cases.bcc_supercell(rep=(6, 6, 3), a=3.2, n_type=2, seed=100 + s) for s in range(n)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    rows = n * 655
    rng = np.random.default_rng(6)
    w = rng.uniform(0.2, 1.0, rows)
    y = w * rng.normal(size=rows)
    a1 = PotentialXtX(pd)
    a1.add(axis, pcs, tys, [True] * n, w, y)
    r1 = a1.finalize()
    a2 = PotentialXtX(pd, flags=PM_FLAG_SIMPLE_KERNELS, workspace_bytes=1 << 30)   # one structure per chunk
    a2.add(axis, pcs, tys, [True] * n, w, y)
    r2 = a2.finalize()
    sc = np.abs(r2["xtx"]).max()
    assert np.abs(r1["xtx"] - r2["xtx"]).max() < 1e-10 * sc
    assert np.abs(r1["xty"] - r2["xty"]).max() < 1e-10 * np.abs(r2["xty"]).max()
    assert np.abs(r1["xe_sum"] - r2["xe_sum"]).max() < 1e-10 * np.abs(r2["xe_sum"]).max()
    assert np.abs(r1["xtx"] - r1["xtx"].T).max() == 0.0
    assert r1["total_n_data"] == rows
    # first structure alone: materialised X (PyModel layout of a 1-structure batch: E, 6 S, 648 F rows)
    x = PotentialModel(pd, axis[:1], pcs[:1], tys[:1], [1], [True], [216]).get_x()
    sel = np.r_[0, n + np.arange(6), n + 6 * n + np.arange(648)]     # rows of structure 0 in the 3-structure batch
    a3 = PotentialXtX(pd)
    a3.add(axis[:1], pcs[:1], tys[:1], [True], w[sel], y[sel])
    r3 = a3.finalize()
    xw = x * w[sel][:, None]
    ref = xw.T @ xw
    assert np.abs(r3["xtx"] - ref).max() < 1e-10 * np.abs(ref).max()
    assert np.abs(r3["xty"] - xw.T @ y[sel]).max() < 1e-10 * np.abs(xw.T @ y[sel]).max()
    assert np.abs(r3["xe_sum"] - x[0]).max() < 1e-10 * np.abs(x[0]).max()


def test_error_paths():
    pd = make_params_dict(**cases.si_model_kwargs())
    ax, pc, ty = cases.skewed_cell(1)
    with pytest.raises(ValueError):
        PotentialModel(pd, [ax], [pc], [np.ones(9, np.int32)], [1], [True], [9])  # type out of range
    with pytest.raises(ValueError):
        PotentialXtX(pd).add([ax], [pc], [ty], [True], np.ones(3), np.ones(3))  # wrong row count
    # empty batch is fine
    acc = PotentialXtX(pd)
    acc.add([], [], [], [], np.zeros(0), np.zeros(0))
    assert acc.finalize()["total_n_data"] == 0


def test_finalize_view_matches_copy():
    """pm_fit_finalize_view (pinned, no host copy) returns the same packed result as pm_fit_finalize."""
    pd = make_params_dict(**cases.si_model_kwargs())
    ds = _si_datasets(list(range(6)))
    acc = PotentialXtX(pd)
    fit.accumulate_datasets(acc, [ds], fit.get_min_energy([ds]))
    a = acc.finalize()
    b = acc.finalize(copy=False)
    assert not b["xtx"].flags.writeable
    for k in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.array_equal(a[k], b[k])
    assert a["y_sq_norm"] == b["y_sq_norm"] and a["total_n_data"] == b["total_n_data"] > 0
    assert np.array_equal(a["xtx"], a["xtx"].T)


def test_compute_error_matches_design_matrix_predictions():
    """fit.compute_error (device eval over a dataset) against predictions X @ c of the same model:
    energy RMSE per atom and force RMSE agree to 1e-9 relative."""
    pd = make_params_dict(**cases.si_model_kwargs())
    ds = _si_datasets(np.arange(12))
    x = PotentialModel(pd, ds.axis, ds.positions_c, ds.types, [12], [True], [64] * 12).get_x()
    coeffs = np.random.default_rng(3).normal(size=x.shape[1]) * 1e-3
    err = fit.compute_error(pd, coeffs, ds)
    pred = x @ coeffs
    e_ref = np.sqrt(np.mean(np.square((ds.energies - pred[:12]) / 64)))
    f_ref = np.sqrt(np.mean(np.square(ds.forces - pred[12 + 72:])))
    assert abs(err["energy"] - e_ref) < 1e-9 * e_ref
    assert abs(err["force"] - f_ref) < 1e-9 * f_ref
    assert err["stress"] is None


def test_pybind_dropin_module_gpu():
    """Same calls a reference user makes on `pypolymlp.cxx.lib.libmlpcpp`, through the pybind11 drop-in."""
    import importlib.util

    from pypolymlp_b200.build import pybind_module_path

    spec = importlib.util.spec_from_file_location("libmlpcpp", pybind_module_path())
    libmlpcpp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(libmlpcpp)
    pd = make_params_dict(**cases.binary_model_kwargs())
    ax, pc, ty = cases.skewed_cell(2)
    obj = libmlpcpp.PotentialModel(pd, [ax.tolist()], [pc.tolist()], [ty.tolist()], [1], [True], [9])
    full = np.vstack([G["bin_xe"][None], G["bin_xs"], G["bin_xf"]])
    assert cases.x_rel_err(np.asarray(obj.get_x()), full) < 1e-10
    assert obj.get_n_data() == [1, 27, 6] and obj.get_fbegin() == [7] and obj.get_sbegin() == [1]
    prop = libmlpcpp.PotentialPropertiesFast(pd, G["bin_coeffs"].tolist())
    prop.eval(ax.tolist(), pc.tolist(), ty.tolist(), True)
    assert abs(prop.get_e() - G["bin_e"][0]) < 1e-10 * abs(G["bin_e"][0])
    assert np.abs(np.array(prop.get_f()) - G["bin_f"]).max() < 1e-10 * np.abs(G["bin_f"]).max()
    prop.eval_multiple([ax.tolist()] * 2, [pc.tolist()] * 2, [ty.tolist()] * 2)
    assert np.abs(np.array(prop.get_s_array())[1] - G["bin_s"]).max() < 1e-10 * np.abs(G["bin_s"]).max()
    with pytest.raises(ValueError):
        libmlpcpp.PotentialPropertiesFast(pd, [0.0, 1.0])
    # the producers of the reference hand over lists of NumPy arrays (mlp_dev/core/features.py:13-44): same result
    # through the buffer protocol, and a shape error instead of a crash for a malformed structure
    obj2 = libmlpcpp.PotentialModel(pd, [ax], [pc], [ty], [1], [True], [9])
    assert np.array_equal(np.asarray(obj2.get_x()), np.asarray(obj.get_x()))
    prop.eval_multiple([ax, ax], [pc, pc[:, ::-1]], [ty, ty[::-1]])
    e2, f2 = np.array(prop.get_e_array()), [np.array(f) for f in prop.get_f_array()]
    assert abs(e2[0] - e2[1]) < 1e-12 * abs(e2[0]) and np.abs(f2[0] - f2[1][::-1]).max() < 1e-10
    with pytest.raises(ValueError):
        libmlpcpp.PotentialModel(pd, [ax], [pc[:, :5]], [ty], [1], [True], [9])
    # hybrid drop-in class against the Python mirror
    from pypolymlp_b200.libmlpcpp import PotentialHybridModel

    pd1 = dict(pd, type_full=True, type_indices=[0, 1])
    pd2 = dict(make_params_dict(**cases.pair_model_kwargs(1)), type_full=False, type_indices=[1])
    args = ([pd1, pd2], [ax, ax], [pc, pc], [ty, ty], [1, 1], [True, False], [9, 9])
    hm, hm_py = libmlpcpp.PotentialHybridModel(*args), PotentialHybridModel(*args)
    assert np.array_equal(np.asarray(hm.get_x()), hm_py.get_x())
    assert hm.get_cumulative_n_features() == hm_py.get_cumulative_n_features() and hm.get_n_data() == hm_py.get_n_data()
    assert hm.get_fbegin() == hm_py.get_fbegin() and hm.get_sbegin() == hm_py.get_sbegin()


def test_device_ridge_solve_matches_host():
    """pm_fit_solve_ridge (cuSOLVER potrf/potrs on the resident accumulator) against the host path that follows
    the reference (scipy posv): scales, coefficients, RMSE from X^T X (solvers.py:48-84, utils_model_selection.py:37-72)."""
    pd = make_params_dict(**cases.si_model_kwargs())
    train_ids, test_ids = cases.split_ids_train_test(200, 0.9)
    train, test = _si_datasets(train_ids[:60]), _si_datasets(test_ids)
    alphas = [10.0 ** a for a in np.linspace(-3, 1, 5)]
    host = fit.fit(pd, [train], [test], alphas)
    dev = fit.fit_device(pd, [train], [test], alphas)
    # both sides take X^T X / X^T y / the energy-row sums from the same deterministic device accumulation (no floating-point
    # atomics since round 2: k_syrk_sk2, k_xe_reduce), so the scales are identical; the solvers differ (scipy posv on the
    # host copy vs cuSOLVER potrf/potrs on the resident accumulator)
    assert np.abs(dev["scales"] - host["scales"]).max() <= 1e-12 * np.abs(host["scales"]).max()
    assert dev["alpha"] == host["alpha"]
    # c^T A c - 2 c^T b + y^T y cancels ~9 digits and the alpha = 1e-3 system is ill conditioned (measured 1.1e-4 relative)
    assert np.abs(dev["rmse_train_array"] - host["rmse_train_array"]).max() < 5e-4 * host["rmse_train_array"].max()
    acc = PotentialXtX(pd)
    fit.accumulate_datasets(acc, [train], fit.get_min_energy([train]))
    _, coefs_dev, rmse_dev = acc.solve_ridge(alphas, len(train.energies), scales=host["scales"])
    assert np.abs(rmse_dev - host["rmse_train_array"]).max() < 5e-4 * host["rmse_train_array"].max()
    dev = dict(dev, coefs_array=coefs_dev, scales=host["scales"])
    # ill-conditioned at the smallest alpha: compare predictions, not raw coefficients
    x = PotentialModel(pd, test.axis, test.positions_c, test.types, [20], [True], [64] * 20).get_x()
    e_t = np.asarray(test.energies)
    f_t = np.asarray(test.forces).reshape(-1)
    for k in range(len(alphas)):
        ph = x @ (host["coefs_array"][:, k] / host["scales"])
        pdv = x @ (dev["coefs_array"][:, k] / dev["scales"])
        # north-star gate: fitted-model energy / force RMSE within 1e-6 eV/atom and 1e-5 eV/A of the reference fit
        rmse_e = [np.sqrt(np.mean(np.square((p[:20] - e_t) / 64))) for p in (ph, pdv)]
        rmse_f = [np.sqrt(np.mean(np.square(p[140:] - f_t))) for p in (ph, pdv)]
        assert abs(rmse_e[0] - rmse_e[1]) < 1e-6
        assert abs(rmse_f[0] - rmse_f[1]) < 1e-5
        # ... and the predictions themselves at the same gates for every alpha (measured: 1e-8 eV/atom, 4.6e-6 eV/A at
        # alpha = 1e-3, below 1e-6 eV/A from alpha = 1e-2 on)
        assert np.sqrt(np.mean(np.square(ph[:20] - pdv[:20]))) / 64 < 1e-7
        assert np.sqrt(np.mean(np.square(ph[140:] - pdv[140:]))) < 1e-5


# ---- feature_type = "pair" (compute/local_pair.cpp) and the reference's published MgO answers ---------------
@pytest.mark.parametrize("flags", FLAVOURS)
def test_pair_features_x_xtx_eval(flags):
    pd = make_params_dict(**cases.pair_model_kwargs(2))
    ax, pc, ty = cases.skewed_cell(2)
    pm = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [9], flags=flags)
    full = np.vstack([G["pair_xe"][None], G["pair_xs"], G["pair_xf"]])
    assert pm.get_x().shape == full.shape
    assert cases.x_rel_err(pm.get_x(), full) < 1e-10
    # energy-only structure next to a force structure, then the fused accumulation against numpy on the oracle's X
    tab = po.Tables(pd)
    sts = [cases.skewed_cell(2, n_atom=6, seed=s) for s in (4, 5, 6)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    force = [True, False, True]
    X = po.build_x(tab, axis, pcs, tys, force)
    rng = np.random.default_rng(2)
    w, t = rng.uniform(0.1, 1.0, X.shape[0]), rng.normal(size=X.shape[0])
    xtx, xty, ysq, xe_sum, xe_sq = po.accumulate(X, 3, w, t)
    acc = PotentialXtX(pd, flags=flags)
    acc.add(axis, pcs, tys, force, w, w * t)
    res = acc.finalize()
    assert np.abs(res["xtx"] - xtx).max() < 1e-10 * np.abs(xtx).max()
    assert np.abs(res["xty"] - xty).max() < 1e-10 * np.abs(xty).max()
    assert np.abs(res["xe_sum"] - xe_sum).max() < 1e-10 * np.abs(xe_sum).max()
    prop = PotentialPropertiesFast(pd, G["pair_coeffs"], flags=flags)
    prop.eval(ax, pc, ty)
    assert abs(prop.get_e() - G["pair_e"][0]) < 1e-10 * abs(G["pair_e"][0])
    assert np.abs(prop.get_f() - G["pair_f"]).max() < 1e-10 * np.abs(G["pair_f"]).max()
    assert np.abs(prop.get_s() - G["pair_s"]).max() < 1e-10 * np.abs(G["pair_s"]).max()


@pytest.mark.parametrize("kw", [cases.pair_model_kwargs(1), cases.pair_model_kwargs(3, max_p=3),
                                cases.pair_model_kwargs(2, model_type=1, max_p=1)])
def test_pair_features_other_shapes_vs_oracle(kw):
    pd = make_params_dict(**kw)
    tab = po.Tables(pd)
    ax, pc, ty = cases.skewed_cell(kw["n_type"], n_atom=7, seed=3)
    X = po.build_x(tab, [ax], [pc], [ty], [True])
    pm = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [7])
    assert cases.x_rel_err(pm.get_x(), X) < 1e-10


@pytest.mark.parametrize("kind", ["pair", "gtinv"])
def test_mgo_published_answers_gpu(kind):
    """The reference's own known answers for its bundled MgO potentials: feature sums
    (tests/test_calc/test_compute_features.py:45-58) and E / F / stress of the displaced rocksalt cell
    (tests/test_calc/test_properties_MgO.py:9-97: E rel 1e-8, F atol 1e-6 eV/A, stress atol 1e-5 GPa)."""
    from test_oracle_golden import MGO_FEATURES, check_mgo_eval

    M = cases.load_mgo()
    pd = make_params_dict(**cases.mgo_model_kwargs(kind))
    pm = PotentialModel(pd, [M["st1_axis"], M["st2_axis"]], [M["st1_pos"], M["st2_pos"]],
                        [M["st1_types"], M["st2_types"]], [2], [False], [64, 64])
    x = pm.get_x()
    ncol, total, diff = MGO_FEATURES[kind]
    assert x.shape == (2, ncol)
    assert np.sum(x) == pytest.approx(total, rel=1e-6)
    assert np.sum(x[0] - x[1]) == pytest.approx(diff, rel=1e-6)
    pmf = PotentialModel(pd, [M["st1_axis"], M["st2_axis"]], [M["st1_pos"], M["st2_pos"]],
                         [M["st1_types"], M["st2_types"]], [2], [True], [64, 64])
    assert pmf.get_x().shape == (398, ncol)  # test_compute_features.py:68-72
    assert cases.x_rel_err(pmf.get_x()[:2], x) < 1e-12
    prop = PotentialPropertiesFast(pd, M[kind + "_coeffs"])
    vol = np.linalg.det(M["rs_axis"])
    prop.eval(M["rs_axis"], M["rs_pos"], M["rs_types"])
    check_mgo_eval(kind, prop.get_e(), prop.get_f(), prop.get_s(), vol)
    prop.eval_multiple([M["rs_axis"]] * 2, [M["rs_pos"]] * 2, [M["rs_types"]] * 2)  # test_properties_MgO.py:89-97
    for k in range(2):
        check_mgo_eval(kind, prop.get_e_array()[k], prop.get_f_array()[k], prop.get_s_array()[k], vol)
    # and against the oracle at full precision
    e, f, s = po.eval_structure(po.Tables(pd), M[kind + "_coeffs"], M["rs_axis"], M["rs_pos"], M["rs_types"])
    assert abs(prop.get_e() - e) < 1e-10 * abs(e)
    assert np.abs(prop.get_f() - f).max() < 1e-10 * np.abs(f).max()
    assert np.abs(prop.get_s() - s).max() < 1e-10 * np.abs(s).max()


def test_hybrid_model_vs_oracle():
    """PotentialHybridModel (compute/py_hybrid_model.cpp): a full binary gtinv model next to a pair model that only
    sees element 1, on a mixed force / energy-only batch (one structure has a single atom of element 1)."""
    from pypolymlp_b200.libmlpcpp import PotentialHybridModel

    pd1 = dict(make_params_dict(**cases.binary_model_kwargs()), type_full=True, type_indices=[0, 1])
    pd2 = dict(make_params_dict(**cases.pair_model_kwargs(1)), type_full=False, type_indices=[1])
    sts = [cases.skewed_cell(2, n_atom=n, seed=s) for n, s in ((7, 1), (5, 2), (6, 3))]
    sts[1][2][:] = 0
    sts[1][2][3] = 1
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    hm = PotentialHybridModel([pd1, pd2], axis, pcs, tys, [2, 1], [True, False], [7, 5, 6])
    tabs = [po.Tables(pd1), po.Tables(pd2)]
    X = po.build_x_hybrid(tabs, [True, False], [[0, 1], [1]], axis, pcs, tys, [True, True, False])
    assert hm.get_x().shape == X.shape
    assert hm.get_cumulative_n_features() == [tabs[0].n_variables, tabs[0].n_variables + tabs[1].n_variables]
    assert hm.get_n_data() == [3, 36, 12] and hm.get_sbegin() == [3, -1] and hm.get_fbegin() == [15, -1]
    assert cases.x_rel_err(hm.get_x(), X) < 1e-10
    # a sub-model alone equals the plain model on the same atoms
    pm = PotentialModel(pd1, axis, pcs, tys, [2, 1], [True, False], [7, 5, 6])
    assert np.array_equal(hm.get_x()[:, : tabs[0].n_variables], pm.get_x())


def test_neighbor_test_hooks_reference_known_answers():
    """The reference's own neighbour-list tests (tests/test_cxx/test_neighbor.py:14-186, POSCAR-rocksalt, cutoff 6)
    replayed on the pybind drop-in's Neighbor / NeighborFull / NeighborHalf hooks, which run the device kernels (K1)."""
    import importlib.util

    from pypolymlp_b200.build import pybind_module_path

    spec = importlib.util.spec_from_file_location("libmlpcpp", pybind_module_path())
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    axis = np.eye(3) * 4.0
    frac = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [0, 0, .5], [0, .5, 0], [.5, 0, 0], [.5, .5, .5]]).T
    pc = (axis @ frac).tolist()
    types = [0, 0, 0, 0, 1, 1, 1, 1]
    full = m.NeighborFull(axis.tolist(), pc, 6.0)
    plain = m.Neighbor(axis.tolist(), pc, types, 2, 6.0)
    for dist, diff, nbr in ((plain.get_distances(), plain.get_differences(), plain.get_neighbor_indices()),
                            (full.get_distances(2, types), full.get_differences(2, types),
                             full.get_neighbor_indices(2, types))):
        assert len(dist) == 8 and len(dist[0][0]) == 54 and len(dist[0][1]) == 38
        d00, d01 = np.array(dist[0][0]), np.array(dist[0][1])
        assert np.isclose(d00, 2.8284271247461903).sum() == 12 and np.isclose(d00, 4.0).sum() == 6
        assert np.isclose(d00, 4.898979485566356).sum() == 24 and np.isclose(d00, 5.656854249492381).sum() == 12
        assert np.isclose(d01, 2.0).sum() == 6 and np.isclose(d01, 3.46410162).sum() == 8
        assert np.isclose(d01, 4.47213595).sum() == 24
        assert np.sum(np.square(diff[0][0])) == pytest.approx(1152) and np.sum(np.square(diff[0][1])) == pytest.approx(600)
        assert np.sum(np.square(diff[4][0])) == pytest.approx(600) and np.sum(np.square(diff[4][1])) == pytest.approx(1152)
        sums = [(np.sum(nbr[i][0]), np.sum(nbr[i][1])) for i in range(8)]
        assert sums == [(72, 206), (78, 208), (84, 210), (90, 212), (54, 288), (56, 294), (58, 300), (60, 306)]
    half = m.NeighborHalf(axis.tolist(), pc, 6.0, True)
    diffs, nb = half.get_differences(), half.get_neighbor_indices()
    assert [len(x) for x in nb] == [9, 21, 33, 45, 47, 59, 71, 83]
    assert [np.array(x).shape for x in diffs] == [(n, 3) for n in (9, 21, 33, 45, 47, 59, 71, 83)]
    assert diffs[2][5] == pytest.approx([-2.0, 4.0, 2.0])
    assert [int(np.sum(x)) for x in nb] == [0, 9, 30, 63, 90, 149, 220, 303]


def test_get_fn_get_ylm_hooks_reference_known_answers():
    """The reference's tests of the radial functions and spherical harmonics (tests/test_cxx/test_functions.py:14-50)
    replayed on the device pair-basis kernel (K2a) through the get_fn / get_ylm hooks."""
    from pypolymlp_b200.libmlpcpp import get_fn, get_ylm

    fn, fn_d = get_fn(1.2, [[1.0, 0.0], [1.0, 1.0], [1.0, 2.0]], 6.0)
    np.testing.assert_allclose(fn, [0.21430317094756238, 0.8690422117212636, 0.4769404780495180], rtol=1e-12)
    np.testing.assert_allclose(fn_d, [-0.5507864848009192, -0.495464911460655, 0.6819640474131319], rtol=1e-12)
    x, y, z = 0.173723561607389, 0.446843340790007, 0.877582561890373
    ylm, ylm_dx, ylm_dy, ylm_dz = get_ylm(x, y, z, 10)
    assert len(ylm) == len(ylm_dx) == len(ylm_dy) == len(ylm_dz) == 66
    assert ylm.sum() == pytest.approx(-3.094632553138235 + 0.09510814961092404j, rel=1e-6)
    assert ylm_dx.sum() == pytest.approx(-6.916830463136405 - 1.910490992920798j, rel=1e-6)
    assert ylm_dy.sum() == pytest.approx(12.686387327323297 + 23.645024627283863j, rel=1e-6)
    assert ylm_dz.sum() == pytest.approx(-5.090360117438893 - 11.661266919165213j, rel=1e-6)
    # against the oracle's restatement of polymlp_spherical_harmonics.cpp, element by element
    o = po.ylm_der(np.array([x]), np.array([y]), np.array([z]), 10)
    for a, b_ in zip((ylm, ylm_dx, ylm_dy, ylm_dz), o):
        assert np.abs(a - b_[:, 0]).max() < 1e-12 * max(1.0, np.abs(b_).max())

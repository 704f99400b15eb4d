"""Round-2 GPU tests (python -m pytest tests -m gpu): second-generation front end (k_lrows_v4 / k_xrows_v6) against
the first generation and the reference goldens, run-to-run bitwise determinism of the fused products, the
deterministic stream-K SYRK against the old RED.F64 one, zero-atom chunks."""

import os

import numpy as np
import pytest

import cases
from pypolymlp_b200._capi import PM_FLAG_SIMPLE_KERNELS
from pypolymlp_b200.libmlpcpp import PotentialModel, PotentialXtX
from pypolymlp_b200.params import make_params_dict

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(cases.GOLDEN, "ref_vectors.npz"))


def _fcc_batch(n, seed0=20240, rep=(4, 4, 4)):
    sts = [cases.fcc_supercell(rep=rep, seed=seed0 + s) for s in range(n)]
    return [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]


def _rows(n_st, n_atom):
    return n_st * (1 + 6 + 3 * n_atom)


@pytest.mark.parametrize("model_type", [4, 3])
def test_front_v2_matches_v1_and_golden(model_type, monkeypatch):
    """Config-2 model: X from the second-generation K4a/K4b == X from the first generation (1e-12 of the column
    scale: same arithmetic, different summation order) and == the reference golden (model_type 4)."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(model_type))
    axis, pcs, tys = _fcc_batch(3)
    x2 = PotentialModel(pd, axis, pcs, tys, [3], [True], [256] * 3).get_x()
    monkeypatch.setenv("PM_FRONT_V1", "1")
    x1 = PotentialModel(pd, axis, pcs, tys, [3], [True], [256] * 3).get_x()
    monkeypatch.delenv("PM_FRONT_V1")
    assert x1.shape == x2.shape
    assert cases.x_rel_err(x2, x1) < 1e-12
    if model_type == 4:
        x = PotentialModel(pd, axis[:1], pcs[:1], tys[:1], [1], [True], [256]).get_x()
        assert cases.x_rel_err(x[0], G["fcc_xe"]) < 1e-10
        assert cases.x_rel_err(x[1:7], G["fcc_xs"]) < 1e-10
        xf = x[7:]
        scale = np.abs(xf).max(axis=0)
        assert (np.abs(xf[G["fcc_rows"]] - G["fcc_xf_rows"]) / np.maximum(scale, 1e-8 * scale.max())).max() < 1e-10


def test_front_v2_ragged_and_mixed_force_flags():
    """Different atom counts per structure, an energy-only structure in the middle, a structure with no neighbours
    inside the cutoff: the persistent K4a must skip / size its work per centre, K4b per row atom."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    sts = [cases.fcc_supercell(rep=(2, 2, 2), seed=1), cases.fcc_supercell(rep=(3, 2, 2), seed=2),
           cases.fcc_supercell(rep=(2, 2, 2), seed=3), (np.eye(3) * 40.0, np.zeros((3, 1)), np.zeros(1, np.int32)),
           cases.fcc_supercell(rep=(3, 3, 2), seed=4)]
    axis, pcs, tys = [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts]
    force = [True, True, False, True, True]
    na = [len(t) for t in tys]
    x2 = PotentialModel(pd, axis, pcs, tys, [1] * 5, force, na).get_x()
    xs = PotentialModel(pd, axis, pcs, tys, [1] * 5, force, na, flags=PM_FLAG_SIMPLE_KERNELS).get_x()
    assert x2.shape == xs.shape
    assert cases.x_rel_err(x2, xs) < 1e-10


def test_fused_products_bitwise_deterministic():
    """Two runs over the same structures give bit-identical X^T X, X^T y, y^T y, xe_sum, xe_sq_sum (fixed-order
    stream-K fix-up, fixed-order energy-row sums): the reference is deterministic, so is the drop-in."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    n = 12
    axis, pcs, tys = _fcc_batch(n)
    rng = np.random.default_rng(11)
    w = rng.uniform(0.2, 1.0, _rows(n, 256))
    y = w * rng.normal(size=w.size)
    res = []
    for ws in (0, 0, 700 << 20):   # default chunking twice, then several chunks
        acc = PotentialXtX(pd, workspace_bytes=ws)
        acc.add(axis, pcs, tys, [True] * n, w, y)
        res.append(acc.finalize())
    for key in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.array_equal(res[0][key], res[1][key]), key
    assert res[0]["y_sq_norm"] == res[1]["y_sq_norm"]
    # different chunking changes the summation order but not the result beyond rounding
    sc = np.abs(res[0]["xtx"]).max()
    assert np.abs(res[2]["xtx"] - res[0]["xtx"]).max() < 1e-12 * sc
    # the same accumulator object, reset and refilled, reproduces itself too
    acc = PotentialXtX(pd)
    acc.add(axis, pcs, tys, [True] * n, w, y)
    r1 = {k: np.array(v, copy=True) if isinstance(v, np.ndarray) else v for k, v in acc.finalize().items()}
    acc.reset()
    acc.add(axis, pcs, tys, [True] * n, w, y)
    r2 = acc.finalize()
    assert np.array_equal(r1["xtx"], r2["xtx"]) and np.array_equal(r1["xty"], r2["xty"])


def test_device_ridge_solve_bitwise_reproducible():
    """Accumulate + cuSOLVER ridge solve twice: identical scales, coefficients and RMSE (the accumulator is deterministic,
    so is everything downstream of it)."""
    from pypolymlp_b200 import fit
    from test_gpu_parity import _si_datasets

    pd = make_params_dict(**cases.si_model_kwargs())
    train_ids, _ = cases.split_ids_train_test(200, 0.9)
    train = _si_datasets(train_ids[:40])
    alphas = [1e-3, 1e-1, 10.0]
    out = []
    for _ in range(2):
        acc = PotentialXtX(pd)
        fit.accumulate_datasets(acc, [train], fit.get_min_energy([train]))
        out.append(acc.solve_ridge(alphas, len(train.energies)))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_syrk_v2_matches_v1(monkeypatch):
    """Triangle-only diagonal tiles + parked partial tiles == the first stream-K kernel (full diagonal tiles, RED.F64)
    within rounding; X^T X stays exactly symmetric after packing."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    n = 5
    axis, pcs, tys = _fcc_batch(n, seed0=777)
    rng = np.random.default_rng(3)
    w = rng.uniform(0.2, 1.0, _rows(n, 256))
    y = w * rng.normal(size=w.size)
    acc = PotentialXtX(pd)
    acc.add(axis, pcs, tys, [True] * n, w, y)
    r2 = acc.finalize()
    monkeypatch.setenv("PM_SYRK_V1", "1")
    acc1 = PotentialXtX(pd)
    acc1.add(axis, pcs, tys, [True] * n, w, y)
    r1 = acc1.finalize()
    monkeypatch.delenv("PM_SYRK_V1")
    sc = np.abs(r1["xtx"]).max()
    assert np.abs(r2["xtx"] - r1["xtx"]).max() < 1e-12 * sc
    assert np.abs(r2["xty"] - r1["xty"]).max() < 1e-12 * np.abs(r1["xty"]).max()
    assert abs(r2["y_sq_norm"] - r1["y_sq_norm"]) < 1e-12 * r1["y_sq_norm"]
    assert np.array_equal(r2["xtx"], r2["xtx"].T)
    assert np.array_equal(r2["xe_sum"], r1["xe_sum"])   # same fixed-order energy-row sums in both


def test_zero_atom_structures_features_x_and_fit():
    """a chunk whose structures hold no atoms (a hybrid sub-model that sees none of its elements) returns
    zero rows from pm_features_x on a fresh context and counts its rows / targets in the fused fit."""
    pd = make_params_dict(**cases.cfg2_model_kwargs(3))
    axis = [np.eye(3) * 8.0, np.eye(3) * 9.0]
    pcs = [np.zeros((3, 0)), np.zeros((3, 0))]
    tys = [np.zeros(0, np.int32), np.zeros(0, np.int32)]
    x = PotentialModel(pd, axis, pcs, tys, [2], [True], [0, 0]).get_x()
    assert x.shape == (2 + 12, 255) and not x.any()
    acc = PotentialXtX(pd)
    w = np.full(14, 0.5)
    y = np.arange(14, dtype=float)
    acc.add(axis, pcs, tys, [True, True], w, y)
    res = acc.finalize()
    assert res["total_n_data"] == 14
    assert res["y_sq_norm"] == pytest.approx(float(y @ y), rel=1e-15)
    assert not res["xtx"].any() and not res["xty"].any()


def test_compiled_dropin_potential_xtx_reference_goldens():
    """The compiled pybind11 module's PotentialXtX (what dropin.install_fused_products() hands the reference's
    calc_xtx_xty) on config 1, the reference's bundled Si set: the reference's own goldens xty[56], scales[56],
    total_n_data = 35820 (tests/test_mlp_dev/test_core_data.py:39-57) and the RMSE goldens of
    tests/test_mlp_dev_api/test_mlp_devel_phono3py.py:22-27; bit-identical to the ctypes mirror (same C ABI underneath).
    The weights / targets follow the reference's apply_weights rules (fit.py restates them; tests/test_dropin_fused_fit.py
    runs the reference's own Python around the same call on the CPU box, where /root/reference exists)."""
    from pypolymlp_b200 import dropin, fit
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast
    from test_gpu_parity import _si_datasets

    ext = dropin.load_extension()
    pd = make_params_dict(**cases.si_model_kwargs())
    train_ids, test_ids = cases.split_ids_train_test(200, 0.9)
    train, test = _si_datasets(train_ids), _si_datasets(test_ids)
    min_e = fit.get_min_energy([train])
    acc = ext.PotentialXtX(pd)
    assert acc.n_devices() == 1
    fit.accumulate_datasets(acc, [train], min_e, batch_size=50)
    res = acc.finalize()
    ref = PotentialXtX(pd)
    fit.accumulate_datasets(ref, [train], min_e, batch_size=50)
    res_ct = ref.finalize()
    for key in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.array_equal(np.asarray(res[key]), res_ct[key]), key
    data_xy = fit.finalize_xtx_xty({k: np.asarray(v) if not np.isscalar(v) else v for k, v in res.items()}, [train],
                                   min_energy=min_e)
    assert data_xy.total_n_data == 35820
    assert data_xy.xty[56] == pytest.approx(6.899130774433e5, rel=1e-6)
    assert data_xy.scales[56] == pytest.approx(0.0032488632685359524)
    assert min_e == pytest.approx(-5.737324395625)
    alphas = [10.0 ** a for a in np.linspace(-1, 1, 3)]
    coefs_array = fit.solver_ridge(data_xy.xtx, data_xy.xty, alphas)
    acc_t = ext.PotentialXtX(pd)
    fit.accumulate_datasets(acc_t, [test], min_e, batch_size=50)
    res_t = acc_t.finalize()
    test_xy = fit.finalize_xtx_xty({k_: np.asarray(v) if not np.isscalar(v) else v for k_, v in res_t.items()}, [test],
                                   scales=data_xy.scales, min_energy=min_e)
    k = int(np.argmin(fit.compute_rmse(coefs_array, test_xy)))   # model selection on the test set, as standard/fit.py:12-63
    prop = PotentialPropertiesFast(pd, coefs_array[:, k] / data_xy.scales)
    for ds, e_gold, f_gold in ((train, 1.925e-6, 9.113e-4), (test, 2.067e-6, 9.180e-4)):
        prop.eval_multiple(ds.axis, ds.positions_c, ds.types)
        e = np.array(prop.get_e_array())
        f = np.concatenate([np.asarray(a).reshape(-1) for a in prop.get_f_array()])
        assert np.sqrt(np.mean(np.square((e - ds.energies) / 64))) == pytest.approx(e_gold, rel=1e-2)
        assert np.sqrt(np.mean(np.square(f - ds.forces))) == pytest.approx(f_gold, rel=1e-2)
    # get_x() of the compiled PotentialModel is Fortran-ordered like the reference's Eigen view (pybind11_mlp.cpp:20-21)
    x = ext.PotentialModel(pd, train.axis[:2], train.positions_c[:2], train.types[:2], [2], [True], [64, 64]).get_x()
    assert x.flags["F_CONTIGUOUS"] and x.shape == (2 + 12 + 384, 168)


G2 = np.load(os.path.join(cases.GOLDEN, "ref_vectors_r02.npz"))


@pytest.mark.parametrize("mrt", ["", "8", "9", "12", "16"])
def test_x_config4_ternary_order4_vs_reference(mrt, monkeypatch):
    """BASELINE config 4 model (ternary, F = 45090, three neighbour-type segments per centre, six type pairs) through the
    large-model kernels -- every k_lrows_big row-tile variant -- against the reference (oracle/_ref) goldens: the
    energy row column by column, every stress / force row through 16 seeded projections of the column-scaled row."""
    if mrt:
        monkeypatch.setenv("PM_LROWS_MRT", mrt)
    pd = make_params_dict(**cases.cfg4_model_kwargs())
    ax, pc, ty = cases.cfg4_small_cell()
    assert np.array_equal(ty, G2["cfg4_types"])
    x = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [16]).get_x()
    assert x.shape == (1 + 6 + 48, 45090)
    assert cases.x_rel_err(x[0], G2["cfg4_xe"]) < 1e-10
    proj = (x / G2["cfg4_colscale"]) @ cases.projection_matrix(45090)
    # entries of the scaled rows are <= 1 and agree to ~1e-13; a 1e-9 error in any single column moves a projection by ~1e-9
    assert np.abs(proj - G2["cfg4_proj"]).max() < 2e-10
    assert np.abs(x.sum(axis=1) - G2["cfg4_rowsum"]).max() < 1e-10 * np.abs(G2["cfg4_rowsum"]).max()


@pytest.mark.parametrize("env", [{"PM_FEAT_V3": "1"}, {"PM_FEAT_NB": "2"}, {"PM_FEAT_NB": "1"}, {"PM_FEAT_NO_RADIAL": "1"}])
def test_x_config4_radial_batched_k3_variants(env, monkeypatch):
    """Large-model K3: the radial-batched kernel (k_features_v4r, default for big a_nlm arrays; the test above pins it to the
    reference goldens) against the slice kernel k_features_v3 and against its own other radial block sizes."""
    pd = make_params_dict(**cases.cfg4_model_kwargs())
    ax, pc, ty = cases.cfg4_small_cell()
    x0 = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [16]).get_x()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    x1 = PotentialModel(pd, [ax], [pc], [ty], [1], [True], [16]).get_x()   # (a new model / context: tables are built per context)
    scale = np.abs(x0).max(axis=0) + 1e-300
    assert np.abs((x1 - x0) / scale).max() < 1e-12


def test_eval_config4_model_vs_reference():
    pd = make_params_dict(**cases.cfg4_model_kwargs())
    ax, pc, ty = cases.cfg4_small_cell()
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast

    coeffs = np.random.default_rng(14).normal(size=45090) * 1e-3
    prop = PotentialPropertiesFast(pd, coeffs)
    prop.eval(ax, pc, ty, True)
    assert abs(prop.get_e() - G2["cfg4_e"][0]) < 1e-10 * abs(G2["cfg4_e"][0])
    assert np.abs(np.asarray(prop.get_f()) - G2["cfg4_f"]).max() < 1e-10 * np.abs(G2["cfg4_f"]).max()
    assert np.abs(np.asarray(prop.get_s()) - G2["cfg4_s"]).max() < 1e-10 * np.abs(G2["cfg4_s"]).max()


def test_eval_config5_shape_vs_reference():
    """BASELINE config 5 shape: the F = 2030 model on an elongated fcc 4x4x8 cell (512 atoms, sigma 0.03 A; a different
    translation set from the cubic cells) -- neighbour list checksums and E / F / S against RefEval, alone and inside a
    batch (eval_multiple) with other cells around it."""
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast, _Context, _Model

    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    ax, pc, ty = cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777)
    off, nb, dx, dy, dz = _Context(_Model(pd)).neighbor_full(ax, pc, ty)
    assert off[-1] == G2["cfg5_nbr_count"][0]
    assert float(nb.sum()) == G2["cfg5_nbr_checksum"][0]
    assert np.square(dx).sum() + np.square(dy).sum() + np.square(dz).sum() == pytest.approx(G2["cfg5_nbr_checksum"][1], rel=1e-14)
    coeffs = np.random.default_rng(12).normal(size=2030) * 1e-3
    prop = PotentialPropertiesFast(pd, coeffs)
    prop.eval(ax, pc, ty, True)
    for e, f, s_ in ((prop.get_e(), prop.get_f(), prop.get_s()),):
        assert abs(e - G2["cfg5_e"][0]) < 1e-10 * abs(G2["cfg5_e"][0])
        assert np.abs(np.asarray(f) - G2["cfg5_f"]).max() < 1e-10 * np.abs(G2["cfg5_f"]).max()
        assert np.abs(np.asarray(s_) - G2["cfg5_s"]).max() < 1e-10 * np.abs(G2["cfg5_s"]).max()
    other = cases.fcc_supercell(rep=(4, 4, 4), seed=5)
    prop.eval_multiple([other[0], ax, other[0]], [other[1], pc, other[1]], [other[2], ty, other[2]])
    assert abs(prop.get_e_array()[1] - G2["cfg5_e"][0]) < 1e-10 * abs(G2["cfg5_e"][0])
    assert np.abs(np.asarray(prop.get_f_array()[1]) - G2["cfg5_f"]).max() < 1e-10 * np.abs(G2["cfg5_f"]).max()
    assert np.abs(np.asarray(prop.get_s_array()[1]) - G2["cfg5_s"]).max() < 1e-10 * np.abs(G2["cfg5_s"]).max()
    assert prop.get_e_array()[0] == pytest.approx(prop.get_e_array()[2], rel=1e-13)   # atomics: order varies


_EVAL_VARIANTS = {
    "default": {},
    "generic_lane_atom": {"PM_EVAL_LB": "0"},
    "slice_kernel": {"PM_EVAL_LA": "0"},
    "stored_pair_records": {"PM_EVAL_STORED_PB": "1"},
    "fit_k2_kernel": {"PM_EVAL_K2_FIT": "1"},
    "one_lane": {"PM_EVAL_LANES": "1"},
    "k2_without_dmma": {"PM_EVAL_K2_NO_DMMA": "1"},
    "radial_items_from_records": {"PM_EVAL_RC_RADS": "1"},
}


@pytest.mark.parametrize("variant", sorted(_EVAL_VARIANTS))
def test_eval_large_batch_kernel_variants(variant, monkeypatch):
    """The large-batch eval path -- two lanes (sibling context + host thread), a_nlm-only K2 kernel, lane = atom feature /
    adjoint kernels (radial-batched and generic), pair pass that recomputes the basis records -- and each of its A/B
    switches against (a) the reference golden of the config-5 cell and (b) the one-structure-at-a-time path, which uses
    the first-generation eval kernels (small batches never reach the lane = atom kernels)."""
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast

    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    coeffs = np.random.default_rng(12).normal(size=2030) * 1e-3
    sts = [cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777 + k) for k in range(34)]   # 17 408 atoms: two lanes
    sts.append(cases.fcc_supercell(rep=(2, 2, 2), sigma=0.02, seed=3))                        # ragged tail
    prop = PotentialPropertiesFast(pd, coeffs)
    singles = []
    for k in (0, 16, 17, 33, 34):    # both sides of the lane split
        prop.eval(*sts[k], True)
        singles.append((k, prop.get_e(), np.array(prop.get_f()), np.array(prop.get_s())))
    monkeypatch.setenv("PM_EVAL_LA_MIN", "1")
    for name, val in _EVAL_VARIANTS[variant].items():
        monkeypatch.setenv(name, val)
    prop.eval_multiple([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts])
    e, f, s_ = prop.get_e_array(), prop.get_f_array(), prop.get_s_array()
    assert abs(e[0] - G2["cfg5_e"][0]) < 1e-10 * abs(G2["cfg5_e"][0])
    assert np.abs(np.asarray(f[0]) - G2["cfg5_f"]).max() < 1e-10 * np.abs(G2["cfg5_f"]).max()
    assert np.abs(np.asarray(s_[0]) - G2["cfg5_s"]).max() < 1e-10 * np.abs(G2["cfg5_s"]).max()
    for k, e1, f1, s1 in singles:
        assert e[k] == pytest.approx(e1, rel=1e-12)
        assert np.abs(np.asarray(f[k]) - f1).max() < 1e-11 * np.abs(f1).max()
        assert np.abs(np.asarray(s_[k]) - s1).max() < 1e-11 * np.abs(s1).max()


def test_eval_lanes_ragged_batches_and_errors():
    """Lane splitting of pm_eval: batches with fewer structures than lanes, one dominant structure, an empty structure in the
    middle of a large batch, and an invalid structure in a later lane (the error surfaces, the context stays usable)."""
    from pypolymlp_b200.libmlpcpp import PotentialPropertiesFast

    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    prop = PotentialPropertiesFast(pd, np.random.default_rng(12).normal(size=2030) * 1e-3)
    big = cases.fcc_supercell(rep=(4, 4, 8), sigma=0.03, seed=777)
    small = cases.fcc_supercell(rep=(2, 2, 2), sigma=0.02, seed=3)
    prop.eval(*big, True)
    e_big, f_big = prop.get_e(), np.array(prop.get_f())
    prop.eval(*small, True)
    e_small, f_small = prop.get_e(), np.array(prop.get_f())
    empty = (small[0], np.zeros((3, 0)), np.zeros(0, dtype=np.int32))
    for sts in ([big] * 40 + [empty] + [small] * 3 + [big] * 2,          # 21 504 atoms, empty structure inside
                [big] * 33 + [small],                                     # 2 lanes' worth, ragged tail
                [small] * 2 + [big] * 36):
        prop.eval_multiple([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts])
        e, f = prop.get_e_array(), prop.get_f_array()
        for k, st in enumerate(sts):
            if st is big:
                assert e[k] == pytest.approx(e_big, rel=1e-12)
                assert np.abs(np.asarray(f[k]) - f_big).max() < 1e-11 * np.abs(f_big).max()
            elif st is small:
                assert e[k] == pytest.approx(e_small, rel=1e-12)
                assert np.abs(np.asarray(f[k]) - f_small).max() < 1e-11 * np.abs(f_small).max()
            else:
                assert e[k] == 0.0 and np.asarray(f[k]).size == 0
    bad = (big[0], big[1], np.full(512, 7, dtype=np.int32))   # atom type outside the model
    sts = [big] * 39 + [bad]
    with pytest.raises((ValueError, RuntimeError)):
        prop.eval_multiple([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts])
    prop.eval(*small, True)
    assert prop.get_e() == pytest.approx(e_small, rel=1e-12)


def _n_devices():
    import ctypes

    from pypolymlp_b200._capi import lib

    n = ctypes.c_int(0)
    lib().pm_device_count(ctypes.byref(n))
    return n.value


needs_two_gpus = pytest.mark.skipif("_n_devices() < 2", reason="needs two GPUs (gpurun --gpus 2)")


@needs_two_gpus
def test_multi_gpu_sharded_xtx_equals_single_gpu():
    """One process, two GPUs (pm_multi_*: one host thread per device, one grouped ncclReduce onto device 0): X^T X, X^T y,
    xe sums and the row count of 64 structures sharded 2-ways == the same 64 structures on one GPU, 1e-10."""
    from pypolymlp_b200.libmlpcpp import PotentialXtXMulti

    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    n = 64
    axis, pcs, tys = _fcc_batch(n, seed0=4100)
    rng = np.random.default_rng(21)
    w = rng.uniform(0.2, 1.0, _rows(n, 256))
    y = w * rng.normal(size=w.size)
    one = PotentialXtX(pd, device=0)
    one.add(axis, pcs, tys, [True] * n, w, y)
    r1 = one.finalize()
    two = PotentialXtXMulti(pd, devices=[0, 1])
    two.add(axis, pcs, tys, [True] * n, w, y)
    r2 = two.finalize()
    assert r2["total_n_data"] == r1["total_n_data"] == _rows(n, 256)
    for key in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.abs(r2[key] - r1[key]).max() < 1e-10 * np.abs(r1[key]).max(), key
    assert abs(r2["y_sq_norm"] - r1["y_sq_norm"]) < 1e-10 * r1["y_sq_norm"]
    # second round on the same object: the non-root accumulators were cleared by finalize()
    two.reset()
    two.add(axis[:10], pcs[:10], tys[:10], [True] * 10, *_sub_rows(w, y, n, 10))
    r3 = two.finalize()
    one.reset()
    one.add(axis[:10], pcs[:10], tys[:10], [True] * 10, *_sub_rows(w, y, n, 10))
    r4 = one.finalize()
    assert np.abs(r3["xtx"] - r4["xtx"]).max() < 1e-10 * np.abs(r4["xtx"]).max()
    # the compiled drop-in module shards the same way
    from pypolymlp_b200 import dropin

    mod = dropin.load_extension()
    acc = mod.PotentialXtX(pd, devices=[0, 1])
    assert acc.n_devices() == 2
    acc.add(axis, pcs, tys, [True] * n, w, y)
    r5 = acc.finalize()
    assert np.abs(r5["xtx"] - r1["xtx"]).max() < 1e-10 * np.abs(r1["xtx"]).max()
    assert r5["total_n_data"] == r1["total_n_data"]


def _sub_rows(w, y, n_st, k, n_atom=256):
    """rows of the first k structures of an n_st-structure batch, regrouped into a k-structure PyModel layout"""
    e = np.arange(k)
    s = n_st + np.arange(6 * k)
    f = n_st + 6 * n_st + np.arange(3 * n_atom * k)
    idx = np.concatenate([e, s, f])
    return w[idx], y[idx]


_RANK_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases
from pypolymlp_b200 import fit
from pypolymlp_b200.libmlpcpp import PotentialXtX
from pypolymlp_b200.params import make_params_dict
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
acc = PotentialXtX(pd, device=rank)
assert fit.comm_init_from_env(acc) == (rank, world)
n = 16
lo, hi = fit.shard_range(n, rank, world)
sts = [cases.fcc_supercell(seed=4100 + s) for s in range(lo, hi)]
rng = np.random.default_rng(100 + rank)
rows = (hi - lo) * 775
w = rng.uniform(0.2, 1.0, rows); y = w * rng.normal(size=rows)
acc.add([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts], [True] * (hi - lo), w, y)
np.save({out!r} + ".w%d.npy" % rank, np.stack([w, y]))
acc.reduce(0)
assert "torch" not in sys.modules
if rank == 0:
    r = acc.finalize()
    np.savez({out!r} + ".npz", **{{k: v for k, v in r.items()}})
acc.barrier()
"""


@needs_two_gpus
def test_two_processes_nccl_reduce_inside_library(tmp_path):
    """One process per GPU (the bench / torchrun layout): file rendezvous of the NCCL id, pm_comm_init_rank, pm_fit_reduce.
    The reduced X^T X of 2 x 8 structures equals a single-GPU accumulation of the same 16; torch is never imported."""
    import subprocess
    import sys

    out = str(tmp_path / "res")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=cases.HERE + "/..", out=out))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_PORT="29517")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    red = np.load(out + ".npz")
    pd = make_params_dict(**cases.cfg2_model_kwargs(4))
    one = PotentialXtX(pd, device=0)
    for rank in range(2):
        w, y = np.load(out + ".w%d.npy" % rank)
        sts = [cases.fcc_supercell(seed=4100 + s) for s in range(8 * rank, 8 * rank + 8)]
        one.add([s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts], [True] * 8, w, y)
    r1 = one.finalize()
    assert int(red["total_n_data"]) == r1["total_n_data"] == 16 * 775
    for key in ("xtx", "xty", "xe_sum", "xe_sq_sum"):
        assert np.abs(red[key] - r1[key]).max() < 1e-10 * np.abs(r1[key]).max(), key

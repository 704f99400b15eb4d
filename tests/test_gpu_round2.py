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

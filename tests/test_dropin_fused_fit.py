"""`dropin.install_fused_products()`: the reference's `calc_xtx_xty` (mlp_dev/core/data_sequential.py:25-156) switched to
the fused device product `PotentialXtX`, with the reference's own batching, `apply_weights`, `compute_scales` and
scaling tail around it.  CPU check of that host glue: the compiled module's `PotentialModel` and `PotentialXtX` are
replaced IN THE TEST by stand-ins that take X from oracle/_ref, so the reference's original flow ("X -> apply_weights ->
x.T @ x" per batch) and the fused flow (weights and targets without X, one accumulator) see the same X and must give the
same X^T X, X^T y, scales, y^T y and row count.  Inputs: the reference's own parameter file for Si
(tests/files/polymlp.in.phono3py.Si) and structures of its bundled dataset, as reference `Dataset` objects with mixed
force / energy-only sets, dataset weights and a batch size that splits them.  Skipped without the reference checkout."""

import copy
import os

import numpy as np
import pytest

import cases
from oracle import ref
from pypolymlp_b200 import dropin
from pypolymlp_b200.libmlpcpp import _set_index

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not (os.path.isdir(REF_SRC + "/pypolymlp") and ref.available()),
                                reason="reference checkout / oracle/_ref not present")


class OracleModel:
    """Stand-in for libmlpcpp.PotentialModel (pybind11_mlp.cpp:12-27) on oracle/_ref."""

    def __init__(self, pd, axis, pos, types, n_st_dataset, force_dataset, n_atoms_all):
        force_st = [bool(f) for n, f in zip(n_st_dataset, force_dataset) for _ in range(n)]
        self._x = ref.RefModel(pd).build_x(axis, pos, [np.asarray(t, np.int32) for t in types], force_st)
        self._fbegin, self._sbegin, self._n_data = _set_index(n_st_dataset, force_dataset, n_atoms_all)

    def get_x(self):
        return self._x

    def get_fbegin(self):
        return self._fbegin

    def get_sbegin(self):
        return self._sbegin

    def get_n_data(self):
        return self._n_data


class OracleXtX:
    """Stand-in for libmlpcpp.PotentialXtX: the sums the device accumulates, from the oracle's X."""

    def __init__(self, pd, devices=()):
        self.rm = ref.RefModel(pd)
        n = self.rm.n_features
        self.xtx, self.xty, self.xe, self.xe2 = np.zeros((n, n)), np.zeros(n), np.zeros(n), np.zeros(n)
        self.ysq, self.rows = 0.0, 0

    def add(self, axis, pos, types, force_st, w, y):
        x = self.rm.build_x(axis, pos, [np.asarray(t, np.int32) for t in types], [bool(f) for f in force_st])
        assert len(w) == len(y) == x.shape[0]
        ne = len(axis)
        self.xe += x[:ne].sum(0)
        self.xe2 += np.square(x[:ne]).sum(0)
        xw = x * np.asarray(w)[:, None]
        self.xtx += xw.T @ xw
        self.xty += xw.T @ np.asarray(y)
        self.ysq += float(np.dot(y, y))
        self.rows += x.shape[0]

    def finalize(self):
        return {"xtx": self.xtx, "xty": self.xty, "xe_sum": self.xe, "xe_sq_sum": self.xe2, "y_sq_norm": self.ysq,
                "total_n_data": self.rows}


def test_fused_calc_xtx_xty_equals_reference_flow(monkeypatch):
    fused = dropin.install_fused_products(REF_SRC)
    ext = dropin.install(REF_SRC)
    from pypolymlp.core.data_format import PolymlpStructure
    from pypolymlp.core.dataset import Dataset, DatasetList
    from pypolymlp.core.dataset_utils import DatasetDFT
    from pypolymlp.core.parser_polymlp_params import ParamsParser
    from pypolymlp.mlp_dev.core import data_sequential as ds

    assert ds.calc_xtx_xty is fused and fused._b200_fused
    assert dropin.install_fused_products(REF_SRC) is fused  # idempotent
    params = ParamsParser("/root/reference/tests/files/polymlp.in.phono3py.Si", parse_dft=False).params
    axis, positions_c, forces, energies = cases.load_si_dataset()
    inv = np.linalg.inv(axis)

    def dataset(ids, include_force, weight, name):
        sts = [PolymlpStructure(axis=axis, positions=inv @ positions_c[i], n_atoms=[64], elements=["Si"] * 64,
                                types=np.zeros(64, int), volume=np.linalg.det(axis)) for i in ids]
        dft = DatasetDFT(sts, energies[ids], forces=[forces[i].T for i in ids], elements=["Si"])
        return Dataset(dataset_type="vasp", files=name, include_force=include_force, include_stress=False,
                       weight=weight, name=name, dft=dft)

    sets = DatasetList([dataset([0, 3, 5, 7, 11], True, 1.0, "forces"), dataset([2, 4, 6], False, 0.5, "energies")])
    monkeypatch.setattr(ext, "PotentialModel", OracleModel, raising=False)
    monkeypatch.setattr(ext, "PotentialXtX", OracleXtX, raising=False)
    for batch_size in (2, 64):
        a = fused(params, copy.deepcopy(sets), batch_size=batch_size, scale_threshold=1e-10)
        b = fused._b200_original(params, copy.deepcopy(sets), batch_size=batch_size, scale_threshold=1e-10)
        assert a.total_n_data == b.total_n_data == 5 * (1 + 6 + 192) + 3
        assert a.min_energy == b.min_energy
        assert a.y_sq_norm == pytest.approx(b.y_sq_norm, rel=1e-14)
        for key in ("xtx", "xty", "scales", "xe_sum", "xe_sq_sum"):
            u, v = getattr(a, key), getattr(b, key)
            assert np.abs(u - v).max() <= 1e-13 * np.abs(v).max(), key
    # a given scale vector (the test-set pass of the reference's fit) is used as is
    c = fused(params, copy.deepcopy(sets), scales=a.scales, min_energy=a.min_energy, batch_size=3)
    d = fused._b200_original(params, copy.deepcopy(sets), scales=a.scales, min_energy=a.min_energy, batch_size=3)
    assert np.array_equal(c.scales, a.scales)
    assert np.abs(c.xtx - d.xtx).max() <= 1e-13 * np.abs(d.xtx).max()
    assert np.abs(c.xty - d.xty).max() <= 1e-13 * np.abs(d.xty).max()

"""The UNMODIFIED reference Python package (src/pypolymlp of /root/reference) running on top of the compiled drop-in
`pypolymlp_b200/lib/libmlpcpp.*.so`, registered under the reference's module name `pypolymlp.cxx.lib.libmlpcpp`
(pypolymlp_b200/dropin.py).  CPU part: everything the reference does on the extension before a device is
needed — Readgtinv at parameter construction (cxx/wrapper/api_gtinv_list.py:6-12), FeaturesAttr for the feature count
and attributes (mlp_dev/core/features_attr.py:11-49), the potential loaders — with the reference's own known answers.
Skipped when the reference checkout is not present (the GPU box)."""

import os
import sys

import numpy as np
import pytest

import cases
from pypolymlp_b200 import dropin
from pypolymlp_b200.io_legacy import load_mlp as our_load_mlp

REF_SRC = "/root/reference/src"
MLPS = "/root/reference/tests/test_calc/files/mlps/"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC + "/pypolymlp"), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ext():
    return dropin.install(REF_SRC)


def test_module_is_ours_and_exports_the_boundary(ext):
    from pypolymlp.cxx.lib import libmlpcpp  # the reference's import statement (mlp_dev/core/features.py:10)

    assert libmlpcpp is ext and "pypolymlp_b200" in ext.__file__
    for name in ("PotentialModel", "PotentialHybridModel", "PotentialPropertiesFast", "Readgtinv", "FeaturesAttr",
                 "Neighbor", "NeighborHalf", "NeighborFull", "NeighborCell",   # pybind11_mlp.cpp:12-142
                 "FeatureParams", "get_fn", "get_ylm"):                         # :145-181
        assert hasattr(libmlpcpp, name), name


def test_reference_loader_and_feature_attributes_on_the_dropin(ext):
    """tests/test_calc/test_compute_features.py:45-90 and test_cxx/test_polymlp_api.py:17-50 publish 324 / 1899
    features for the MgO pair / gtinv potentials; the reference computes them through FeaturesAttr."""
    from pypolymlp.core.io_polymlp import load_mlp
    from pypolymlp.mlp_dev.core.features_attr import _get_num_features, get_features_attr

    for name, n_features in [("polymlp.yaml.pair.MgO", 324), ("polymlp.yaml.gtinv.MgO", 1899),
                             ("polymlp.lammps.pair.Ag", 285), ("polymlp.lammps.gtinv.cond.SrTiO3", 5452)]:
        params, coeffs = load_mlp(MLPS + name)
        assert _get_num_features(params) == n_features == len(coeffs)
        features_attr, polynomial_attr, pair_dict = get_features_attr(params)
        assert len(features_attr) + len(polynomial_attr) == n_features
        assert sorted(pair_dict) == list(range(params.n_type * (params.n_type + 1) // 2))
        # our loaders build the same boundary dict the reference hands to the extension (PolymlpParamsSingle.as_dict)
        pd, our_coeffs, _ = our_load_mlp(MLPS + name)
        ref_pd = params.as_dict()
        assert np.array_equal(our_coeffs, coeffs)
        assert pd["n_type"] == ref_pd["n_type"]
        for key in ("cutoff", "pair_type", "feature_type", "model_type", "max_p", "max_l"):
            assert pd["model"][key] == ref_pd["model"][key], (name, key)
        assert [list(map(float, p)) for p in ref_pd["model"]["pair_params"]] == pd["model"]["pair_params"]
        assert {tuple(k): list(v) for k, v in ref_pd["model"]["pair_params_conditional"].items()} == \
            pd["model"]["pair_params_conditional"]
        for key in ("lm_seq", "l_comb", "lm_coeffs"):
            assert ref_pd["model"]["gtinv"][key] == pd["model"]["gtinv"][key], (name, key)


def test_reference_gtinv_known_answers_on_the_dropin(ext):
    """tests/test_cxx/test_gtinv.py:8-38 through the reference's own wrapper."""
    from pypolymlp.cxx.wrapper.api_gtinv_list import get_gtinv_attrs

    l_comb, lm_seq, lm_coeffs = get_gtinv_attrs(3, (4, 4), 1)
    assert len(l_comb) == len(lm_seq) == len(lm_coeffs) == 20  # cfg 1 has 20 l-combinations
    assert l_comb[0] == [0] and lm_coeffs[0] == [1.0]
    l_comb, _, _ = get_gtinv_attrs(4, (12, 8, 2), 1)
    assert len(l_comb) == 82  # cfg 3 / 4


def test_compute_entry_points_need_a_device(ext):
    """No CPU fallback behind the boundary: without a CUDA device the constructors raise (RuntimeError, as C++
    exceptions map in the reference, pybind11_mlp.cpp) instead of computing on the host."""
    import ctypes

    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    from pypolymlp_b200.params import make_params_dict

    pd = make_params_dict(**cases.si_model_kwargs())
    ax, pc, ty = cases.skewed_cell(1)
    with pytest.raises(RuntimeError):
        ext.PotentialModel(pd, [ax], [pc], [ty], [1], [True], [len(ty)])
    with pytest.raises(RuntimeError):
        ext.PotentialPropertiesFast(pd, [0.0] * 168)
    # the radial / spherical-harmonic hooks run the device pair-basis kernel: through the reference's own wrapper
    from pypolymlp.cxx.wrapper.api_functions import get_fn, get_ylm

    with pytest.raises(RuntimeError):
        get_fn(1.2, [[1.0, 0.0]], 6.0)
    with pytest.raises(RuntimeError):
        get_ylm(0.3, -0.5, 0.81, 4)


def test_feature_params_record(ext):
    """FeatureParams is the reference's plain feature_params record (pybind11_mlp.cpp:145-160): every field
    readable and writable."""
    fp = ext.FeatureParams()
    values = dict(n_type=2, force=True, params=[[1.0, 0.5]], params_conditional=[[[0]], [[0]]], cutoff=6.5,
                  pair_type="gaussian", feature_type="gtinv", model_type=3, maxp=2, maxl=4,
                  lm_array=[[[0]]], l_comb=[[0]], lm_coeffs=[[1.0]])
    for key, val in values.items():
        setattr(fp, key, val)
        assert getattr(fp, key) == val, key


def test_reference_call_sites_reach_the_device_step(ext):
    """The reference's own producers hand their objects to the three compute classes
    (mlp_dev/core/features.py:120,193 Features / FeaturesHybrid; calculator/properties_single.py:41): with no CUDA
    device the calls must get past argument conversion and table construction and stop at context creation
    (RuntimeError 'no CPU fallback'), never at a TypeError of the binding."""
    import ctypes
    import glob

    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    from pypolymlp.calculator.properties import Properties
    from pypolymlp.core.interface_vasp import Poscar
    from pypolymlp.core.io_polymlp import load_mlp, load_mlps
    from pypolymlp.core.params import PolymlpParams
    from pypolymlp.mlp_dev.core.features import Features, FeaturesHybrid

    files = "/root/reference/tests/test_calc/files/"
    mgo = Poscar(files + "poscars/POSCAR-00001.MgO").structure
    srtio3 = Poscar(files + "poscars/POSCAR.perovskite.SrTiO3").structure
    params, _ = load_mlp(MLPS + "polymlp.yaml.gtinv.MgO")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Features(PolymlpParams(params), structures=[mgo], print_memory=False)
    hybrid, _ = load_mlps(sorted(glob.glob(MLPS + "polymlp.yaml.flexible.*.SrTiO3")))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FeaturesHybrid(hybrid, structures=[srtio3], print_memory=False)
    for pot in ("polymlp.lammps.gtinv.SrTiO3", "polymlp.lammps.pair.cond.SrTiO3"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Properties(pot=MLPS + pot)


def test_reference_neighbor_cell_test_on_the_dropin(ext):
    """tests/test_cxx/test_neighbor_variants.py:50-66 verbatim in spirit: the reference's own wrapper
    (cxx/wrapper/api_neighbor.py NeighborCell) and Poscar reader, our NeighborCell underneath (host code)."""
    from pypolymlp.core.interface_vasp import Poscar
    from pypolymlp.cxx.wrapper.api_neighbor import NeighborCell

    str1 = Poscar("/root/reference/tests/files/POSCAR-BiGd2").structure
    neigh = NeighborCell(str1, cutoff=6.0)
    np.testing.assert_allclose(neigh.axis, str1.axis)
    np.testing.assert_allclose(neigh.positions_cartesian, str1.axis @ str1.positions)
    assert len(neigh.translations) == 91
    assert len(NeighborCell(str1, cutoff=8.0).translations) == 117
    assert len(NeighborCell(str1, cutoff=16.0).translations) == 281

"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, the host model tables agree with the oracle / reference, and the host logic (weights,
row bookkeeping, sharding, multi-rank reduction) behaves like the reference's Python."""

import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
from oracle import polymlp_oracle as po
from pypolymlp_b200 import _capi, fit
from pypolymlp_b200.libmlpcpp import Readgtinv, _Model, _set_index
from pypolymlp_b200.params import make_params_dict, set_gaussian_params

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "polymlp_b200.h")).read()
    declared = set(re.findall(r"\b(pm_[a-z_0-9]+)\s*\(", header))
    declared -= {"pm_model", "pm_context"}
    lib = _capi.lib()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_capi.EXPORTS) <= declared
    assert b"sm_100a" in lib.pm_version()


def test_no_device_means_error_not_fallback():
    lib = _capi.lib()
    n = C.c_int(-1)
    assert lib.pm_device_count(C.byref(n)) == 0
    if n.value == 0:
        m = _Model(make_params_dict(**cases.si_model_kwargs()))
        h = C.c_void_p()
        st = lib.pm_context_create(m.handle, 0, 0, 0, C.byref(h))
        assert st == 3 and b"no CPU fallback" in lib.pm_last_error()


def test_readgtinv_matches_oracle_reader_and_reference_shapes():
    # reference: tests/test_cxx/test_gtinv.py (v1 tables): order-3 maxl [4,4] -> 20 l-combs
    rg = Readgtinv(3, [4, 4], 1)
    assert len(rg.get_l_comb()) == 20
    refdir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(refdir, "polymlp_gtinv_data_v1_order1.bin")):
        for order, maxl in ((2, [20]), (3, [12, 12]), (4, [12, 8, 2]), (6, [2, 2, 2, 2, 2])):
            a = Readgtinv(order, maxl, 1)
            b = po.readgtinv(order, maxl, refdir, 1)
            assert a.get_l_comb() == b[0] and a.get_lm_seq() == b[1] and a.get_lm_coeffs() == b[2]
    with pytest.raises(ValueError):
        Readgtinv(7, [1] * 6, 1)
    with pytest.raises(RuntimeError):
        Readgtinv(3, [4, 4], 2)  # v2 order-3 table is absent from the reference checkout


def test_gaussian_params():
    pp = set_gaussian_params(n_gaussians=10, cutoff=6.0)
    assert len(pp) == 10 and pp[-1] == [0.0, 0.0] and pp[0] == [1.0, 0.0] and pp[8] == [1.0, 5.0]


@pytest.mark.parametrize("kwargs,n_features", [
    (cases.si_model_kwargs(), 168),            # reference tests/test_mlp_dev/test_core_features.py:18
    (cases.cfg2_model_kwargs(4), 2030),        # config 1/2/5
    (cases.cfg2_model_kwargs(3), 255),
    (cases.cfg3_model_kwargs(), 9385),         # config 3
    (cases.cfg4_model_kwargs(), 45090),        # config 4
])
def test_model_sizes(kwargs, n_features):
    m = _Model(make_params_dict(**kwargs))
    assert m.n_features == n_features


@pytest.mark.parametrize("kwargs", [cases.si_model_kwargs(), cases.binary_model_kwargs(),
                                    cases.ternary_p3_model_kwargs()])
def test_tables_match_oracle(kwargs):
    pd = make_params_dict(**kwargs)
    m, tab = _Model(pd), po.Tables(pd)
    info = m.info()
    assert info["n_variables"] == tab.n_variables and info["n_linear"] == tab.n_linear
    assert info["n_comb2"] == len(tab.comb2) and info["n_comb3"] == len(tab.comb3)
    for t in range(tab.n_type):
        ti = m.type_info(t)
        assert ti["n_full"] == len(tab.local[t]["gids"]) and ti["n_feat"] == len(tab.features[t])
        col, order, ids = m.polynomial(t)
        assert [int(c) for c in col] == [c for c, _ in tab.poly[t]]
        for k, (_, lids) in enumerate(tab.poly[t]):
            assert [int(x) for x in ids[k][: len(lids)]] == list(lids)


def test_invalid_model_arguments():
    pd = make_params_dict(**cases.si_model_kwargs())
    pd["model"]["model_type"] = 5
    with pytest.raises(ValueError):
        _Model(pd)
    pd = make_params_dict(**cases.si_model_kwargs())
    pd["model"]["feature_type"] = "triple"
    with pytest.raises(ValueError):
        _Model(pd)
    # pair models stop at model_type 2 (polymlp_model_params_polynomial.cpp:42-45)
    with pytest.raises(ValueError):
        _Model(make_params_dict(**cases.pair_model_kwargs(2, model_type=3)))


@pytest.mark.parametrize("kw", [cases.pair_model_kwargs(1), cases.pair_model_kwargs(2),
                                cases.pair_model_kwargs(3, max_p=3), cases.mgo_model_kwargs("pair")])
def test_pair_model_tables_match_oracle(kw):
    """feature_type = 'pair': column count and, per centre type, the polynomial terms as sets of global linear ids
    (the local feature order differs on purpose: ours is regrouped by radial index)."""
    pd = make_params_dict(**kw)
    tab, m = po.Tables(pd), _Model(pd)
    assert m.n_features == tab.n_variables and m.info()["n_linear"] == tab.n_linear
    for t in range(kw["n_type"]):
        col, order, ids = m.polynomial(t)
        l2g = {int(ids[k][0]): int(col[k]) for k in range(len(col)) if order[k] == 1}
        mine = {int(col[k]): sorted(l2g[int(x)] for x in ids[k][: order[k]]) for k in range(len(col))}
        o_l2g = {lids[0]: c for c, lids in tab.poly[t] if len(lids) == 1}
        assert mine == {c: sorted(o_l2g[x] for x in lids) for c, lids in tab.poly[t]}


def test_polymlp_yaml_round_trip(tmp_path):
    """save_mlp_yaml -> load_mlp_yaml (format of src/pypolymlp/core/io_polymlp_yaml.py:17-161)."""
    from pypolymlp_b200.io_yaml import load_mlp_yaml, save_mlp_yaml

    for kw, elements in ((cases.binary_model_kwargs(), ["Mg", "O"]), (cases.pair_model_kwargs(2), ["Sr", "Ti"])):
        pd = make_params_dict(**kw)
        n = _Model(pd).n_features
        rng = np.random.default_rng(3)
        coeffs, scales = rng.normal(size=n), rng.uniform(0.5, 2.0, n)
        path = str(tmp_path / "polymlp.yaml")
        save_mlp_yaml(pd, coeffs, scales, elements, path)
        pd2, c2, meta = load_mlp_yaml(path)
        assert meta["elements"] == elements and meta["type_full"]
        np.testing.assert_allclose(c2, coeffs / scales, rtol=1e-15)
        assert pd2["model"] == pd["model"] and pd2["n_type"] == pd["n_type"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/test_calc/files/mlps"), reason="reference tree not mounted")
def test_polymlp_yaml_reads_reference_files():
    from pypolymlp_b200.io_yaml import load_mlp_yaml

    M = cases.load_mgo()
    for kind in ("pair", "gtinv"):
        pd, coeffs, meta = load_mlp_yaml("/root/reference/tests/test_calc/files/mlps/polymlp.yaml.%s.MgO" % kind)
        assert meta["elements"] == ["Mg", "O"]
        assert np.array_equal(coeffs, M[kind + "_coeffs"])
        want = make_params_dict(**cases.mgo_model_kwargs(kind))["model"]
        want["pair_conditional"] = True  # the loader always passes explicit per-pair lists (io_polymlp_yaml.py:139)
        assert pd["model"] == want


def test_flop_count_config2():
    # config 2, 256 atoms x 54 neighbours, F = 2030
    m = _Model(make_params_dict(**cases.cfg2_model_kwargs(4)))
    w = m.count_flops([256], [[13824]], True)
    assert w["syrk"] == 775 * 2030 * 2031
    assert w["anlm"] == 13824 * 10 * 15 * 32
    assert w["poly"] == 256 * 172 * 7520
    assert 0.5e9 < w["deriv"] < 2e9


def test_set_index_row_layout():
    # compute/py_model.cpp:58-106: energies | stress | forces
    fb, sb, nd = _set_index([2, 1, 2], [True, False, True], [4, 4, 3, 2, 5])
    assert nd == [5, 3 * (4 + 4 + 2 + 5), 24]
    assert sb == [5, -1, 17] and fb == [29, -1, 53]


def test_apply_weights_matches_reference_rules():
    rng = np.random.default_rng(0)
    n = 3
    ds = fit.Dataset([np.eye(3) * 5] * n, [rng.random((3, 2))] * n, [[0, 0]] * n, energies=np.array([-10.0, -7.0, 1.0]),
                     forces=np.array([0.5, -2.0, 1e-15] * 6), stresses=np.array([1e-3, -20.0, 0.0, 0.3, 5.0, -1e-13] * n),
                     include_force=True, include_stress=True, weight=2.0)
    w, y = fit.apply_weights(ds, weight_stress=0.1, min_e=-5.0)
    we = po.weights_energy(ds.energies, ds.total_n_atoms, -5.0, 2.0)
    ws = po.weights_stress(ds.stresses, 0.1 * 2.0)
    wf = po.weights_force(ds.forces, 2.0)
    np.testing.assert_array_equal(w, np.concatenate([we, ws, wf]))
    np.testing.assert_array_equal(y, np.concatenate([we * ds.energies, ws * ds.stresses, wf * ds.forces]))
    ds.include_stress = False
    w, y = fit.apply_weights(ds, min_e=-5.0)
    assert np.all(w[n:n + 6 * n] == 0.0) and np.all(y[n:n + 6 * n] == 0.0)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 10000):
        for ws in (1, 2, 3, 8):
            parts = [fit.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[k][1] == parts[k + 1][0] for k in range(ws - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    X = rng.normal(size=(40, 6))
    y = rng.normal(size=40)
    b, e = fit.shard_range(40, rank, world)
    # packed accumulator of this rank's shard, same layout as pm_fit_accumulator: [C | xe_sum | xe_sq | n]
    Xt = np.hstack([X[b:e], y[b:e, None]])
    acc = np.concatenate([(Xt.T @ Xt).ravel(), X[b:e].sum(0), np.square(X[b:e]).sum(0), [e - b]])
    t = torch.from_numpy(acc)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        Xa = np.hstack([X, y[:, None]])
        ok = np.allclose(t.numpy()[:49], (Xa.T @ Xa).ravel(), rtol=1e-12) and t.numpy()[-1] == 40
        q.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_reduce_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def _pybind_module():
    import importlib.util

    from pypolymlp_b200.build import pybind_module_path

    spec = importlib.util.spec_from_file_location("libmlpcpp", pybind_module_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pybind_dropin_module_host_side():
    """The pybind11 `libmlpcpp` drop-in exposes the reference's class names and getters
    (python/pybind11_mlp.cpp:10-94) and maps C-ABI errors to ValueError / RuntimeError."""
    m = _pybind_module()
    for cls, methods in (("PotentialModel", ("get_x", "get_fbegin", "get_sbegin", "get_n_data")),
                         ("PotentialPropertiesFast", ("eval", "eval_multiple", "get_e", "get_f", "get_s",
                                                      "get_e_array", "get_f_array", "get_s_array")),
                         ("Readgtinv", ("get_lm_seq", "get_l_comb", "get_lm_coeffs")),
                         ("PotentialHybridModel", ("get_x", "get_fbegin", "get_sbegin", "get_cumulative_n_features",
                                                   "get_n_data")),
                         ("FeaturesAttr", ("get_n_features",)), ("PotentialXtX", ("add", "finalize"))):
        assert all(hasattr(getattr(m, cls), k) for k in methods), cls
    rg = m.Readgtinv(3, [4, 4], 1)
    ours = Readgtinv(3, [4, 4], 1)
    assert rg.get_l_comb() == ours.get_l_comb() and rg.get_lm_seq() == ours.get_lm_seq()
    assert rg.get_lm_coeffs() == ours.get_lm_coeffs()
    pd = make_params_dict(**cases.si_model_kwargs())
    assert m.FeaturesAttr(pd).get_n_features() == 168
    assert m.FeaturesAttr(make_params_dict(**cases.mgo_model_kwargs("pair"))).get_n_features() == 324
    with pytest.raises(ValueError):
        m.Readgtinv(9, [1], 1)
    bad = make_params_dict(**cases.si_model_kwargs())
    bad["model"]["max_p"] = 7
    with pytest.raises(ValueError):
        m.FeaturesAttr(bad)


def test_neighbor_cell_hook_reference_known_answers():
    """NeighborCell test hook (host code, pm_cell_translations): the reference's translation counts for
    POSCAR-rocksalt (tests/test_cxx/test_neighbor.py:189-207), the unrefined cell handed back unchanged, and the
    skewed-cell reduction against the oracle's restatement of compute/neighbor_cell.cpp."""
    m = _pybind_module()
    axis = np.eye(3) * 4.0
    frac = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [0, 0, .5], [0, .5, 0], [.5, 0, 0], [.5, .5, .5]]).T
    pc = axis @ frac
    for cutoff, n in ((6.0, 147), (8.0, 203), (16.0, 751)):
        nc = m.NeighborCell(axis.tolist(), pc.tolist(), cutoff)
        assert len(nc.get_translations()) == n
        np.testing.assert_allclose(np.array(nc.get_positions_cartesian()), pc)
        np.testing.assert_allclose(np.array(nc.get_axis()), axis)
    ax, pcs, _ = cases.skewed_cell(1, n_atom=5, seed=3)
    ref = po.NeighborCell(ax, pcs, 6.0)
    nc = m.NeighborCell(ax.tolist(), pcs.tolist(), 6.0)
    assert np.array_equal(np.array(nc.get_translations()), np.asarray(ref.trans))
    assert np.array_equal(np.array(nc.get_axis()), np.asarray(ref.axis))
    assert np.array_equal(np.array(nc.get_positions_cartesian()), np.asarray(ref.pos))
    for cls, methods in (("Neighbor", ("get_distances", "get_differences", "get_neighbor_indices")),
                         ("NeighborFull", ("get_distances", "get_differences", "get_neighbor_indices")),
                         ("NeighborHalf", ("get_differences", "get_neighbor_indices"))):
        assert all(hasattr(getattr(m, cls), k) for k in methods), cls


# ---- FeaturesAttr: the reference's five getters (compute/py_features_attr.cpp:11-63) -------------------------------
@pytest.mark.parametrize("kw", [cases.si_model_kwargs(), cases.cfg2_model_kwargs(), cases.binary_model_kwargs(),
                                cases.ternary_p3_model_kwargs(), cases.pair_model_kwargs(2),
                                cases.pair_model_kwargs(3, max_p=3), cases.mgo_model_kwargs("gtinv"),
                                cases.cfg3_model_kwargs()])
def test_features_attr_getters_match_reference(kw):
    """radial / gtinv / type-pair-combination ids per linear feature, polynomial combinations and the type-pair table,
    from the ctypes mirror (pm_model_feature_attrs) and from the compiled pybind11 drop-in, against the reference's
    own FeaturesAttr construction run in oracle/_ref."""
    from oracle import ref
    from pypolymlp_b200 import dropin
    from pypolymlp_b200.libmlpcpp import FeaturesAttr

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    pd = make_params_dict(**kw)
    radial, gtinv, tcomb, poly, type_pairs = ref.RefModel(pd).feature_attrs()
    assert len(radial) + len(poly) == FeaturesAttr(pd).get_n_features()
    for impl in (FeaturesAttr(pd), dropin.load_extension().FeaturesAttr(pd)):
        assert list(impl.get_radial_ids()) == radial
        assert list(impl.get_gtinv_ids()) == gtinv  # empty for pair models, as in the reference
        assert [list(v) for v in impl.get_tcomb_ids()] == tcomb
        assert [list(v) for v in impl.get_polynomial_ids()] == poly
        assert [list(v) for v in impl.get_type_pairs()] == type_pairs


def test_count_flops_tool_config2():
    """tools/count_flops.py reproduces SURVEY 8(d)'s config-2 figures: 13 824 ordered pairs, 775 rows,
    W = 4.33 GFLOP per structure (SYRK 3.195e9)."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "count_flops.py"), "2"], capture_output=True,
                         text=True, check=True).stdout.strip().splitlines()[-1]
    rec = json.loads(out)
    assert (rec["n_features"], rec["atoms"], rec["ordered_pairs"], rec["rows"]) == (2030, 256, 13824, 775)
    assert rec["flops"]["syrk"] == pytest.approx(775 * 2030 * 2031, rel=1e-3)
    assert rec["flops_per_structure"] == pytest.approx(4.327e9, rel=1e-3)

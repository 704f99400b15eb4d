"""Makes the compiled drop-in (`pypolymlp_b200/lib/libmlpcpp.*.so`) importable under the reference's own module
name `pypolymlp.cxx.lib.libmlpcpp`, so that an UNMODIFIED checkout of the reference's Python package
(`src/pypolymlp`, whose `cxx/lib/` directory is a build product and absent from a source tree) runs on top of it.

The reference imports the extension as `from pypolymlp.cxx.lib import libmlpcpp`
(src/pypolymlp/mlp_dev/core/features.py:10, src/pypolymlp/cxx/wrapper/api_gtinv_list.py:3,
src/pypolymlp/calculator/properties_single.py).  INTEGRATION.md describes the permanent variant (copy the two
shared objects into `src/pypolymlp/cxx/lib/`); `install()` is the in-process variant used by the tests."""

import glob
import importlib.util
import os
import sys
import types

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def load_extension():
    """Imports the compiled pybind11 module itself (not the ctypes mirror `pypolymlp_b200.libmlpcpp`)."""
    hits = sorted(glob.glob(os.path.join(_LIB_DIR, "libmlpcpp.*.so")))
    if not hits:
        raise ImportError("pypolymlp_b200/lib/libmlpcpp.*.so is missing: run `python __graft_entry__.py` (build())")
    spec = importlib.util.spec_from_file_location("libmlpcpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install(reference_src=None):
    """Registers the extension as `pypolymlp.cxx.lib.libmlpcpp`.  `reference_src` (the directory holding the
    reference's `pypolymlp/` package) is prepended to sys.path when given.  Returns the extension module."""
    if reference_src is not None and reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    name = "pypolymlp.cxx.lib.libmlpcpp"
    if name in sys.modules:
        return sys.modules[name]
    ext = load_extension()
    import pypolymlp.cxx  # the reference's package (namespace or regular); fails loudly when it is not importable

    pkg = types.ModuleType("pypolymlp.cxx.lib")
    pkg.__path__ = [_LIB_DIR]
    pkg.libmlpcpp = ext
    sys.modules["pypolymlp.cxx.lib"] = pkg
    sys.modules[name] = ext
    pypolymlp.cxx.lib = pkg
    return ext


def install_fused_products(reference_src=None):
    """Switches the reference's fit to the fused device product: replaces `calc_xtx_xty` of
    `pypolymlp.mlp_dev.core.data_sequential` (data_sequential.py:25-94; its inner step
    `_compute_products_single_batch`, :97-156, is "compute X -> apply_weights -> x.T @ x" on the host) by a version that
    hands every batch of structures with its row weights and weighted targets to `libmlpcpp.PotentialXtX` and reads
    X^T X / X^T y / xe_sum / xe_sq_sum / y^T y back once.  Everything else is the reference's own code: batching
    (`get_batch_slice`, `slice_dft`), weights and targets (`apply_weights`, run on a one-column stand-in for X, which
    is how the weights come out without X), scales (`compute_scales`) and the scaling tail.  Hybrid models
    (len(params) > 1) keep the reference's original path.  Call before importing the reference's fit modules
    (they bind `calc_xtx_xty` by name at import).  Returns the replacement function."""
    ext = install(reference_src)
    import numpy as np
    from pypolymlp.mlp_dev.core import data_sequential as ds
    from pypolymlp.mlp_dev.core.features import _init_features

    if getattr(ds.calc_xtx_xty, "_b200_fused", False):
        return ds.calc_xtx_xty
    original = ds.calc_xtx_xty

    def batch_weights(dataset_sliced, n_atoms_sum, weight_stress, min_energy):
        """Row weights w and weighted targets y of one batch in the PotentialModel row layout
        (energies | 6 stress rows per structure | 3N force rows per structure; compute/py_model.cpp:58-106)."""
        n_st = len(dataset_sliced.energies)
        if dataset_sliced.include_force:
            n_rows = n_st + 6 * n_st + 3 * int(sum(n_atoms_sum))
            first_indices = (0, 7 * n_st, n_st)          # (ebegin, fbegin, sbegin)
        else:
            n_rows, first_indices = n_st, (0, -1, -1)
        x1 = np.ones((n_rows, 1))
        x1, y, w = ds.apply_weights(x1, np.zeros(n_rows), np.ones(n_rows), dataset_sliced, first_indices,
                                    weight_stress=weight_stress, min_e=min_energy)
        return x1[:, 0].copy(), y                         # the scaled column IS the applied row weight

    def calc_xtx_xty(params, datasets, scales=None, min_energy=None, weight_stress=0.1, batch_size=None,
                     use_gradient=False, n_features_threshold=50000, scale_threshold=1e-10, verbose=False):
        if len(params) > 1:
            return original(params, datasets, scales=scales, min_energy=min_energy, weight_stress=weight_stress,
                            batch_size=batch_size, use_gradient=use_gradient,
                            n_features_threshold=n_features_threshold, scale_threshold=scale_threshold, verbose=verbose)
        if batch_size is None:
            batch_size = 256      # structures per call; the device chunks further by its own workspace
        if min_energy is None:
            min_energy = ds.get_min_energy(datasets)
        # POLYMLP_B200_DEVICES="0,1,..": shard every batch over these GPUs (one NCCL reduce in finalize)
        devs = [int(v) for v in os.environ.get("POLYMLP_B200_DEVICES", "").split(",") if v.strip()]
        acc = ext.PotentialXtX(params.as_dict(), devices=devs)
        for data in datasets:
            if verbose:
                print("----- Dataset:", data.name, "-----", flush=True)
            data.sort_dft()
            n_str = len(data.structures)
            begin_ids, end_ids = ds.get_batch_slice(n_str, batch_size)
            for begin, end in zip(begin_ids, end_ids):
                sliced = data.slice_dft(begin, end)
                axis, positions_c, types, n_atoms_sum, _, _ = _init_features(sliced, None, params)
                w, y = batch_weights(sliced, n_atoms_sum, weight_stress, min_energy)
                acc.add(axis, positions_c, types, [bool(sliced.include_force)] * len(axis), w, y)
        res = acc.finalize()
        data_xy = ds.PolymlpDataXY()
        data_xy.xtx, data_xy.xty = res["xtx"], res["xty"]
        data_xy.y_sq_norm, data_xy.total_n_data = float(res["y_sq_norm"]), int(res["total_n_data"])
        if scales is None:
            data_xy.xe_sum, data_xy.xe_sq_sum = res["xe_sum"], res["xe_sq_sum"]
        n_data = sum([len(d.energies) for d in datasets])
        scales, zero_ids = ds.compute_scales(scales, data_xy.xe_sum, data_xy.xe_sq_sum, n_data,
                                             include_force=datasets.include_force, threshold=scale_threshold)
        data_xy.xtx[zero_ids] = 0.0
        data_xy.xtx[:, zero_ids] = 0.0
        data_xy.xty[zero_ids] = 0.0
        data_xy.xtx /= scales[:, np.newaxis]
        data_xy.xtx /= scales[np.newaxis, :]
        data_xy.xty /= scales
        data_xy.scales = scales
        data_xy.min_energy = min_energy
        return data_xy

    calc_xtx_xty._b200_fused = True
    calc_xtx_xty._b200_original = original
    calc_xtx_xty._b200_batch_weights = batch_weights
    ds.calc_xtx_xty = calc_xtx_xty
    return calc_xtx_xty

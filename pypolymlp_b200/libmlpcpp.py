"""Drop-in mirror of the reference's pybind11 module `pypolymlp.cxx.lib.libmlpcpp` for the hot path.

Same class names, constructor arguments, getters and error behaviour as
src/pypolymlp/cxx/src/python/pybind11_mlp.cpp:10-94 of the reference, implemented on top of the
C ABI (include/polymlp_b200.h) -> CUDA.  Additive entry point: `PotentialXtX` (fused feature +
X^T X accumulation, replaces get_x() + apply_weights + x.T @ x of data_sequential.py:97-156).
"""

import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import FeatureParamsC, StructureBatch, as_d, as_i, check, lib, pd, pi


def default_device():
    for key in ("POLYMLP_B200_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


class Readgtinv:
    """Readgtinv(order, maxl, version) (reference: polymlp_read_gtinv.cpp:10-63)."""

    def __init__(self, gtinv_order, gtinv_maxl, version=1, datadir=None):
        L = lib()
        ml = as_i(list(gtinv_maxl) + [0])
        sizes = (C.c_int64 * 4)()
        dd = (datadir or _capi.DATA_DIR).encode()
        n_ml = len(gtinv_maxl)
        check(L.pm_gtinv_read(dd, gtinv_order, pi(ml), n_ml, version, sizes, None, None, None, None, None))
        n, s1, s2, s3 = (int(x) for x in sizes)
        lo, lc, nt = np.zeros(n, np.int32), np.zeros(max(s1, 1), np.int32), np.zeros(n, np.int32)
        lm, cf = np.zeros(max(s3, 1), np.int32), np.zeros(max(s2, 1), np.float64)
        check(L.pm_gtinv_read(dd, gtinv_order, pi(ml), n_ml, version, sizes, pi(lo), pi(lc), pi(nt), pi(lm), pd(cf)))
        self._l_comb, self._lm_seq, self._lm_coeffs = [], [], []
        p1 = p2 = p3 = 0
        for i in range(n):
            o, t = int(lo[i]), int(nt[i])
            self._l_comb.append([int(x) for x in lc[p1:p1 + o]])
            self._lm_coeffs.append([float(x) for x in cf[p2:p2 + t]])
            self._lm_seq.append(lm[p3:p3 + t * o].reshape(t, o).tolist())
            p1, p2, p3 = p1 + o, p2 + t, p3 + t * o

    def get_lm_seq(self):
        return self._lm_seq

    def get_l_comb(self):
        return self._l_comb

    def get_lm_coeffs(self):
        return self._lm_coeffs


class _Model:
    """Host tables built from `params.as_dict()` (keys read: compute/py_params.cpp:14-43)."""

    def __init__(self, params_dict):
        model = params_dict["model"]
        if model["feature_type"] not in ("gtinv", "pair"):
            raise ValueError("feature_type must be 'gtinv' or 'pair'")
        pair = model["feature_type"] == "pair"
        if model.get("pair_type", "gaussian") != "gaussian":
            raise ValueError("pypolymlp_b200 implements pair_type='gaussian' only")
        n_type = int(params_dict["n_type"])
        pp = as_d(model["pair_params"]).reshape(-1, 2)
        cond = model.get("pair_params_conditional")
        off, val = [0], []
        for i in range(n_type):
            for j in range(i, n_type):
                lst = list(cond[(i, j)]) if cond else list(range(len(pp)))
                val.extend(lst)
                off.append(len(val))
        g = model.get("gtinv", {})
        l_comb, lm_seq, lm_coeffs = ([], [], []) if pair else (g["l_comb"], g["lm_seq"], g["lm_coeffs"])
        lo = as_i([len(x) for x in l_comb])
        lc = as_i([v for x in l_comb for v in x])
        nt = as_i([len(x) for x in lm_seq])
        lm = as_i([v for x in lm_seq for term in x for v in term])
        cf = as_d([v for x in lm_coeffs for v in x])
        off, val = as_i(off), as_i(val if val else [0])
        self._keep = (pp, off, val, lo, lc, nt, lm, cf)
        fp = FeatureParamsC(n_type, len(pp), pd(pp), pi(off), pi(val), float(model["cutoff"]),
                            int(model["model_type"]), int(model["max_p"]), int(model["max_l"]), len(l_comb),
                            pi(lo), pi(lc), pi(nt), pi(lm), pd(cf), 1 if pair else 0)
        h = C.c_void_p()
        check(lib().pm_model_create(C.byref(fp), C.byref(h)))
        self.handle = h
        self.n_type = n_type
        self.n_features = lib().pm_model_n_features(self.handle)

    def __del__(self):
        if getattr(self, "handle", None):
            lib().pm_model_destroy(self.handle)
            self.handle = None

    def info(self):
        out = (C.c_int64 * 8)()
        check(lib().pm_model_info(self.handle, out))
        keys = ["n_type", "n_linear", "n_comb2", "n_comb3", "n_variables", "n_polyvars", "n_type_pairs", "n_lm_half"]
        return dict(zip(keys, (int(x) for x in out)))

    def type_info(self, t):
        out = (C.c_int64 * 12)()
        check(lib().pm_model_type_info(self.handle, t, out))
        keys = ["n_full", "n_head", "n_feat", "n_fpad", "n_terms", "n_G_entries", "n_contributions", "n_blocks",
                "n_poly", "n_deriv_pairs"]
        return dict(zip(keys, (int(x) for x in out)))

    def polynomial(self, t):
        n = C.c_int(0)
        check(lib().pm_model_polynomial(self.handle, t, C.byref(n), None, None, None))
        col, order, ids = np.zeros(n.value, np.int32), np.zeros(n.value, np.int32), np.zeros((n.value, 3), np.int32)
        check(lib().pm_model_polynomial(self.handle, t, C.byref(n), pi(col), pi(order), pi(ids)))
        return col, order, ids

    def count_flops(self, atoms_t, pairs_tt, force=True):
        a = np.ascontiguousarray(atoms_t, dtype=np.int64)
        p = np.ascontiguousarray(pairs_tt, dtype=np.int64).reshape(-1)
        out = (C.c_double * 5)()
        check(lib().pm_model_count_flops(self.handle, a.ctypes.data_as(_capi._i64p), p.ctypes.data_as(_capi._i64p),
                                         int(force), out))
        return dict(zip(["syrk", "xty", "poly", "deriv", "anlm"], (float(x) for x in out)))


class FeaturesAttr:
    """FeaturesAttr(params_dict) with the reference's five getters (compute/py_features_attr.cpp:11-63,
    pybind11_mlp.cpp:70-82) plus `get_n_features`; host tables only, no device needed."""

    def __init__(self, params_dict):
        self._model = m = _Model(params_dict)
        sizes = (C.c_int64 * 6)()
        check(lib().pm_model_feature_attrs(m.handle, sizes, None, None, None, None, None, None, None))
        n_lin, n_gt, n_tc, n_poly, n_pv, nt = (int(v) for v in sizes)
        radial, gt = np.zeros(max(n_lin, 1), np.int32), np.zeros(max(n_gt, 1), np.int32)
        tc_off, tc = np.zeros(n_lin + 1, np.int32), np.zeros(max(n_tc, 1), np.int32)
        p_off, pv = np.zeros(n_poly + 1, np.int32), np.zeros(max(n_pv, 1), np.int32)
        tps = np.zeros(nt * nt, np.int32)
        check(lib().pm_model_feature_attrs(m.handle, sizes, pi(radial), pi(gt), pi(tc_off), pi(tc), pi(p_off), pi(pv),
                                           pi(tps)))
        self._radial_ids = radial[:n_lin].tolist()
        self._gtinv_ids = gt[:n_gt].tolist()
        self._tcomb_ids = [tc[tc_off[k]:tc_off[k + 1]].tolist() for k in range(n_lin)]
        self._polynomial_ids = [pv[p_off[k]:p_off[k + 1]].tolist() for k in range(n_poly)]
        self._type_pairs = tps.reshape(nt, nt).tolist()

    def get_n_features(self):
        return self._model.n_features

    def get_radial_ids(self):
        return self._radial_ids

    def get_gtinv_ids(self):
        return self._gtinv_ids

    def get_tcomb_ids(self):
        return self._tcomb_ids

    def get_polynomial_ids(self):
        return self._polynomial_ids

    def get_type_pairs(self):
        return self._type_pairs

    def get_polynomial_terms(self, t=0):
        """(column, order, local ids) of centre type t's polynomial terms (additive; not in the reference class)."""
        return self._model.polynomial(t)


class _Context:
    def __init__(self, model, device=None, workspace_bytes=0, flags=0):
        self.model = model
        self.device = default_device() if device is None else device
        h = C.c_void_p()
        check(lib().pm_context_create(model.handle, self.device, workspace_bytes, flags, C.byref(h)))
        self.handle = h

    def __del__(self):
        if getattr(self, "handle", None):
            lib().pm_context_destroy(self.handle)
            self.handle = None

    def profile(self, on=True):
        lib().pm_profile_enable(self.handle, int(on))

    def profile_get(self):
        n = C.c_int(0)
        ms = (C.c_double * 32)()
        ln = (C.c_int64 * 32)()
        lib().pm_profile_get(self.handle, C.byref(n), ms, ln)
        return {lib().pm_stage_name(i).decode(): (ms[i], int(ln[i])) for i in range(n.value)}

    def timer_start(self):
        check(lib().pm_timer_start(self.handle))

    def timer_stop(self):
        """Milliseconds since timer_start on the context's stream (CUDA events)."""
        ms = C.c_double(0.0)
        check(lib().pm_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().pm_launch_count(self.handle))

    def synchronize(self):
        check(lib().pm_synchronize(self.handle))

    def microbench(self, which, n=8192):
        t = C.c_double(0.0)
        check(lib().pm_microbench(self.handle, which, n, C.byref(t)))
        return t.value

    def neighbor_full(self, axis, positions_c, types):
        axis, pos, ty = as_d(axis), as_d(positions_c), as_i(types)
        n = pos.shape[1]
        off = np.zeros(n + 1, np.int32)
        check(lib().pm_neighbor_full(self.handle, pd(axis), pd(pos), pi(ty), n, pi(off), None, None, None, None))
        P = int(off[-1])
        nb, dx, dy, dz = np.zeros(max(P, 1), np.int32), np.zeros(max(P, 1)), np.zeros(max(P, 1)), np.zeros(max(P, 1))
        check(lib().pm_neighbor_full(self.handle, pd(axis), pd(pos), pi(ty), n, pi(off), pi(nb), pd(dx), pd(dy), pd(dz)))
        return off, nb[:P], dx[:P], dy[:P], dz[:P]

    def debug_fetch(self, what):
        n = C.c_size_t(0)
        check(lib().pm_debug_fetch(self.handle, what, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1))
        check(lib().pm_debug_fetch(self.handle, what, pd(out), out.size, C.byref(n)))
        return out[: n.value]


def _force_per_structure(n_st_dataset, force_dataset):
    flags = []
    for n, f in zip(n_st_dataset, force_dataset):
        flags.extend([bool(f)] * int(n))
    return flags


def _set_index(n_st_dataset, force_dataset, n_atoms_all):
    """Row bookkeeping of PyModel::set_index (compute/py_model.cpp:58-106)."""
    n_st = int(sum(n_st_dataset))
    fbegin, sbegin = [-1] * len(n_st_dataset), [-1] * len(n_st_dataset)
    n_data = [n_st, 0, 0]
    ist, k = n_st, 0
    for i, (n, f) in enumerate(zip(n_st_dataset, force_dataset)):
        if f:
            sbegin[i] = ist
            n_data[2] += 6 * n
            ist += 6 * n
        k += n
    ifo, k = ist, 0
    for i, (n, f) in enumerate(zip(n_st_dataset, force_dataset)):
        if f:
            fbegin[i] = ifo
        for _ in range(n):
            if f:
                ifo += 3 * int(n_atoms_all[k])
                n_data[1] += 3 * int(n_atoms_all[k])
            k += 1
    return fbegin, sbegin, n_data


class PotentialModel:
    """PotentialModel(params_dict, axis, positions_c, types, n_st_dataset, force_dataset, n_atoms_all)
    (reference: pybind11_mlp.cpp:12-27, compute/py_model.cpp:10-54).  All work happens in the
    constructor; get_x() returns the (n_rows, n_features) design matrix."""

    def __init__(self, params_dict, axis, positions_c, types, n_st_dataset, force_dataset, n_atoms_all,
                 device=None, flags=0):
        if len(axis) != sum(n_st_dataset):
            raise ValueError("n_st_dataset does not match the number of structures")
        self._model = _Model(params_dict)
        self._ctx = _Context(self._model, device, flags=flags)
        batch = StructureBatch(axis, positions_c, types, _force_per_structure(n_st_dataset, force_dataset))
        self._fbegin, self._sbegin, self._n_data = _set_index(n_st_dataset, force_dataset, n_atoms_all)
        self._x = np.zeros((batch.n_rows, self._model.n_features))
        if params_dict.get("print_memory", False):
            print(" Matrix shape (X):", self._x.shape, flush=True)
        check(lib().pm_features_x(self._ctx.handle, C.byref(batch.c), pd(self._x)))

    def get_x(self):
        return self._x

    def get_fbegin(self):
        return self._fbegin

    def get_sbegin(self):
        return self._sbegin

    def get_n_data(self):
        return self._n_data


def find_active_atoms(type_full, type_indices, types, positions_c):
    """Atoms a hybrid sub-model sees (PyHybridModel::find_active_atoms, compute/py_hybrid_model.cpp:116-156):
    all of them when type_full, else those whose type is listed in type_indices, renumbered by position in it."""
    types = np.asarray(types, dtype=np.int32)
    positions_c = np.asarray(positions_c, dtype=np.float64)
    if type_full:
        return np.arange(len(types)), types, positions_c
    type_indices = list(type_indices)
    active = np.array([a for a, t in enumerate(types) if t in type_indices], dtype=np.int64)
    # std::replace runs in list order on the already partly renumbered array: reproduce it literally
    ta = types[active].copy()
    for t_rep, t in enumerate(type_indices):
        ta[ta == t] = t_rep
    return active, ta, positions_c[:, active]


class PotentialHybridModel:
    """PotentialHybridModel(params_dict_array, axis, positions_c, types, n_st_dataset, force_dataset, n_atoms_all)
    (reference: pybind11_mlp.cpp:30-49, compute/py_hybrid_model.cpp:11-114): the design matrices of the
    sub-models side by side, each computed on the atoms of its own element subset (type_indices / type_full)
    and scattered back to the rows of the full structure.  Every sub-model runs through the same CUDA path."""

    def __init__(self, params_dict_array, axis, positions_c, types, n_st_dataset, force_dataset, n_atoms_all,
                 device=None, flags=0):
        if len(axis) != sum(n_st_dataset):
            raise ValueError("n_st_dataset does not match the number of structures")
        force_st = _force_per_structure(n_st_dataset, force_dataset)
        self._fbegin, self._sbegin, self._n_data = _set_index(n_st_dataset, force_dataset, n_atoms_all)
        n_st = len(axis)
        # first X row of every structure's stress / force block in the full layout
        xs_begin, xf_begin = [-1] * n_st, [-1] * n_st
        ist = n_st
        for k in range(n_st):
            if force_st[k]:
                xs_begin[k] = ist
                ist += 6
        for k in range(n_st):
            if force_st[k]:
                xf_begin[k] = ist
                ist += 3 * int(n_atoms_all[k])
        models = [_Model(p) for p in params_dict_array]
        self._cumulative = list(np.cumsum([m.n_features for m in models]).astype(int))
        self._x = np.zeros((sum(self._n_data), self._cumulative[-1] if models else 0))
        if params_dict_array and params_dict_array[0].get("print_memory", False):
            print(" matrix shape (X):", self._x.shape, flush=True)
        for n, (p, model) in enumerate(zip(params_dict_array, models)):
            first = 0 if n == 0 else self._cumulative[n - 1]
            cols = slice(first, self._cumulative[n])
            act = [find_active_atoms(p["type_full"], p["type_indices"], types[k], positions_c[k]) for k in range(n_st)]
            ctx = _Context(model, device, flags=flags)
            batch = StructureBatch(axis, [a[2] for a in act], [a[1] for a in act], force_st)
            xn = np.zeros((batch.n_rows, model.n_features))
            check(lib().pm_features_x(ctx.handle, C.byref(batch.c), pd(xn)))
            self._x[:n_st, cols] = xn[:n_st]
            isb = n_st
            ifb = n_st + 6 * sum(force_st)
            for k in range(n_st):
                if not force_st[k]:
                    continue
                self._x[xs_begin[k]:xs_begin[k] + 6, cols] = xn[isb:isb + 6]
                isb += 6
                active = act[k][0]
                rows = (xf_begin[k] + 3 * active[:, None] + np.arange(3)[None, :]).reshape(-1)
                self._x[rows, cols] = xn[ifb:ifb + 3 * len(active)]
                ifb += 3 * len(active)

    def get_x(self):
        return self._x

    def get_fbegin(self):
        return self._fbegin

    def get_sbegin(self):
        return self._sbegin

    def get_cumulative_n_features(self):
        return self._cumulative

    def get_n_data(self):
        return self._n_data


class PotentialXtX:
    """Fused feature + X^T X / X^T y accumulation on the GPU (additive entry point).

    add(axis, positions_c, types, force_flags, w, y): w and y are the per-row weights and weighted
    targets of the batch in the PyModel row layout (what apply_weights produces)."""

    def __init__(self, params_dict, device=None, workspace_bytes=0, flags=0):
        self._model = _Model(params_dict)
        self._ctx = _Context(self._model, device, workspace_bytes, flags)
        self.n_features = self._model.n_features
        check(lib().pm_fit_reset(self._ctx.handle))
        self._staged = None

    @property
    def context(self):
        return self._ctx

    @property
    def model(self):
        return self._model

    def reset(self):
        check(lib().pm_fit_reset(self._ctx.handle))

    def add(self, axis, positions_c, types, force_flags, w, y):
        batch = StructureBatch(axis, positions_c, types, force_flags)
        w, y = as_d(w), as_d(y)
        if len(w) != batch.n_rows or len(y) != batch.n_rows:
            raise ValueError("w and y must have one entry per row of the batch")
        check(lib().pm_fit_accumulate(self._ctx.handle, C.byref(batch.c), pd(w), pd(y)))
        return batch

    def add_batch(self, batch, w, y):
        check(lib().pm_fit_accumulate(self._ctx.handle, C.byref(batch.c), pd(w), pd(y)))

    def stage(self, axis, positions_c, types, force_flags, w, y):
        batch = StructureBatch(axis, positions_c, types, force_flags)
        w, y = as_d(w), as_d(y)
        check(lib().pm_fit_stage(self._ctx.handle, C.byref(batch.c), pd(w), pd(y)))
        self._staged = batch
        return batch

    def add_staged(self):
        check(lib().pm_fit_accumulate_staged(self._ctx.handle))

    # ---- multi-GPU (one process per GPU): NCCL communicator owned by the library -----------------------------------
    def comm_init_rank(self, n_ranks, rank, unique_id):
        """Join an NCCL communicator (pm_comm_init_rank); unique_id = the 128 bytes of comm_unique_id() of rank 0."""
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        check(lib().pm_comm_init_rank(self._ctx.handle, int(n_ranks), int(rank), C.c_char_p(bytes(unique_id))))

    def reduce(self, root=0):
        """One ncclReduce(sum) of the packed upper tiles + xe_sum / xe_sq_sum / n_data onto rank `root` (collective)."""
        check(lib().pm_fit_reduce(self._ctx.handle, int(root)))

    def reduce_bytes(self):
        return int(lib().pm_fit_reduce_bytes(self._ctx.handle))

    def barrier(self):
        check(lib().pm_comm_barrier(self._ctx.handle))

    def allreduce(self, values, op="sum"):
        """Small host-value collective (timing: max over ranks).  op: sum | max | min."""
        v = as_d(np.atleast_1d(values)).copy()
        check(lib().pm_comm_allreduce(self._ctx.handle, pd(v), len(v), {"sum": 0, "max": 1, "min": 2}[op]))
        return v

    def accumulator(self):
        """(device pointer, n_doubles) of the packed accumulator, for a cross-GPU reduction."""
        p, n = C.c_void_p(), C.c_size_t(0)
        check(lib().pm_fit_accumulator(self._ctx.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def finalize(self, want_xtx=True, copy=True):
        """X^T X (F, F), X^T y, energy-row sums, y^T y and the row count on the host.

        copy=False returns read-only views of the context's pinned result buffer (no extra host copy of the
        33 MB matrix); they are overwritten by the next finalize() of this accumulator."""
        F = self.n_features
        if not copy:
            p, n = C.POINTER(C.c_double)(), C.c_size_t(0)
            check(lib().pm_fit_finalize_view(self._ctx.handle, C.byref(p), C.byref(n)))
            flat = np.ctypeslib.as_array(p, shape=(n.value,))
            flat.flags.writeable = False
            tail = flat[F * F:]
            return {"xtx": flat[:F * F].reshape(F, F) if want_xtx else None, "xty": tail[:F], "xe_sum": tail[F:2 * F],
                    "xe_sq_sum": tail[2 * F:3 * F], "y_sq_norm": float(tail[3 * F]), "total_n_data": int(tail[3 * F + 1])}
        xtx = np.empty((F, F)) if want_xtx else None
        xty, xe_sum, xe_sq = np.zeros(F), np.zeros(F), np.zeros(F)
        ysq, nd = C.c_double(0.0), C.c_int64(0)
        check(lib().pm_fit_finalize(self._ctx.handle, pd(xtx), pd(xty), pd(xe_sum), pd(xe_sq), C.byref(ysq), C.byref(nd)))
        return {"xtx": xtx, "xty": xty, "xe_sum": xe_sum, "xe_sq_sum": xe_sq, "y_sq_norm": ysq.value,
                "total_n_data": int(nd.value)}

    def solve_ridge(self, alphas, n_energy, scales=None, include_force=True, scale_threshold=1e-10):
        """Ridge solve on the device-resident accumulator (cuSOLVER Cholesky per alpha).
        Returns scales (F,), coefs_array (F, n_alpha) in the scaled basis, rmse (n_alpha,)."""
        F = self.n_features
        al = as_d(alphas)
        sc_in = as_d(scales) if scales is not None else None
        sc_out, coefs, rmse = np.zeros(F), np.zeros((len(al), F)), np.zeros(len(al))
        check(lib().pm_fit_solve_ridge(self._ctx.handle, pd(al), len(al), pd(sc_in), C.c_int64(int(n_energy)),
                                       int(bool(include_force)), C.c_double(scale_threshold), pd(sc_out), pd(coefs),
                                       pd(rmse)))
        return sc_out, coefs.T.copy(), rmse


def comm_unique_id():
    """128-byte NCCL unique id (pm_comm_unique_id); rank 0 creates it, every rank passes it to comm_init_rank."""
    buf = C.create_string_buffer(128)
    check(lib().pm_comm_unique_id(buf))
    return buf.raw


class PotentialXtXMulti:
    """Fused feature + X^T X accumulation over several GPUs of ONE process (C ABI: pm_multi_*): a batch is sharded over
    the devices, each accumulates its part on its own host thread, finalize() sums the partial accumulators with one
    grouped ncclReduce onto the first device.  Same add() / finalize() contract as PotentialXtX."""

    def __init__(self, params_dict, devices, workspace_bytes=0, flags=0):
        self._model = _Model(params_dict)
        self.n_features = self._model.n_features
        self.devices = [int(d) for d in devices]
        dev = as_i(self.devices)
        h = C.c_void_p()
        check(lib().pm_multi_create(self._model.handle, pi(dev), len(dev), C.c_size_t(workspace_bytes), int(flags), C.byref(h)))
        self._h = h
        check(lib().pm_multi_fit_reset(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pm_multi_destroy(self._h)
            self._h = None

    def reset(self):
        check(lib().pm_multi_fit_reset(self._h))

    def add(self, axis, positions_c, types, force_flags, w, y):
        batch = StructureBatch(axis, positions_c, types, force_flags)
        w, y = as_d(w), as_d(y)
        if len(w) != batch.n_rows or len(y) != batch.n_rows:
            raise ValueError("w and y must have one entry per row of the batch")
        check(lib().pm_multi_fit_accumulate(self._h, C.byref(batch.c), pd(w), pd(y)))
        return batch

    def finalize(self):
        F = self.n_features
        xtx, xty, xe_sum, xe_sq = np.empty((F, F)), np.zeros(F), np.zeros(F), np.zeros(F)
        ysq, nd = C.c_double(0.0), C.c_int64(0)
        check(lib().pm_multi_fit_finalize(self._h, pd(xtx), pd(xty), pd(xe_sum), pd(xe_sq), C.byref(ysq), C.byref(nd)))
        return {"xtx": xtx, "xty": xty, "xe_sum": xe_sum, "xe_sq_sum": xe_sq, "y_sq_norm": ysq.value,
                "total_n_data": int(nd.value)}


class PotentialPropertiesFast:
    """PotentialPropertiesFast(params_dict, coeffs) (reference: pybind11_mlp.cpp:51-67,
    compute/py_properties_fast.cpp:10-76).  Energies in eV/cell, forces (N, 3) in eV/A,
    stress as 6 virial components (xx, yy, zz, xy, yz, zx) in eV/cell."""

    def __init__(self, params_dict, coeffs, device=None, flags=0):
        self._model = _Model(params_dict)
        self._ctx = _Context(self._model, device, flags=flags)
        c = as_d(coeffs)
        check(lib().pm_eval_set_coeffs(self._ctx.handle, pd(c), len(c)))
        self._e = self._f = self._s = None
        self._e_array = self._f_array = self._s_array = None

    def _run(self, axis_array, positions_c_array, types_array):
        batch = StructureBatch(axis_array, positions_c_array, types_array, [True] * len(axis_array))
        e = np.zeros(batch.n_st)
        f = np.zeros((int(batch.n_atoms.sum()), 3))
        s = np.zeros((batch.n_st, 6))
        check(lib().pm_eval(self._ctx.handle, C.byref(batch.c), pd(e), pd(f), pd(s)))
        offs = np.concatenate([[0], np.cumsum(batch.n_atoms)])
        return e, [f[offs[k]:offs[k + 1]] for k in range(batch.n_st)], s

    def eval(self, axis, positions_c, types, use_openmp=True):
        e, f, s = self._run([axis], [positions_c], [types])
        self._e, self._f, self._s = float(e[0]), f[0], s[0]

    def eval_multiple(self, axis_array, positions_c_array, types_array):
        self._e_array, self._f_array, self._s_array = self._run(axis_array, positions_c_array, types_array)

    def get_e(self):
        return self._e

    def get_f(self):
        return self._f

    def get_s(self):
        return self._s

    def get_e_array(self):
        return self._e_array

    def get_f_array(self):
        return self._f_array

    def get_s_array(self):
        return self._s_array


def _pair_basis_record(params_dict, dvec):
    """Pair-basis record (device kernel k_pair_basis, K2a) of the single pair 0 -> 1 of a two-atom cell whose lattice
    is far larger than the cutoff: items dx dy dz 1/r | f_n | f_n' | Y (re, im) | dY/dx | dY/dy | dY/dz."""
    model = _Model(params_dict)
    ctx = _Context(model)
    rc = float(params_dict["model"]["cutoff"])
    axis = np.eye(3) * (4.0 * rc + 10.0)
    pos = np.zeros((3, 2))
    pos[:, 1] = dvec
    st = StructureBatch([axis], [pos], [np.zeros(2, np.int32)], [True])
    x = np.zeros((int(lib().pm_batch_rows(C.byref(st.c))), model.n_features))
    check(lib().pm_features_x(ctx.handle, C.byref(st.c), pd(x)))
    raw = ctx.debug_fetch(2)
    if raw.size == 0:
        raise ValueError("the pair is not inside the cutoff")
    n_fn = len(params_dict["model"]["pair_params"])
    L = int(params_dict["model"]["max_l"])
    nh = (L + 1) * (L + 2) // 2
    stride = 4 + 2 * n_fn + 8 * nh
    return raw.reshape(-1, stride, 32)[0, :, 0], n_fn, nh     # pair 0 of block 0 = (atom 0 -> atom 1)


def get_fn(dis, params, cutoff):
    """Test hook of the reference (pybind11_mlp.cpp:158-167, get_fn_): Gaussian x cosine-cutoff radial functions and
    their derivatives at distance `dis`, computed by the device pair-basis kernel."""
    from .params import make_params_dict

    pd_ = make_params_dict(n_type=1, cutoff=cutoff, model_type=2, max_p=1, gtinv_order=0, gtinv_maxl=[],
                           pair_params=[list(map(float, p)) for p in params], feature_type="pair")
    rec, n_fn, _ = _pair_basis_record(pd_, np.array([0.0, 0.0, float(dis)]))
    return rec[4:4 + n_fn].copy(), rec[4 + n_fn:4 + 2 * n_fn].copy()


def get_ylm(x, y, z, lmax, r=1.0):
    """Test hook of the reference (pybind11_mlp.cpp:169-181, get_ylm_): Y_lm(m <= 0) of the direction (x, y, z) / r and
    its Cartesian gradient at distance r, computed by the device pair-basis kernel."""
    from .params import make_params_dict

    pd_ = make_params_dict(n_type=1, cutoff=max(6.0, 2.0 * r), model_type=2, max_p=1, gtinv_order=2, gtinv_maxl=[lmax],
                           pair_params=[[0.0, 0.0]])
    rec, n_fn, nh = _pair_basis_record(pd_, np.array([x, y, z], dtype=float))
    o = 4 + 2 * n_fn
    out = [rec[o + k * 2 * nh:o + (k + 1) * 2 * nh].copy().view(np.complex128) for k in range(4)]
    return tuple(out)

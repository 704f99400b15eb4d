"""ctypes wrapper around oracle/_ref/libpolymlp_ref.so (the UNMODIFIED reference C++).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(pypolymlp_b200) never imports this module.

The library is built by `make -C oracle ref` from the sources under
/root/reference (see oracle/Makefile); the built .so and the reference's gtinv
.bin tables live in oracle/_ref/ (git-ignored, shipped to the GPU box).
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libpolymlp_ref.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                "oracle/_ref/libpolymlp_ref.so is missing: run `make -C oracle ref` "
                "in the build container (needs /root/reference)."
            )
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_model_create.restype = C.c_void_p
        _lib.ref_eval_create.restype = C.c_void_p
        _lib.ref_model_n_rows.restype = C.c_long
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference oracle: " + lib().ref_last_error().decode())


def num_threads() -> int:
    return lib().ref_num_threads()


def set_num_threads(n: int) -> None:
    """OpenMP threads of the reference library (torchrun exports OMP_NUM_THREADS=1)."""
    L = lib()
    if hasattr(L, "ref_set_num_threads"):
        L.ref_set_num_threads(int(n))


def make_params(n_type, cutoff, model_type, max_p, gtinv_order, gtinv_maxl, n_gaussians=None, pair_params=None,
                gtinv_version=1, feature_type="gtinv", **_ignored):
    """The subset of PolymlpParams.as_dict() that RefModel / RefEval read (the reference library calls its own Readgtinv,
    so no coupling-coefficient tables are needed here).  Radial parameters as in src/pypolymlp/core/params_utils.py:84-94.
    Exists so that bench.py's reference arm does not have to import the product package."""
    if pair_params is None:
        g2 = np.linspace(0.0, 1.0, n_gaussians - 1) * (cutoff - 1.0)
        width = max(1.0, g2[1] - g2[0])
        pair_params = [[float(width), float(p2)] for p2 in g2] + [[0.0, 0.0]]
    cond = {(i, j): list(range(len(pair_params))) for i in range(n_type) for j in range(i, n_type)}
    return {"n_type": int(n_type),
            "model": {"cutoff": float(cutoff), "feature_type": feature_type, "model_type": int(model_type),
                      "max_p": int(max_p), "max_l": int(max(gtinv_maxl)) if len(gtinv_maxl) else 0,
                      "pair_params": pair_params, "pair_params_conditional": cond,
                      "gtinv": {"order": int(gtinv_order), "max_l": list(gtinv_maxl), "version": int(gtinv_version)}}}


def _cond_arrays(params_dict):
    n_type = params_dict["n_type"]
    model = params_dict["model"]
    n_fn = len(model["pair_params"])
    cond = model.get("pair_params_conditional")
    offsets, values = [0], []
    for i in range(n_type):
        for j in range(i, n_type):
            lst = list(cond[(i, j)]) if cond else list(range(n_fn))
            values.extend(lst)
            offsets.append(len(values))
    return _i(offsets), _i(values)


def _fp_args(params_dict):
    model = params_dict["model"]
    g = model["gtinv"] if model["feature_type"] != "pair" else {"order": 0, "max_l": []}
    params = _d(model["pair_params"]).reshape(-1, 2)
    off, val = _cond_arrays(params_dict)
    maxl = _i(list(g["max_l"]) + [0])
    keep = (params, off, val, maxl)
    args = (
        C.c_int(params_dict["n_type"]), C.c_int(params.shape[0]), _pd(params),
        _pi(off), _pi(val), C.c_double(model["cutoff"]), C.c_int(model["model_type"]),
        C.c_int(model["max_p"]), C.c_int(model["max_l"]), C.c_int(g["order"]),
        _pi(maxl), C.c_int(g.get("version", 1)),
    )
    return args, keep


def readgtinv(order, maxl, version=1):
    """Reference Readgtinv(order, maxl, version) -> (l_comb, lm_seq, lm_coeffs) nested lists."""
    L = lib()
    ml = _i(list(maxl) + [0])
    sizes = (C.c_long * 4)()
    _check(L.ref_readgtinv(order, _pi(ml), version, sizes, None, None, None, None, None))
    n, s1, s2, s3 = (int(x) for x in sizes)
    lo, lc, nt = np.zeros(n, np.int32), np.zeros(s1, np.int32), np.zeros(n, np.int32)
    lm, cf = np.zeros(s3, np.int32), np.zeros(s2, np.float64)
    _check(L.ref_readgtinv(order, _pi(ml), version, sizes, _pi(lo), _pi(lc), _pi(nt), _pi(lm), _pd(cf)))
    l_comb, lm_seq, lm_coeffs = [], [], []
    p1 = p2 = p3 = 0
    for i in range(n):
        o = int(lo[i])
        l_comb.append([int(x) for x in lc[p1:p1 + o]])
        p1 += o
        t = int(nt[i])
        lm_coeffs.append([float(x) for x in cf[p2:p2 + t]])
        p2 += t
        lm_seq.append(lm[p3:p3 + t * o].reshape(t, o).tolist())
        p3 += t * o
    return l_comb, lm_seq, lm_coeffs


def get_fn(dis, cutoff, pair_params):
    p = _d(pair_params).reshape(-1, 2)
    fn, fnd = np.zeros(len(p)), np.zeros(len(p))
    lib().ref_get_fn(C.c_double(dis), C.c_double(cutoff), len(p), _pd(p), _pd(fn), _pd(fnd))
    return fn, fnd


def get_ylm(r, x, y, z, lmax):
    n = (lmax + 1) * (lmax + 2) // 2
    out = [np.zeros(2 * n) for _ in range(4)]
    lib().ref_get_ylm(C.c_double(r), C.c_double(x), C.c_double(y), C.c_double(z), lmax,
                      *[_pd(o) for o in out])
    return tuple(o.view(np.complex128) for o in out)


def neighbor(kind, axis, positions_c, cutoff):
    """kind: 'full' | 'half' | 'half_full'. Returns offsets, neigh, dx, dy, dz."""
    k = {"full": 0, "half": 1, "half_full": 2}[kind]
    axis, pos = _d(axis), _d(positions_c)
    n = pos.shape[1]
    off = np.zeros(n + 1, np.int32)
    _check(lib().ref_neighbor(k, _pd(axis), _pd(pos), n, C.c_double(cutoff), _pi(off), None, None, None, None))
    P = int(off[-1])
    nb, dx, dy, dz = np.zeros(P, np.int32), np.zeros(P), np.zeros(P), np.zeros(P)
    _check(lib().ref_neighbor(k, _pd(axis), _pd(pos), n, C.c_double(cutoff), _pi(off), _pi(nb), _pd(dx), _pd(dy), _pd(dz)))
    return off, nb, dx, dy, dz


def neighbor_cell(axis, positions_c, cutoff, max_trans=100000):
    axis, pos = _d(axis), _d(positions_c)
    n = pos.shape[1]
    nt = C.c_int(0)
    trans = np.zeros((max_trans, 3))
    ax, po = np.zeros((3, 3)), np.zeros((3, n))
    lib().ref_neighbor_cell(_pd(axis), _pd(pos), n, C.c_double(cutoff), C.byref(nt), _pd(trans),
                            max_trans, _pd(ax), _pd(po))
    return trans[: nt.value].copy(), ax, po


class RefModel:
    """Reference `Model` (compute/model.cpp) for one feature_params."""

    def __init__(self, params_dict):
        args, self._keep = _fp_args(params_dict)
        self._h = lib().ref_model_create(*args)
        if not self._h:
            raise RuntimeError("reference oracle: " + lib().ref_last_error().decode())
        self._h = C.c_void_p(self._h)
        self.n_type = params_dict["n_type"]
        self.n_features = lib().ref_model_n_features(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_model_destroy(self._h)
            self._h = None

    def table_sizes(self, t):
        out = (C.c_long * 7)()
        lib().ref_model_table_sizes(self._h, t, out)
        keys = ["n_nlmtp", "n_noconj", "n_linear", "n_prod", "n_prod_deriv", "n_deriv_terms", "n_poly"]
        return dict(zip(keys, (int(x) for x in out)))

    def polynomial(self, t):
        n = lib().ref_model_polynomial(self._h, t, None, None)
        g, l3 = np.zeros(n, np.int32), np.zeros((n, 3), np.int32)
        lib().ref_model_polynomial(self._h, t, _pi(g), _pi(l3))
        return g, l3

    def feature_attrs(self):
        """(radial_ids, gtinv_ids, tcomb_ids, polynomial_ids, type_pairs) of the reference's FeaturesAttr."""
        L = lib()
        L.ref_model_feature_attrs.restype = C.c_long
        n = L.ref_model_feature_attrs(self._h, None)
        v = np.zeros(n, np.int32)
        L.ref_model_feature_attrs(self._h, _pi(v))
        v = v.tolist()
        pos, radial, gtinv, tcomb, poly = 1, [], [], [], []
        for _ in range(v[0]):
            radial.append(v[pos])
            if v[pos + 1] >= 0:
                gtinv.append(v[pos + 1])
            tcomb.append(v[pos + 3:pos + 3 + v[pos + 2]])
            pos += 3 + v[pos + 2]
        n_poly, pos = v[pos], pos + 1
        for _ in range(n_poly):
            poly.append(v[pos + 1:pos + 1 + v[pos]])
            pos += 1 + v[pos]
        nt, pos = v[pos], pos + 1
        return radial, gtinv, tcomb, poly, [v[pos + i * nt:pos + (i + 1) * nt] for i in range(nt)]

    def run(self, axis, positions_c, types, force=True):
        axis, pos, ty = _d(axis), _d(positions_c), _i(types)
        n, F = pos.shape[1], self.n_features
        xe = np.zeros(F)
        xf = np.zeros((3 * n, F)) if force else np.zeros((0, F))
        xs = np.zeros((6, F)) if force else np.zeros((0, F))
        _check(lib().ref_model_run(self._h, _pd(axis), _pd(pos), _pi(ty), n, int(force), _pd(xe), _pd(xf), _pd(xs)))
        return xe, xf, xs

    def atom(self, axis, positions_c, types, atom):
        axis, pos, ty = _d(axis), _d(positions_c), _i(types)
        n = pos.shape[1]
        a = np.zeros(2 * 200000)
        d = np.zeros(200000)
        na, nd = C.c_int(0), C.c_int(0)
        _check(lib().ref_model_atom(self._h, _pd(axis), _pd(pos), _pi(ty), n, atom, _pd(a), C.byref(na), _pd(d), C.byref(nd)))
        return a[: 2 * na.value].view(np.complex128).copy(), d[: nd.value].copy()

    def build_x(self, axis_list, positions_c_list, types_list, force_st, n_threads=0):
        """PyModel-layout X (row-major) for a batch: energies | stress | forces."""
        n_st = len(axis_list)
        axes = _d(np.array(axis_list)).reshape(n_st, 9)
        n_atoms = _i([np.asarray(p).shape[1] for p in positions_c_list])
        pos = _d(np.concatenate([_d(p).reshape(-1) for p in positions_c_list]))
        ty = _i(np.concatenate([_i(t) for t in types_list]))
        fs = _i([int(bool(f)) for f in force_st])
        rows = lib().ref_model_n_rows(n_st, _pi(n_atoms), _pi(fs))
        x = np.zeros((rows, self.n_features))
        _check(lib().ref_model_build_x(self._h, n_st, _pd(axes), _pd(pos), _pi(ty), _pi(n_atoms), _pi(fs), n_threads, _pd(x)))
        return x


class RefEval:
    """Reference `PolymlpEval` (compute/polymlp_eval.cpp) for one model + coefficients."""

    def __init__(self, params_dict, coeffs):
        args, self._keep = _fp_args(params_dict)
        c = _d(coeffs)
        self._h = lib().ref_eval_create(*args, _pd(c), len(c))
        if not self._h:
            raise RuntimeError("reference oracle: " + lib().ref_last_error().decode())
        self._h = C.c_void_p(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_eval_destroy(self._h)
            self._h = None

    def eval_multiple(self, axis_list, positions_c_list, types_list):
        """PyPropertiesFast::eval_multiple semantics (compute/py_properties_fast.cpp:30-76): serial over structures,
        OpenMP over atoms inside each."""
        return [self.eval(a, p, t, use_openmp=True) for a, p, t in zip(axis_list, positions_c_list, types_list)]

    def eval(self, axis, positions_c, types, use_openmp=False):
        axis, pos, ty = _d(axis), _d(positions_c), _i(types)
        n = pos.shape[1]
        e = C.c_double(0.0)
        f, s = np.zeros((n, 3)), np.zeros(6)
        _check(lib().ref_eval(self._h, _pd(axis), _pd(pos), _pi(ty), n, int(use_openmp), C.byref(e), _pd(f), _pd(s)))
        return e.value, f, s

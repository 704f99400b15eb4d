"""numpy restatement of pypolymlp's gtinv hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product (pypolymlp_b200) never
does.  Parity status: PINNED -- checked against the reference's golden vectors
(tests/test_oracle_golden.py: get_fn, Y_lm sums, translation counts, Si design
matrix column sums, the MgO potentials' published E / F / stress and feature
sums, the PolymlpAPI sums of tests/test_cxx/test_polymlp_api.py, the skewed
BiGd2 neighbour lists, the 25-cell shape-invariance energy of
tests/test_calc/test_check_neighbors.py; tests/test_legacy_io.py: the published
answers of the bundled legacy SrTiO3 / Ag / MgO potentials) and against the
unmodified reference C++ compiled into oracle/_ref (tests/test_oracle_vs_ref.py).

The restatement is deliberately written in a different algebraic form from the
reference (neighbour-sparse derivatives through G = d(feature)/d(a_nlm)), so
that agreement with oracle/_ref is a real check and not a transcription.
Every function cites the reference file:line whose behaviour it restates
(paths relative to /root/reference/src/pypolymlp/cxx/src unless noted).
"""

import itertools
import struct

import numpy as np

# --------------------------------------------------------------------------
# gtinv coupling-coefficient tables  (polymlp/polymlp_gtinv_binary.cpp:11-90,
# polymlp/polymlp_gtinv_data.cpp:35-66, polymlp/polymlp_read_gtinv.cpp:23-58)
# --------------------------------------------------------------------------


def read_gtinv_bin(path):
    """Parse one reference `.bin` table -> (l_array, coeffs, m_array)."""
    with open(path, "rb") as f:
        buf = f.read()
    pos = [4]  # skip magic "DATA"

    def i32():
        v = struct.unpack_from("<i", buf, pos[0])[0]
        pos[0] += 4
        return v

    def f64():
        v = struct.unpack_from("<d", buf, pos[0])[0]
        pos[0] += 8
        return v

    i32()  # number of blocks
    i32()  # block type
    l_array = [[i32() for _ in range(i32())] for _ in range(i32())]
    i32()
    coeffs = [[f64() for _ in range(i32())] for _ in range(i32())]
    i32()
    m_array = [[[i32() for _ in range(i32())] for _ in range(i32())] for _ in range(i32())]
    return l_array, coeffs, m_array


def readgtinv(order, maxl, datadir, version=1):
    """Readgtinv::screening (polymlp_read_gtinv.cpp:23-58) -> l_comb, lm_seq, lm_coeffs."""
    l_comb, lm_seq, lm_coeffs = [], [], []
    for o in range(1, order + 1):
        la, cf, ma = read_gtinv_bin(f"{datadir}/polymlp_gtinv_data_v{version}_order{o}.bin")
        ml = maxl[o - 2] if o > 1 else 0
        for i, lc in enumerate(la):
            if ml < lc[-1]:
                continue
            l_comb.append(list(lc))
            lm_seq.append([[l * l + l + m for l, m in zip(lc, mc)] for mc in ma[i]])
            lm_coeffs.append(list(cf[i]))
    return l_comb, lm_seq, lm_coeffs


# --------------------------------------------------------------------------
# radial and angular functions
# --------------------------------------------------------------------------

PI = 3.1415926535897932384626433832795  # polymlp/polymlp_basis_function.h:14


def radial(r, params, cutoff):
    """get_fn_ (polymlp_functions_interface.cpp:41-62) + basis_function.h:16-42.

    r: (P,), params: (n_fn, 2) of (beta, mu).  Returns fn, fn_d of shape (P, n_fn).
    """
    r = np.asarray(r, dtype=np.float64)
    inside = r < cutoff
    v1 = PI / cutoff
    v2 = v1 * r
    fc = np.where(inside, 0.5 * (np.cos(v2) + 1.0), 0.0)
    fcd = np.where(inside, -0.5 * v1 * np.sin(v2), 0.0)
    beta, mu = params[:, 0][None, :], params[:, 1][None, :]
    dx = r[:, None] - mu
    bf = np.exp(-beta * dx * dx)
    bfd = -2.0 * beta * dx * bf
    return bf * fc[:, None], bfd * fc[:, None] + bf * fcd[:, None]


def _lm2i(l, m):
    return l * (l + 1) // 2 + m


def _legendre(ct, lmax):
    """normalized_associated_legendre with q = P/sin(theta)
    (polymlp_spherical_harmonics.cpp:162-209).  Returns p, q of shape (n_half, P)."""
    nh = (lmax + 1) * (lmax + 2) // 2
    P = ct.shape[0]
    p = np.zeros((nh, P))
    q = np.zeros((nh, P))
    s2pi = 0.39894228040143267794
    p[0] = s2pi
    if lmax == 0:
        return p, q
    st = np.sqrt(1.0 - ct * ct)
    p[_lm2i(1, 0)] = ct * 1.7320508075688772935 * s2pi
    p[_lm2i(1, 1)] = -st * 1.2247448713915890491 * s2pi
    q[_lm2i(1, 1)] = -1.2247448713915890491 * s2pi
    for l in range(2, lmax + 1):
        c1 = -np.sqrt(1.0 + 0.5 / l) * st
        p[_lm2i(l, l)] = c1 * p[_lm2i(l - 1, l - 1)]
        q[_lm2i(l, l)] = c1 * q[_lm2i(l - 1, l - 1)]
        c2 = np.sqrt(2.0 * (l - 1.0) + 3.0) * ct
        p[_lm2i(l, l - 1)] = c2 * p[_lm2i(l - 1, l - 1)]
        q[_lm2i(l, l - 1)] = c2 * q[_lm2i(l - 1, l - 1)]
    for l in range(2, lmax + 1):
        ls, lm1s = float(l * l), float((l - 1) * (l - 1))
        for m in range(0, l - 1):
            ms = float(m * m)
            alm = np.sqrt((4.0 * ls - 1.0) / (ls - ms))
            blm = -np.sqrt((lm1s - ms) / (4.0 * lm1s - 1.0))
            p[_lm2i(l, m)] = alm * (ct * p[_lm2i(l - 1, m)] + blm * p[_lm2i(l - 2, m)])
            q[_lm2i(l, m)] = alm * (ct * q[_lm2i(l - 1, m)] + blm * q[_lm2i(l - 2, m)])
    return p, q


def ylm_der(dx, dy, dz, lmax):
    """get_ylm_ + SphericalHarmonics::compute_ylm_der
    (polymlp_functions_interface.cpp:114-142, polymlp_spherical_harmonics.cpp:60-121).

    Returns ylm, ylm_dx, ylm_dy, ylm_dz, complex arrays of shape (n_half, P); only m <= 0
    is stored, at key l(l+3)/2 + m."""
    dx, dy, dz = (np.asarray(v, dtype=np.float64) for v in (dx, dy, dz))
    r = np.sqrt(dx * dx + dy * dy + dz * dz)
    ct = dz / r
    rho = np.hypot(dx, dy)
    safe = np.where(rho > 0.0, rho, 1.0)
    cp = np.where(rho > 0.0, dx / safe, 1.0)
    sp = np.where(rho > 0.0, dy / safe, 0.0)
    nh = (lmax + 1) * (lmax + 2) // 2
    P = r.shape[0]
    p, q = _legendre(ct, lmax)
    st = np.sqrt(1.0 - ct * ct)
    invr = 1.0 / r
    hs2 = 0.5 * np.sqrt(2.0)
    y = np.zeros((nh, P), complex)
    yx, yy, yz = np.zeros_like(y), np.zeros_like(y), np.zeros_like(y)
    for l in range(lmax + 1):
        idx = _lm2i(l, 0) + l
        y[idx] = p[_lm2i(l, 0)] * hs2
        if l >= 1:
            common = q[_lm2i(l, 1)] * st * invr * np.sqrt(0.5 * l * (l + 1))
            yx[idx] = common * ct * cp
            yy[idx] = common * ct * sp
            yz[idx] = -common * st
    c1, c2 = np.ones(P), cp.copy()
    s1, s2 = np.zeros(P), -sp
    tc = 2.0 * c2
    sign = -1.0
    for mp in range(1, lmax + 1):
        s = tc * s1 - s2
        c = tc * c1 - c2
        c2, c1, s2, s1 = c1, c, s1, s
        eim = c + 1j * s
        for l in range(mp, lmax + 1):
            idx = _lm2i(l, -mp) + l
            y[idx] = (sign * p[_lm2i(l, mp)] * hs2) * (c - 1j * s)
            common = eim * hs2 * invr
            dth = mp * ct * q[_lm2i(l, mp)]
            if mp != l:
                dth = dth + np.sqrt((l - mp) * (l + mp + 1)) * q[_lm2i(l, mp + 1)] * st
            dphi = 1j * (mp * q[_lm2i(l, mp)])
            yx[idx] = sign * np.conj(common * (dth * ct * cp - dphi * sp))
            yy[idx] = sign * np.conj(common * (dth * ct * sp + dphi * cp))
            yz[idx] = sign * np.conj(-common * dth * st)
        sign = -sign
    return y, yx, yy, yz


# --------------------------------------------------------------------------
# neighbour lists
# --------------------------------------------------------------------------


class NeighborCell:
    """compute/neighbor_cell.cpp:11-19,24-44,126-168,203-256 (bit-exact restatement:
    every expression keeps the reference's left-to-right operation order)."""

    def __init__(self, axis, positions_c, cutoff):
        self.axis = [[float(axis[i][j]) for j in range(3)] for i in range(3)]
        self.pos = np.array(positions_c, dtype=np.float64).copy()
        self.cutoff = float(cutoff)
        self._metric()
        self._find_trans()

    def _col(self, c):
        return [self.axis[0][c], self.axis[1][c], self.axis[2][c]]

    @staticmethod
    def _dot(a, b):
        return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]

    def _metric(self):
        a0, a1, a2 = self._col(0), self._col(1), self._col(2)
        d = self._dot
        self.d00, self.d11, self.d22 = d(a0, a0), d(a1, a1), d(a2, a2)
        self.d01, self.d02, self.d12 = d(a0, a1), d(a0, a2), d(a1, a2)
        # NB: neighbor_cell.cpp:36-38 writes `abs(dot01)` where only ::abs(int) is visible,
        # so the dot products are truncated to int before the comparison.  Kept on purpose.
        ia = lambda x: abs(int(x))
        self.r01 = ia(self.d01) > 0.5 * self.d00 or ia(self.d01) > 0.5 * self.d11
        self.r02 = ia(self.d02) > 0.5 * self.d00 or ia(self.d02) > 0.5 * self.d22
        self.r12 = ia(self.d12) > 0.5 * self.d11 or ia(self.d12) > 0.5 * self.d22
        self.ref_sum = int(self.r01) + int(self.r02) + int(self.r12)

    def _cart(self, i, j, k):
        a = self.axis
        return [a[r][0] * i + a[r][1] * j + a[r][2] * k for r in range(3)]

    def _dist(self, i, j, k):
        v = self._cart(i, j, k)
        return np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])

    def _replace(self, vec, col):
        a0 = self._cart(*vec)
        for i in range(3):
            self.axis[i][col] = a0[i]
        self._metric()

    @staticmethod
    def _round(x):  # C round(): half away from zero
        return int(np.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))

    def _refine(self):
        it = 0
        while self.ref_sum > 0 and it < 100:
            if self.r01:
                if self.d00 > self.d11:
                    i = self._round(-(self.d01 * 1 + self.d12 * 0) / self.d11)
                    self._replace([1, i, 0], 0)
                else:
                    i = self._round(-(self.d01 * 1 + self.d02 * 0) / self.d00)
                    self._replace([i, 1, 0], 1)
            if self.r02:
                if self.d00 > self.d22:
                    i = self._round(-(self.d02 * 1 + self.d12 * 0) / self.d22)
                    self._replace([1, 0, i], 0)
                else:
                    i = self._round(-(self.d01 * 0 + self.d02 * 1) / self.d00)
                    self._replace([i, 0, 1], 2)
            if self.r12:
                if self.d11 > self.d22:
                    i = self._round(-(self.d02 * 0 + self.d12 * 1) / self.d22)
                    self._replace([0, 1, i], 1)
                else:
                    i = self._round(-(self.d01 * 0 + self.d12 * 1) / self.d11)
                    self._replace([0, i, 1], 2)
            it += 1

    def _inverse(self):
        a = self.axis
        det = (a[0][0] * a[1][1] * a[2][2] + a[0][1] * a[1][2] * a[2][0] + a[0][2] * a[1][0] * a[2][1]
               - a[0][2] * a[1][1] * a[2][0] - a[0][1] * a[1][0] * a[2][2] - a[0][0] * a[1][2] * a[2][1])
        inv = [[0.0] * 3 for _ in range(3)]
        inv[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1]
        inv[0][1] = -(a[0][1] * a[2][2] - a[0][2] * a[2][1])
        inv[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1]
        inv[1][0] = -(a[1][0] * a[2][2] - a[1][2] * a[2][0])
        inv[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0]
        inv[1][2] = -(a[0][0] * a[1][2] - a[0][2] * a[1][0])
        inv[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0]
        inv[2][1] = -(a[0][0] * a[2][1] - a[0][1] * a[2][0])
        inv[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0]
        return [[inv[i][j] / det for j in range(3)] for i in range(3)]

    def _find_trans(self):
        if self.ref_sum > 0:
            self._refine()
            inv = self._inverse()
            pc = self.pos
            frac = np.empty_like(pc)
            for r in range(3):
                frac[r] = inv[r][0] * pc[0] + inv[r][1] * pc[1] + inv[r][2] * pc[2]
            frac = frac - np.floor(frac)
            a = self.axis
            for r in range(3):
                self.pos[r] = a[r][0] * frac[0] + a[r][1] * frac[1] + a[r][2] * frac[2]
        mx = [int(np.ceil(self.cutoff / self._dist(1, 0, 0)) + 1),
              int(np.ceil(self.cutoff / self._dist(0, 1, 0)) + 1),
              int(np.ceil(self.cutoff / self._dist(0, 0, 1)) + 1)]
        max_len = 0.0
        for i, j, k in itertools.product((-1, 0, 1), repeat=3):
            max_len = max(max_len, self._dist(i, j, k))
        trans = []
        for i in range(-mx[0], mx[0] + 1):
            for j in range(-mx[1], mx[1] + 1):
                for k in range(-mx[2], mx[2] + 1):
                    v = self._cart(i, j, k)
                    if np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) < max_len + self.cutoff:
                        trans.append(v)
        self.trans = np.array(trans, dtype=np.float64).reshape(-1, 3)


def neighbor_full(axis, positions_c, cutoff):
    """NeighborFull (compute/neighbor_full.cpp:10-76): CSR list ordered i, j, translation."""
    nc = NeighborCell(axis, positions_c, cutoff)
    pos, tr = nc.pos, nc.trans
    n = pos.shape[1]
    csq, tsq = cutoff * cutoff, 1e-10 * 1e-10
    off, nb, ddx, ddy, ddz = [0], [], [], [], []
    for i in range(n):
        dxij = pos[0] - pos[0, i]
        dyij = pos[1] - pos[1, i]
        dzij = pos[2] - pos[2, i]
        dx = dxij[:, None] + tr[None, :, 0]
        dy = dyij[:, None] + tr[None, :, 1]
        dz = dzij[:, None] + tr[None, :, 2]
        r2 = dx * dx + dy * dy + dz * dz
        jj, tt = np.nonzero((r2 < csq) & (r2 > tsq))  # row-major: j then translation
        nb.append(jj)
        ddx.append(dx[jj, tt])
        ddy.append(dy[jj, tt])
        ddz.append(dz[jj, tt])
        off.append(off[-1] + len(jj))
    cat = lambda v, dt: np.concatenate(v).astype(dt) if v else np.zeros(0, dt)
    return (np.array(off, np.int32), cat(nb, np.int32), cat(ddx, np.float64),
            cat(ddy, np.float64), cat(ddz, np.float64))


def neighbor_half(axis, positions_c, cutoff):
    """NeighborHalf (compute/neighbor_half.cpp:10-100): j < i all images; j == i only images
    whose translation is lexicographically positive in (z, y, x)."""
    nc = NeighborCell(axis, positions_c, cutoff)
    pos, tr = nc.pos, nc.trans
    n = pos.shape[1]
    tol = 1e-10
    csq, tsq = cutoff * cutoff, tol * tol
    off, nb, ddx, ddy, ddz = [0], [], [], [], []
    r2t = tr[:, 0] * tr[:, 0] + tr[:, 1] * tr[:, 1] + tr[:, 2] * tr[:, 2]
    keep = (tr[:, 2] >= tol) | ((np.abs(tr[:, 2]) < tol) & (tr[:, 1] >= tol)) | (
        (np.abs(tr[:, 2]) < tol) & (np.abs(tr[:, 1]) < tol) & (tr[:, 0] >= tol))
    self_t = np.nonzero((r2t < csq) & (r2t > tsq) & keep)[0]
    for i in range(n):
        dx = (pos[0, :i] - pos[0, i])[:, None] + tr[None, :, 0]
        dy = (pos[1, :i] - pos[1, i])[:, None] + tr[None, :, 1]
        dz = (pos[2, :i] - pos[2, i])[:, None] + tr[None, :, 2]
        r2 = dx * dx + dy * dy + dz * dz
        jj, tt = np.nonzero((r2 < csq) & (r2 > tsq))
        nb.append(np.concatenate([jj, np.full(len(self_t), i)]))
        ddx.append(np.concatenate([dx[jj, tt], tr[self_t, 0]]))
        ddy.append(np.concatenate([dy[jj, tt], tr[self_t, 1]]))
        ddz.append(np.concatenate([dz[jj, tt], tr[self_t, 2]]))
        off.append(off[-1] + len(nb[-1]))
    cat = lambda v, dt: np.concatenate(v).astype(dt) if v else np.zeros(0, dt)
    return (np.array(off, np.int32), cat(nb, np.int32), cat(ddx, np.float64),
            cat(ddy, np.float64), cat(ddz, np.float64))


# --------------------------------------------------------------------------
# model tables
# --------------------------------------------------------------------------


class Tables:
    """Index tables of a gtinv polymlp.

    Restates Mapping (polymlp/polymlp_mapping.cpp:36-279), uniq_gtinv_type
    (polymlp_model_params_gtinv.cpp:53-151), ModelParamsPoly
    (polymlp_model_params_polynomial.cpp:58-110,185-251), set_linear_features_gtinv
    (polymlp_features_utils.cpp:48-76) and FeaturesPoly::set_polynomial
    (polymlp_features_polynomial.cpp:22-58)."""

    def __init__(self, params_dict):
        model = params_dict["model"]
        self.feature_type = model["feature_type"]
        if self.feature_type not in ("gtinv", "pair"):
            raise ValueError("feature_type must be 'gtinv' or 'pair'")
        self.n_type = nt = params_dict["n_type"]
        self.cutoff = float(model["cutoff"])
        self.maxl = int(model["max_l"])
        self.model_type = int(model["model_type"])
        self.maxp = int(model["max_p"])
        self.params = np.array(model["pair_params"], dtype=np.float64).reshape(-1, 2)
        self.n_fn = n_fn = len(self.params)
        cond = model.get("pair_params_conditional")
        if self.feature_type == "pair":
            # pair features d_{n,tp} = sum_j f_n(r_ij) (compute/local_pair.cpp); the table code below only needs a
            # single order-1 entry per (n, tp) -- its lm / coefficient content is not used by the pair numerics
            if self.model_type > 2:
                raise ValueError("Polymlp: Model type error.")  # polymlp_model_params_polynomial.cpp:42-45
            self.maxl = 0
            l_comb, lm_seq, lm_coeffs = [[0]], [[[0]]], [[1.0]]
        else:
            g = model["gtinv"]
            l_comb, lm_seq, lm_coeffs = g["l_comb"], g["lm_seq"], g["lm_coeffs"]

        # type pairs (polymlp_mapping.cpp:36-56)
        self.type_pairs = np.zeros((nt, nt), int)
        self.tp_to_types, self.tp_to_n = [], []
        for i in range(nt):
            for j in range(i, nt):
                self.type_pairs[i, j] = self.type_pairs[j, i] = len(self.tp_to_types)
                self.tp_to_types.append((i, j))
                self.tp_to_n.append(list(cond[(i, j)]) if cond else list(range(n_fn)))
        self.n_tp = ntp = len(self.tp_to_types)
        self.tp_params = [self.params[ns] for ns in self.tp_to_n]
        n_to_tp = [[tp for tp in range(ntp) if n in self.tp_to_n[tp]] for n in range(n_fn)]
        self.tpn_to_nid = [{n: k for k, n in enumerate(self.tp_to_n[tp])} for tp in range(ntp)]

        # lm attributes (polymlp_mapping.cpp:248-279)
        self.lm = []  # (l, m, ylm_key, conj, cc)
        for l in range(self.maxl + 1):
            for m in range(-l, l + 1):
                key = (l + 3) * l // 2 + m if m < 1 else (l + 3) * l // 2 - m
                self.lm.append((l, m, key, m > 0, 1.0 if m % 2 == 0 else -1.0))
        n_lm = len(self.lm)

        # global (n, lm, tp) list: n-major, lm, tp (polymlp_mapping.cpp:136-168)
        self.g_attr, self.g_index, self.g_conj = [], {}, []
        for n in range(n_fn):
            for lm in range(n_lm):
                sub = 2 * self.lm[lm][1] * len(n_to_tp[n])
                for tp in n_to_tp[n]:
                    gid = len(self.g_attr)
                    self.g_attr.append((n, lm, tp))
                    self.g_index[(n, lm, tp)] = gid
                    self.g_conj.append(gid - sub)

        # local lists per centre type (polymlp_mapping.cpp:170-208)
        self.local = []
        for t in range(nt):
            gids = [gid for gid, (n, lm, tp) in enumerate(self.g_attr) if t in self.tp_to_types[tp]]
            g2l = {gid: k for k, gid in enumerate(gids)}
            self.local.append({"gids": gids, "g2l": g2l})

        # linear terms (polymlp_model_params_gtinv.cpp:92-151)
        order_max = len(l_comb[-1])
        tp_combs = {}
        for order in range(1, order_max + 1):
            tp_combs[order] = [
                list(p) for p in itertools.product(range(ntp), repeat=order)
                if any(all(tp in self.type_pairs[t] for tp in p) for t in range(nt))]
        lin_by_n = [[] for _ in range(n_fn)]
        for lcid, lc in enumerate(l_comb):
            order = len(lc)
            uniq = sorted({tuple(sorted(zip(lc, tpc))) for tpc in tp_combs[order]})
            for lt in uniq:
                tpc = [x[1] for x in lt]
                n_list = sorted(set.intersection(*[set(self.tp_to_n[tp]) for tp in tpc]))
                t1 = [t for t in range(nt) if all(tp in self.type_pairs[t] for tp in tpc)]
                for n in n_list:
                    lin_by_n[n].append((n, lcid, tpc, order, t1))
        self.linear = [x for sub in lin_by_n for x in sub]
        if self.feature_type == "pair":
            # (tp, n) enumerated tp-major (polymlp_mapping.cpp:75-88 set_ntp_global_attrs)
            self.linear = [x for tp in range(ntp) for n in self.tp_to_n[tp]
                           for x in lin_by_n[n] if x[2][0] == tp]
        self.n_linear = len(self.linear)

        # per-type linear features: term lists over type-local ids into the full +-m array
        # (polymlp_features_utils.cpp:48-76)
        self.features = [[] for _ in range(nt)]  # (global_feature_id, coeffs (T,), ids (T, order))
        for fid, (n, lcid, tpc, order, t1) in enumerate(self.linear):
            for t in t1:
                ids = []
                for lmseq in lm_seq[lcid]:
                    gl = sorted(self.g_index[(n, lm, tp)] for lm, tp in zip(lmseq, tpc))
                    ids.append([self.local[t]["g2l"][gid] for gid in gl])
                self.features[t].append((fid, np.array(lm_coeffs[lcid], float), np.array(ids, int)))

        # polynomial terms (polymlp_model_params_polynomial.cpp:58-86,195-251)
        if self.model_type == 2:
            pidx = list(range(self.n_linear))
        elif self.model_type > 2:
            mo = 1 if self.model_type == 3 else 2
            pidx = [k for k, lt in enumerate(self.linear) if lt[3] <= mo]
        else:
            pidx = []
        self.comb2, self.comb3 = [], []
        comb2_t, comb3_t = [[] for _ in range(nt)], [[] for _ in range(nt)]
        if self.model_type > 1 and self.maxp > 1:
            for i1 in range(len(pidx)):
                for i2 in range(i1 + 1):
                    inter = sorted(set(self.linear[pidx[i1]][4]) & set(self.linear[pidx[i2]][4]))
                    if inter:
                        for t in inter:
                            comb2_t[t].append(len(self.comb2))
                        self.comb2.append((pidx[i2], pidx[i1]))
        if self.model_type > 1 and self.maxp > 2:
            for i1 in range(len(pidx)):
                for i2 in range(i1 + 1):
                    for i3 in range(i2 + 1):
                        inter = sorted(set(self.linear[pidx[i1]][4]) & set(self.linear[pidx[i2]][4])
                                       & set(self.linear[pidx[i3]][4]))
                        if inter:
                            for t in inter:
                                comb3_t[t].append(len(self.comb3))
                            self.comb3.append((pidx[i3], pidx[i2], pidx[i1]))
        self.n_variables = self.n_linear + len(self.comb2) + len(self.comb3)

        # per-type polynomial: (global column, local feature ids) (polymlp_features_polynomial.cpp:22-58)
        self.poly = []
        for t in range(nt):
            map1 = {fid: k for k, (fid, _, _) in enumerate(self.features[t])}
            terms = [(fid, [k]) for fid, k in map1.items()]
            terms += [(self.n_linear + i, [map1[c] for c in self.comb2[i]]) for i in comb2_t[t]]
            b = self.n_linear + len(self.comb2)
            terms += [(b + i, [map1[c] for c in self.comb3[i]]) for i in comb3_t[t]]
            self.poly.append(terms)

        # conjugation helpers per type: full local id -> (noconj local position, is_conj, cc)
        for t in range(nt):
            loc = self.local[t]
            full = []
            for gid in loc["gids"]:
                n, lm, tp = self.g_attr[gid]
                l, m, key, conj, cc = self.lm[lm]
                src = loc["g2l"][self.g_conj[gid]] if conj else loc["g2l"][gid]
                full.append((n, l, m, key, tp, conj, cc, src))
            loc["full"] = full


# --------------------------------------------------------------------------
# numerics
# --------------------------------------------------------------------------


def atom_anlm(tab, types, off, nb, dx, dy, dz, i, deriv):
    """a_nlmtp(i) for the full +-m local array of type(i), and (if deriv) the pair
    derivative blocks.  Restates Local::compute_anlmtp{,_d} (compute/local.cpp:62-113,
    116-217) pair by pair instead of through dense (n_nlmtp x N) arrays.

    Returns a (n_local,), and v (3, n_local, M) = d a / d Delta_alpha per neighbour."""
    t1 = types[i]
    loc = tab.local[t1]
    full = loc["full"]
    nl = len(full)
    sl = slice(off[i], off[i + 1])
    js, x, y, z = nb[sl], dx[sl], dy[sl], dz[sl]
    M = len(js)
    a = np.zeros(nl, complex)
    v = np.zeros((3, nl, M), complex) if deriv else None
    if M == 0:
        return a, v
    r = np.sqrt(x * x + y * y + z * z)
    ok = r < tab.cutoff
    Y, Yx, Yy, Yz = ylm_der(x, y, z, tab.maxl)
    tps = tab.type_pairs[t1, types[js]]
    fn_all = {tp: radial(r, tab.tp_params[tp], tab.cutoff) for tp in set(tps.tolist())}
    for k, (n, l, m, key, tp, conj, cc, src) in enumerate(full):
        if conj or tp not in fn_all:
            continue
        nid = tab.tpn_to_nid[tp][n]
        fn, fnd = fn_all[tp][0][:, nid], fn_all[tp][1][:, nid]
        mask = ok & (tps == tp) & ~(fn < 1e-20)  # local.cpp:164 skip rule
        val = np.where(mask, fn * Y[key], 0.0)
        a[k] = val.sum()
        if deriv:
            d1 = fnd * Y[key] / r
            v[0, k] = np.where(mask, d1 * x + fn * Yx[key], 0.0)
            v[1, k] = np.where(mask, d1 * y + fn * Yy[key], 0.0)
            v[2, k] = np.where(mask, d1 * z + fn * Yz[key], 0.0)
    for k, (n, l, m, key, tp, conj, cc, src) in enumerate(full):
        if conj:  # local.cpp:200-216
            a[k] = cc * np.conj(a[src])
            if deriv:
                v[:, k] = cc * np.conj(v[:, src])
    return a, v


def atom_features(tab, t1, a, deriv):
    """Linear invariants d_f and G[f, id] = d d_f / d a_id (as complex row so that
    delta d_f = Re(sum_id G[f,id] delta a_id)).  Restates Features::compute_features
    (polymlp_features.cpp:173-206) and compute_features_deriv (:208-248)."""
    feats = tab.features[t1]
    d = np.zeros(len(feats))
    G = np.zeros((len(feats), len(a)), complex) if deriv else None
    for f, (fid, c, ids) in enumerate(feats):
        vals = a[ids]  # (T, order)
        d[f] = np.sum(c * np.real(np.prod(vals, axis=1)))
        if deriv:
            order = ids.shape[1]
            for k in range(order):
                others = np.prod(np.delete(vals, k, axis=1), axis=1) if order > 1 else np.ones(len(c))
                np.add.at(G[f], ids[:, k], c * others)
    return d, G


def atom_pair_features(tab, types, off, nb, dx, dy, dz, i, deriv):
    """Pair features d_{n,tp}(i) = sum_j f_n(r_ij) and their derivatives w.r.t. each pair vector,
    g (3, F_local, M).  Restates LocalPair::pair / pair_d (compute/local_pair.cpp:17-55, 57-122):
    no spherical harmonics, no 1e-20 skip rule."""
    t1 = types[i]
    feats = tab.features[t1]
    sl = slice(off[i], off[i + 1])
    js, x, y, z = nb[sl], dx[sl], dy[sl], dz[sl]
    M = len(js)
    d = np.zeros(len(feats))
    g = np.zeros((3, len(feats), M)) if deriv else None
    if M == 0:
        return d, g
    r = np.sqrt(x * x + y * y + z * z)
    ok = r < tab.cutoff
    tps = tab.type_pairs[t1, types[js]]
    fn_all = {tp: radial(r, tab.tp_params[tp], tab.cutoff) for tp in set(tps.tolist())}
    for f, (fid, _, _) in enumerate(feats):
        n, _, tpc, _, _ = tab.linear[fid]
        tp = tpc[0]
        if tp not in fn_all:
            continue
        nid = tab.tpn_to_nid[tp][n]
        mask = ok & (tps == tp)
        d[f] = np.where(mask, fn_all[tp][0][:, nid], 0.0).sum()
        if deriv:
            fnd = np.where(mask, fn_all[tp][1][:, nid], 0.0)
            g[0, f], g[1, f], g[2, f] = fnd * x / r, fnd * y / r, fnd * z / r
    return d, g


def structure_x(tab, axis, positions_c, types, force=True):
    """X rows of one structure: xe (F,), xf (3N, F), xs (6, F).

    Restates Model::run/gtinv/model_polynomial (compute/model.cpp:39-59,90-116,159-267)
    with the sign/row conventions of compute/local.cpp:181-196."""
    types = np.asarray(types, int)
    n = len(types)
    F = tab.n_variables
    off, nb, dx, dy, dz = neighbor_full(axis, positions_c, tab.cutoff)
    xe = np.zeros(F)
    xf = np.zeros((3 * n, F)) if force else np.zeros((0, F))
    xs = np.zeros((6, F)) if force else np.zeros((0, F))
    for i in range(n):
        t1 = types[i]
        if tab.feature_type == "pair":
            d, g = atom_pair_features(tab, types, off, nb, dx, dy, dz, i, force)
        else:
            a, v = atom_anlm(tab, types, off, nb, dx, dy, dz, i, force)
            d, G = atom_features(tab, t1, a, force)
        terms = tab.poly[t1]
        if force:
            sl = slice(off[i], off[i + 1])
            js = nb[sl]
            M = len(js)
            # derivative of each linear feature w.r.t. each pair vector: (3, Floc, M)
            if tab.feature_type != "pair":
                g = np.real(np.einsum("fh,ahm->afm", G, v)) if M else np.zeros((3, len(d), 0))
            Df = np.zeros((len(d), 3 * n))  # rows of X^T restricted to this centre
            Ds = np.zeros((len(d), 6))
            dl = np.stack([dx[sl], dy[sl], dz[sl]])
            for al in range(3):
                np.add.at(Df, (slice(None), 3 * js + al), -g[al])
                Df[:, 3 * i + al] += g[al].sum(axis=1)
            for s, (al, be) in enumerate(((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (2, 0))):
                Ds[:, s] = -(g[al] * dl[be][None, :]).sum(axis=1)
        for col, lids in terms:
            vals = d[lids]
            xe[col] += np.prod(vals)
            if force:
                for k, c in enumerate(lids):
                    w = np.prod(np.delete(vals, k))
                    xf[:, col] += w * Df[c]
                    xs[:, col] += w * Ds[c]
    return xe, xf, xs


def build_x(tab, axis_list, positions_c_list, types_list, force_st):
    """Batch X in the PyModel row layout (compute/py_model.cpp:58-106):
    energies | stress, 6 per force structure | forces, 3N per force structure."""
    n_st = len(axis_list)
    rows_s = 6 * sum(bool(f) for f in force_st)
    rows_f = sum(3 * np.asarray(p).shape[1] for p, f in zip(positions_c_list, force_st) if f)
    X = np.zeros((n_st + rows_s + rows_f, tab.n_variables))
    isb, ifb = n_st, n_st + rows_s
    for s in range(n_st):
        xe, xf, xs = structure_x(tab, axis_list[s], positions_c_list[s], types_list[s], bool(force_st[s]))
        X[s] = xe
        if force_st[s]:
            X[isb:isb + 6] = xs
            isb += 6
            X[ifb:ifb + len(xf)] = xf
            ifb += len(xf)
    return X


def build_x_hybrid(tabs, type_full, type_indices, axis_list, positions_c_list, types_list, force_st):
    """Hybrid design matrix: sub-model blocks side by side, each evaluated on its own subset of atoms.

    Restates PyHybridModel (compute/py_hybrid_model.cpp:58-114) and find_active_atoms (:116-156): a sub-model
    with type_full = False sees only atoms whose type is in its type_indices, renumbered to the position in
    that list; its force rows go back to rows 3 * (original atom) + alpha of the full structure."""
    n_st = len(axis_list)
    n_atoms = [np.asarray(p).shape[1] for p in positions_c_list]
    rows_s = 6 * sum(bool(f) for f in force_st)
    rows_f = sum(3 * n for n, f in zip(n_atoms, force_st) if f)
    X = np.zeros((n_st + rows_s + rows_f, sum(t.n_variables for t in tabs)))
    c0 = 0
    for tab, full, tind in zip(tabs, type_full, type_indices):
        isb, ifb = n_st, n_st + rows_s
        for s in range(n_st):
            ty = np.asarray(types_list[s], int)
            pc = np.asarray(positions_c_list[s], float)
            if full:
                act, ty_a = np.arange(len(ty)), ty
            else:
                act = np.array([a for a in range(len(ty)) if ty[a] in tind], int)
                ty_a = ty[act].copy()
                for rep, t in enumerate(tind):  # std::replace, in list order
                    ty_a[ty_a == t] = rep
            xe, xf, xs = structure_x(tab, axis_list[s], pc[:, act], ty_a, bool(force_st[s]))
            X[s, c0:c0 + tab.n_variables] = xe
            if force_st[s]:
                X[isb:isb + 6, c0:c0 + tab.n_variables] = xs
                isb += 6
                for k, a in enumerate(act):
                    X[ifb + 3 * a:ifb + 3 * a + 3, c0:c0 + tab.n_variables] = xf[3 * k:3 * k + 3]
                ifb += 3 * n_atoms[s]
        c0 += tab.n_variables
    return X


def eval_structure(tab, coeffs, axis, positions_c, types):
    """E (eV/cell), F (N,3), stress (6: xx,yy,zz,xy,yz,zx; eV/cell) of a trained model.

    Restates PolymlpEval::eval_gtinv/collect_properties (compute/polymlp_eval.cpp:182-335):
    the reference folds the coefficients into per-head sums and walks a half list; the
    result is, term by term, the design-matrix rows contracted with the coefficients."""
    xe, xf, xs = structure_x(tab, axis, positions_c, types, True)
    c = np.asarray(coeffs, float)
    return float(xe @ c), (xf @ c).reshape(-1, 3), xs @ c


# --------------------------------------------------------------------------
# regression accumulation (reference Python, restated)
# --------------------------------------------------------------------------


def weights_energy(energies, total_n_atoms, min_e, weight=1.0):
    """_set_weight_energy_data (src/pypolymlp/mlp_dev/core/utils_weights.py:10-22)."""
    e = np.asarray(energies, float) / np.asarray(total_n_atoms, float)
    w = np.ones(len(e))
    w[e > min_e * 0.75] = 0.5
    w[e > min_e * 0.50] = 0.3
    w[e > 0.0] = 0.1
    return w * weight


def weights_force(forces, weight=1.0, tol=1e-12):
    """_set_weight_force_data (utils_weights.py:25-31)."""
    w = np.abs(np.asarray(forces, float))
    w[w < tol] = tol
    w = np.reciprocal(w)
    w[w > 1.0] = 1.0
    return w * weight


def weights_stress(stress, weight_stress, tol=1e-12):
    """_set_weight_stress_data (utils_weights.py:34-47)."""
    s = np.asarray(stress, float)
    nz = np.abs(s) > tol
    lg = np.ones(len(s)) * np.log10(tol)
    lg[nz] = np.log10(np.abs(s)[nz])
    w = np.power(5, -lg)
    w[w > 1.0] = 1.0
    return w * weight_stress


def accumulate(X, n_energy, w, y):
    """One batch of _compute_products_single_batch (data_sequential.py:97-156):
    returns xtx, xty, y_sq_norm, xe_sum, xe_sq_sum with X rows scaled by w and y = w * target."""
    xe = X[:n_energy]
    xe_sum, xe_sq = xe.sum(axis=0), np.square(xe).sum(axis=0)
    Xw = X * w[:, None]
    yw = w * y
    return Xw.T @ Xw, Xw.T @ yw, float(yw @ yw), xe_sum, xe_sq

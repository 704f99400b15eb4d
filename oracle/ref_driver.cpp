// oracle/ref_driver.cpp -- C-ABI driver around the UNMODIFIED reference C++.
//
// TEST INFRASTRUCTURE ONLY.  This file is new code; it only *calls* the
// reference classes, whose sources are compiled in place from
// /root/reference/src/pypolymlp/cxx/src by oracle/Makefile (outputs go to
// oracle/_ref/, git-ignored).  Nothing here is used by the product path
// (pypolymlp_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting library.
//
// What it drives (reference file:line):
//   NeighborFull / NeighborHalf / NeighborCell   compute/neighbor_*.cpp
//   Model::run (X rows of one structure)         compute/model.cpp:39-59
//   PyModel semantics (batch X, OpenMP loop)     compute/py_model.cpp:10-106
//   PolymlpEval::eval (E/F/S)                    compute/polymlp_eval.cpp:24-42
//   Readgtinv                                    polymlp/polymlp_read_gtinv.cpp:23-58
//   get_fn_ / get_ylm_                           polymlp/polymlp_functions_interface.cpp
// Compiled with -fno-access-control so private tables can be exported for
// table-parity tests.

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "compute/model.h"
#include "compute/neighbor_cell.h"
#include "compute/neighbor_full.h"
#include "compute/neighbor_half.h"
#include "compute/polymlp_eval.h"
#include "polymlp/polymlp_functions_interface.h"
#include "polymlp/polymlp_read_gtinv.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct RefModel {
    feature_params fp;
    std::unique_ptr<Model> model;
};

struct RefEval {
    feature_params fp;
    std::unique_ptr<PolymlpEval> eval;
};

feature_params make_fp(int n_type, int n_fn, const double* params,
                       const int* cond_offsets, const int* cond_values,
                       double cutoff, int model_type, int maxp, int maxl,
                       int gtinv_order, const int* gtinv_maxl, int gtinv_version) {
    feature_params fp;
    fp.n_type = n_type;
    fp.force = false;
    fp.params.resize(n_fn);
    for (int n = 0; n < n_fn; ++n) fp.params[n] = {params[2 * n], params[2 * n + 1]};
    fp.params_conditional.assign(n_type, vector2i(n_type));
    int tp = 0;
    for (int i = 0; i < n_type; ++i)
        for (int j = i; j < n_type; ++j) {
            vector1i v(cond_values + cond_offsets[tp], cond_values + cond_offsets[tp + 1]);
            fp.params_conditional[i][j] = v;
            ++tp;
        }
    fp.cutoff = cutoff;
    fp.pair_type = "gaussian";
    fp.model_type = model_type;
    fp.maxp = maxp;
    fp.maxl = maxl;
    if (gtinv_order <= 0) {   // feature_type "pair": the reference passes empty gtinv tables (params_utils.py:57-60)
        fp.feature_type = "pair";
        return fp;
    }
    fp.feature_type = "gtinv";
    vector1i ml(gtinv_maxl, gtinv_maxl + (gtinv_order > 1 ? gtinv_order - 1 : 0));
    Readgtinv rg(gtinv_order, ml, gtinv_version);
    fp.lm_array = rg.get_lm_seq();
    fp.l_comb = rg.get_l_comb();
    fp.lm_coeffs = rg.get_lm_coeffs();
    return fp;
}

void to_vec(const double* axis9, const double* pos3n, int n_atom, vector2d& axis, vector2d& pos) {
    axis.assign(3, vector1d(3));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) axis[i][j] = axis9[3 * i + j];
    pos.assign(3, vector1d(n_atom));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < n_atom; ++j) pos[i][j] = pos3n[i * n_atom + j];
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// torchrun exports OMP_NUM_THREADS=1; the reference arm of bench.py asks for all host cores explicitly
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- Readgtinv --------------------------------------------------------
// Two-pass: call with out pointers null to get sizes.
// sizes[0]=n_lcomb, sizes[1]=sum(order), sizes[2]=sum(n_terms), sizes[3]=sum(n_terms*order)
int ref_readgtinv(int order, const int* maxl, int version, long* sizes,
                  int* lcomb_order, int* l_comb, int* n_terms, int* lm_array, double* coeffs) {
    try {
        vector1i ml(maxl, maxl + (order > 1 ? order - 1 : 0));
        Readgtinv rg(order, ml, version);
        const auto& lm = rg.get_lm_seq();
        const auto& lc = rg.get_l_comb();
        const auto& cf = rg.get_lm_coeffs();
        long s1 = 0, s2 = 0, s3 = 0;
        for (size_t i = 0; i < lc.size(); ++i) {
            if (lcomb_order) lcomb_order[i] = (int)lc[i].size();
            if (n_terms) n_terms[i] = (int)lm[i].size();
            for (size_t k = 0; k < lc[i].size(); ++k) {
                if (l_comb) l_comb[s1] = lc[i][k];
                ++s1;
            }
            for (size_t t = 0; t < lm[i].size(); ++t) {
                if (coeffs) coeffs[s2] = cf[i][t];
                ++s2;
                for (size_t k = 0; k < lm[i][t].size(); ++k) {
                    if (lm_array) lm_array[s3] = lm[i][t][k];
                    ++s3;
                }
            }
        }
        sizes[0] = (long)lc.size(); sizes[1] = s1; sizes[2] = s2; sizes[3] = s3;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// ---- radial / angular functions --------------------------------------
int ref_get_fn(double dis, double cutoff, int n_fn, const double* params, double* fn, double* fn_d) {
    feature_params fp;
    fp.cutoff = cutoff;
    fp.pair_type = "gaussian";
    vector2d p(n_fn);
    for (int n = 0; n < n_fn; ++n) p[n] = {params[2 * n], params[2 * n + 1]};
    vector1d f, fd;
    get_fn_(dis, fp, p, f, fd);
    for (int n = 0; n < n_fn; ++n) { fn[n] = f[n]; fn_d[n] = fd[n]; }
    return 0;
}

// out arrays: complex interleaved (re,im), length 2*(lmax+1)(lmax+2)/2 each
int ref_get_ylm(double r, double x, double y, double z, int lmax,
                double* ylm, double* ylm_dx, double* ylm_dy, double* ylm_dz) {
    vector1dc a, b, c, d;
    get_ylm_(r, x, y, z, lmax, a, b, c, d);
    for (size_t i = 0; i < a.size(); ++i) {
        ylm[2 * i] = a[i].real(); ylm[2 * i + 1] = a[i].imag();
        ylm_dx[2 * i] = b[i].real(); ylm_dx[2 * i + 1] = b[i].imag();
        ylm_dy[2 * i] = c[i].real(); ylm_dy[2 * i + 1] = c[i].imag();
        ylm_dz[2 * i] = d[i].real(); ylm_dz[2 * i + 1] = d[i].imag();
    }
    return 0;
}

// ---- neighbour lists ----------------------------------------------------
// kind: 0 = NeighborFull, 1 = NeighborHalf, 2 = NeighborHalf::get_full_list
// Two-pass: offsets[n_atom+1] always filled; neigh/dx/dy/dz filled if non-null.
int ref_neighbor(int kind, const double* axis9, const double* pos3n, int n_atom, double cutoff,
                 int* offsets, int* neigh, double* dx, double* dy, double* dz) {
    try {
        vector2d axis, pos;
        to_vec(axis9, pos3n, n_atom, axis, pos);
        vector1i off, nb; vector1d vx, vy, vz;
        if (kind == 0) {
            NeighborFull nf(axis, pos, cutoff);
            off = nf.offset; nb = nf.neigh; vx = nf.dx; vy = nf.dy; vz = nf.dz;
        } else {
            NeighborHalf nh(axis, pos, cutoff, false);
            if (kind == 1) { off = nh.offset; nb = nh.neigh; vx = nh.dx; vy = nh.dy; vz = nh.dz; }
            else nh.get_full_list(nb, vx, vy, vz, off);
        }
        for (int i = 0; i <= n_atom; ++i) offsets[i] = off[i];
        if (neigh) {
            for (size_t k = 0; k < nb.size(); ++k) {
                neigh[k] = nb[k]; dx[k] = vx[k]; dy[k] = vy[k]; dz[k] = vz[k];
            }
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// NeighborCell: translations (n_trans x 3), possibly modified axis (9) and positions (3 x n_atom).
int ref_neighbor_cell(const double* axis9, const double* pos3n, int n_atom, double cutoff,
                      int* n_trans, double* trans, int max_trans, double* axis_out, double* pos_out) {
    vector2d axis, pos;
    to_vec(axis9, pos3n, n_atom, axis, pos);
    NeighborCell nc(axis, pos, cutoff);
    const auto& tr = nc.get_translations();
    *n_trans = (int)tr.size();
    if (trans) {
        for (int t = 0; t < (int)tr.size() && t < max_trans; ++t)
            for (int k = 0; k < 3; ++k) trans[3 * t + k] = tr[t][k];
    }
    if (axis_out) {
        const auto& a = nc.get_axis();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) axis_out[3 * i + j] = a[i][j];
    }
    if (pos_out) {
        const auto& p = nc.get_positions_cartesian();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < n_atom; ++j) pos_out[i * n_atom + j] = p[i][j];
    }
    return 0;
}

// ---- fit path -------------------------------------------------------------
void* ref_model_create(int n_type, int n_fn, const double* params,
                       const int* cond_offsets, const int* cond_values,
                       double cutoff, int model_type, int maxp, int maxl,
                       int gtinv_order, const int* gtinv_maxl, int gtinv_version) {
    try {
        auto* m = new RefModel;
        m->fp = make_fp(n_type, n_fn, params, cond_offsets, cond_values, cutoff, model_type,
                        maxp, maxl, gtinv_order, gtinv_maxl, gtinv_version);
        m->model.reset(new Model(m->fp));
        return m;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

void ref_model_destroy(void* h) { delete static_cast<RefModel*>(h); }

int ref_model_n_features(void* h) { return static_cast<RefModel*>(h)->model->get_n_features(); }

// table sizes for type t: out[0]=n_nlmtp, [1]=n_noconj, [2]=n_linear_features(type),
// [3]=n_prod, [4]=n_prod_deriv, [5]=n_deriv_terms, [6]=n_poly_terms
int ref_model_table_sizes(void* h, int t, long* out) {
    auto* m = static_cast<RefModel*>(h);
    auto& f = m->model->polymlp.features;
    auto& maps = f.mapping.get_maps();
    auto& mt = maps.maps_type[t];
    out[0] = (long)mt.nlmtp_attrs.size();
    out[1] = (long)mt.nlmtp_attrs_noconj.size();
    out[2] = (long)f.feature_sizes[t];
    out[3] = (long)f.prod[t].size();
    out[4] = (long)f.prod_deriv[t].size();
    long nd = 0;
    for (const auto& mf : f.mapped_features_deriv[t]) nd += (long)mf.size();
    out[5] = nd;
    out[6] = (long)mt.polynomial.size();
    return 0;
}

// polynomial terms of type t: global column + up to 3 local ids (-1 padded); returns count
int ref_model_polynomial(void* h, int t, int* global_id, int* local_ids3) {
    auto* m = static_cast<RefModel*>(h);
    auto& maps = m->model->polymlp.features.mapping.get_maps();
    auto& poly = maps.maps_type[t].polynomial;
    int i = 0;
    for (const auto& term : poly) {
        if (global_id) {
            global_id[i] = term.global_id;
            for (int k = 0; k < 3; ++k)
                local_ids3[3 * i + k] = k < (int)term.local_ids.size() ? term.local_ids[k] : -1;
        }
        ++i;
    }
    return i;
}

// Feature attributes as the reference's FeaturesAttr reports them (pybind11_mlp.cpp:70-82): flattened to one int
// stream so the test can compare with pm_model_feature_attrs.  Layout of `out` (returned length; out may be NULL):
//   n_linear, then per linear feature: radial id, gtinv id (-1 for pair models), len, tp ids...;
//   n_poly, then per polynomial column (comb2 then comb3): len, linear ids...;  n_type, type_pairs row-major.
long ref_model_feature_attrs(void* h, int* out) {
    auto* m = static_cast<RefModel*>(h);
    PolymlpAPI api;                       // as compute/py_features_attr.cpp:16-19 does: a fresh API object,
    api.set_model_parameters(m->fp);      // model parameters only (Model itself goes through set_features)
    const auto& mp = api.get_model_params();
    auto& maps = api.get_maps();
    std::vector<int> v;
    if (m->fp.feature_type == "pair") {
        v.push_back((int)maps.ntp_attrs.size());
        for (const auto& a : maps.ntp_attrs) { v.push_back(a.n); v.push_back(-1); v.push_back(1); v.push_back(a.tp); }
    } else {
        const auto& lin = mp.get_linear_terms();
        const auto& tpc = mp.get_tp_combs();
        v.push_back((int)lin.size());
        for (const auto& t : lin) {
            const auto& c = tpc[t.order][t.tp_comb_id];
            v.push_back(t.n); v.push_back(t.lm_comb_id); v.push_back((int)c.size());
            v.insert(v.end(), c.begin(), c.end());
        }
    }
    v.push_back((int)(mp.get_comb2().size() + mp.get_comb3().size()));
    for (const auto* combs : {&mp.get_comb2(), &mp.get_comb3()})
        for (const auto& c : *combs) { v.push_back((int)c.size()); v.insert(v.end(), c.begin(), c.end()); }
    const int nt = (int)maps.type_pairs.size();
    v.push_back(nt);
    for (const auto& row : maps.type_pairs) v.insert(v.end(), row.begin(), row.end());
    if (out) std::copy(v.begin(), v.end(), out);
    return (long)v.size();
}

// X rows of one structure: xe[F], xf[3N*F] (row-major rows 3*atom+alpha), xs[6*F]
int ref_model_run(void* h, const double* axis9, const double* pos3n, const int* types, int n_atom,
                  int force, double* xe, double* xf, double* xs) {
    try {
        auto* m = static_cast<RefModel*>(h);
        vector2d axis, pos;
        to_vec(axis9, pos3n, n_atom, axis, pos);
        vector1i ty(types, types + n_atom);
        NeighborFull neigh(axis, pos, m->fp.cutoff);
        Eigen::VectorXd e; Eigen::MatrixXd f, s;
        m->model->run(neigh, ty, force != 0, e, f, s);
        const int F = (int)e.size();
        for (int c = 0; c < F; ++c) xe[c] = e(c);
        if (force) {
            for (int r = 0; r < 3 * n_atom; ++r) for (int c = 0; c < F; ++c) xf[(size_t)r * F + c] = f(r, c);
            for (int r = 0; r < 6; ++r) for (int c = 0; c < F; ++c) xs[(size_t)r * F + c] = s(r, c);
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// per-atom intermediates: a_nlmtp (full, complex interleaved, size out n) and linear features d
int ref_model_atom(void* h, const double* axis9, const double* pos3n, const int* types, int n_atom,
                   int atom, double* anlmtp_ri, int* n_anlmtp, double* dn, int* n_dn) {
    try {
        auto* m = static_cast<RefModel*>(h);
        vector2d axis, pos;
        to_vec(axis9, pos3n, n_atom, axis, pos);
        vector1i ty(types, types + n_atom);
        NeighborFull neigh(axis, pos, m->fp.cutoff);
        Local local(n_atom);
        vector1dc a;
        local.compute_anlmtp(m->model->polymlp, neigh, ty, atom, a);
        vector1d d;
        m->model->polymlp.compute_features(a, ty[atom], d);
        *n_anlmtp = (int)a.size();
        *n_dn = (int)d.size();
        for (size_t i = 0; i < a.size(); ++i) { anlmtp_ri[2 * i] = a[i].real(); anlmtp_ri[2 * i + 1] = a[i].imag(); }
        for (size_t i = 0; i < d.size(); ++i) dn[i] = d[i];
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// Batch X with PyModel row layout (py_model.cpp:58-106): energies | stress (6/str) | forces (3N/str).
// X is ROW-major (n_rows x F) here.  n_threads <= 0 -> OpenMP default.
long ref_model_n_rows(int n_st, const int* n_atoms, const int* force_st) {
    long r = n_st;
    for (int i = 0; i < n_st; ++i) if (force_st[i]) r += 6 + 3L * n_atoms[i];
    return r;
}

int ref_model_build_x(void* h, int n_st, const double* axis9s, const double* pos_concat,
                      const int* types_concat, const int* n_atoms, const int* force_st,
                      int n_threads, double* x) {
    try {
        auto* m = static_cast<RefModel*>(h);
        const int F = m->model->get_n_features();
        std::vector<long> aoff(n_st + 1, 0), sbeg(n_st, -1), fbeg(n_st, -1);
        for (int i = 0; i < n_st; ++i) aoff[i + 1] = aoff[i] + n_atoms[i];
        long is = n_st;
        for (int i = 0; i < n_st; ++i) if (force_st[i]) { sbeg[i] = is; is += 6; }
        long ifo = is;
        for (int i = 0; i < n_st; ++i) if (force_st[i]) { fbeg[i] = ifo; ifo += 3L * n_atoms[i]; }
#ifdef _OPENMP
        if (n_threads > 0) omp_set_num_threads(n_threads);
        #pragma omp parallel for schedule(guided, 1)
#endif
        for (int i = 0; i < n_st; ++i) {
            const int na = n_atoms[i];
            vector2d axis, pos;
            to_vec(axis9s + 9 * i, pos_concat + 3 * aoff[i], na, axis, pos);
            vector1i ty(types_concat + aoff[i], types_concat + aoff[i] + na);
            NeighborFull neigh(axis, pos, m->fp.cutoff);
            Eigen::VectorXd e; Eigen::MatrixXd f, s;
            Model& mod = *m->model;
            mod.run(neigh, ty, force_st[i] != 0, e, f, s);
            for (int c = 0; c < F; ++c) x[(size_t)i * F + c] = e(c);
            if (force_st[i]) {
                for (int r = 0; r < 6; ++r) for (int c = 0; c < F; ++c) x[(size_t)(sbeg[i] + r) * F + c] = s(r, c);
                for (int r = 0; r < 3 * na; ++r) for (int c = 0; c < F; ++c) x[(size_t)(fbeg[i] + r) * F + c] = f(r, c);
            }
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// ---- eval path ----------------------------------------------------------
void* ref_eval_create(int n_type, int n_fn, const double* params,
                      const int* cond_offsets, const int* cond_values,
                      double cutoff, int model_type, int maxp, int maxl,
                      int gtinv_order, const int* gtinv_maxl, int gtinv_version,
                      const double* coeffs, int n_coeffs) {
    try {
        auto* m = new RefEval;
        m->fp = make_fp(n_type, n_fn, params, cond_offsets, cond_values, cutoff, model_type,
                        maxp, maxl, gtinv_order, gtinv_maxl, gtinv_version);
        vector1d c(coeffs, coeffs + n_coeffs);
        m->eval.reset(new PolymlpEval(m->fp, c));
        return m;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

void ref_eval_destroy(void* h) { delete static_cast<RefEval*>(h); }

// forces out: (n_atom x 3) row-major; stress out: 6 (xx,yy,zz,xy,yz,zx)
int ref_eval(void* h, const double* axis9, const double* pos3n, const int* types, int n_atom,
             int use_openmp, double* energy, double* forces, double* stress) {
    try {
        auto* m = static_cast<RefEval*>(h);
        vector2d axis, pos;
        to_vec(axis9, pos3n, n_atom, axis, pos);
        vector1i ty(types, types + n_atom);
        NeighborHalf neigh(axis, pos, m->fp.cutoff, use_openmp != 0);
        double e; vector2d f; vector1d s;
        m->eval->eval(ty, neigh, use_openmp != 0, e, f, s);
        *energy = e;
        for (int i = 0; i < n_atom; ++i) for (int k = 0; k < 3; ++k) forces[3 * i + k] = f[i][k];
        for (int k = 0; k < 6; ++k) stress[k] = s[k];
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"

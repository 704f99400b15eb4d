// pybind_module.cpp -- drop-in `libmlpcpp` module for the hot path.
//
// Exposes the same classes, constructor signatures and getters as the reference's pybind11 module
// (src/pypolymlp/cxx/src/python/pybind11_mlp.cpp:10-94) for the hot path, implemented in host C++ on top of
// the C ABI (include/polymlp_b200.h) -> CUDA:
//   PotentialModel(params_dict, axis, positions_c, types, n_st_dataset, force_dataset, n_atoms_all)
//       .get_x() .get_fbegin() .get_sbegin() .get_n_data()                      (pybind11_mlp.cpp:12-27)
//   PotentialPropertiesFast(params_dict, coeffs)
//       .eval(axis, positions_c, types, use_openmp) .eval_multiple(axis[], positions_c[], types[])
//       .get_e() .get_f() .get_s() .get_e_array() .get_f_array() .get_s_array() (pybind11_mlp.cpp:51-67)
//   Readgtinv(order, maxl, version) .get_lm_seq() .get_l_comb() .get_lm_coeffs() (pybind11_mlp.cpp:84-94)
//   FeaturesAttr(params_dict) .get_n_features()                                  (subset of :70-82)
// Additive: PotentialXtX(params_dict) .add(...) .finalize() -- the fused feature + X^T X accumulation.
// Errors: C-ABI status PM_ERR_INVALID -> ValueError, anything else -> RuntimeError (as pybind11 maps
// std::invalid_argument / std::runtime_error in the reference).  No CPU fallback.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/polymlp_b200.h"

namespace py = pybind11;
using vector1i = std::vector<int>;
using vector2i = std::vector<vector1i>;
using vector3i = std::vector<vector2i>;
using vector1d = std::vector<double>;
using vector2d = std::vector<vector1d>;
using vector3d = std::vector<vector2d>;

namespace {

void check(int status) {
    if (status == PM_OK) return;
    const std::string msg = pm_last_error();
    if (status == PM_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

int default_device() {
    for (const char* key : {"POLYMLP_B200_DEVICE", "LOCAL_RANK"})
        if (const char* v = std::getenv(key)) return std::atoi(v);
    return 0;
}

// feature_params from params.as_dict() (keys as in compute/py_params.cpp:14-43)
struct Model {
    pm_model* h = nullptr;
    explicit Model(const py::dict& params_dict) {
        const int n_type = params_dict["n_type"].cast<int>();
        const py::dict model = params_dict["model"].cast<py::dict>();
        const std::string feature_type = model["feature_type"].cast<std::string>();
        if (feature_type != "gtinv" && feature_type != "pair")
            throw std::invalid_argument("feature_type must be 'gtinv' or 'pair'");
        const bool pair = feature_type == "pair";
        const auto pair_params = model["pair_params"].cast<vector2d>();
        vector1d pp;
        for (const auto& p : pair_params) { pp.push_back(p.at(0)); pp.push_back(p.at(1)); }
        const py::dict cond = model["pair_params_conditional"].cast<py::dict>();
        vector1i off{0}, val;
        for (int i = 0; i < n_type; ++i)
            for (int j = i; j < n_type; ++j) {
                const auto lst = cond[py::make_tuple(i, j)].cast<vector1i>();
                val.insert(val.end(), lst.begin(), lst.end());
                off.push_back((int)val.size());
            }
        if (val.empty()) val.push_back(0);
        vector3i lm_seq;
        vector2i l_comb;
        vector2d lm_coeffs;
        if (!pair) {
            const py::dict gtinv = model["gtinv"].cast<py::dict>();
            lm_seq = gtinv["lm_seq"].cast<vector3i>();
            l_comb = gtinv["l_comb"].cast<vector2i>();
            lm_coeffs = gtinv["lm_coeffs"].cast<vector2d>();
        }
        vector1i lo, lc, nt, lm;
        vector1d cf;
        for (size_t i = 0; i < l_comb.size(); ++i) {
            lo.push_back((int)l_comb[i].size());
            lc.insert(lc.end(), l_comb[i].begin(), l_comb[i].end());
            nt.push_back((int)lm_seq[i].size());
            for (const auto& term : lm_seq[i]) lm.insert(lm.end(), term.begin(), term.end());
            cf.insert(cf.end(), lm_coeffs[i].begin(), lm_coeffs[i].end());
        }
        pm_feature_params fp{};
        fp.n_type = n_type; fp.n_fn = (int)pair_params.size(); fp.pair_params = pp.data();
        fp.cond_offsets = off.data(); fp.cond_values = val.data();
        fp.cutoff = model["cutoff"].cast<double>(); fp.model_type = model["model_type"].cast<int>();
        fp.max_p = model["max_p"].cast<int>(); fp.max_l = model["max_l"].cast<int>();
        fp.n_lcomb = (int)l_comb.size(); fp.lcomb_order = lo.data(); fp.l_comb = lc.data();
        fp.n_terms = nt.data(); fp.lm_seq = lm.data(); fp.lm_coeffs = cf.data();
        fp.feature_type = pair ? PM_FEATURE_PAIR : PM_FEATURE_GTINV;
        check(pm_model_create(&fp, &h));
    }
    ~Model() { pm_model_destroy(h); }
    Model(const Model&) = delete;
};

struct Batch {
    vector1d axis, pos;
    vector1i types, n_atoms, force;
    pm_structures st{};
    Batch(const vector3d& axis_a, const vector3d& pos_a, const vector2i& types_a, const std::vector<bool>& force_st) {
        const size_t n = axis_a.size();
        if (pos_a.size() != n || types_a.size() != n || force_st.size() != n)
            throw std::invalid_argument("inconsistent number of structures");
        for (size_t s = 0; s < n; ++s) {
            if (axis_a[s].size() != 3 || pos_a[s].size() != 3) throw std::invalid_argument("axis/positions_c must be 3 x .");
            for (int i = 0; i < 3; ++i) {
                if (axis_a[s][i].size() != 3) throw std::invalid_argument("axis must be 3 x 3");
                axis.insert(axis.end(), axis_a[s][i].begin(), axis_a[s][i].end());
            }
            const size_t na = types_a[s].size();
            for (int i = 0; i < 3; ++i) {
                if (pos_a[s][i].size() != na) throw std::invalid_argument("positions_c must be (3, N) with N == len(types)");
                pos.insert(pos.end(), pos_a[s][i].begin(), pos_a[s][i].end());
            }
            types.insert(types.end(), types_a[s].begin(), types_a[s].end());
            n_atoms.push_back((int)na);
            force.push_back(force_st[s] ? 1 : 0);
        }
        st.n_st = (int)n; st.axis = axis.data(); st.positions_c = pos.data(); st.types = types.data();
        st.n_atoms = n_atoms.data(); st.force = force.data();
    }
};

class PyModel {
    Model model;
    pm_context* ctx = nullptr;
    py::array_t<double> x;
    vector1i fbegin, sbegin, n_data;

  public:
    PyModel(const py::dict& params_dict, const vector3d& axis, const vector3d& positions_c, const vector2i& types,
            const vector1i& n_st_dataset, const std::vector<bool>& force_dataset, const vector1i& n_atoms_all)
        : model(params_dict) {
        std::vector<bool> force_st;
        for (size_t i = 0; i < n_st_dataset.size(); ++i)
            for (int k = 0; k < n_st_dataset[i]; ++k) force_st.push_back(force_dataset[i]);
        // row bookkeeping of PyModel::set_index (compute/py_model.cpp:58-106)
        const int n_st = (int)force_st.size();
        fbegin.assign(n_st_dataset.size(), -1);
        sbegin.assign(n_st_dataset.size(), -1);
        n_data = {n_st, 0, 0};
        int ist = n_st;
        for (size_t i = 0; i < n_st_dataset.size(); ++i)
            if (force_dataset[i]) { sbegin[i] = ist; ist += 6 * n_st_dataset[i]; n_data[2] += 6 * n_st_dataset[i]; }
        int ifo = ist, k = 0;
        for (size_t i = 0; i < n_st_dataset.size(); ++i) {
            if (force_dataset[i]) fbegin[i] = ifo;
            for (int j = 0; j < n_st_dataset[i]; ++j, ++k)
                if (force_dataset[i]) { ifo += 3 * n_atoms_all.at(k); n_data[1] += 3 * n_atoms_all.at(k); }
        }
        Batch b(axis, positions_c, types, force_st);
        check(pm_context_create(model.h, default_device(), 0, 0, &ctx));
        const py::ssize_t rows = (py::ssize_t)pm_batch_rows(&b.st), F = pm_model_n_features(model.h);
        x = py::array_t<double>({rows, F});
        check(pm_features_x(ctx, &b.st, x.mutable_data()));
    }
    ~PyModel() { pm_context_destroy(ctx); }
    py::array_t<double> get_x() { return x; }
    const vector1i& get_fbegin() const { return fbegin; }
    const vector1i& get_sbegin() const { return sbegin; }
    const vector1i& get_n_data() const { return n_data; }
};

class PyXtX {
    Model model;
    pm_context* ctx = nullptr;
    int F;

  public:
    explicit PyXtX(const py::dict& params_dict) : model(params_dict) {
        check(pm_context_create(model.h, default_device(), 0, 0, &ctx));
        check(pm_fit_reset(ctx));
        F = pm_model_n_features(model.h);
    }
    ~PyXtX() { pm_context_destroy(ctx); }
    void add(const vector3d& axis, const vector3d& positions_c, const vector2i& types, const std::vector<bool>& force_st,
             py::array_t<double, py::array::c_style | py::array::forcecast> w,
             py::array_t<double, py::array::c_style | py::array::forcecast> y) {
        Batch b(axis, positions_c, types, force_st);
        const int64_t rows = pm_batch_rows(&b.st);
        if (w.size() != rows || y.size() != rows) throw std::invalid_argument("w and y must have one entry per row of the batch");
        check(pm_fit_accumulate(ctx, &b.st, w.data(), y.data()));
    }
    py::dict finalize() {
        py::array_t<double> xtx({(py::ssize_t)F, (py::ssize_t)F}), xty(F), xe_sum(F), xe_sq(F);
        double ysq = 0.0;
        int64_t nd = 0;
        check(pm_fit_finalize(ctx, xtx.mutable_data(), xty.mutable_data(), xe_sum.mutable_data(), xe_sq.mutable_data(), &ysq, &nd));
        py::dict d;
        d["xtx"] = xtx; d["xty"] = xty; d["xe_sum"] = xe_sum; d["xe_sq_sum"] = xe_sq;
        d["y_sq_norm"] = ysq; d["total_n_data"] = nd;
        return d;
    }
};

class PyPropertiesFast {
    Model model;
    pm_context* ctx = nullptr;
    double energy = 0.0;
    vector2d force;
    vector1d stress;
    vector1d e_array;
    vector3d f_array;
    vector2d s_array;

    void run(const vector3d& axis, const vector3d& pos, const vector2i& types, vector1d& e, vector3d& f, vector2d& s) {
        Batch b(axis, pos, types, std::vector<bool>(axis.size(), true));
        const size_t n = axis.size();
        size_t na = 0;
        for (int v : b.n_atoms) na += v;
        vector1d eb(n), fb(3 * na + 1), sb(6 * n + 1);
        check(pm_eval(ctx, &b.st, eb.data(), fb.data(), sb.data()));
        e = eb;
        f.assign(n, {});
        s.assign(n, vector1d(6));
        size_t off = 0;
        for (size_t k = 0; k < n; ++k) {
            f[k].assign(b.n_atoms[k], vector1d(3));
            for (int a = 0; a < b.n_atoms[k]; ++a)
                for (int c = 0; c < 3; ++c) f[k][a][c] = fb[3 * (off + a) + c];
            off += b.n_atoms[k];
            for (int c = 0; c < 6; ++c) s[k][c] = sb[6 * k + c];
        }
    }

  public:
    PyPropertiesFast(const py::dict& params_dict, const vector1d& coeffs) : model(params_dict) {
        check(pm_context_create(model.h, default_device(), 0, 0, &ctx));
        check(pm_eval_set_coeffs(ctx, coeffs.data(), (int)coeffs.size()));
    }
    ~PyPropertiesFast() { pm_context_destroy(ctx); }
    void eval(const vector2d& axis, const vector2d& positions_c, const vector1i& types, const bool) {
        vector1d e; vector3d f; vector2d s;
        run({axis}, {positions_c}, {types}, e, f, s);
        energy = e[0]; force = f[0]; stress = s[0];
    }
    void eval_multiple(const vector3d& axis, const vector3d& positions_c, const vector2i& types) {
        run(axis, positions_c, types, e_array, f_array, s_array);
    }
    const double& get_e() const { return energy; }
    const vector2d& get_f() const { return force; }
    const vector1d& get_s() const { return stress; }
    const vector1d& get_e_array() const { return e_array; }
    const vector3d& get_f_array() const { return f_array; }
    const vector2d& get_s_array() const { return s_array; }
};

class PyReadgtinv {
    vector3i lm_array;
    vector2i l_array;
    vector2d coeffs;

  public:
    PyReadgtinv(const int order, const vector1i& maxl, const int version) {
        int64_t sz[4];
        const char* dir = std::getenv("POLYMLP_B200_GTINV_DIR");
        check(pm_gtinv_read(dir, order, maxl.data(), (int)maxl.size(), version, sz, nullptr, nullptr, nullptr, nullptr, nullptr));
        vector1i lo(sz[0]), lc(sz[1] + 1), nt(sz[0]), lm(sz[3] + 1);
        vector1d cf(sz[2] + 1);
        check(pm_gtinv_read(dir, order, maxl.data(), (int)maxl.size(), version, sz, lo.data(), lc.data(), nt.data(), lm.data(), cf.data()));
        size_t p1 = 0, p2 = 0, p3 = 0;
        for (int64_t i = 0; i < sz[0]; ++i) {
            l_array.emplace_back(lc.begin() + p1, lc.begin() + p1 + lo[i]);
            coeffs.emplace_back(cf.begin() + p2, cf.begin() + p2 + nt[i]);
            vector2i seq(nt[i]);
            for (int t = 0; t < nt[i]; ++t) { seq[t].assign(lm.begin() + p3, lm.begin() + p3 + lo[i]); p3 += lo[i]; }
            lm_array.push_back(seq);
            p1 += lo[i]; p2 += nt[i];
        }
    }
    const vector3i& get_lm_seq() const { return lm_array; }
    const vector2i& get_l_comb() const { return l_array; }
    const vector2d& get_lm_coeffs() const { return coeffs; }
};

class PyFeaturesAttr {
    Model model;

  public:
    explicit PyFeaturesAttr(const py::dict& params_dict) : model(params_dict) {}
    int get_n_features() const { return pm_model_n_features(model.h); }
};

}  // namespace

PYBIND11_MODULE(libmlpcpp, m) {
    m.doc() = "B200-native drop-in for pypolymlp.cxx.lib.libmlpcpp (hot path only)";
    py::class_<PyModel>(m, "PotentialModel")
        .def(py::init<const py::dict&, const vector3d&, const vector3d&, const vector2i&, const vector1i&,
                      const std::vector<bool>&, const vector1i&>())
        .def("get_x", &PyModel::get_x)
        .def("get_fbegin", &PyModel::get_fbegin, py::return_value_policy::reference_internal)
        .def("get_sbegin", &PyModel::get_sbegin, py::return_value_policy::reference_internal)
        .def("get_n_data", &PyModel::get_n_data, py::return_value_policy::reference_internal);
    py::class_<PyXtX>(m, "PotentialXtX")
        .def(py::init<const py::dict&>())
        .def("add", &PyXtX::add)
        .def("finalize", &PyXtX::finalize);
    py::class_<PyPropertiesFast>(m, "PotentialPropertiesFast")
        .def(py::init<const py::dict&, const vector1d&>())
        .def("eval", &PyPropertiesFast::eval)
        .def("eval_multiple", &PyPropertiesFast::eval_multiple)
        .def("get_e", &PyPropertiesFast::get_e, py::return_value_policy::reference_internal)
        .def("get_f", &PyPropertiesFast::get_f, py::return_value_policy::reference_internal)
        .def("get_s", &PyPropertiesFast::get_s, py::return_value_policy::reference_internal)
        .def("get_e_array", &PyPropertiesFast::get_e_array, py::return_value_policy::reference_internal)
        .def("get_f_array", &PyPropertiesFast::get_f_array, py::return_value_policy::reference_internal)
        .def("get_s_array", &PyPropertiesFast::get_s_array, py::return_value_policy::reference_internal);
    py::class_<PyReadgtinv>(m, "Readgtinv")
        .def(py::init<const int, const vector1i&, const int>())
        .def("get_lm_seq", &PyReadgtinv::get_lm_seq, py::return_value_policy::reference_internal)
        .def("get_l_comb", &PyReadgtinv::get_l_comb, py::return_value_policy::reference_internal)
        .def("get_lm_coeffs", &PyReadgtinv::get_lm_coeffs, py::return_value_policy::reference_internal);
    py::class_<PyFeaturesAttr>(m, "FeaturesAttr")
        .def(py::init<const py::dict&>())
        .def("get_n_features", &PyFeaturesAttr::get_n_features);
}

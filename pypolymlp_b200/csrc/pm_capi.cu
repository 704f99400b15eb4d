// pm_capi.cu -- C ABI (include/polymlp_b200.h): host tables, device context, chunked pipeline.
#include "../../include/polymlp_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "pm_kernels.cuh"
#include "pm_nccl.hpp"

namespace pm {
void set_features_smem(size_t smem_bytes);
void set_lrows_kmax(int kmax, int runmax, int tilemax);
size_t lrows_mma_smem(const DevModel& m);
bool lrows_big_supported(const DevModel& m);
size_t xrows_mma_smem(const DevModel& m);
double microbench_dgemm(int n, cudaStream_t s);
double microbench_red(int n_rows, cudaStream_t s);
double microbench_dmma_chain(int code, cudaStream_t s);
void solve_ridge_device(const double* C, int fpad, int F, const double* xe_sum_h, const double* xe_sq_h,
                        double y_sq_norm, long n_data, const double* alphas, int n_alpha, const double* scales_in,
                        long n_energy, bool include_force, double threshold, double* scales_out, double* coefs,
                        double* rmse, cudaStream_t stream);
}  // namespace pm

using namespace pm;

static thread_local std::string g_err;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
    } while (0)

template <typename F> static int guarded(F&& f) {
    try {
        f();
        return PM_OK;
    } catch (const CudaError& e) {
        g_err = e.what();
        return PM_ERR_CUDA;
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        return PM_ERR_INVALID;
    } catch (const std::exception& e) {
        g_err = e.what();
        return PM_ERR_RUNTIME;
    }
}

struct pm_model {
    HostModel hm;
};

enum Stage { ST_H2D = 0, ST_NEIGH, ST_BASIS, ST_ANLM, ST_FEAT, ST_LROWS, ST_XROWS, ST_SYRK, ST_EVAL, ST_D2H, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"h2d", "neighbor", "pair_basis", "anlm", "features_G", "lrows", "xrows",
                                            "syrk", "eval", "d2h"};

template <typename T> struct DevVec {
    T* p = nullptr;
    size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        size_t want = n + n / 8 + 64;
        CK(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostChunk {
    int n_st = 0, n_atoms = 0, n_rows = 0, max_trans = 0, max_atoms = 0;
    std::vector<int> atom_off, st_of_atom, types, trans_off, force, erow, srow, frow;
    std::vector<double> x, y, z, trans, w, yv;
    // device cell list (K1): per-structure parameters, see DevBatch
    std::vector<int> cl_int, cl_tmap;
    std::vector<double> cl_ainv;
    int n_bins = 0, cl_nmax = 1, cl_tmax = 1;
    bool use_cl = false;
    std::vector<double> we;   // fit: true weights of the energy rows (w[erow] is 1 on the device, see k_xe_reduce)
    std::vector<long> brow_e, brow_s, brow_f;  // rows in the caller's batch layout
};

struct pm_context {
    const pm_model* model = nullptr;
    int device = 0;
    int flags = 0;
    size_t ws_cap = 0;
    size_t ws_cap_max = 0;   // > ws_cap: plan_chunks may grow the workspace for large models (default workspace only)
    cudaStream_t stream = nullptr;
    DevModel dm{};
    std::vector<void*> table_allocs;
    bool simple_l = false, simple_x = false, simple_s = false, scatter = false;
    size_t feat_smem = 0;
    // accumulators
    double* acc = nullptr;
    size_t acc_n = 0;
    int64_t n_data = 0;
    // chunk device buffers
    DevVec<int> d_atom_off, d_st_of_atom, d_types, d_trans_off, d_force, d_erow, d_srow, d_frow, d_counts, d_seg_off,
        d_nbr, d_centre, d_rev, d_err, d_cl_int, d_cl_tmap, d_bin_start, d_bin_atoms, d_atom_bin, d_atom_img, d_bin_count, d_cl_hkey;
    DevVec<ulonglong2> d_masks;
    DevVec<double> d_x, d_y, d_z, d_trans, d_w, d_yv, d_PB, d_dfeat, d_dpv, d_G, d_L, d_Lpv, d_Xown, d_S, d_X, d_Ah, d_coeffs, d_cmat, d_e,
        d_f, d_s, d_we;
    DevVec<double> d_cl_ainv, d_cl_hd;
    DevVec<double> d_clin;      // [n_type][fl] coefficient of the linear column of each padded feature (eval, with d_cmat)
    DevVec<double2> d_anc, d_agg;
    DevVec<unsigned char> d_scan_tmp;
    DevBatch last_batch{};
    int last_pairs = 0;
    // staged batch (bench: inputs resident in HBM)
    std::vector<HostChunk> staged;
    std::vector<std::vector<void*>> staged_dev;  // per chunk device copies
    // stats
    int64_t launches = 0;
    bool profile = false;
    double stage_ms[ST_COUNT] = {0};
    int64_t stage_launches[ST_COUNT] = {0};
    cudaEvent_t ev[ST_COUNT + 1] = {nullptr};
    bool has_coeffs = false;
    bool has_cmat = false;
    // K5 runs on its own stream: the host-side work and the first kernels of the next chunk overlap the SYRK of the
    // previous chunk.  (Measured: the FP64-heavy small kernels make little progress next to the DMMA-saturating SYRK
    // CTA -- the FP64 pipe is the shared resource -- so the gain is ~1 % on the device and ~2 % end to end.)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_x = nullptr, ev_s = nullptr;
    bool two_stream = true, syrk_pending = false;
    // pm_profile_enable(c, 2): only K5 is timed, with event pairs on its own stream and no host synchronisation
    bool time_syrk = false;
    std::vector<cudaEvent_t> syrk_ev;   // pairs (before, after), created lazily
    size_t syrk_ev_used = 0;
    size_t g_zero_cap = 0;              // capacity of d_G for which the buffer has been cleared (single-type models)
    double* pinned = nullptr;   // host (pinned) copy of the packed result of pm_fit_finalize
    size_t pinned_n = 0;
    double* packed = nullptr;   // device: [xtx F*F | xty F | xe_sum F | xe_sq F | y_sq_norm | n_data]
    SyrkScratch syrk_scr{};     // parked partial tiles + per-tile arrival counters of the deterministic SYRK fix-up
    // multi-GPU: NCCL communicator (one rank per context) and the packed upper-triangle reduce buffer
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    bool comm_owned = true;     // false: the communicator belongs to a pm_multi (destroyed there)
    DevVec<double> d_red;
    double* d_small = nullptr;  // 64 doubles of device scratch for pm_comm_allreduce
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // pm_timer_start / pm_timer_stop
    // pm_eval on a large batch runs two lanes: this context and a sibling context on the same device (own stream and
    // chunk buffers) work on the two halves of the batch from two host threads, so that the copies, the host-side chunk
    // preparation and the K1 round trip of one lane overlap the kernels of the other
    pm_context* sibling = nullptr;
    size_t ws_arg = 0;          // workspace_bytes as given to pm_context_create
    uint64_t coeff_version = 0; // bumped by pm_eval_set_coeffs (the sibling copies the coefficients when it is behind)
};

// ------------------------------------------------------------------------------------------------
template <typename T> static T* upload(pm_context* c, const std::vector<T>& v) {
    T* p = nullptr;
    const size_t n = std::max<size_t>(v.size(), 1);
    CK(cudaMalloc(&p, n * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    c->table_allocs.push_back(p);
    return p;
}

static void build_device_model(pm_context* c) {
    const HostModel& hm = c->model->hm;
    DevModel& d = c->dm;
    if (hm.fp.n_type > MAXT) throw std::invalid_argument("device path supports at most 4 atom types");
    if (hm.n_lm_half > MAX_NH) throw std::invalid_argument("max_l > 20 is not supported");
    d.n_type = hm.fp.n_type;
    d.n_fn = hm.fp.n_fn;
    d.n_tp = hm.n_tp;
    d.maxl = hm.fp.maxl;
    d.nh = hm.n_lm_half;
    d.n_variables = hm.n_variables;
    d.n_linear = hm.n_linear;
    d.fpad = (hm.n_variables + 1 + 127) / 128 * 128;
    d.cutoff = hm.fp.cutoff;
    d.pbstride = 4 + 2 * d.n_fn + 8 * d.nh;
    d.fl = 8;
    d.hmax = 1;
    d.gstride = 32;
    int kmax = 4;
    for (const auto& T : hm.types) {
        d.fl = std::max(d.fl, T.n_fpad);
        d.hmax = std::max(d.hmax, T.n_head);
        d.gstride = std::max(d.gstride, T.g_size);
        for (int u = 0; u < d.n_type; ++u)
            for (int n = 0; n < d.n_fn; ++n) kmax = std::max(kmax, 2 * (T.seg_n_off[u][n + 1] - T.seg_n_off[u][n]));
    }
    {   // longest run of G blocks / most feature tiles of one (type, neighbour type, radial index): table sizes of k_lrows_big
        int runmax = 1, tilemax = 1;
        for (const auto& T : hm.types) {
            const int ntile = T.n_fpad / 8;
            std::vector<int> first(d.n_fn + 1, ntile);
            for (int tile = ntile - 1; tile >= 0; --tile) first[T.tile_n[tile]] = tile;
            for (int n = d.n_fn - 1; n >= 0; --n) first[n] = std::min(first[n], first[n + 1]);
            for (int n = 0; n < d.n_fn; ++n) {
                tilemax = std::max(tilemax, first[n + 1] - first[n]);
                for (int u = 0; u < d.n_type; ++u)
                    runmax = std::max(runmax, T.tile_blk_off[u][first[n + 1]] - T.tile_blk_off[u][first[n]]);
            }
        }
        set_lrows_kmax(kmax, runmax, tilemax);
    }
    {   // fast-path template selection for K4a: (tiles per radial index, k-chunks per radial group)
        int tpn = 1;
        for (const auto& T : hm.types) {
            std::vector<int> cnt(d.n_fn, 0);
            for (int n_ : T.tile_n) cnt[n_]++;
            for (int v : cnt) tpn = std::max(tpn, v);
        }
        const int kc = kmax / 4;
        d.kpn = kc <= 2 ? 2 : (kc <= 4 ? 4 : (kc <= 8 ? 8 : 0));
        d.tpn = tpn <= 4 ? tpn : 0;
        if (d.tpn == 0) d.kpn = 0;
        d.dense = 1;
        for (const auto& T : hm.types) d.dense = d.dense && T.dense_blocks;
        // k_lrows_v4 forms its A operands from the Y_lm arrays directly: position inside a radial group == Y_lm key
        d.front2 = d.n_type == 1 ? 1 : 0;
        if (d.front2) {
            const TypeTables& T = hm.types[0];
            for (int n = 0; n < d.n_fn && d.front2; ++n) {
                const int o0 = T.seg_n_off[0][n], o1 = T.seg_n_off[0][n + 1];
                int real = 0;
                for (int pos = o0; pos < o1; ++pos) {
                    const int h = T.seg_heads[0][pos];
                    if (h < 0) continue;
                    ++real;
                    if (T.head_key[h] != pos - o0) d.front2 = 0;
                }
                if (real != 0 && real != hm.n_lm_half) d.front2 = 0;
            }
        }
    }
    std::vector<double> tpp((size_t)d.n_tp * d.n_fn * 2, 0.0);
    std::vector<int> tpn(d.n_tp, 0), tpairs((size_t)d.n_type * d.n_type);
    for (int tp = 0; tp < d.n_tp; ++tp) {
        tpn[tp] = (int)hm.fp.cond[tp].size();
        for (size_t k = 0; k < hm.fp.cond[tp].size(); ++k) {
            tpp[((size_t)tp * d.n_fn + k) * 2] = hm.fp.params[hm.fp.cond[tp][k]][0];
            tpp[((size_t)tp * d.n_fn + k) * 2 + 1] = hm.fp.params[hm.fp.cond[tp][k]][1];
        }
    }
    for (int i = 0; i < d.n_type; ++i)
        for (int j = 0; j < d.n_type; ++j) tpairs[(size_t)i * d.n_type + j] = hm.type_pairs[i][j];
    d.tp_params = upload(c, tpp);
    d.tp_nfn = upload(c, tpn);
    d.type_pairs = upload(c, tpairs);
    d.npv = (int)hm.pv_gid.size();
    d.npv_pad = std::max(8, (d.npv + 7) / 8 * 8);
    std::vector<int> pvfp((size_t)d.n_type * d.npv_pad, -1);
    for (int t = 0; t < d.n_type; ++t)
        for (int a = 0; a < d.npv; ++a) pvfp[(size_t)t * d.npv_pad + a] = hm.pv_fp[t][a];
    d.pv_fp = upload(c, pvfp);
    d.n_pair_terms = (int)hm.pair_terms.size();
    std::vector<int> pt;
    for (const auto& x : hm.pair_terms) { pt.push_back(x[0]); pt.push_back(x[1]); pt.push_back(x[2]); }
    d.pair_terms = upload(c, pt);
    {
        std::vector<int> colof(64 * 64, -1);
        bool ok = hm.pv_gid.size() <= 64;
        for (const auto& x : hm.pair_terms) {
            if (!ok) break;
            const int a = std::min(x[1], x[2]), b2 = std::max(x[1], x[2]);
            if (colof[a * 64 + b2] >= 0) ok = false;   // two columns for one pair: not expected
            colof[a * 64 + b2] = x[0];
        }
        d.pair_colof = ok ? upload(c, colof) : nullptr;
    }
    {
        std::vector<int> linfp((size_t)d.n_type * hm.n_linear, -1), pvl(std::max(hm.n_linear, 1), -1);
        for (int t = 0; t < d.n_type; ++t)
            for (int f = 0; f < hm.types[t].n_feat; ++f)
                linfp[(size_t)t * hm.n_linear + hm.types[t].feat_gid[f]] = hm.types[t].feat_pad[f];
        for (int a = 0; a < d.npv; ++a) pvl[hm.pv_gid[a]] = a;
        d.lin_fp = upload(c, linfp);
        d.pv_of_lin = upload(c, pvl);
    }

    size_t max_full = 1;
    for (int t = 0; t < d.n_type; ++t) {
        const TypeTables& T = hm.types[t];
        DevType& D = d.types[t];
        D.n_full = T.n_full; D.n_head = T.n_head; D.n_feat = T.n_feat; D.n_fpad = T.n_fpad;
        D.n_tiles = T.n_fpad / 8; D.max_order = T.max_order; D.n_ent = (int)T.ent_off.size() - 1;
        D.n_blocks = (int)T.blocks.size(); D.g_size = T.g_size;
        max_full = std::max<size_t>(max_full, T.n_full);
        D.full_head = upload(c, T.full_head);
        std::vector<signed char> fc(T.full_conj.begin(), T.full_conj.end());
        D.full_conj = upload(c, fc);
        D.full_cc = upload(c, T.full_cc);
        D.head_nid = upload(c, T.head_nid);
        D.head_key = upload(c, T.head_key);
        std::vector<int> head_seg(T.n_head, 0);
        for (int h = 0; h < T.n_head; ++h)
            for (int u = 0; u < d.n_type; ++u)
                if (T.seg_tp[u] == T.head_tp[h]) { head_seg[h] = u; break; }
        D.head_seg = upload(c, head_seg);
        std::vector<int> tile_n_off(d.n_fn + 1, 0);
        {
            int tile = 0;
            for (int n = 0; n < d.n_fn; ++n) {
                tile_n_off[n] = tile;
                while (tile < D.n_tiles && T.tile_n[tile] == n) ++tile;
            }
            tile_n_off[d.n_fn] = tile;
            if (tile != D.n_tiles) throw std::runtime_error("feature tiles are not ordered by radial index");
        }
        D.tile_n_off = upload(c, tile_n_off);
        for (int u = 0; u < MAXT; ++u) {
            D.seg_heads[u] = nullptr; D.seg_len[u] = 0; D.seg_key[u] = nullptr; D.seg_n_off[u] = nullptr;
            D.seg_nid[u] = nullptr; D.tile_blk_off[u] = nullptr; D.blkmap[u] = nullptr; D.tile_order[u] = nullptr;
        }
        for (int u = 0; u < d.n_type; ++u) {
            D.seg_heads[u] = upload(c, T.seg_heads[u]);
            D.seg_len[u] = (int)T.seg_heads[u].size();
            std::vector<int> key(T.seg_heads[u].size(), -1);
            for (size_t k = 0; k < key.size(); ++k)
                if (T.seg_heads[u][k] >= 0) key[k] = T.head_key[T.seg_heads[u][k]];
            D.seg_key[u] = upload(c, key);
            D.seg_n_off[u] = upload(c, T.seg_n_off[u]);
            std::vector<int> nid(d.n_fn, -1);
            for (int n = 0; n < d.n_fn; ++n) nid[n] = hm.tp_nid[T.seg_tp[u]][n];
            D.seg_nid[u] = upload(c, nid);
            D.tile_blk_off[u] = upload(c, T.tile_blk_off[u]);
            {
                std::vector<int> ord(D.n_tiles);
                for (int tile = 0; tile < D.n_tiles; ++tile) ord[tile] = tile;
                const auto& tb = T.tile_blk_off[u];
                for (int n = 0; n < d.n_fn; ++n)
                    std::stable_sort(ord.begin() + tile_n_off[n], ord.begin() + tile_n_off[n + 1],
                                     [&](int a, int b2) { return tb[a + 1] - tb[a] > tb[b2 + 1] - tb[b2]; });
                D.tile_order[u] = upload(c, ord);
            }
            D.blkmap[u] = nullptr;
            if (d.kpn > 0) {
                std::vector<int> bm((size_t)D.n_tiles * d.kpn, -1);
                for (int tile = 0; tile < D.n_tiles; ++tile) {
                    const int kc0 = T.seg_n_off[u][T.tile_n[tile]] / 2;
                    for (int bk = T.tile_blk_off[u][tile]; bk < T.tile_blk_off[u][tile + 1]; ++bk)
                        bm[(size_t)tile * d.kpn + (T.blocks[bk].kchunk - kc0)] = bk;
                }
                D.blkmap[u] = upload(c, bm);
            }
        }
        D.term_off = upload(c, T.term_off);
        D.term_coeff = upload(c, T.term_coeff);
        D.term_order = upload(c, T.term_order);
        D.term_ids = upload(c, T.term_ids);
        D.feat_pad = upload(c, T.feat_pad);
        D.ent_pos_re = upload(c, T.ent_pos_re);
        D.ent_pos_im = upload(c, T.ent_pos_im);
        D.ent_off = upload(c, T.ent_off);
        std::vector<DevContribution> cb(T.contribs.size());
        for (size_t k = 0; k < cb.size(); ++k) {
            cb[k].coeff = T.contribs[k].coeff; cb[k].conj = T.contribs[k].conj; cb[k].n_ids = T.contribs[k].n_ids;
            for (int q = 0; q < 5; ++q) cb[k].ids[q] = T.contribs[k].ids[q];
            cb[k].pad = 0;
        }
        D.contribs = upload(c, cb);
        {   // plain tables of the lane = atom eval kernel: packed terms, entries grouped by head
            const int mo = std::max(T.max_order, 1);
            bool ok = mo <= 4 && T.n_full < 32768;
            std::vector<LaItem> lt(T.term_coeff.size());
            std::vector<int> forder(T.n_feat, 1);
            for (int f = 0; f < T.n_feat && ok; ++f)
                for (int ti = T.term_off[f]; ti < T.term_off[f + 1]; ++ti) {
                    const int o = T.term_order[ti];
                    if (ti > T.term_off[f] && o != forder[f]) ok = false;
                    forder[f] = o;
                    unsigned id[4] = {0, 0, 0, 0};
                    for (int q = 0; q < o && q < 4; ++q) id[q] = (unsigned)T.term_ids[(size_t)ti * mo + q];
                    lt[ti].coeff = T.term_coeff[ti]; lt[ti].w0 = id[0] | id[1] << 16; lt[ti].w1 = id[2] | id[3] << 16;
                }
            // head-major flat item lists: every contribution carries its entry's padded feature id, its id count and a
            // "last contribution of the entry" flag, so that a warp streams one contiguous list per head
            const int ne = (int)T.ent_pos_re.size();
            if (c->dm.fl > 4096) ok = false;
            std::vector<LaItem> hitems;
            std::vector<int> hoff{0}, hpos;
            std::map<int, std::vector<LaItem>> head_items;   // head position key -> its contributions, entry by entry
            if (ok) {
                for (int e = 0; e < ne && ok; ++e) {
                    const int pos = T.ent_pos_re[e];
                    const auto& blk = T.blocks[pos / 32];
                    const unsigned fpad_id = (unsigned)(blk.tile * 8 + (pos % 32) / 4);
                    auto& dst = head_items[(blk.seg << 20) | (4 * blk.kchunk + pos % 4)];
                    for (int q = T.ent_off[e]; q < T.ent_off[e + 1]; ++q) {
                        const auto& cbq = T.contribs[q];
                        if (cbq.n_ids > 3) { ok = false; break; }
                        unsigned id[3] = {0, 0, 0};
                        for (int z = 0; z < cbq.n_ids; ++z) id[z] = (unsigned)cbq.ids[z];
                        LaItem it;
                        it.coeff = cbq.coeff;
                        it.w0 = id[0] | id[1] << 16 | (cbq.conj ? 0x80000000u : 0u);
                        it.w1 = id[2] | fpad_id << 15 | (unsigned)cbq.n_ids << 27 | (q + 1 == T.ent_off[e + 1] ? 1u << 30 : 0u);
                        dst.push_back(it);
                    }
                }
            }
            auto item_cost = [](const std::vector<LaItem>& v) {
                long n = 0;
                for (const auto& it : v) n += std::max<int>((it.w1 >> 27) & 7, 1) + 1;
                return n;
            };
            if (ok) {
                std::vector<std::pair<long, int>> cost;   // (-work, key): most expensive head first
                for (const auto& kv : head_items) cost.push_back({-item_cost(kv.second), kv.first});
                std::sort(cost.begin(), cost.end());
                for (const auto& ck : cost) {
                    const auto& v = head_items[ck.second];
                    hitems.insert(hitems.end(), v.begin(), v.end());
                    hoff.push_back((int)hitems.size());
                    hpos.push_back(ck.second);
                }
            }
            D.la_ok = ok ? 1 : 0;
            D.n_la_heads = ok ? (int)hpos.size() : 0;
            D.la_terms = upload(c, lt); D.la_forder = upload(c, forder); D.la_hitems = upload(c, hitems);
            D.la_hoff = upload(c, hoff); D.la_hpos = upload(c, hpos);

            // ---- radial replication: the term lists of radial index n are those of n = 0 with every full id shifted by
            // n * S, the padded feature id by n * Fs and the head position by n * Ps (the usual gtinv model: every
            // product shares one radial index).  Verified item by item; k_eval_features_lb then decodes each item once
            // for a block of radial indices.
            int nr = 0, S = 0, Fs = 0, Ps = 0;
            std::vector<LaItem> bterms, bhitems;
            std::vector<int> bfoff{0}, bforder, bfpad;
            std::vector<int4> bwork, bfwork;
            if (ok && T.n_feat > 0 && getenv("PM_EVAL_NO_RADIAL_BATCH") == nullptr) {
                bool reg = true;
                std::vector<std::vector<int>> fn_(hm.fp.n_fn);
                for (int f = 0; f < T.n_feat; ++f) fn_[T.tile_n[T.feat_pad[f] / 8]].push_back(f);
                nr = 0;
                while (nr < hm.fp.n_fn && !fn_[nr].empty()) ++nr;
                for (int n = nr; n < hm.fp.n_fn; ++n) if (!fn_[n].empty()) reg = false;
                const size_t G = fn_[0].size();
                for (int n = 0; n < nr; ++n) if (fn_[n].size() != G) reg = false;
                if (nr < 2 || G == 0) reg = false;
                auto idk = [&](int ti, int k) { return T.term_ids[(size_t)ti * mo + k]; };
                if (reg) {
                    Fs = T.feat_pad[fn_[1][0]] - T.feat_pad[fn_[0][0]];
                    const int ta = T.term_off[fn_[0][0]], tb = T.term_off[fn_[1][0]];
                    if (T.term_off[fn_[0][0] + 1] == ta || T.term_off[fn_[1][0] + 1] == tb) reg = false;
                    else S = idk(tb, 0) - idk(ta, 0);
                    if (S <= 0 || Fs <= 0) reg = false;
                }
                for (int n = 1; n < nr && reg; ++n)
                    for (size_t g = 0; g < G && reg; ++g) {
                        const int f0 = fn_[0][g], f1 = fn_[n][g];
                        const int t0 = T.term_off[f0], t1 = T.term_off[f1], cnt = T.term_off[f0 + 1] - t0;
                        if (T.feat_pad[f1] - T.feat_pad[f0] != n * Fs || T.term_off[f1 + 1] - t1 != cnt) { reg = false; break; }
                        for (int k = 0; k < cnt && reg; ++k) {
                            if (T.term_coeff[t0 + k] != T.term_coeff[t1 + k] || T.term_order[t0 + k] != T.term_order[t1 + k]) reg = false;
                            for (int z = 0; z < T.term_order[t0 + k] && reg; ++z)
                                if (idk(t1 + k, z) - idk(t0 + k, z) != n * S) reg = false;
                        }
                    }
                // heads of radial index 0 and their images
                auto head_n = [&](int key) {
                    const int u = key >> 20, hp = (key & 0xfffff) / 2;
                    for (int n = 0; n < hm.fp.n_fn; ++n)
                        if (hp >= T.seg_n_off[u][n] && hp < T.seg_n_off[u][n + 1]) return n;
                    return -1;
                };
                std::vector<int> keys0;
                if (reg) {
                    for (const auto& kv : head_items) {
                        const int n = head_n(kv.first);
                        if (n < 0 || n >= nr) { reg = false; break; }
                        if (n == 0) keys0.push_back(kv.first);
                    }
                    if (keys0.empty()) reg = false;
                }
                if (reg) {
                    const int u0 = keys0[0] >> 20;
                    Ps = 2 * (T.seg_n_off[u0][1] - T.seg_n_off[u0][0]);
                    if (Ps <= 0 || head_items.size() != keys0.size() * (size_t)nr) reg = false;
                }
                for (size_t k = 0; k < keys0.size() && reg; ++k)
                    for (int n = 1; n < nr && reg; ++n) {
                        const auto itn = head_items.find(keys0[k] + n * Ps);
                        const auto& v0 = head_items[keys0[k]];
                        if (itn == head_items.end() || itn->second.size() != v0.size()) { reg = false; break; }
                        for (size_t q = 0; q < v0.size() && reg; ++q) {
                            const LaItem &a0 = v0[q], &a1 = itn->second[q];
                            const int cn = (a0.w1 >> 27) & 7;
                            if (a0.coeff != a1.coeff || (a0.w0 >> 31) != (a1.w0 >> 31) || (a0.w1 >> 27) != (a1.w1 >> 27)) reg = false;
                            if ((int)((a1.w1 >> 15) & 0xfffu) - (int)((a0.w1 >> 15) & 0xfffu) != n * Fs) reg = false;
                            const int d0 = (int)(a1.w0 & 0xffffu) - (int)(a0.w0 & 0xffffu);
                            const int d1 = (int)((a1.w0 >> 16) & 0x7fffu) - (int)((a0.w0 >> 16) & 0x7fffu);
                            const int d2 = (int)(a1.w1 & 0x7fffu) - (int)(a0.w1 & 0x7fffu);
                            if ((cn > 0 && d0 != n * S) || (cn > 1 && d1 != n * S) || (cn > 2 && d2 != n * S)) reg = false;
                        }
                    }
                if (reg) {
                    // features of radial index 0, most terms first
                    std::vector<int> gs(G);
                    for (size_t g = 0; g < G; ++g) gs[g] = (int)g;
                    std::stable_sort(gs.begin(), gs.end(), [&](int x, int y) {
                        return T.term_off[fn_[0][x] + 1] - T.term_off[fn_[0][x]] > T.term_off[fn_[0][y] + 1] - T.term_off[fn_[0][y]];
                    });
                    const int tchunk = getenv("PM_LB_TCHUNK") ? atoi(getenv("PM_LB_TCHUNK")) : 16;
                    for (int g : gs) {
                        const int f0 = fn_[0][g];
                        const int tb0 = (int)bterms.size();
                        for (int ti = T.term_off[f0]; ti < T.term_off[f0 + 1]; ++ti) bterms.push_back(lt[ti]);
                        bfoff.push_back((int)bterms.size());
                        bforder.push_back(forder[f0]);
                        bfpad.push_back(T.feat_pad[f0]);
                        // stage-1 work items: chunks of <= tchunk terms of one feature (partial sums are added in shared memory)
                        for (int q = tb0; q < (int)bterms.size(); q += tchunk)
                            bfwork.push_back(make_int4(q, std::min(q + tchunk, (int)bterms.size()), (int)bforder.size() - 1, 0));
                    }
                    std::stable_sort(bfwork.begin(), bfwork.end(), [](const int4& x, const int4& y) { return x.y - x.x > y.y - y.x; });
                    // head work list: chunks of >= hchunk items that end on entry boundaries, most expensive head first
                    const int hchunk = getenv("PM_LB_HCHUNK") ? atoi(getenv("PM_LB_HCHUNK")) : 24;
                    std::vector<std::pair<long, int>> cost;
                    for (int key : keys0) cost.push_back({-item_cost(head_items[key]), key});
                    std::sort(cost.begin(), cost.end());
                    for (const auto& ck : cost) {
                        const auto& v = head_items[ck.second];
                        int q0 = (int)bhitems.size(), cnt = 0;
                        for (size_t q = 0; q < v.size(); ++q) {
                            bhitems.push_back(v[q]);
                            ++cnt;
                            const bool last = (v[q].w1 >> 30) & 1u;
                            if ((last && cnt >= hchunk) || q + 1 == v.size()) {
                                bwork.push_back(make_int4(q0, (int)bhitems.size(), ck.second, 0));
                                q0 = (int)bhitems.size();
                                cnt = 0;
                            }
                        }
                    }
                    std::stable_sort(bwork.begin(), bwork.end(), [](const int4& x, const int4& y) { return x.y - x.x > y.y - y.x; });
                } else {
                    nr = 0;
                }
            }
            D.lb_nr = nr; D.lb_S = S; D.lb_Fs = Fs; D.lb_Ps = Ps;
            D.lb_G = (int)bforder.size(); D.lb_nwork = (int)bwork.size();
            D.lb_terms = upload(c, bterms); D.lb_foff = upload(c, bfoff); D.lb_forder = upload(c, bforder);
            D.lb_fpad = upload(c, bfpad); D.lb_hitems = upload(c, bhitems); D.lb_work = upload(c, bwork);
            D.lb_fwork = upload(c, bfwork); D.lb_nfwork = (int)bfwork.size();
            if (getenv("PM_DEBUG_TABLES"))
                fprintf(stderr, "[pm] type %d: la_ok %d, radial batch nr %d S %d Fs %d Ps %d, %zu features (%zu terms, %zu chunks), %zu head items in %zu chunks\n", t, D.la_ok,
                        nr, S, Fs, Ps, bforder.size(), bterms.size(), bfwork.size(), bhitems.size(), bwork.size());
        }
        // sliced tables (k_features_v3): `feats` / `ents` select the features and G entries that are sliced
        struct Sliced {
            std::vector<double> scoef;
            std::vector<unsigned> flat;
            std::vector<int4> fmeta, emeta;
            std::vector<int> fout;
            std::vector<int2> eout, efh;
            bool ids_fit = true;
        };
        auto build_slices = [&](const std::vector<int>& feats, const std::vector<int>& ents, Sliced& out) {
            const int mo = std::max(T.max_order, 1);
            const int nw = (mo + 1) / 2;
            std::vector<double> scoef;
            std::vector<std::vector<unsigned>> sw(nw);
            auto grow = [&](size_t n) {
                scoef.resize(n, 0.0);
                for (auto& w : sw) w.resize(n, 0u);
            };
            std::vector<int4> fmeta, emeta;
            std::vector<int> fout;
            std::vector<int2> eout, efh;
            bool ids_fit = T.n_full < 32768;
            // features: rows of 8, 4 k-lanes
            {
                const int n_feat_sel = (int)feats.size();
                std::vector<int> order_of(T.n_feat, 1), idx(feats);
                for (int f : feats) {
                    int o = 1;
                    for (int ti = T.term_off[f]; ti < T.term_off[f + 1]; ++ti) o = std::max(o, T.term_order[ti]);
                    order_of[f] = o;
                }
                std::sort(idx.begin(), idx.end(), [&](int a, int b2) {
                    if (order_of[a] != order_of[b2]) return order_of[a] < order_of[b2];
                    const int na = T.term_off[a + 1] - T.term_off[a], nb = T.term_off[b2 + 1] - T.term_off[b2];
                    if (na != nb) return na > nb;
                    return a < b2;
                });
                for (int p0 = 0; p0 < n_feat_sel;) {
                    int p1 = p0;
                    while (p1 < n_feat_sel && p1 - p0 < 8 && order_of[idx[p1]] == order_of[idx[p0]]) ++p1;
                    int mx = 0;
                    for (int p = p0; p < p1; ++p) mx = std::max(mx, T.term_off[idx[p] + 1] - T.term_off[idx[p]]);
                    const int iters = (mx + 3) / 4;
                    const size_t base = scoef.size();
                    grow(base + (size_t)iters * 32);
                    fmeta.push_back(make_int4((int)base, iters, order_of[idx[p0]], 0));
                    for (int r = 0; r < 8; ++r) fout.push_back(p0 + r < p1 ? T.feat_pad[idx[p0 + r]] : -1);
                    for (int p = p0; p < p1; ++p) {
                        const int f = idx[p], row = p - p0;
                        for (int ti = T.term_off[f], k = 0; ti < T.term_off[f + 1]; ++ti, ++k) {
                            // slot of k-th term: iteration k / 4, k-lane k % 4 -> lane = (k % 4) * 8 + row
                            const size_t slot = base + (size_t)(k / 4) * 32 + (k % 4) * 8 + row;
                            scoef[slot] = T.term_coeff[ti];
                            const int o = T.term_order[ti];
                            if (o != order_of[f]) ids_fit = false;   // mixed orders inside a feature: keep v2
                            for (int q = 0; q < o && q < mo; ++q)
                                sw[q / 2][slot] |= (unsigned)T.term_ids[(size_t)ti * mo + q] << (16 * (q & 1));
                        }
                    }
                    p0 = p1;
                }
            }
            // G entries: 32 per slice
            {
                std::vector<int> idx(ents), cn_of(T.ent_pos_re.size(), 0);
                for (int e : ents) {
                    int cn = 0;
                    for (int q = T.ent_off[e]; q < T.ent_off[e + 1]; ++q) {
                        if (q > T.ent_off[e] && T.contribs[q].n_ids != cn) ids_fit = false;
                        cn = T.contribs[q].n_ids;
                    }
                    cn_of[e] = cn;
                }
                std::sort(idx.begin(), idx.end(), [&](int a, int b2) {
                    if (cn_of[a] != cn_of[b2]) return cn_of[a] < cn_of[b2];
                    const int na = T.ent_off[a + 1] - T.ent_off[a], nb = T.ent_off[b2 + 1] - T.ent_off[b2];
                    if (na != nb) return na > nb;
                    return a < b2;
                });
                const int ne = (int)idx.size();
                for (int p0 = 0; p0 < ne;) {
                    int p1 = p0;
                    while (p1 < ne && p1 - p0 < 32 && cn_of[idx[p1]] == cn_of[idx[p0]]) ++p1;
                    int mx = 0;
                    for (int p = p0; p < p1; ++p) mx = std::max(mx, T.ent_off[idx[p] + 1] - T.ent_off[idx[p]]);
                    const size_t base = scoef.size();
                    grow(base + (size_t)mx * 32);
                    emeta.push_back(make_int4((int)base, mx, cn_of[idx[p0]], 0));
                    for (int r = 0; r < 32; ++r) {
                        eout.push_back(p0 + r < p1 ? make_int2(T.ent_pos_re[idx[p0 + r]], T.ent_pos_im[idx[p0 + r]])
                                                   : make_int2(-1, -1));
                        int2 fh = make_int2(-1, -1);
                        if (p0 + r < p1) {   // pos_re = 32 * block + 4 * (feature inside the tile) + 2 * (head inside the block)
                            const int pos = T.ent_pos_re[idx[p0 + r]];
                            const auto& blk = T.blocks[pos / 32];
                            fh = make_int2(blk.tile * 8 + (pos % 32) / 4, (blk.seg << 20) | (4 * blk.kchunk + pos % 4));
                        }
                        efh.push_back(fh);
                    }
                    for (int p = p0; p < p1; ++p) {
                        const int e = idx[p], row = p - p0;
                        for (int q = T.ent_off[e], k = 0; q < T.ent_off[e + 1]; ++q, ++k) {
                            const size_t slot = base + (size_t)k * 32 + row;
                            scoef[slot] = T.contribs[q].coeff;
                            for (int z = 0; z < T.contribs[q].n_ids && z < 2 * nw; ++z)
                                sw[z / 2][slot] |= (unsigned)T.contribs[q].ids[z] << (16 * (z & 1));
                            if (T.contribs[q].conj) sw[0][slot] |= 0x80000000u;
                        }
                    }
                    p0 = p1;
                }
            }
            // slice order: longest first so that the round-robin over warps is balanced
            auto by_len = [](const int4& a, const int4& b2) { return a.y > b2.y; };
            {
                std::vector<int> perm(fmeta.size());
                for (size_t k = 0; k < perm.size(); ++k) perm[k] = (int)k;
                std::stable_sort(perm.begin(), perm.end(), [&](int a, int b2) { return by_len(fmeta[a], fmeta[b2]); });
                std::vector<int4> fm2; std::vector<int> fo2;
                for (int k : perm) { fm2.push_back(fmeta[k]); fo2.insert(fo2.end(), fout.begin() + 8 * k, fout.begin() + 8 * k + 8); }
                fmeta.swap(fm2); fout.swap(fo2);
                perm.resize(emeta.size());
                for (size_t k = 0; k < perm.size(); ++k) perm[k] = (int)k;
                std::stable_sort(perm.begin(), perm.end(), [&](int a, int b2) { return by_len(emeta[a], emeta[b2]); });
                std::vector<int4> em2; std::vector<int2> eo2, ef2;
                for (int k : perm) {
                    em2.push_back(emeta[k]);
                    eo2.insert(eo2.end(), eout.begin() + 32 * k, eout.begin() + 32 * k + 32);
                    ef2.insert(ef2.end(), efh.begin() + 32 * k, efh.begin() + 32 * k + 32);
                }
                emeta.swap(em2); eout.swap(eo2); efh.swap(ef2);
            }
            out.ids_fit = ids_fit;
            out.scoef.swap(scoef);
            for (auto& w : sw) out.flat.insert(out.flat.end(), w.begin(), w.end());
            out.fmeta.swap(fmeta); out.emeta.swap(emeta); out.fout.swap(fout); out.eout.swap(eout); out.efh.swap(efh);
        };
        {
            std::vector<int> feats(T.n_feat), ents(T.ent_pos_re.size());
            for (int f = 0; f < T.n_feat; ++f) feats[f] = f;
            for (size_t e = 0; e < ents.size(); ++e) ents[e] = (int)e;
            Sliced sl;
            build_slices(feats, ents, sl);
            D.n_fsl = sl.ids_fit ? (int)sl.fmeta.size() : 0;
            D.n_esl = sl.ids_fit ? (int)sl.emeta.size() : 0;
            D.sl_words = (std::max(T.max_order, 1) + 1) / 2;
            D.n_slots = (long)sl.scoef.size();
            D.fsl_meta = upload(c, sl.fmeta); D.fsl_out = upload(c, sl.fout);
            D.esl_meta = upload(c, sl.emeta); D.esl_out = upload(c, sl.eout); D.esl_fh = upload(c, sl.efh);
            D.sl_coeff = upload(c, sl.scoef); D.sl_ids = upload(c, sl.flat);
            if (getenv("PM_DEBUG_TABLES"))
                fprintf(stderr, "[pm] type %d: %d feature slices, %d entry slices, %ld slots (terms %zu, contribs %zu)\n",
                        t, D.n_fsl, D.n_esl, D.n_slots, T.term_coeff.size(), T.contribs.size());
        }
        {   // ---- radial replication for the sliced K3 tables of LARGE models (k_features_v4r) ------------------------
            // The term lists and G entries of radial index n are those of radial index 0 with every full id shifted by
            // n * S, the padded feature id by n * Fs, the head position by n * Ps(segment) and the G position by n * Gs.
            // Verified entry by entry; then only the radial-0 features / entries are sliced (tables 1 / n_radial of the
            // size: the big K3 kernel is bound by streaming its ~15 MB of tables per type from L2 for every CTA).
            int nr = 0, S = 0, Fs = 0, Gs = 0;
            std::vector<int> r0_feats, r0_ents, gs_u;
            const int mo = std::max(T.max_order, 1);
            const int n_fn = hm.fp.n_fn;
            const int ne = (int)T.ent_pos_re.size();
            bool reg = T.n_feat > 0 && mo <= 4 && getenv("PM_FEAT_NO_RADIAL") == nullptr;
            int why = 0;   // first failed check (PM_DEBUG_TABLES)
#define PM_RFAIL(code) { if (reg) why = code; reg = false; }
            std::vector<std::vector<int>> fn_(n_fn);
            if (reg) {
                for (int f = 0; f < T.n_feat; ++f) fn_[T.tile_n[T.feat_pad[f] / 8]].push_back(f);
                while (nr < n_fn && !fn_[nr].empty()) ++nr;
                for (int n = nr; n < n_fn; ++n) if (!fn_[n].empty()) PM_RFAIL(1)
                for (int n = 1; n < nr; ++n) if (fn_[n].size() != fn_[0].size()) PM_RFAIL(2)
                if (nr < 2) PM_RFAIL(3)
            }
            auto idk = [&](int ti, int k) { return T.term_ids[(size_t)ti * mo + k]; };
            if (reg) {
                Fs = T.feat_pad[fn_[1][0]] - T.feat_pad[fn_[0][0]];
                const int ta = T.term_off[fn_[0][0]], tb = T.term_off[fn_[1][0]];
                if (T.term_off[fn_[0][0] + 1] == ta || T.term_off[fn_[1][0] + 1] == tb) PM_RFAIL(4)
                else S = idk(tb, 0) - idk(ta, 0);
                if (S <= 0 || Fs <= 0) PM_RFAIL(5)
            }
            for (int n = 1; n < nr && reg; ++n)
                for (size_t g = 0; g < fn_[0].size() && reg; ++g) {
                    const int f0 = fn_[0][g], f1 = fn_[n][g];
                    const int t0 = T.term_off[f0], t1 = T.term_off[f1], cnt = T.term_off[f0 + 1] - t0;
                    if (T.feat_pad[f1] - T.feat_pad[f0] != n * Fs || T.term_off[f1 + 1] - t1 != cnt) { PM_RFAIL(6) break; }
                    for (int k = 0; k < cnt && reg; ++k) {
                        if (T.term_coeff[t0 + k] != T.term_coeff[t1 + k] || T.term_order[t0 + k] != T.term_order[t1 + k]) PM_RFAIL(7)
                        for (int z = 0; z < T.term_order[t0 + k] && reg; ++z)
                            if (idk(t1 + k, z) - idk(t0 + k, z) != n * S) PM_RFAIL(8)
                    }
                }
            if (reg) {
                auto ent_key = [&](int e) {   // (padded feature id, segment, head position inside the segment)
                    const int pos = T.ent_pos_re[e];
                    const auto& blk = T.blocks[pos / 32];
                    return std::array<int, 3>{blk.tile * 8 + (pos % 32) / 4, blk.seg, 2 * blk.kchunk + (pos % 4) / 2};
                };
                std::map<std::array<int, 3>, int> ent_of;
                for (int e = 0; e < ne; ++e) ent_of[ent_key(e)] = e;
                gs_u.assign(d.n_type, 0);
                for (int e = 0; e < ne && reg; ++e) {
                    const auto k0 = ent_key(e);
                    if (T.tile_n[k0[0] / 8] != 0) continue;
                    r0_ents.push_back(e);
                    const int u = k0[1];
                    const int ps = T.seg_n_off[u][1] - T.seg_n_off[u][0];   // heads per radial group of this segment
                    for (int n = 1; n < nr && reg; ++n) {
                        if (T.seg_n_off[u][n + 1] - T.seg_n_off[u][n] != ps) { PM_RFAIL(9) break; }
                        const auto it = ent_of.find({k0[0] + n * Fs, u, k0[2] + n * ps});
                        if (it == ent_of.end()) { PM_RFAIL(10) break; }
                        const int e1 = it->second;
                        // blocks are sorted by (segment, tile, k-chunk): the G stride per radial index is per segment
                        const int d = T.ent_pos_re[e1] - T.ent_pos_re[e];
                        if (gs_u[u] == 0 && n == 1) gs_u[u] = d;
                        if (d != n * gs_u[u] || T.ent_pos_im[e1] - T.ent_pos_im[e] != n * gs_u[u]) PM_RFAIL(11)
                        const int c0 = T.ent_off[e], c1 = T.ent_off[e1], cnt = T.ent_off[e + 1] - c0;
                        if (T.ent_off[e1 + 1] - c1 != cnt) { PM_RFAIL(12) break; }
                        for (int q = 0; q < cnt && reg; ++q) {
                            const auto &a0 = T.contribs[c0 + q], &a1 = T.contribs[c1 + q];
                            if (a0.coeff != a1.coeff || a0.conj != a1.conj || a0.n_ids != a1.n_ids || a0.n_ids > 3) PM_RFAIL(13)
                            for (int z = 0; z < a0.n_ids && reg; ++z)
                                if (a1.ids[z] - a0.ids[z] != n * S) PM_RFAIL(14)
                        }
                    }
                }
                for (int e : r0_ents) if (gs_u[T.blocks[T.ent_pos_re[e] / 32].seg] <= 0) PM_RFAIL(16)
                if (!gs_u.empty()) Gs = gs_u[0];
                if ((size_t)r0_ents.size() * nr != (size_t)ne) PM_RFAIL(15)
            }
            Sliced sl;
            if (reg) build_slices(fn_[0], r0_ents, sl);
            const bool okr = reg && sl.ids_fit && !sl.fmeta.empty();
            D.r_nr = okr ? nr : 0; D.r_S = S; D.r_Fs = Fs; D.r_Gs = Gs;
            D.r_n_fsl = okr ? (int)sl.fmeta.size() : 0;
            D.r_n_esl = okr ? (int)sl.emeta.size() : 0;
            D.r_n_slots = (long)sl.scoef.size();
            D.r_fsl_meta = upload(c, sl.fmeta); D.r_fsl_out = upload(c, sl.fout);
            std::vector<int4> eout4(sl.eout.size());   // (pos_re, pos_im, G stride per radial index of the entry's segment, 0)
            for (size_t k = 0; k < eout4.size(); ++k) {
                const int2 pz = sl.eout[k];
                eout4[k] = make_int4(pz.x, pz.y, pz.x >= 0 && okr ? gs_u[T.blocks[pz.x / 32].seg] : 0, 0);
            }
            D.r_esl_meta = upload(c, sl.emeta); D.r_esl_out = upload(c, eout4);
            D.r_sl_coeff = upload(c, sl.scoef); D.r_sl_ids = upload(c, sl.flat);
            if (getenv("PM_DEBUG_TABLES"))
                fprintf(stderr, "[pm] type %d: radial-0 slices: nr %d S %d Fs %d Gs %d, %d feature slices, %d entry slices, %ld slots (failed check %d, slices fit %d)\n",
                        t, D.r_nr, S, Fs, Gs, D.r_n_fsl, D.r_n_esl, D.r_n_slots, why, (int)sl.ids_fit);
#undef PM_RFAIL
        }
        std::vector<int> bk(T.blocks.size());
        for (size_t k = 0; k < bk.size(); ++k) bk[k] = T.blocks[k].kchunk;
        D.blk_kchunk = upload(c, bk);
        D.pad_gid = upload(c, T.pad_gid);
        {
            std::vector<int> gid2pv(std::max(hm.n_linear, 1), -1), padpv(T.n_fpad, -1);
            for (int a = 0; a < (int)hm.pv_gid.size(); ++a) gid2pv[hm.pv_gid[a]] = a;
            for (int fp_ = 0; fp_ < T.n_fpad; ++fp_)
                if (T.pad_gid[fp_] >= 0) padpv[fp_] = gid2pv[T.pad_gid[fp_]];
            D.pad_pv = upload(c, padpv);
        }
        std::vector<DevPolyTerm> ct(hm.n_variables);
        for (int col = 0; col < hm.n_variables; ++col) {
            const PolyTerm& p = hm.colterm[t][col];
            ct[col] = {p.order, p.fp[0], p.fp[1], p.fp[2]};
        }
        D.colterm = upload(c, ct);
    }
    c->feat_smem = max_full * sizeof(double2);
    if (c->feat_smem > 200 * 1024) throw std::invalid_argument("model too large: a_nlm array exceeds shared memory");
    if (c->feat_smem > 48 * 1024) set_features_smem(c->feat_smem);
    const bool force_simple = (c->flags & PM_FLAG_SIMPLE_KERNELS) != 0;
    c->simple_s = force_simple;
    c->simple_l = force_simple || (lrows_mma_smem(d) > 200 * 1024 && !lrows_big_supported(d));
    c->simple_x = force_simple || hm.has_order3 || xrows_mma_smem(d) > 200 * 1024;
    c->scatter = !c->simple_l && !c->simple_x && scatter_mode_supported(d) && (c->flags & PM_FLAG_SCATTER);
}

// ------------------------------------------------------------------------------------------------
static void validate_structures(const pm_context* c, const pm_structures* st) {
    if (!st || st->n_st < 0) throw std::invalid_argument("invalid structure batch");
    if (st->n_st > 0 && (!st->axis || !st->positions_c || !st->types || !st->n_atoms))
        throw std::invalid_argument("null pointer in structure batch");
    size_t off = 0;
    for (int s = 0; s < st->n_st; ++s) {
        if (st->n_atoms[s] < 0) throw std::invalid_argument("negative atom count");
        for (int a = 0; a < st->n_atoms[s]; ++a) {
            const int t = st->types[off + a];
            if (t < 0 || t >= c->dm.n_type) throw std::invalid_argument("atom type out of range");
        }
        off += st->n_atoms[s];
    }
}

static double est_bytes_per_structure(const pm_context* c, const double* axis, int n_atoms, bool force) {
    const DevModel& d = c->dm;
    const double* a = axis;
    const double vol = std::fabs(a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                                 a[2] * (a[3] * a[7] - a[4] * a[6]));
    const double rc = d.cutoff;
    const double nb = vol > 0 ? n_atoms / vol * 4.18879 * rc * rc * rc * 1.1 + 4 : 64;
    const double pairs = nb * n_atoms;
    double bytes = pairs * (d.pbstride * 8.0 + 12.0);
    bytes += n_atoms * (d.hmax * 16.0 * 10 + d.fl * 8.0 * 10 + 64);
    if (force) {
        bytes += (pairs + n_atoms) * 3.0 * (c->scatter ? d.npv_pad : d.fl) * 8.0;
        bytes += (double)n_atoms * d.gstride * 8.0;
        bytes += (7.0 + 3.0 * n_atoms) * d.fpad * 8.0;
    } else {
        bytes += d.fpad * 8.0;
    }
    return bytes;
}

// Prepares structures [s0, s1) of the batch on the host (translations, cell reduction, row maps).
static void prepare_chunk(const pm_context* c, const pm_structures* st, const std::vector<size_t>& aoff,
                          const std::vector<long>& be, const std::vector<long>& bs, const std::vector<long>& bf,
                          int s0, int s1, const double* w, const double* y, HostChunk& h) {
    h = HostChunk();
    h.n_st = s1 - s0;
    h.atom_off.push_back(0);
    h.trans_off.push_back(0);
    // lattice translations / cell reduction of every structure: independent, a few host threads for large chunks (the
    // E/F/S path spends as long here as the GPU needs for the chunk otherwise)
    const int nst = s1 - s0;
    std::vector<CellTranslations> cts(nst);
    std::vector<std::vector<double>> poss(nst);
    {
        auto work = [&](int k0, int k1) {
            for (int k = k0; k < k1; ++k) {
                const int s = s0 + k, na = st->n_atoms[s];
                poss[k].assign(st->positions_c + 3 * aoff[s], st->positions_c + 3 * aoff[s] + 3 * (size_t)na);
                find_translations(st->axis + 9 * (size_t)s, poss[k].data(), na, c->dm.cutoff, cts[k]);
            }
        };
        const int nthr = nst >= 16 ? (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
        if (nthr <= 1) {
            work(0, nst);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; ++t) th.emplace_back(work, (int)((long)nst * t / nthr), (int)((long)nst * (t + 1) / nthr));
            for (auto& t : th) t.join();
        }
    }
    {
        const size_t natom = aoff[s1] - aoff[s0];
        h.x.reserve(natom); h.y.reserve(natom); h.z.reserve(natom); h.types.reserve(natom); h.st_of_atom.reserve(natom);
    }
    for (int s = s0; s < s1; ++s) {
        const int na = st->n_atoms[s];
        const std::vector<double>& pos = poss[s - s0];
        const CellTranslations& ct = cts[s - s0];
        h.x.insert(h.x.end(), pos.begin(), pos.begin() + na);
        h.y.insert(h.y.end(), pos.begin() + na, pos.begin() + 2 * (size_t)na);
        h.z.insert(h.z.end(), pos.begin() + 2 * (size_t)na, pos.end());
        h.types.insert(h.types.end(), st->types + aoff[s], st->types + aoff[s] + na);
        h.st_of_atom.insert(h.st_of_atom.end(), na, s - s0);
        h.trans.insert(h.trans.end(), ct.trans.begin(), ct.trans.end());
        h.trans_off.push_back((int)(h.trans.size() / 3));
        {   // cell-list parameters of the structure: inverse axis, bins of >= r_c / 2 per lattice direction, search range
            const double* A = ct.axis;
            const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                               A[2] * (A[3] * A[7] - A[4] * A[6]);
            double inv[9] = {0};
            bool ok = std::fabs(det) > 1e-300 && na > 0;
            if (ok) {
                inv[0] = (A[4] * A[8] - A[5] * A[7]) / det; inv[1] = (A[2] * A[7] - A[1] * A[8]) / det; inv[2] = (A[1] * A[5] - A[2] * A[4]) / det;
                inv[3] = (A[5] * A[6] - A[3] * A[8]) / det; inv[4] = (A[0] * A[8] - A[2] * A[6]) / det; inv[5] = (A[2] * A[3] - A[0] * A[5]) / det;
                inv[6] = (A[3] * A[7] - A[4] * A[6]) / det; inv[7] = (A[1] * A[6] - A[0] * A[7]) / det; inv[8] = (A[0] * A[4] - A[1] * A[3]) / det;
            }
            const double rc = c->dm.cutoff;
            int nb[3] = {1, 1, 1}, R[3] = {1, 1, 1};
            double hgt[3] = {1, 1, 1};
            for (int d = 0; d < 3 && ok; ++d) {
                const double nrm = std::sqrt(inv[3 * d] * inv[3 * d] + inv[3 * d + 1] * inv[3 * d + 1] + inv[3 * d + 2] * inv[3 * d + 2]);
                hgt[d] = 1.0 / nrm;   // spacing of the lattice planes along direction d
                nb[d] = std::max(1, std::min(1024, (int)std::floor(hgt[d] / (0.5 * rc * (1.0 + 1e-6)))));
            }
            while (ok && (long)nb[0] * nb[1] * nb[2] > 8L * na + 64) {   // sparse cells: do not build more bins than atoms
                const int dmax = nb[0] >= nb[1] && nb[0] >= nb[2] ? 0 : (nb[1] >= nb[2] ? 1 : 2);
                nb[dmax] = std::max(1, nb[dmax] / 2);
            }
            for (int d = 0; d < 3 && ok; ++d) R[d] = (int)std::ceil(rc / (hgt[d] / nb[d]) * (1.0 + 1e-6));
            const int vals[CL_NI] = {nb[0], nb[1], nb[2], R[0], R[1], R[2], ct.mx[0], ct.mx[1], ct.mx[2], (int)h.cl_tmap.size(),
                                     h.n_bins, ok ? 1 : 0};
            h.cl_int.insert(h.cl_int.end(), vals, vals + CL_NI);
            h.cl_ainv.insert(h.cl_ainv.end(), inv, inv + 9);
            h.cl_tmap.insert(h.cl_tmap.end(), ct.tmap.begin(), ct.tmap.end());
            h.n_bins += nb[0] * nb[1] * nb[2];
        }
        h.max_trans = std::max(h.max_trans, (int)(ct.trans.size() / 3));
        h.atom_off.push_back(h.atom_off.back() + na);
        h.max_atoms = std::max(h.max_atoms, na);
        h.force.push_back(st->force ? (st->force[s] != 0) : 0);
    }
    h.n_atoms = h.atom_off.back();
    {   // cell list or masked sweep: estimated distance tests of either, and the sort key must fit 31 bits
        double cost_cl = 0.0, cost_sweep = 0.0;
        bool ok = getenv("PM_NO_CELL_LIST") == nullptr;
        for (int k = 0; k < h.n_st; ++k) {
            const int* ci = &h.cl_int[(size_t)CL_NI * k];
            const double na = h.atom_off[k + 1] - h.atom_off[k];
            const double nt_ = h.trans_off[k + 1] - h.trans_off[k];
            const double bins = (double)ci[0] * ci[1] * ci[2];
            cost_cl += na * (2.0 * ci[3] + 1) * (2.0 * ci[4] + 1) * (2.0 * ci[5] + 1) * (na / bins + 2.0);
            cost_sweep += na * na * nt_;
            ok = ok && (ci[11] != 0 || na == 0);
        }
        h.cl_nmax = std::max(1, h.max_atoms);
        h.cl_tmax = std::max(1, h.max_trans);
        const double keys = (double)c->dm.n_type * h.cl_nmax * h.cl_tmax;
        // a cell-list candidate (table lookups, scattered loads, two passes) costs ~20 sweep tests (regular, one pass with
        // stored hit masks): measured break-even between config 2 (256 atoms x 33 translations: sweep 5.0 vs 5.6 us) and
        // config 5 (512 atoms x 57: cell list 2.7 vs 3.9 ms per 131072 atoms)
        h.use_cl = ok && keys < 2.0e9 && 20.0 * cost_cl < cost_sweep;
        if (const char* e = getenv("PM_CELL_LIST")) h.use_cl = ok && keys < 2.0e9 && atoi(e) != 0;
    }
    // chunk-local PyModel layout
    int ns_force = 0;
    for (int f : h.force) ns_force += f;
    int is = h.n_st, ifo = h.n_st + 6 * ns_force;
    for (int k = 0; k < h.n_st; ++k) {
        h.erow.push_back(k);
        if (h.force[k]) {
            h.srow.push_back(is); is += 6;
            h.frow.push_back(ifo); ifo += 3 * (h.atom_off[k + 1] - h.atom_off[k]);
        } else {
            h.srow.push_back(-1); h.frow.push_back(-1);
        }
        h.brow_e.push_back(be[s0 + k]); h.brow_s.push_back(bs[s0 + k]); h.brow_f.push_back(bf[s0 + k]);
    }
    h.n_rows = ifo;
    h.w.assign(h.n_rows, 1.0);
    h.yv.assign(h.n_rows, 0.0);
    h.we.assign(h.n_st, 1.0);
    if (w && y) {
        for (int k = 0; k < h.n_st; ++k) {
            // the energy rows leave K4b unweighted (device weight 1): k_xe_reduce takes the unweighted column sums
            // xe_sum / xe_sq_sum from them in structure order and applies the weight afterwards
            h.we[k] = w[h.brow_e[k]]; h.yv[h.erow[k]] = y[h.brow_e[k]];
            if (h.force[k]) {
                for (int r = 0; r < 6; ++r) { h.w[h.srow[k] + r] = w[h.brow_s[k] + r]; h.yv[h.srow[k] + r] = y[h.brow_s[k] + r]; }
                const int nf = 3 * (h.atom_off[k + 1] - h.atom_off[k]);
                for (int r = 0; r < nf; ++r) { h.w[h.frow[k] + r] = w[h.brow_f[k] + r]; h.yv[h.frow[k] + r] = y[h.brow_f[k] + r]; }
            }
        }
    }
}

static void batch_layout(const pm_structures* st, std::vector<size_t>& aoff, std::vector<long>& be,
                         std::vector<long>& bs, std::vector<long>& bf, long& n_rows) {
    const int n = st->n_st;
    aoff.assign(n + 1, 0);
    for (int s = 0; s < n; ++s) aoff[s + 1] = aoff[s] + st->n_atoms[s];
    be.assign(n, 0); bs.assign(n, -1); bf.assign(n, -1);
    long is = n;
    for (int s = 0; s < n; ++s) { be[s] = s; if (st->force && st->force[s]) { bs[s] = is; is += 6; } }
    long ifo = is;
    for (int s = 0; s < n; ++s) if (st->force && st->force[s]) { bf[s] = ifo; ifo += 3L * st->n_atoms[s]; }
    n_rows = ifo;
}

static std::vector<std::pair<int, int>> plan_chunks(const pm_context* c, const pm_structures* st, double bytes_scale = 1.0) {
    std::vector<std::pair<int, int>> chunks;
    int s0 = 0;
    double bytes = 0.0;
    // Large models (MBs of G and L per atom) would get chunks of a few hundred atoms out of the default workspace:
    // too few CTAs for 148 SMs.  Unless the caller fixed the workspace, grow it to ~2400 atoms per chunk (bounded).
    double cap = (double)c->ws_cap;
    if (c->ws_cap_max > c->ws_cap && st->n_st > 0 && st->n_atoms[0] > 0) {
        const double b0 = est_bytes_per_structure(c, st->axis, st->n_atoms[0], st->force && st->force[0]);
        cap = std::min((double)c->ws_cap_max, std::max(cap, b0 / st->n_atoms[0] * 2400.0));
    }
    // balanced chunks: the greedy count of chunks first, then the same walk with the capacity lowered to total / count
    // (no tiny tail chunk: a 2-structure tail costs as many launches and round trips as a full chunk)
    std::vector<double> bs_of(st->n_st);
    double total = 0.0;
    int n_chunks = 1;
    for (int s = 0; s < st->n_st; ++s) {
        bs_of[s] = bytes_scale * est_bytes_per_structure(c, st->axis + 9 * (size_t)s, st->n_atoms[s], st->force && st->force[s]);
        total += bs_of[s];
        if (s > s0 && bytes + bs_of[s] > cap) { ++n_chunks; s0 = s; bytes = 0.0; }
        bytes += bs_of[s];
    }
    const double cap_bal = std::min(cap, total / n_chunks * 1.02 + 1.0);
    s0 = 0; bytes = 0.0;
    for (int s = 0; s < st->n_st; ++s) {
        const double bs = bs_of[s];
        if (s > s0 && (bytes + bs > cap || (bytes + bs > cap_bal && (int)chunks.size() + 1 < n_chunks))) {
            chunks.push_back({s0, s});
            s0 = s;
            bytes = 0.0;
        }
        bytes += bs;
    }
    if (st->n_st > s0) chunks.push_back({s0, st->n_st});
    return chunks;
}

template <typename T> static void h2d(DevVec<T>& d, const std::vector<T>& h, cudaStream_t s) {
    d.ensure(std::max<size_t>(h.size(), 1));
    if (!h.empty()) CK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

struct StageTimer {
    pm_context* c;
    int idx = 0;
    explicit StageTimer(pm_context* c_) : c(c_) {
        if (c->profile) CK(cudaEventRecord(c->ev[0], c->stream));
    }
    void mark(int stage, int launches) {
        c->launches += launches;
        c->stage_launches[stage] += launches;
        if (!c->profile) return;
        CK(cudaEventRecord(c->ev[idx + 1], c->stream));
        CK(cudaEventSynchronize(c->ev[idx + 1]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev[idx], c->ev[idx + 1]));
        c->stage_ms[stage] += ms;
        idx = (idx + 1) % ST_COUNT;
        if (idx == 0) CK(cudaEventRecord(c->ev[0], c->stream));
    }
};

enum Mode { MODE_FIT = 0, MODE_X = 1, MODE_EVAL = 2, MODE_NEIGH = 3 };

__global__ void k_add_scalar(double* dst, double v) { *dst += v; }

// xe_sum / xe_sq_sum (src/pypolymlp/mlp_dev/core/data_sequential.py:128-132: sums of the UNWEIGHTED energy rows and of
// their squares): one thread per column walks the chunk's energy rows in structure order -- a fixed summation
// order, unlike the atomics this replaces -- and then applies the energy-row weight that K4b left out.
__global__ void __launch_bounds__(128) k_xe_reduce(double* __restrict__ X, int fpad, int F, const int* __restrict__ erow,
                                                   const double* __restrict__ we, int n_st, double* __restrict__ xe_sum,
                                                   double* __restrict__ xe_sq) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= F) return;
    double s1 = 0.0, s2 = 0.0;
    for (int s = 0; s < n_st; ++s) {
        double* px = X + (size_t)erow[s] * fpad + col;
        const double v = *px;
        s1 += v;
        s2 += v * v;
        *px = we[s] * v;
    }
    xe_sum[col] += s1;
    xe_sq[col] += s2;
}

// makes the main stream wait for the SYRK that is still running on stream2 (no host sync)
static void join_syrk(pm_context* c) {
    if (!c->syrk_pending) return;
    CK(cudaStreamWaitEvent(c->stream, c->ev_s, 0));
    c->syrk_pending = false;
}

// Runs the device pipeline on one prepared chunk.  `upload_inputs` false -> inputs already on the device
// (staged); dev_in then holds the device pointers in the order of upload below.
static void run_chunk(pm_context* c, const HostChunk& h, int mode, bool upload_inputs, StageTimer& tm) {
    cudaStream_t s = c->stream;
    const DevModel& d = c->dm;
    const int nt = d.n_type;
    if (upload_inputs) {
        h2d(c->d_atom_off, h.atom_off, s); h2d(c->d_st_of_atom, h.st_of_atom, s); h2d(c->d_types, h.types, s);
        h2d(c->d_trans_off, h.trans_off, s); h2d(c->d_force, h.force, s); h2d(c->d_erow, h.erow, s);
        h2d(c->d_srow, h.srow, s); h2d(c->d_frow, h.frow, s);
        h2d(c->d_x, h.x, s); h2d(c->d_y, h.y, s); h2d(c->d_z, h.z, s); h2d(c->d_trans, h.trans, s);
        h2d(c->d_w, h.w, s); h2d(c->d_yv, h.yv, s); h2d(c->d_we, h.we, s);
        h2d(c->d_cl_int, h.cl_int, s); h2d(c->d_cl_ainv, h.cl_ainv, s); h2d(c->d_cl_tmap, h.cl_tmap, s);
    }
    DevBatch b{};
    b.n_st = h.n_st; b.n_atoms = h.n_atoms; b.n_pairs = 0; b.n_rows = h.n_rows; b.max_trans = h.max_trans;
    b.need_agg = mode == MODE_EVAL ? 0 : 1;
    b.atom_off = c->d_atom_off.p; b.st_of_atom = c->d_st_of_atom.p; b.types = c->d_types.p;
    b.x = c->d_x.p; b.y = c->d_y.p; b.z = c->d_z.p; b.trans_off = c->d_trans_off.p; b.trans = c->d_trans.p;
    b.force = c->d_force.p; b.erow = c->d_erow.p; b.srow = c->d_srow.p; b.frow = c->d_frow.p;
    b.w = c->d_w.p; b.yv = c->d_yv.p;
    tm.mark(ST_H2D, 0);
    if (h.n_atoms == 0) {
        // structures without (active) atoms: every X row is zero.  The fit still counts the rows and their targets
        // (y^T y); pm_features_x zero-fills the caller's rows in process_batch.
        c->last_batch = b; c->last_pairs = 0;
        if (mode == MODE_FIT && h.n_rows > 0) {
            double ysq = 0.0;
            for (double v : h.yv) ysq += v * v;
            k_add_scalar<<<1, 1, 0, s>>>(c->acc + (size_t)d.n_variables * d.fpad + d.n_variables, ysq);
            c->n_data += h.n_rows;
        }
        return;
    }

    // ---- K1 ------------------------------------------------------------------------------------
    const size_t nseg = (size_t)h.n_atoms * nt;
    c->d_counts.ensure(nseg + 1);
    c->d_seg_off.ensure(nseg + 1);
    c->d_err.ensure(2);
    CK(cudaMemsetAsync(c->d_counts.p + nseg, 0, sizeof(int), s));
    CK(cudaMemsetAsync(c->d_err.p, 0, 2 * sizeof(int), s));
    // the count pass keeps its per-(i, j) hit masks (<= 128 translations) so that the fill pass does not repeat the
    // distance sweep; skipped when the table would be large (very ragged or very big cells)
    b.mask_stride = 0; b.masks = nullptr;
    {
        const size_t n_mask = (size_t)h.n_atoms * (size_t)h.max_atoms;
        if (h.max_trans <= 128 && h.max_atoms > 0 && n_mask <= ((size_t)32 << 20)) {
            c->d_masks.ensure(n_mask);
            b.mask_stride = h.max_atoms;
            b.masks = c->d_masks.p;
        }
    }
    b.use_cl = h.use_cl ? 1 : 0;
    b.cl_nmax = h.cl_nmax; b.cl_tmax = h.cl_tmax;
    int* d_maxc = c->d_err.p + 1;   // second int of the error buffer: largest neighbour count of an atom (cell-list fill capacity)
    if (b.use_cl) {
        c->d_bin_start.ensure((size_t)h.n_bins + 2); c->d_bin_count.ensure((size_t)h.n_bins + 2);
        c->d_bin_atoms.ensure(h.n_atoms); c->d_atom_bin.ensure(h.n_atoms); c->d_atom_img.ensure(3 * (size_t)h.n_atoms);
        b.cl_int = c->d_cl_int.p; b.cl_ainv = c->d_cl_ainv.p; b.cl_tmap = c->d_cl_tmap.p;
        b.cl_bin_start = c->d_bin_start.p; b.cl_bin_atoms = c->d_bin_atoms.p; b.cl_atom_bin = c->d_atom_bin.p;
        b.cl_atom_img = c->d_atom_img.p;
        c->d_cl_hkey.ensure((size_t)h.n_atoms * CL_CAP); c->d_cl_hd.ensure((size_t)h.n_atoms * CL_CAP * 3);
        b.cl_hkey = c->d_cl_hkey.p; b.cl_hd = c->d_cl_hd.p;
        b.mask_stride = 0; b.masks = nullptr;
        launch_cl_bins(d, b, h.n_bins, c->d_bin_count.p, s);
        launch_neighbor_cl_count(d, b, c->d_counts.p, d_maxc, s);
    } else {
        launch_neighbor_count(d, b, c->d_counts.p, s);
    }
    launch_scan_exclusive(c->d_counts.p, c->d_seg_off.p, (int)(nseg + 1), s);
    int n_pairs = 0, max_count = 0;
    CK(cudaMemcpyAsync(&n_pairs, c->d_seg_off.p + nseg, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (b.use_cl) CK(cudaMemcpyAsync(&max_count, d_maxc, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    b.n_pairs = n_pairs;
    b.seg_off = c->d_seg_off.p;
    const size_t np1 = std::max<size_t>(n_pairs, 1);
    c->d_nbr.ensure(np1); c->d_centre.ensure(np1); c->d_rev.ensure(np1);
    c->d_PB.ensure(pb_doubles(np1, d.pbstride));
    b.nbr = c->d_nbr.p; b.centre = c->d_centre.p; b.rev = c->d_rev.p;
    if (b.use_cl && max_count <= CL_CAP) launch_neighbor_cl_fill(d, b, c->d_PB.p, s);
    else launch_neighbor_fill(d, b, c->d_PB.p, s);   // (an atom with > CL_CAP neighbours: the sweep's fill orders any count)
    // reverse-pair index: K4a's target-major rows need it; it is also the symmetry check of the list (error flag).  The
    // eval kernels use neither (forces are scattered to both atoms of a pair).
    if (mode != MODE_EVAL) launch_neighbor_rev(d, b, c->d_PB.p, c->d_err.p, s);
    tm.mark(ST_NEIGH, (b.use_cl ? 8 : 5) - (mode == MODE_EVAL ? 1 : 0));
    c->last_batch = b;
    c->last_pairs = n_pairs;
    if (mode == MODE_NEIGH) return;

    bool any_force = false;
    for (int f : h.force) any_force = any_force || f;
    if (mode == MODE_EVAL) any_force = true;

    // ---- K2 ------------------------------------------------------------------------------------
    c->d_anc.ensure((size_t)h.n_atoms * d.hmax);
    c->d_agg.ensure(any_force ? (size_t)h.n_atoms * d.hmax * 9 : 1);
    // eval through the fused front end: the pair pass recomputes the basis records, K2 does not store them
    const bool eval_fused = mode == MODE_EVAL && !c->simple_s && eval_fused_supported(d, c->feat_smem);
    const bool pairs_rc = eval_fused && eval_pairs_rc_supported(d);
    bool pairs_rads = getenv("PM_EVAL_RC_RADS") != nullptr;   // A/B: the pair pass reads f_n, f_n' from the records
    if (pairs_rc && launch_anlm_eval(d, b, c->d_PB.p, c->d_anc.p, s, pairs_rads)) {
        tm.mark(ST_ANLM, 1);
    } else if (!c->simple_s && launch_pair_anlm(d, b, c->d_PB.p, c->d_anc.p, c->d_agg.p, s, !pairs_rc)) {
        if (pairs_rc) pairs_rads = false;   // (ran without storing anything)
        tm.mark(ST_ANLM, 1);   // fused pair basis + a_nlm kernel: its time is booked under "anlm"
    } else {
        launch_pair_basis(d, b, c->d_PB.p, s);
        tm.mark(ST_BASIS, 1);
        launch_anlm(d, b, c->d_PB.p, c->d_anc.p, c->d_agg.p, s);
        tm.mark(ST_ANLM, 1);
    }

    // ---- K3 ------------------------------------------------------------------------------------
    c->d_dfeat.ensure((size_t)h.n_atoms * d.fl);
    c->d_G.ensure(any_force && !eval_fused ? (size_t)h.n_atoms * d.gstride : 1);
    // single-type models: every atom slot of G has the same zero pattern, so the buffer is cleared once per
    // allocation and the kernel only writes the non-zero entries afterwards
    bool zero_g = true;
    if (d.n_type == 1 && any_force && !eval_fused) {
        if (c->g_zero_cap != c->d_G.cap) {
            CK(cudaMemsetAsync(c->d_G.p, 0, c->d_G.cap * sizeof(double), s));
            c->g_zero_cap = c->d_G.cap;
        }
        zero_g = false;
    }
    // compact polynomial-variable rows for K4b v5 (single-type models); the padding entries [npv, 64) stay zero
    double* dpv = nullptr;
    if (d.n_type == 1 && d.npv_pad <= 64 && d.npv > 0) {
        const size_t cap0 = c->d_dpv.cap;
        c->d_dpv.ensure((size_t)h.n_atoms * 64);
        if (c->d_dpv.cap != cap0) CK(cudaMemsetAsync(c->d_dpv.p, 0, c->d_dpv.cap * sizeof(double), s));
        dpv = c->d_dpv.p;
    }
    if (!eval_fused) launch_features(d, b, c->d_anc.p, c->d_dfeat.p, c->d_G.p, c->feat_smem, s, zero_g, dpv);
    tm.mark(ST_FEAT, eval_fused ? 0 : 1);

    Workspace ws;
    ws.PB = c->d_PB.p; ws.anc = c->d_anc.p; ws.agg = c->d_agg.p; ws.dfeat = c->d_dfeat.p; ws.dpv = dpv; ws.Gbuf = c->d_G.p;
    c->d_Xown.ensure((size_t)h.n_atoms * 3 * d.fl);
    c->d_S.ensure((size_t)h.n_atoms * 6 * d.fl);
    ws.Xown = c->d_Xown.p; ws.Sbuf = c->d_S.p; ws.errflag = c->d_err.p;

    if (mode == MODE_EVAL) {
        int maxseg = 0;
        for (int t = 0; t < nt; ++t)
            for (int u = 0; u < nt; ++u) maxseg = std::max(maxseg, d.types[t].seg_len[u]);
        c->d_Ah.ensure((size_t)h.n_atoms * nt * 2 * maxseg + 1);
        ws.Ah = c->d_Ah.p;
        c->d_e.ensure(h.n_st); c->d_f.ensure((size_t)h.n_atoms * 3 + 1); c->d_s.ensure((size_t)h.n_st * 6);
        CK(cudaMemsetAsync(c->d_e.p, 0, h.n_st * sizeof(double), s));
        CK(cudaMemsetAsync(c->d_f.p, 0, ((size_t)h.n_atoms * 3 + 1) * sizeof(double), s));
        CK(cudaMemsetAsync(c->d_s.p, 0, (size_t)h.n_st * 6 * sizeof(double), s));
        ws.cmat = c->has_cmat ? c->d_cmat.p : nullptr;
        ws.clin = c->has_cmat ? c->d_clin.p : nullptr;
        ws.pairs_rc = pairs_rc;
        ws.pairs_rads = pairs_rads;
        launch_eval_adjoint(d, b, ws, c->d_coeffs.p, c->d_e.p, c->d_f.p, c->d_s.p, s, eval_fused ? c->feat_smem : 0);
        tm.mark(ST_EVAL, 5);
        return;
    }

    // X-tilde chunk: zeroed before K4a in scatter mode (K4a adds into it with RED.F64), else right before K4b so
    // that K4a of this chunk does not have to wait for the previous chunk's SYRK (which still reads X)
    const bool fit = mode == MODE_FIT;
    ws.scatter = c->scatter;
    ws.lt = !c->simple_l && !c->simple_x && !c->scatter && dpv != nullptr && front_v2_supported(d);
    auto prep_X = [&] {
        join_syrk(c);
        c->d_X.ensure((size_t)std::max(h.n_rows, 1) * d.fpad);
        ws.X = c->d_X.p;
        if (c->simple_x || !xrows_fills_rows(d, ws.scatter))
            CK(cudaMemsetAsync(ws.X, 0, (size_t)h.n_rows * d.fpad * sizeof(double), s));
    };
    if (ws.scatter) prep_X();
    // ---- K4a -------------------------------------------------------------------------------------
    if (any_force) {
        if (ws.scatter) {
            c->d_Lpv.ensure(np1 * 3 * d.npv_pad);
            ws.Lpv = c->d_Lpv.p;
        } else {
            c->d_L.ensure((np1 + (ws.lt ? (size_t)h.n_atoms : 0)) * 3 * d.fl);
            ws.Lbuf = c->d_L.p;
        }
        launch_lrows(d, b, ws, c->simple_l, fit, s);
        tm.mark(ST_LROWS, 1);
    }
    // ---- K4b -------------------------------------------------------------------------------------
    if (!ws.scatter) prep_X();
    launch_xrows(d, b, ws, nullptr, nullptr, c->simple_x, fit, s);
    if (fit) {
        double* xe_sum = c->acc + (size_t)d.fpad * d.fpad;
        k_xe_reduce<<<(d.n_variables + 127) / 128, 128, 0, s>>>(ws.X, d.fpad, d.n_variables, b.erow, c->d_we.p, h.n_st,
                                                               xe_sum, xe_sum + d.fpad);
    }
    tm.mark(ST_XROWS, fit ? 4 : 3);
    // ---- K5 --------------------------------------------------------------------------------------
    if (fit) {
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (c->time_syrk) {
            if (c->syrk_ev.size() < 2 * (c->syrk_ev_used + 1)) {
                cudaEvent_t a, bq;
                CK(cudaEventCreate(&a)); CK(cudaEventCreate(&bq));
                c->syrk_ev.push_back(a); c->syrk_ev.push_back(bq);
            }
            t0 = c->syrk_ev[2 * c->syrk_ev_used]; t1 = c->syrk_ev[2 * c->syrk_ev_used + 1];
            ++c->syrk_ev_used;
        }
        if (c->two_stream && !c->profile) {
            CK(cudaEventRecord(c->ev_x, s));
            CK(cudaStreamWaitEvent(c->stream2, c->ev_x, 0));
            if (t0) CK(cudaEventRecord(t0, c->stream2));
            launch_syrk(ws.X, h.n_rows, d.fpad, c->acc, c->simple_s, c->stream2, &c->syrk_scr);
            if (t1) CK(cudaEventRecord(t1, c->stream2));
            CK(cudaEventRecord(c->ev_s, c->stream2));
            c->syrk_pending = true;
        } else {
            if (t0) CK(cudaEventRecord(t0, s));
            launch_syrk(ws.X, h.n_rows, d.fpad, c->acc, c->simple_s, s, &c->syrk_scr);
            if (t1) CK(cudaEventRecord(t1, s));
        }
        tm.mark(ST_SYRK, syrk_launches(h.n_rows, d.fpad, c->simple_s));
        c->n_data += h.n_rows;
    }
}

static void check_device_error(pm_context* c) {
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    if (err) throw std::runtime_error("neighbour list is not symmetric (reverse pair not found)");
}

// ================================================================================================
extern "C" {

const char* pm_last_error(void) { return g_err.c_str(); }
const char* pm_version(void) { return "pypolymlp_b200 0.1 (sm_100a, fp64)"; }

int pm_gtinv_read(const char* datadir, int order, const int* maxl, int n_maxl, int version, int64_t sizes[4],
                  int* lcomb_order, int* l_comb, int* n_terms, int* lm_seq, double* lm_coeffs) {
    return guarded([&] {
        std::vector<int> ml(maxl, maxl + std::max(0, n_maxl));
        GtinvTables g = read_gtinv(datadir ? datadir : "", order, ml, version);
        int64_t s1 = 0, s2 = 0, s3 = 0;
        for (size_t i = 0; i < g.l_comb.size(); ++i) {
            if (lcomb_order) lcomb_order[i] = (int)g.l_comb[i].size();
            if (n_terms) n_terms[i] = (int)g.lm_seq[i].size();
            for (int v : g.l_comb[i]) { if (l_comb) l_comb[s1] = v; ++s1; }
            for (size_t t = 0; t < g.lm_seq[i].size(); ++t) {
                if (lm_coeffs) lm_coeffs[s2] = g.lm_coeffs[i][t];
                ++s2;
                for (int v : g.lm_seq[i][t]) { if (lm_seq) lm_seq[s3] = v; ++s3; }
            }
        }
        sizes[0] = (int64_t)g.l_comb.size(); sizes[1] = s1; sizes[2] = s2; sizes[3] = s3;
    });
}

int pm_model_create(const pm_feature_params* p, pm_model** out) {
    return guarded([&] {
        if (!p || !out) throw std::invalid_argument("null argument");
        FeatureParams fp;
        fp.n_type = p->n_type; fp.n_fn = p->n_fn;
        if (fp.n_type < 1 || fp.n_fn < 1) throw std::invalid_argument("invalid n_type / n_fn");
        for (int n = 0; n < p->n_fn; ++n) fp.params.push_back({p->pair_params[2 * n], p->pair_params[2 * n + 1]});
        const int ntp = p->n_type * (p->n_type + 1) / 2;
        for (int tp = 0; tp < ntp; ++tp)
            fp.cond.emplace_back(p->cond_values + p->cond_offsets[tp], p->cond_values + p->cond_offsets[tp + 1]);
        fp.cutoff = p->cutoff; fp.model_type = p->model_type; fp.maxp = p->max_p; fp.maxl = p->max_l;
        fp.feature_type = p->feature_type;
        size_t o1 = 0, o2 = 0, o3 = 0;
        for (int i = 0; i < (fp.feature_type == PM_FEATURE_PAIR ? 0 : p->n_lcomb); ++i) {
            const int o = p->lcomb_order[i], nt_ = p->n_terms[i];
            fp.l_comb.emplace_back(p->l_comb + o1, p->l_comb + o1 + o);
            o1 += o;
            fp.lm_coeffs.emplace_back(p->lm_coeffs + o2, p->lm_coeffs + o2 + nt_);
            o2 += nt_;
            std::vector<std::vector<int>> seq(nt_);
            for (int t = 0; t < nt_; ++t) { seq[t].assign(p->lm_seq + o3, p->lm_seq + o3 + o); o3 += o; }
            fp.lm_seq.push_back(std::move(seq));
        }
        auto m = std::make_unique<pm_model>();
        m->hm.build(fp);
        *out = m.release();
    });
}

void pm_model_destroy(pm_model* m) { delete m; }
int pm_model_n_features(const pm_model* m) { return m ? m->hm.n_variables : -1; }

int pm_model_info(const pm_model* m, int64_t info[8]) {
    return guarded([&] {
        if (!m) throw std::invalid_argument("null model");
        info[0] = m->hm.fp.n_type; info[1] = m->hm.n_linear; info[2] = (int64_t)m->hm.comb2.size();
        info[3] = (int64_t)m->hm.comb3.size(); info[4] = m->hm.n_variables; info[5] = (int64_t)m->hm.pv_gid.size();
        info[6] = m->hm.n_tp; info[7] = m->hm.n_lm_half;
    });
}

int pm_model_type_info(const pm_model* m, int type, int64_t out[12]) {
    return guarded([&] {
        if (!m || type < 0 || type >= m->hm.fp.n_type) throw std::invalid_argument("invalid type");
        const TypeTables& T = m->hm.types[type];
        out[0] = T.n_full; out[1] = T.n_head; out[2] = T.n_feat; out[3] = T.n_fpad; out[4] = (int64_t)T.term_coeff.size();
        out[5] = (int64_t)T.ent_off.size() - 1; out[6] = (int64_t)T.contribs.size(); out[7] = (int64_t)T.blocks.size();
        out[8] = (int64_t)T.poly.size(); out[9] = T.n_deriv_pairs; out[10] = 0; out[11] = 0;
    });
}

int pm_model_polynomial(const pm_model* m, int type, int* n, int* col, int* order, int* local_ids3) {
    return guarded([&] {
        if (!m || type < 0 || type >= m->hm.fp.n_type) throw std::invalid_argument("invalid type");
        const TypeTables& T = m->hm.types[type];
        *n = (int)T.poly.size();
        if (!col) return;
        for (size_t i = 0; i < T.poly.size(); ++i) {
            col[i] = T.poly[i].col; order[i] = T.poly[i].order;
            for (int k = 0; k < 3; ++k) {
                const int fp_ = T.poly[i].fp[k];
                local_ids3[3 * i + k] = fp_ >= 0 ? T.pad_feat[fp_] : -1;
            }
        }
    });
}

int pm_model_feature_attrs(const pm_model* m, int64_t sizes[6], int* radial_ids, int* gtinv_ids, int* tcomb_off,
                           int* tcomb_ids, int* poly_off, int* poly_ids, int* type_pairs) {
    return guarded([&] {
        if (!m || !sizes) throw std::invalid_argument("null model");
        const HostModel& hm = m->hm;
        const bool gtinv = hm.fp.feature_type == 0;
        const int nt = hm.fp.n_type;
        int64_t n_tc = 0;
        for (const auto& lt : hm.linear) n_tc += (int64_t)lt.tp_comb.size();
        sizes[0] = hm.n_linear; sizes[1] = gtinv ? hm.n_linear : 0; sizes[2] = n_tc;
        sizes[3] = (int64_t)(hm.comb2.size() + hm.comb3.size());
        sizes[4] = (int64_t)(2 * hm.comb2.size() + 3 * hm.comb3.size()); sizes[5] = nt;
        int pos = 0;
        for (int k = 0; k < hm.n_linear; ++k) {
            const LinearTerm& lt = hm.linear[k];
            if (radial_ids) radial_ids[k] = lt.n;
            if (gtinv && gtinv_ids) gtinv_ids[k] = lt.lcid;
            if (tcomb_off) tcomb_off[k] = pos;
            for (int tp : lt.tp_comb) {
                if (tcomb_ids) tcomb_ids[pos] = tp;
                ++pos;
            }
        }
        if (tcomb_off) tcomb_off[hm.n_linear] = pos;
        int row = 0;
        pos = 0;
        for (const auto& c : hm.comb2) {
            if (poly_off) poly_off[row] = pos;
            for (int v : c) {
                if (poly_ids) poly_ids[pos] = v;
                ++pos;
            }
            ++row;
        }
        for (const auto& c : hm.comb3) {
            if (poly_off) poly_off[row] = pos;
            for (int v : c) {
                if (poly_ids) poly_ids[pos] = v;
                ++pos;
            }
            ++row;
        }
        if (poly_off) poly_off[row] = pos;
        if (type_pairs)
            for (int i = 0; i < nt; ++i)
                for (int j = 0; j < nt; ++j) type_pairs[i * nt + j] = hm.type_pairs[i][j];
    });
}

int pm_model_count_flops(const pm_model* m, const int64_t* atoms_t, const int64_t* pairs_tt, int force, double out[5]) {
    return guarded([&] {
        if (!m) throw std::invalid_argument("null model");
        const HostModel& hm = m->hm;
        const int nt = hm.fp.n_type;
        double n_atoms = 0, pairs = 0;
        for (int t = 0; t < nt; ++t) n_atoms += (double)atoms_t[t];
        for (int k = 0; k < nt * nt; ++k) pairs += (double)pairs_tt[k];
        const double F = hm.n_variables;
        const double R = 1.0 + (force ? 3.0 * n_atoms + 6.0 : 0.0);
        out[0] = R * F * (F + 1.0);
        out[1] = 2.0 * R * F;
        double wpoly = 0.0, wderiv = 0.0;
        for (int t = 0; t < nt; ++t) {
            const TypeTables& T = hm.types[t];
            double tp_ = 0.0;
            for (const auto& p : T.poly) tp_ += p.order == 1 ? 1.0 : (p.order == 2 ? 4.0 : 9.0);
            double pairs_t = 0.0;
            for (int u = 0; u < nt; ++u) pairs_t += (double)pairs_tt[t * nt + u];
            // sum_i (r_i + 1) T_p, r_i = 3 (M_i + 1) + 6
            if (force) wpoly += (3.0 * (pairs_t + atoms_t[t]) + 7.0 * atoms_t[t]) * tp_;
            else wpoly += atoms_t[t] * tp_;
            if (force) {
                // T_d(type, tp): derivative terms (merged by head and product, as the reference's
                // mapped_features_deriv) whose head belongs to type pair tp = the contributions of our G entries
                std::vector<double> td(nt, 0.0);
                for (size_t e = 0; e + 1 < T.ent_off.size(); ++e) {
                    // segment of the entry: recover it from the block that holds it
                    const int blk = T.ent_pos_re[e] / 32;
                    td[T.blocks[blk].seg] += (double)(T.ent_off[e + 1] - T.ent_off[e]);
                }
                for (int u = 0; u < nt; ++u) wderiv += 12.0 * (double)pairs_tt[t * nt + u] * td[u];
            }
        }
        out[2] = wpoly;
        out[3] = wderiv;
        out[4] = pairs * hm.fp.n_fn * hm.n_lm_half * (force ? 32.0 : 8.0);
    });
}

int pm_device_count(int* n) {
    return guarded([&] {
        int k = 0;
        cudaError_t e = cudaGetDeviceCount(&k);
        if (e != cudaSuccess) { cudaGetLastError(); k = 0; }
        *n = k;
    });
}

int pm_context_create(const pm_model* m, int device, size_t workspace_bytes, int flags, pm_context** out) {
    return guarded([&] {
        if (!m || !out) throw std::invalid_argument("null argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            throw CudaError("no CUDA device available: pypolymlp_b200 has no CPU fallback");
        }
        if (device < 0 || device >= ndev) throw std::invalid_argument("device index out of range");
        CK(cudaSetDevice(device));
        auto c = std::make_unique<pm_context>();
        c->model = m; c->device = device; c->flags = flags; c->ws_arg = workspace_bytes;
        c->ws_cap = workspace_bytes ? workspace_bytes : (size_t)12 << 30;   // measured sweet spot on B200 (bigger chunks: less SYRK fix-up; smaller: more host/device overlap)
        if (!workspace_bytes && getenv("PM_WORKSPACE_GB")) c->ws_cap = (size_t)atol(getenv("PM_WORKSPACE_GB")) << 30;
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto& e : c->ev) CK(cudaEventCreate(&e));
        CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_s, cudaEventDisableTiming));
        c->two_stream = getenv("PM_ONE_STREAM") == nullptr;
        build_device_model(c.get());
        const DevModel& d = c->dm;
        c->acc_n = (size_t)d.fpad * d.fpad + 2 * (size_t)d.fpad + 1;
        if (!workspace_bytes && !getenv("PM_WORKSPACE_GB")) {
            size_t fr = 0, tot = 0;
            CK(cudaMemGetInfo(&fr, &tot));
            const double avail = 0.5 * ((double)fr - 2.0 * (double)c->acc_n * sizeof(double));
            if (avail > (double)c->ws_cap) c->ws_cap_max = (size_t)std::min(avail, 64.0 * (1ull << 30));
        }
        *out = c.release();
    });
}

void pm_context_destroy(pm_context* c) {
    if (!c) return;
    if (c->sibling) { pm_context_destroy(c->sibling); c->sibling = nullptr; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->ev_x) cudaEventDestroy(c->ev_x);
    for (auto e : c->syrk_ev) cudaEventDestroy(e);
    if (c->ev_s) cudaEventDestroy(c->ev_s);
    for (void* p : c->table_allocs) cudaFree(p);
    if (c->acc) cudaFree(c->acc);
    if (c->syrk_scr.partials) cudaFree(c->syrk_scr.partials);
    if (c->syrk_scr.counters) cudaFree(c->syrk_scr.counters);
    if (c->comm && c->comm_owned) nccl_api().CommDestroy(c->comm);
    c->d_red.release();
    if (c->d_small) cudaFree(c->d_small);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    c->d_atom_off.release(); c->d_st_of_atom.release(); c->d_types.release(); c->d_trans_off.release();
    c->d_force.release(); c->d_erow.release(); c->d_srow.release(); c->d_frow.release(); c->d_counts.release();
    c->d_seg_off.release(); c->d_nbr.release(); c->d_centre.release(); c->d_rev.release(); c->d_err.release();
    c->d_x.release(); c->d_y.release(); c->d_z.release(); c->d_trans.release(); c->d_w.release(); c->d_yv.release(); c->d_we.release();
    c->d_cl_int.release(); c->d_cl_tmap.release(); c->d_bin_start.release(); c->d_bin_atoms.release(); c->d_atom_bin.release();
    c->d_atom_img.release(); c->d_bin_count.release(); c->d_cl_ainv.release(); c->d_cl_hkey.release(); c->d_cl_hd.release(); c->d_clin.release();
    c->d_PB.release(); c->d_dfeat.release(); c->d_dpv.release(); c->d_masks.release(); c->d_G.release(); c->d_L.release(); c->d_Xown.release(); c->d_S.release();
    c->d_X.release(); c->d_Lpv.release(); c->d_Ah.release(); c->d_coeffs.release(); c->d_cmat.release(); c->d_e.release(); c->d_f.release(); c->d_s.release();
    c->d_anc.release(); c->d_agg.release(); c->d_scan_tmp.release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->packed) cudaFree(c->packed);
    for (auto& v : c->staged_dev)
        for (void* p : v) cudaFree(p);
    cudaStreamDestroy(c->stream);
    delete c;
}

int pm_cell_translations(const double* axis, const double* positions_c, int n_atom, double cutoff, double* axis_out,
                         double* positions_out, int* n_trans, double* trans, int cap) {
    return guarded([&] {
        if (!axis || !n_trans || n_atom < 0 || (n_atom > 0 && !positions_c)) throw std::invalid_argument("null argument");
        std::vector<double> pos(positions_c, positions_c + 3 * (size_t)n_atom);
        CellTranslations ct;
        find_translations(axis, pos.data(), n_atom, cutoff, ct);
        *n_trans = (int)(ct.trans.size() / 3);
        if (axis_out) std::copy(ct.axis, ct.axis + 9, axis_out);
        if (positions_out) std::copy(pos.begin(), pos.end(), positions_out);
        if (trans && cap >= *n_trans) std::copy(ct.trans.begin(), ct.trans.end(), trans);
    });
}

int64_t pm_batch_rows(const pm_structures* st) {
    int64_t r = st->n_st;
    for (int s = 0; s < st->n_st; ++s)
        if (st->force && st->force[s]) r += 6 + 3LL * st->n_atoms[s];
    return r;
}

static void ensure_acc(pm_context* c) {
    if (!c->acc) {
        CK(cudaMalloc(&c->acc, c->acc_n * sizeof(double)));
        CK(cudaMemsetAsync(c->acc, 0, c->acc_n * sizeof(double), c->stream));
        c->n_data = 0;
    }
    if (!c->syrk_scr.partials) {
        int n_sm = 0;
        CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
        const long nt = c->dm.fpad / 128;
        c->syrk_scr.n_slots = 2L * std::max(n_sm, 1);
        c->syrk_scr.n_counters = nt * (nt + 1) / 2;
        CK(cudaMalloc(&c->syrk_scr.partials, (size_t)c->syrk_scr.n_slots * 128 * 128 * sizeof(double)));
        CK(cudaMalloc(&c->syrk_scr.counters, (size_t)c->syrk_scr.n_counters * sizeof(int)));
        CK(cudaMemsetAsync(c->syrk_scr.counters, 0, (size_t)c->syrk_scr.n_counters * sizeof(int), c->stream));
    }
}

int pm_fit_reset(pm_context* c) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        ensure_acc(c);
        CK(cudaMemsetAsync(c->acc, 0, c->acc_n * sizeof(double), c->stream));
        c->n_data = 0;
    });
}

static void process_batch(pm_context* c, const pm_structures* st, const double* w, const double* y, int mode,
                          double* x_out, double* e_out, double* f_out, double* s_out) {
    CK(cudaSetDevice(c->device));
    validate_structures(c, st);
    std::vector<size_t> aoff;
    std::vector<long> be, bs, bf;
    long n_rows = 0;
    batch_layout(st, aoff, be, bs, bf, n_rows);
    const DevModel& d = c->dm;
    const int F = d.n_variables;
    StageTimer tm(c);
    // The host prepares chunk k+1 (translations, row maps) while the GPU still works on chunk k.
    const bool dbg = getenv("PM_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    const auto t_start = now();
    // (eval keeps no derivative rows: the per-structure estimate of the fit path overstates its footprint)
    const double eval_scale = getenv("PM_EVAL_CHUNK_SCALE") ? atof(getenv("PM_EVAL_CHUNK_SCALE")) : 1.0;
    const auto chunks = plan_chunks(c, st, mode == MODE_EVAL ? eval_scale : 1.0);
    if (dbg) fprintf(stderr, "[pm] layout + plan (%zu chunks): %.3f ms\n", chunks.size(), ms_since(t_start));
    HostChunk h_next;
    auto prepare = [&](size_t k, HostChunk& out) {
        prepare_chunk(c, st, aoff, be, bs, bf, chunks[k].first, chunks[k].second, w, y, out);
        if (mode == MODE_EVAL)
            for (auto& f : out.force) f = 1;
    };
    {
        const auto t0 = now();
        if (!chunks.empty()) prepare(0, h_next);
        if (dbg) fprintf(stderr, "[pm] prepare chunk 0: %.3f ms\n", ms_since(t0));
    }
    for (size_t ck = 0; ck < chunks.size(); ++ck) {
        const int s0 = chunks[ck].first;
        HostChunk h = std::move(h_next);
        const auto t_run = now();
        run_chunk(c, h, mode, true, tm);
        if (dbg) fprintf(stderr, "[pm] chunk %zu (%d structures): run_chunk host %.3f ms\n", ck, h.n_st, ms_since(t_run));
        // the next chunk is prepared while the device works on this one -- BEFORE the result copies below, which block the
        // host until the chunk's kernels are done (device -> pageable host memory)
        const auto t_prep = now();
        if (ck + 1 < chunks.size()) prepare(ck + 1, h_next);
        const double prep_ms = ms_since(t_prep);
        if (mode == MODE_X && h.n_atoms == 0) {
            for (int k = 0; k < h.n_st; ++k) {   // no atoms: zero rows, as the reference returns them
                std::fill_n(x_out + (size_t)h.brow_e[k] * F, (size_t)F, 0.0);
                if (h.force[k]) std::fill_n(x_out + (size_t)h.brow_s[k] * F, (size_t)6 * F, 0.0);
            }
        } else if (mode == MODE_X) {
            for (int k = 0; k < h.n_st; ++k) {
                CK(cudaMemcpy2DAsync(x_out + (size_t)h.brow_e[k] * F, F * sizeof(double), c->d_X.p + (size_t)h.erow[k] * d.fpad,
                                     d.fpad * sizeof(double), F * sizeof(double), 1, cudaMemcpyDeviceToHost, c->stream));
                if (h.force[k]) {
                    CK(cudaMemcpy2DAsync(x_out + (size_t)h.brow_s[k] * F, F * sizeof(double),
                                         c->d_X.p + (size_t)h.srow[k] * d.fpad, d.fpad * sizeof(double), F * sizeof(double), 6,
                                         cudaMemcpyDeviceToHost, c->stream));
                    const int nf = 3 * (h.atom_off[k + 1] - h.atom_off[k]);
                    if (nf > 0)
                        CK(cudaMemcpy2DAsync(x_out + (size_t)h.brow_f[k] * F, F * sizeof(double),
                                             c->d_X.p + (size_t)h.frow[k] * d.fpad, d.fpad * sizeof(double),
                                             F * sizeof(double), nf, cudaMemcpyDeviceToHost, c->stream));
                }
            }
            tm.mark(ST_D2H, 0);
        } else if (mode == MODE_EVAL && h.n_atoms > 0) {
            CK(cudaMemcpyAsync(e_out + s0, c->d_e.p, h.n_st * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(f_out + 3 * aoff[s0], c->d_f.p, (size_t)h.n_atoms * 3 * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(s_out + 6 * (size_t)s0, c->d_s.p, (size_t)h.n_st * 6 * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
            tm.mark(ST_D2H, 0);
        } else if (mode == MODE_EVAL) {
            for (int k = 0; k < h.n_st; ++k) { e_out[s0 + k] = 0.0; for (int r = 0; r < 6; ++r) s_out[6 * (size_t)(s0 + k) + r] = 0.0; }
        }
        const auto t_wait = now();
        if (h.n_atoms > 0) check_device_error(c);
        else CK(cudaStreamSynchronize(c->stream));
        if (dbg) fprintf(stderr, "[pm] chunk %zu: prepare next %.3f ms, wait for the device %.3f ms\n", ck, prep_ms, ms_since(t_wait));
    }
    join_syrk(c);   // later work on the main stream (reset, finalize, the caller's events) is ordered after K5
}

int pm_fit_accumulate(pm_context* c, const pm_structures* st, const double* w, const double* y) {
    return guarded([&] {
        if (!c || !st || !w || !y) throw std::invalid_argument("null argument");
        CK(cudaSetDevice(c->device));
        ensure_acc(c);
        process_batch(c, st, w, y, MODE_FIT, nullptr, nullptr, nullptr, nullptr);
    });
}

int pm_features_x(pm_context* c, const pm_structures* st, double* x) {
    return guarded([&] {
        if (!c || !st || !x) throw std::invalid_argument("null argument");
        process_batch(c, st, nullptr, nullptr, MODE_X, x, nullptr, nullptr, nullptr);
    });
}

int pm_fit_stage(pm_context* c, const pm_structures* st, const double* w, const double* y) {
    return guarded([&] {
        if (!c || !st || !w || !y) throw std::invalid_argument("null argument");
        CK(cudaSetDevice(c->device));
        validate_structures(c, st);
        ensure_acc(c);
        std::vector<size_t> aoff;
        std::vector<long> be, bs, bf;
        long n_rows = 0;
        batch_layout(st, aoff, be, bs, bf, n_rows);
        for (auto& v : c->staged_dev)
            for (void* p : v) cudaFree(p);
        c->staged_dev.clear();
        c->staged.clear();
        for (auto [s0, s1] : plan_chunks(c, st)) {
            HostChunk h;
            prepare_chunk(c, st, aoff, be, bs, bf, s0, s1, w, y, h);
            std::vector<void*> dev;
            auto put = [&](const void* src, size_t bytes) {
                void* p = nullptr;
                CK(cudaMalloc(&p, std::max<size_t>(bytes, 8)));
                if (bytes) CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
                dev.push_back(p);
            };
            put(h.atom_off.data(), h.atom_off.size() * 4); put(h.st_of_atom.data(), h.st_of_atom.size() * 4);
            put(h.types.data(), h.types.size() * 4); put(h.trans_off.data(), h.trans_off.size() * 4);
            put(h.force.data(), h.force.size() * 4); put(h.erow.data(), h.erow.size() * 4);
            put(h.srow.data(), h.srow.size() * 4); put(h.frow.data(), h.frow.size() * 4);
            put(h.x.data(), h.x.size() * 8); put(h.y.data(), h.y.size() * 8); put(h.z.data(), h.z.size() * 8);
            put(h.trans.data(), h.trans.size() * 8); put(h.w.data(), h.w.size() * 8); put(h.yv.data(), h.yv.size() * 8);
            put(h.we.data(), h.we.size() * 8);
            put(h.cl_int.data(), h.cl_int.size() * 4); put(h.cl_ainv.data(), h.cl_ainv.size() * 8);
            put(h.cl_tmap.data(), h.cl_tmap.size() * 4);
            c->staged_dev.push_back(dev);
            // keep only the metadata on the host
            HostChunk meta;
            meta.n_st = h.n_st; meta.n_atoms = h.n_atoms; meta.n_rows = h.n_rows; meta.force = h.force;
            meta.max_trans = h.max_trans; meta.max_atoms = h.max_atoms;
            meta.n_bins = h.n_bins; meta.cl_nmax = h.cl_nmax; meta.cl_tmax = h.cl_tmax; meta.use_cl = h.use_cl;
            c->staged.push_back(std::move(meta));
        }
    });
}

int pm_fit_accumulate_staged(pm_context* c) {
    return guarded([&] {
        if (!c) throw std::invalid_argument("null argument");
        CK(cudaSetDevice(c->device));
        StageTimer tm(c);
        for (size_t k = 0; k < c->staged.size(); ++k) {
            const auto& dev = c->staged_dev[k];
            // point the chunk buffers at the staged copies (no copy, no ownership change)
            auto swap_in = [&](auto& vec, void* p) { vec.p = reinterpret_cast<decltype(vec.p)>(p); };
            struct Saved { void* p; size_t cap; };
            std::vector<Saved> saved;
            auto save = [&](auto& vec) { saved.push_back({vec.p, vec.cap}); };
            save(c->d_atom_off); save(c->d_st_of_atom); save(c->d_types); save(c->d_trans_off); save(c->d_force);
            save(c->d_erow); save(c->d_srow); save(c->d_frow); save(c->d_x); save(c->d_y); save(c->d_z); save(c->d_trans);
            save(c->d_w); save(c->d_yv); save(c->d_we); save(c->d_cl_int); save(c->d_cl_ainv); save(c->d_cl_tmap);
            swap_in(c->d_atom_off, dev[0]); swap_in(c->d_st_of_atom, dev[1]); swap_in(c->d_types, dev[2]);
            swap_in(c->d_trans_off, dev[3]); swap_in(c->d_force, dev[4]); swap_in(c->d_erow, dev[5]);
            swap_in(c->d_srow, dev[6]); swap_in(c->d_frow, dev[7]); swap_in(c->d_x, dev[8]); swap_in(c->d_y, dev[9]);
            swap_in(c->d_z, dev[10]); swap_in(c->d_trans, dev[11]); swap_in(c->d_w, dev[12]); swap_in(c->d_yv, dev[13]);
            swap_in(c->d_we, dev[14]); swap_in(c->d_cl_int, dev[15]); swap_in(c->d_cl_ainv, dev[16]);
            swap_in(c->d_cl_tmap, dev[17]);
            std::exception_ptr ex;
            try {
                run_chunk(c, c->staged[k], MODE_FIT, false, tm);
            } catch (...) { ex = std::current_exception(); }
            size_t q = 0;
            auto restore = [&](auto& vec) { vec.p = reinterpret_cast<decltype(vec.p)>(saved[q].p); vec.cap = saved[q].cap; ++q; };
            restore(c->d_atom_off); restore(c->d_st_of_atom); restore(c->d_types); restore(c->d_trans_off);
            restore(c->d_force); restore(c->d_erow); restore(c->d_srow); restore(c->d_frow); restore(c->d_x);
            restore(c->d_y); restore(c->d_z); restore(c->d_trans); restore(c->d_w); restore(c->d_yv); restore(c->d_we);
            restore(c->d_cl_int); restore(c->d_cl_ainv); restore(c->d_cl_tmap);
            if (ex) std::rethrow_exception(ex);
        }
        join_syrk(c);
    });
}

int pm_fit_accumulator(pm_context* c, void** dev_ptr, size_t* n_doubles) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        ensure_acc(c);
        // n_data travels inside the accumulator so that a cross-GPU sum reduces it too
        const double nd = (double)c->n_data;
        CK(cudaMemcpyAsync(c->acc + c->acc_n - 1, &nd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        *dev_ptr = c->acc;
        *n_doubles = c->acc_n;
    });
}

int pm_fit_fpad(pm_context* c) { return c ? c->dm.fpad : -1; }

// Packs the accumulator for the host: symmetric dense X^T X (the kernels fill the upper triangle of the padded,
// tiled C), X^T y (the y column), the energy-row sums and y^T y.  One 32 x 32 tile per CTA; the lower triangle is
// written from the transposed upper tile through shared memory so that reads and writes stay coalesced.
__global__ void __launch_bounds__(256) k_pack_result(const double* __restrict__ acc, int fp, int F, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int si = min(bi, bj), sj = max(bi, bj);   // source tile (upper)
    for (int r = ty; r < 32; r += 8) {
        const int i = si * 32 + r, j = sj * 32 + tx;
        tile[r][tx] = (i < F && j < F) ? acc[(size_t)i * fp + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        if (i >= F || j >= F) continue;
        double v;
        if (bi < bj) v = tile[r][tx];
        else if (bi > bj) v = tile[tx][r];
        else v = r <= tx ? tile[r][tx] : tile[tx][r];
        out[(size_t)i * F + j] = v;
    }
    if (bi == 0 && bj == 0) {
        double* tail = out + (size_t)F * F;
        for (int i = threadIdx.x; i < F; i += blockDim.x) {
            tail[i] = acc[(size_t)i * fp + F];
            tail[F + i] = acc[(size_t)fp * fp + i];
            tail[2 * F + i] = acc[(size_t)fp * fp + fp + i];
        }
        if (threadIdx.x == 0) {
            tail[3 * F] = acc[(size_t)F * fp + F];
            tail[3 * F + 1] = acc[(size_t)fp * fp + 2 * (size_t)fp];
        }
    }
}

// pack on the device, one D2H copy into the context's pinned buffer; returns the host pointer
static const double* finalize_packed(pm_context* c, size_t* n_out) {
    CK(cudaSetDevice(c->device));
    ensure_acc(c);
    const DevModel& d = c->dm;
    const int F = d.n_variables, fp = d.fpad;
    const size_t n = (size_t)F * F + 3 * (size_t)F + 2;
    if (c->pinned_n < n) {
        if (c->pinned) cudaFreeHost(c->pinned);
        if (c->packed) cudaFree(c->packed);
        c->pinned = nullptr; c->packed = nullptr; c->pinned_n = 0;
        CK(cudaHostAlloc(&c->pinned, n * sizeof(double), cudaHostAllocDefault));
        CK(cudaMalloc(&c->packed, n * sizeof(double)));
        c->pinned_n = n;
    }
    const int nb = (F + 31) / 32;
    k_pack_result<<<dim3(nb, nb), 256, 0, c->stream>>>(c->acc, fp, F, c->packed);
    c->launches += 1;
    CK(cudaMemcpyAsync(c->pinned, c->packed, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    double* tail = c->pinned + (size_t)F * F;
    const double nd = tail[3 * (size_t)F + 1];
    tail[3 * (size_t)F + 1] = nd > (double)c->n_data ? (double)llround(nd) : (double)c->n_data;
    if (n_out) *n_out = n;
    return c->pinned;
}

// large host copies (the 33 MB X^T X into a caller buffer whose pages are usually untouched): a few threads
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    constexpr size_t MIN_PER_THREAD = 4u << 20;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nthr = std::min<size_t>({(size_t)8, (size_t)hw, bytes / MIN_PER_THREAD});
    if (nthr < 2) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nthr) + 4095) & ~(size_t)4095;
    for (size_t k = 0; k < nthr; ++k) {
        const size_t o = k * per;
        if (o >= bytes) break;
        const size_t len = std::min(per, bytes - o);
        th.emplace_back([=] { std::memcpy((char*)dst + o, (const char*)src + o, len); });
    }
    for (auto& t : th) t.join();
}

int pm_fit_finalize_view(pm_context* c, const double** packed, size_t* n_doubles) {
    return guarded([&] {
        if (!c || !packed) throw std::invalid_argument("null argument");
        *packed = finalize_packed(c, n_doubles);
    });
}

int pm_fit_finalize(pm_context* c, double* xtx, double* xty, double* xe_sum, double* xe_sq_sum, double* y_sq_norm,
                    int64_t* n_data) {
    return guarded([&] {
        if (!c) throw std::invalid_argument("null argument");
        const double* pk = finalize_packed(c, nullptr);
        const size_t F = (size_t)c->dm.n_variables;
        const double* tail = pk + F * F;
        if (xtx) parallel_memcpy(xtx, pk, F * F * sizeof(double));
        if (xty) std::memcpy(xty, tail, F * sizeof(double));
        if (xe_sum) std::memcpy(xe_sum, tail + F, F * sizeof(double));
        if (xe_sq_sum) std::memcpy(xe_sq_sum, tail + 2 * F, F * sizeof(double));
        if (y_sq_norm) *y_sq_norm = tail[3 * F];
        if (n_data) *n_data = (int64_t)tail[3 * F + 1];
    });
}

int pm_fit_solve_ridge(pm_context* c, const double* alphas, int n_alpha, const double* scales_in, int64_t n_energy,
                       int include_force, double scale_threshold, double* scales_out, double* coefs, double* rmse) {
    return guarded([&] {
        if (!c || !alphas || n_alpha < 1 || !coefs || !rmse) throw std::invalid_argument("null argument");
        CK(cudaSetDevice(c->device));
        ensure_acc(c);
        const DevModel& d = c->dm;
        const int F = d.n_variables, fp = d.fpad;
        CK(cudaStreamSynchronize(c->stream));
        std::vector<double> tail(2 * (size_t)fp + 1);
        CK(cudaMemcpy(tail.data(), c->acc + (size_t)fp * fp, tail.size() * sizeof(double), cudaMemcpyDeviceToHost));
        double ysq = 0.0;
        CK(cudaMemcpy(&ysq, c->acc + (size_t)F * fp + F, sizeof(double), cudaMemcpyDeviceToHost));
        const double nd_acc = tail[2 * (size_t)fp];
        const long n_data = nd_acc > (double)c->n_data ? (long)llround(nd_acc) : (long)c->n_data;
        if (!scales_in && n_energy <= 0) throw std::invalid_argument("n_energy is required when scales are not given");
        if (n_data <= 0) throw std::runtime_error("no data accumulated");
        solve_ridge_device(c->acc, fp, F, tail.data(), tail.data() + fp, ysq, n_data, alphas, n_alpha, scales_in,
                           (long)n_energy, include_force != 0, scale_threshold, scales_out, coefs, rmse, c->stream);
    });
}

int pm_synchronize(pm_context* c) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
    });
}

void* pm_stream(pm_context* c) { return c ? (void*)c->stream : nullptr; }

int pm_timer_start(pm_context* c) {
    return guarded([&] {
        if (!c) throw std::invalid_argument("null argument");
        CK(cudaSetDevice(c->device));
        if (!c->ev_t0) { CK(cudaEventCreate(&c->ev_t0)); CK(cudaEventCreate(&c->ev_t1)); }
        join_syrk(c);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaEventRecord(c->ev_t0, c->stream));
    });
}

int pm_timer_stop(pm_context* c, double* ms) {
    return guarded([&] {
        if (!c || !ms || !c->ev_t0) throw std::invalid_argument("pm_timer_stop without pm_timer_start");
        CK(cudaSetDevice(c->device));
        join_syrk(c);
        CK(cudaEventRecord(c->ev_t1, c->stream));
        CK(cudaEventSynchronize(c->ev_t1));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, c->ev_t0, c->ev_t1));
        *ms = t;
    });
}
int64_t pm_launch_count(pm_context* c) { return c ? c->launches : 0; }

int pm_profile_enable(pm_context* c, int on) {
    c->profile = on == 1;
    c->time_syrk = on == 2;
    c->syrk_ev_used = 0;
    for (int k = 0; k < ST_COUNT; ++k) { c->stage_ms[k] = 0.0; c->stage_launches[k] = 0; }
    return PM_OK;
}

int pm_profile_get(pm_context* c, int* n_stages, double* ms, int64_t* launches) {
    *n_stages = ST_COUNT;
    if (c->time_syrk && c->syrk_ev_used > 0) {   // mode 2: sum the K5 launches recorded since pm_profile_enable
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->stream2) cudaStreamSynchronize(c->stream2);
        double tot = 0.0;
        for (size_t k = 0; k < c->syrk_ev_used; ++k) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, c->syrk_ev[2 * k], c->syrk_ev[2 * k + 1]) == cudaSuccess) tot += t;
        }
        c->stage_ms[ST_SYRK] = tot;
        c->stage_launches[ST_SYRK] = (int64_t)c->syrk_ev_used;
        c->syrk_ev_used = 0;
    }
    for (int k = 0; k < ST_COUNT; ++k) { if (ms) ms[k] = c->stage_ms[k]; if (launches) launches[k] = c->stage_launches[k]; }
    return PM_OK;
}

const char* pm_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }

int pm_neighbor_full(pm_context* c, const double* axis, const double* positions_c, const int* types, int n_atom,
                     int* offsets, int* neigh, double* dx, double* dy, double* dz) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        pm_structures st{};
        int na = n_atom, force = 0;
        st.n_st = 1; st.axis = axis; st.positions_c = positions_c; st.types = types; st.n_atoms = &na; st.force = &force;
        validate_structures(c, &st);
        std::vector<size_t> aoff{0, (size_t)n_atom};
        std::vector<long> be{0}, bs{-1}, bf{-1};
        HostChunk h;
        prepare_chunk(c, &st, aoff, be, bs, bf, 0, 1, nullptr, nullptr, h);
        StageTimer tm(c);
        run_chunk(c, h, MODE_NEIGH, true, tm);
        if (n_atom > 0) check_device_error(c);
        const int nt = c->dm.n_type;
        std::vector<int> seg((size_t)n_atom * nt + 1, 0);
        if (n_atom > 0) CK(cudaMemcpy(seg.data(), c->d_seg_off.p, seg.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int i = 0; i <= n_atom; ++i) offsets[i] = seg[(size_t)i * nt];
        if (neigh && c->last_pairs > 0) {
            const int P = c->last_pairs;
            CK(cudaMemcpy(neigh, c->d_nbr.p, P * sizeof(int), cudaMemcpyDeviceToHost));
            const int stride = c->dm.pbstride;
            std::vector<double> raw(pb_doubles(P, stride));
            CK(cudaMemcpy(raw.data(), c->d_PB.p, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int p = 0; p < P; ++p) {
                const double* base = raw.data() + (size_t)(p >> 5) * stride * PB_BLK + (p & 31);
                dx[p] = base[0]; dy[p] = base[PB_BLK]; dz[p] = base[2 * PB_BLK];
            }
        }
    });
}

int pm_eval_set_coeffs(pm_context* c, const double* coeffs, int n) {
    return guarded([&] {
        if (!c || !coeffs) throw std::invalid_argument("null argument");
        if (n != c->dm.n_variables) throw std::invalid_argument("number of coefficients does not match the model");
        CK(cudaSetDevice(c->device));
        c->d_coeffs.ensure(n);
        CK(cudaMemcpy(c->d_coeffs.p, coeffs, n * sizeof(double), cudaMemcpyHostToDevice));
        c->has_coeffs = true;
        ++c->coeff_version;
        // dense order-2 coefficient matrix over the polynomial variables for the fused eval kernel (max_p = 2, <= 64
        // variables, one column per unordered pair): C'[a][b] = C'[b][a] = c_col, diagonal doubled (d E / d d_a)
        const HostModel& hm = c->model->hm;
        c->has_cmat = false;
        if (c->dm.pair_colof && !hm.has_order3 && c->dm.npv_pad <= 64) {
            const int nt = c->dm.n_type;
            std::vector<double> cm((size_t)nt * 4096, 0.0);
            for (const auto& pt : hm.pair_terms) {
                const int col = pt[0], a = pt[1], b2 = pt[2];
                for (int t = 0; t < nt; ++t) {
                    if (hm.colterm[t][col].order != 2) continue;
                    if (a == b2) cm[(size_t)t * 4096 + a * 64 + a] = 2.0 * coeffs[col];
                    else { cm[(size_t)t * 4096 + a * 64 + b2] = coeffs[col]; cm[(size_t)t * 4096 + b2 * 64 + a] = coeffs[col]; }
                }
            }
            c->d_cmat.ensure(cm.size());
            CK(cudaMemcpy(c->d_cmat.p, cm.data(), cm.size() * sizeof(double), cudaMemcpyHostToDevice));
            std::vector<double> cl((size_t)nt * c->dm.fl, 0.0);
            for (int t = 0; t < nt; ++t)
                for (int col = 0; col < c->dm.n_linear; ++col)
                    if (hm.colterm[t][col].order == 1) cl[(size_t)t * c->dm.fl + hm.colterm[t][col].fp[0]] = coeffs[col];
            c->d_clin.ensure(cl.size());
            CK(cudaMemcpy(c->d_clin.p, cl.data(), cl.size() * sizeof(double), cudaMemcpyHostToDevice));
            c->has_cmat = true;
        }
    });
}

// second eval lane: a sibling context on the same device with the same coefficients
static pm_context* eval_sibling(pm_context* c) {
    if (!c->sibling) {
        pm_context* sib = nullptr;
        if (pm_context_create(c->model, c->device, c->ws_arg, c->flags, &sib) != PM_OK) throw std::runtime_error(pm_last_error());
        sib->simple_l = c->simple_l; sib->simple_x = c->simple_x; sib->simple_s = c->simple_s; sib->scatter = c->scatter;
        c->sibling = sib;
    }
    pm_context* sib = c->sibling;
    if (sib->coeff_version != c->coeff_version) {
        CK(cudaSetDevice(c->device));
        sib->d_coeffs.ensure(c->dm.n_variables);
        CK(cudaMemcpy(sib->d_coeffs.p, c->d_coeffs.p, (size_t)c->dm.n_variables * sizeof(double), cudaMemcpyDeviceToDevice));
        if (c->has_cmat) {
            const size_t n = (size_t)c->dm.n_type * 4096;
            sib->d_cmat.ensure(n);
            CK(cudaMemcpy(sib->d_cmat.p, c->d_cmat.p, n * sizeof(double), cudaMemcpyDeviceToDevice));
            const size_t nl = (size_t)c->dm.n_type * c->dm.fl;
            sib->d_clin.ensure(nl);
            CK(cudaMemcpy(sib->d_clin.p, c->d_clin.p, nl * sizeof(double), cudaMemcpyDeviceToDevice));
        }
        sib->has_cmat = c->has_cmat; sib->has_coeffs = true;
        sib->coeff_version = c->coeff_version;
    }
    return sib;
}

int pm_eval(pm_context* c, const pm_structures* st, double* energies, double* forces, double* stresses) {
    return guarded([&] {
        if (!c || !st || !energies || !forces || !stresses) throw std::invalid_argument("null argument");
        if (!c->has_coeffs) throw std::runtime_error("coefficients are not set");
        // several lanes for batches that span several chunks' worth of work (PM_EVAL_LANES=1: one lane): lane k is the
        // k-th context of the sibling chain c -> c->sibling -> ..., each with its own host thread
        int lanes = getenv("PM_EVAL_LANES") ? atoi(getenv("PM_EVAL_LANES")) : 3;   // (measured 2 / 3 / 4 lanes: 16.9 / 17.8 / 17.9 M atoms/s)
        size_t total = 0;
        if (st->n_atoms)
            for (int s = 0; s < st->n_st; ++s) total += (size_t)std::max(st->n_atoms[s], 0);
        lanes = std::min<int>({lanes, 8, st->n_st, (int)(total / 8192)});
        if (lanes < 2 || c->profile) {
            process_batch(c, st, nullptr, nullptr, MODE_EVAL, nullptr, energies, forces, stresses);
            return;
        }
        validate_structures(c, st);
        std::vector<pm_context*> ctx{c};
        for (int k = 1; k < lanes; ++k) ctx.push_back(eval_sibling(ctx.back()));
        // contiguous slices of about total / lanes atoms
        std::vector<int> first(lanes + 1, st->n_st);
        std::vector<size_t> afirst(lanes + 1, total);
        {
            size_t acc = 0;
            int k = 0;
            for (int s = 0; s < st->n_st && k < lanes; ++s) {
                if (acc * lanes >= (size_t)k * total) { first[k] = s; afirst[k] = acc; ++k; }
                acc += (size_t)st->n_atoms[s];
            }
            for (; k < lanes; ++k) { first[k] = st->n_st; afirst[k] = total; }
        }
        std::vector<pm_structures> sub(lanes, *st);
        std::vector<std::exception_ptr> err(lanes);
        auto run_lane = [&](int k) {
            try {
                if (sub[k].n_st > 0)
                    process_batch(ctx[k], &sub[k], nullptr, nullptr, MODE_EVAL, nullptr, energies + first[k], forces + 3 * afirst[k],
                                  stresses + 6 * (size_t)first[k]);
            } catch (...) { err[k] = std::current_exception(); }
        };
        for (int k = 0; k < lanes; ++k) {
            sub[k].n_st = first[k + 1] - first[k];
            sub[k].axis = st->axis + 9 * (size_t)first[k];
            sub[k].positions_c = st->positions_c + 3 * afirst[k];
            sub[k].types = st->types + afirst[k];
            sub[k].n_atoms = st->n_atoms + first[k];
            sub[k].force = st->force ? st->force + first[k] : nullptr;
        }
        std::vector<std::thread> th;
        for (int k = 1; k < lanes; ++k) th.emplace_back(run_lane, k);
        run_lane(0);
        for (auto& t : th) t.join();
        for (int k = 1; k < lanes; ++k) {
            c->launches += ctx[k]->launches;
            ctx[k]->launches = 0;
            for (int q = 0; q < ST_COUNT; ++q) { c->stage_launches[q] += ctx[k]->stage_launches[q]; ctx[k]->stage_launches[q] = 0; }
        }
        for (int k = 0; k < lanes; ++k)
            if (err[k]) std::rethrow_exception(err[k]);
    });
}

int pm_debug_fetch(pm_context* c, int what, double* out, size_t cap, size_t* n) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        const DevModel& d = c->dm;
        const DevBatch& b = c->last_batch;
        const void* src = nullptr;
        size_t cnt = 0;
        if (what == 0) { src = c->d_anc.p; cnt = (size_t)b.n_atoms * d.hmax * 2; }
        else if (what == 1) { src = c->d_dfeat.p; cnt = (size_t)b.n_atoms * d.fl; }
        else if (what == 2) { src = c->d_PB.p; cnt = pb_doubles(c->last_pairs, d.pbstride); }
        else if (what == 3) { src = c->d_X.p; cnt = (size_t)b.n_rows * d.fpad; }
        else throw std::invalid_argument("unknown debug buffer");
        *n = cnt;
        if (out) {
            if (cnt > cap) throw std::invalid_argument("debug buffer too small");
            if (cnt) CK(cudaMemcpy(out, src, cnt * sizeof(double), cudaMemcpyDeviceToHost));
        }
    });
}

int pm_microbench(pm_context* c, int which, int n, double* tflops) {
    return guarded([&] {
        CK(cudaSetDevice(c->device));
        if (which == 3) *tflops = microbench_dgemm(n > 0 ? n : 8192, c->stream);
        else if (which == 4) *tflops = microbench_red(n > 16 ? n : 768, c->stream);  // giga fp64 atomics / s
        else if (which == 5) *tflops = microbench_dmma_chain(n, c->stream);          // cycles per DMMA of one warp; n = 100 * warps + chains
        else *tflops = microbench_fp64(which, c->stream);
        CK(cudaGetLastError());
    });
}

}  // extern "C"

// ================================================================================================
// Multi-GPU: the partial accumulators of the ranks are summed with ONE ncclReduce over NVLink
// (replaces the OpenMP-over-structures loop of compute/py_model.cpp:39-53 and the batch sum of
// src/pypolymlp/mlp_dev/core/data_sequential.py:49-70 across devices).  Only the upper 128 x 128 tiles of C are valid,
// so they are packed into a contiguous buffer first: [tiles (ti <= tj, row-major) x 128 x 128 | xe_sum | xe_sq | n_data].
// ================================================================================================
__global__ void __launch_bounds__(256) k_pack_upper(double* __restrict__ acc, int fpad, double* __restrict__ buf, int unpack) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;
    const int ntile = fpad / 128;
    const long tile = (long)ti * ntile - (long)ti * (ti - 1) / 2 + (tj - ti);
    double2* b2 = reinterpret_cast<double2*>(buf + tile * 128 * 128);
    for (int e = threadIdx.x; e < 128 * 64; e += 256) {
        const int r = e >> 6, c2 = e & 63;
        double2* a2 = reinterpret_cast<double2*>(acc + (size_t)(ti * 128 + r) * fpad + tj * 128) + c2;
        if (unpack) *a2 = b2[e];
        else b2[e] = *a2;
    }
}

static size_t packed_upper_doubles(const pm_context* c) {
    const size_t nt = (size_t)c->dm.fpad / 128;
    return nt * (nt + 1) / 2 * 128 * 128 + 2 * (size_t)c->dm.fpad + 1;
}

// pack this context's accumulator (with its row count in the last slot) into c->d_red
static void pack_for_reduce(pm_context* c) {
    CK(cudaSetDevice(c->device));
    ensure_acc(c);
    const int fp = c->dm.fpad, nt = fp / 128;
    const size_t n = packed_upper_doubles(c);
    c->d_red.ensure(n);
    const double nd = (double)c->n_data;
    CK(cudaMemcpyAsync(c->acc + c->acc_n - 1, &nd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_pack_upper<<<dim3(nt, nt), 256, 0, c->stream>>>(c->acc, fp, c->d_red.p, 0);
    CK(cudaMemcpyAsync(c->d_red.p + n - (2 * (size_t)fp + 1), c->acc + (size_t)fp * fp, (2 * (size_t)fp + 1) * sizeof(double),
                       cudaMemcpyDeviceToDevice, c->stream));
    c->launches += 1;
}

static void unpack_after_reduce(pm_context* c) {
    CK(cudaSetDevice(c->device));
    const int fp = c->dm.fpad, nt = fp / 128;
    const size_t n = packed_upper_doubles(c);
    k_pack_upper<<<dim3(nt, nt), 256, 0, c->stream>>>(c->acc, fp, c->d_red.p, 1);
    CK(cudaMemcpyAsync(c->acc + (size_t)fp * fp, c->d_red.p + n - (2 * (size_t)fp + 1), (2 * (size_t)fp + 1) * sizeof(double),
                       cudaMemcpyDeviceToDevice, c->stream));
    c->launches += 1;
}

struct pm_multi {
    std::vector<pm_context*> ctx;
    std::vector<ncclComm_t> comms;
};

// structures [s0, s1) of a batch as a batch of their own, with the rows of w / y regrouped into its PyModel layout
struct SubBatch {
    pm_structures st{};
    std::vector<double> w, y;
    SubBatch(const pm_structures* all, const std::vector<size_t>& aoff, const std::vector<long>& be,
             const std::vector<long>& bs, const std::vector<long>& bf, int s0, int s1, const double* w_all, const double* y_all) {
        st.n_st = s1 - s0;
        st.axis = all->axis + 9 * (size_t)s0;
        st.positions_c = all->positions_c + 3 * aoff[s0];
        st.types = all->types + aoff[s0];
        st.n_atoms = all->n_atoms + s0;
        st.force = all->force ? all->force + s0 : nullptr;
        auto push = [&](long row, long n) {
            for (long r = 0; r < n; ++r) { w.push_back(w_all[row + r]); y.push_back(y_all[row + r]); }
        };
        for (int s = s0; s < s1; ++s) push(be[s], 1);
        for (int s = s0; s < s1; ++s) if (bs[s] >= 0) push(bs[s], 6);
        for (int s = s0; s < s1; ++s) if (bf[s] >= 0) push(bf[s], 3L * all->n_atoms[s]);
    }
};

extern "C" {

int pm_comm_unique_id(char id[128]) {
    return guarded([&] {
        if (!id) throw std::invalid_argument("null argument");
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId u;
        nccl_check(nccl_api().GetUniqueId(&u), "ncclGetUniqueId");
        std::memcpy(id, &u, 128);
    });
}

int pm_comm_init_rank(pm_context* c, int n_ranks, int rank, const char id[128]) {
    return guarded([&] {
        if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::invalid_argument("invalid communicator arguments");
        if (c->comm) throw std::runtime_error("the context already has a communicator");
        CK(cudaSetDevice(c->device));
        ncclUniqueId u;
        std::memcpy(&u, id, 128);
        nccl_check(nccl_api().CommInitRank(&c->comm, n_ranks, u, rank), "ncclCommInitRank");
        c->comm_rank = rank; c->comm_size = n_ranks; c->comm_owned = true;
    });
}

int pm_comm_size(pm_context* c) { return c ? c->comm_size : -1; }
int pm_comm_rank(pm_context* c) { return c ? c->comm_rank : -1; }

int pm_comm_allreduce(pm_context* c, double* values, int n, int op) {
    return guarded([&] {
        if (!c || !values || n < 1 || n > 64) throw std::invalid_argument("pm_comm_allreduce: 1..64 values");
        if (!c->comm) return;   // a single rank: the values are their own reduction
        CK(cudaSetDevice(c->device));
        if (!c->d_small) CK(cudaMalloc(&c->d_small, 64 * sizeof(double)));
        CK(cudaMemcpyAsync(c->d_small, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const ncclRedOp_t rop = op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum);
        nccl_check(nccl_api().AllReduce(c->d_small, c->d_small, (size_t)n, ncclDouble, rop, c->comm, c->stream), "ncclAllReduce");
        CK(cudaMemcpyAsync(values, c->d_small, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    });
}

int pm_comm_barrier(pm_context* c) {
    double v = 0.0;
    return pm_comm_allreduce(c, &v, 1, 0);
}

int pm_fit_reduce(pm_context* c, int root) {
    return guarded([&] {
        if (!c) throw std::invalid_argument("null argument");
        if (!c->comm || c->comm_size == 1) return;
        if (root < 0 || root >= c->comm_size) throw std::invalid_argument("root out of range");
        join_syrk(c);
        pack_for_reduce(c);
        nccl_check(nccl_api().Reduce(c->d_red.p, c->d_red.p, packed_upper_doubles(c), ncclDouble, ncclSum, root, c->comm, c->stream),
                   "ncclReduce");
        if (c->comm_rank == root) unpack_after_reduce(c);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
    });
}

int64_t pm_fit_reduce_bytes(pm_context* c) { return c ? (int64_t)(packed_upper_doubles(c) * sizeof(double)) : -1; }

// ---- several GPUs driven by one process -----------------------------------------------------------------------------
int pm_multi_create(const pm_model* m, const int* devices, int n_dev, size_t workspace_bytes, int flags, pm_multi** out) {
    return guarded([&] {
        if (!m || !devices || n_dev < 1 || !out) throw std::invalid_argument("invalid argument");
        auto mg = std::make_unique<pm_multi>();
        try {
            for (int k = 0; k < n_dev; ++k) {
                pm_context* c = nullptr;
                if (pm_context_create(m, devices[k], workspace_bytes, flags, &c) != PM_OK) throw std::runtime_error(g_err);
                mg->ctx.push_back(c);
            }
            if (n_dev > 1) {
                mg->comms.resize(n_dev);
                nccl_check(nccl_api().CommInitAll(mg->comms.data(), n_dev, devices), "ncclCommInitAll");
                for (int k = 0; k < n_dev; ++k) {
                    mg->ctx[k]->comm = mg->comms[k]; mg->ctx[k]->comm_rank = k; mg->ctx[k]->comm_size = n_dev;
                    mg->ctx[k]->comm_owned = false;
                }
            }
        } catch (...) {
            for (pm_context* c : mg->ctx) pm_context_destroy(c);
            throw;
        }
        *out = mg.release();
    });
}

void pm_multi_destroy(pm_multi* mg) {
    if (!mg) return;
    for (pm_context* c : mg->ctx) pm_context_destroy(c);
    for (ncclComm_t cm : mg->comms) if (cm) nccl_api().CommDestroy(cm);
    delete mg;
}

int pm_multi_size(const pm_multi* mg) { return mg ? (int)mg->ctx.size() : -1; }
pm_context* pm_multi_context(pm_multi* mg, int k) { return (mg && k >= 0 && k < (int)mg->ctx.size()) ? mg->ctx[k] : nullptr; }

int pm_multi_fit_reset(pm_multi* mg) {
    return guarded([&] {
        if (!mg) throw std::invalid_argument("null argument");
        for (pm_context* c : mg->ctx)
            if (pm_fit_reset(c) != PM_OK) throw std::runtime_error(g_err);
    });
}

int pm_multi_fit_accumulate(pm_multi* mg, const pm_structures* st, const double* w, const double* y) {
    return guarded([&] {
        if (!mg || !st || !w || !y) throw std::invalid_argument("null argument");
        const int nd = (int)mg->ctx.size();
        if (nd == 1) {
            if (pm_fit_accumulate(mg->ctx[0], st, w, y) != PM_OK) throw std::runtime_error(g_err);
            return;
        }
        validate_structures(mg->ctx[0], st);
        std::vector<size_t> aoff;
        std::vector<long> be, bs, bf;
        long n_rows = 0;
        batch_layout(st, aoff, be, bs, bf, n_rows);
        // contiguous slices balanced by atom count (rows and neighbour work both scale with it)
        std::vector<int> cut(nd + 1, st->n_st);
        cut[0] = 0;
        const double total = (double)aoff[st->n_st] + st->n_st;
        for (int k = 1, s = 0; k < nd; ++k) {
            const double want = total * k / nd;
            while (s < st->n_st && (double)aoff[s] + s < want) ++s;
            cut[k] = s;
        }
        std::vector<std::string> errs(nd);
        std::vector<int> status(nd, PM_OK);
        std::vector<std::thread> th;
        for (int k = 0; k < nd; ++k) {
            if (cut[k + 1] <= cut[k]) continue;
            th.emplace_back([&, k] {
                SubBatch sb(st, aoff, be, bs, bf, cut[k], cut[k + 1], w, y);
                status[k] = pm_fit_accumulate(mg->ctx[k], &sb.st, sb.w.data(), sb.y.data());
                if (status[k] != PM_OK) errs[k] = pm_last_error();
            });
        }
        for (auto& t : th) t.join();
        for (int k = 0; k < nd; ++k)
            if (status[k] != PM_OK) {
                if (status[k] == PM_ERR_INVALID) throw std::invalid_argument(errs[k]);
                if (status[k] == PM_ERR_CUDA) throw CudaError(errs[k]);
                throw std::runtime_error(errs[k]);
            }
    });
}

int pm_multi_fit_reduce(pm_multi* mg, int root) {
    return guarded([&] {
        if (!mg) throw std::invalid_argument("null argument");
        const int nd = (int)mg->ctx.size();
        if (nd == 1) return;
        if (root < 0 || root >= nd) throw std::invalid_argument("root out of range");
        for (pm_context* c : mg->ctx) { join_syrk(c); pack_for_reduce(c); }
        nccl_check(nccl_api().GroupStart(), "ncclGroupStart");
        for (pm_context* c : mg->ctx) {
            CK(cudaSetDevice(c->device));
            nccl_check(nccl_api().Reduce(c->d_red.p, c->d_red.p, packed_upper_doubles(c), ncclDouble, ncclSum, root, c->comm, c->stream),
                       "ncclReduce");
        }
        nccl_check(nccl_api().GroupEnd(), "ncclGroupEnd");
        unpack_after_reduce(mg->ctx[root]);
        for (pm_context* c : mg->ctx) {
            CK(cudaSetDevice(c->device));
            CK(cudaStreamSynchronize(c->stream));
        }
        CK(cudaGetLastError());
    });
}

int pm_multi_fit_finalize(pm_multi* mg, double* xtx, double* xty, double* xe_sum, double* xe_sq_sum, double* y_sq_norm,
                          int64_t* n_data) {
    if (!mg) { g_err = "null argument"; return PM_ERR_INVALID; }
    // the reduce ADDS the other ranks' partial sums into device 0's accumulator: the other accumulators are cleared so
    // that a later accumulate / finalize pair does not count them twice
    const int r = pm_multi_fit_reduce(mg, 0);
    if (r != PM_OK) return r;
    for (size_t k = 1; k < mg->ctx.size(); ++k) {
        const int rr = pm_fit_reset(mg->ctx[k]);
        if (rr != PM_OK) return rr;
    }
    return pm_fit_finalize(mg->ctx[0], xtx, xty, xe_sum, xe_sq_sum, y_sq_norm, n_data);
}

}  // extern "C"

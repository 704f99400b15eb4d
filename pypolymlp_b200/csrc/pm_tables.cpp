// pm_tables.cpp -- construction of the host model tables (see pm_tables.hpp).
#include "pm_tables.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>

#include <dlfcn.h>

namespace pm {

namespace {

bool comb_in_type(const std::vector<int>& comb, const std::vector<int>& tps_of_type) {
    for (int tp : comb)
        if (std::find(tps_of_type.begin(), tps_of_type.end(), tp) == tps_of_type.end()) return false;
    return true;
}

// all sequences of length `order` over [0, n) in odometer order (first index slowest)
void sequences(int n, int order, std::vector<std::vector<int>>& out) {
    std::vector<int> cur(order, 0);
    if (n == 0) return;
    while (true) {
        out.push_back(cur);
        int k = order - 1;
        while (k >= 0 && ++cur[k] == n) { cur[k] = 0; --k; }
        if (k < 0) break;
    }
}

}  // namespace

void HostModel::build(const FeatureParams& fp_in) {
    fp = fp_in;
    const int nt = fp.n_type, n_fn = fp.n_fn;
    if (nt < 1 || n_fn < 1) throw std::invalid_argument("invalid n_type / pair_params");
    if (fp.model_type < 1 || fp.model_type > 4) throw std::invalid_argument("Polymlp: Model type error.");
    if (fp.maxp < 1 || fp.maxp > 3) throw std::invalid_argument("Polymlp: maxp must be smaller than or equal to 3.");
    const bool pair_features = fp.feature_type == 1;
    if (pair_features) {
        // feature_type = "pair" (compute/local_pair.cpp:57-122): d_{n,tp}(i) = sum_j f_n(r_ij).  Expressed through the
        // same machinery as an l = 0, order-1 invariant: a_{n,00,tp} = Y_00 sum_j f_n, coefficient 1 / Y_00, with Y_00
        // the constant k_pair_basis writes (s2pi * hs2).  Column order and polynomial rules differ (below).
        if (fp.model_type > 2) throw std::invalid_argument("Polymlp: Model type error.");   // model_params_polynomial.cpp:42-45
        const double y00 = 0.39894228040143267794 * 0.70710678118654752440;
        fp.maxl = 0;
        fp.l_comb = {{0}};
        fp.lm_seq = {{{0}}};
        fp.lm_coeffs = {{1.0 / y00}};
    } else if (fp.feature_type != 0) {
        throw std::invalid_argument("feature_type must be 0 (gtinv) or 1 (pair)");
    }
    if (fp.l_comb.empty()) throw std::invalid_argument("gtinv tables are empty");
    n_lm_half = (fp.maxl + 1) * (fp.maxl + 2) / 2;

    // ---- type pairs, conditional radial sets ------------------------------
    type_pairs.assign(nt, std::vector<int>(nt, 0));
    tp_types.clear();
    for (int i = 0; i < nt; ++i)
        for (int j = i; j < nt; ++j) {
            type_pairs[i][j] = type_pairs[j][i] = (int)tp_types.size();
            tp_types.push_back({i, j});
        }
    n_tp = (int)tp_types.size();
    if ((int)fp.cond.size() != n_tp) throw std::invalid_argument("pair_params_conditional size mismatch");
    tp_nid.assign(n_tp, std::vector<int>(n_fn, -1));
    std::vector<std::vector<int>> n_to_tp(n_fn);
    for (int tp = 0; tp < n_tp; ++tp) {
        int id = 0;
        for (int n : fp.cond[tp]) {
            if (n < 0 || n >= n_fn) throw std::invalid_argument("conditional radial index out of range");
            n_to_tp[n].push_back(tp);
            tp_nid[tp][n] = id++;
        }
    }

    // ---- lm attributes -------------------------------------------------------
    struct Lm { int l, m, key; bool conj; double cc; };
    std::vector<Lm> lms;
    for (int l = 0; l <= fp.maxl; ++l)
        for (int m = -l; m <= l; ++m) {
            int key = m < 1 ? (l + 3) * l / 2 + m : (l + 3) * l / 2 - m;
            lms.push_back({l, m, key, m > 0, (m % 2 == 0) ? 1.0 : -1.0});
        }
    const int n_lm = (int)lms.size();

    // ---- global (n, lm, tp) list: n-major, then lm, then tp --------------------
    struct G { int n, lm, tp, conj_gid; };
    std::vector<G> gl;
    std::vector<int> gindex((size_t)n_fn * n_lm * n_tp, -1);
    for (int n = 0; n < n_fn; ++n)
        for (int lm = 0; lm < n_lm; ++lm) {
            const int sub = 2 * lms[lm].m * (int)n_to_tp[n].size();
            for (int tp : n_to_tp[n]) {
                const int gid = (int)gl.size();
                gindex[((size_t)n * n_lm + lm) * n_tp + tp] = gid;
                gl.push_back({n, lm, tp, gid - sub});
            }
        }

    // ---- per-type local lists ----------------------------------------------------
    types.assign(nt, TypeTables());
    std::vector<std::vector<int>> g2l(nt, std::vector<int>(gl.size(), -1));
    for (int t = 0; t < nt; ++t) {
        TypeTables& T = types[t];
        T.type = t;
        for (int gid = 0; gid < (int)gl.size(); ++gid) {
            const auto& tt = tp_types[gl[gid].tp];
            if (tt[0] != t && tt[1] != t) continue;
            g2l[t][gid] = T.n_full++;
        }
        T.full_head.assign(T.n_full, -1);
        T.full_conj.assign(T.n_full, 0);
        T.full_cc.assign(T.n_full, 1.0);
        for (int gid = 0; gid < (int)gl.size(); ++gid) {
            const int loc = g2l[t][gid];
            if (loc < 0) continue;
            const Lm& a = lms[gl[gid].lm];
            if (!a.conj) {
                T.full_head[loc] = T.n_head++;
                T.head_full.push_back(loc);
                T.head_n.push_back(gl[gid].n);
                T.head_nid.push_back(tp_nid[gl[gid].tp][gl[gid].n]);
                T.head_key.push_back(a.key);
                T.head_l.push_back(a.l);
                T.head_m.push_back(a.m);
                T.head_tp.push_back(gl[gid].tp);
            }
        }
        for (int gid = 0; gid < (int)gl.size(); ++gid) {
            const int loc = g2l[t][gid];
            if (loc < 0) continue;
            const Lm& a = lms[gl[gid].lm];
            if (a.conj) {
                T.full_conj[loc] = 1;
                T.full_cc[loc] = a.cc;
                T.full_head[loc] = T.full_head[g2l[t][gl[gid].conj_gid]];
            }
        }
        // k-space segments per neighbour type
        T.seg_tp.assign(nt, 0);
        T.seg_heads.assign(nt, {});
        T.seg_n_off.assign(nt, std::vector<int>(n_fn + 1, 0));
        for (int u = 0; u < nt; ++u) {
            const int tp = type_pairs[t][u];
            T.seg_tp[u] = tp;
            for (int n = 0; n < n_fn; ++n) {
                T.seg_n_off[u][n] = (int)T.seg_heads[u].size();
                for (int h = 0; h < T.n_head; ++h)
                    if (T.head_tp[h] == tp && T.head_n[h] == n) T.seg_heads[u].push_back(h);
                if (T.seg_heads[u].size() % 2) T.seg_heads[u].push_back(-1);
            }
            T.seg_n_off[u][n_fn] = (int)T.seg_heads[u].size();
        }
    }

    // ---- linear terms ----------------------------------------------------------------
    const int order_max = (int)fp.l_comb.back().size();
    std::vector<std::vector<int>> tps_of_type(nt);
    for (int t = 0; t < nt; ++t) tps_of_type[t] = type_pairs[t];
    std::vector<std::vector<std::vector<int>>> tp_combs(order_max + 1);
    for (int order = 1; order <= order_max; ++order) {
        std::vector<std::vector<int>> all;
        sequences(n_tp, order, all);
        for (const auto& p : all)
            for (int t = 0; t < nt; ++t)
                if (comb_in_type(p, tps_of_type[t])) { tp_combs[order].push_back(p); break; }
    }
    std::vector<std::vector<LinearTerm>> lin_by_n(n_fn);
    for (int lcid = 0; lcid < (int)fp.l_comb.size(); ++lcid) {
        const auto& lc = fp.l_comb[lcid];
        const int order = (int)lc.size();
        if (order > order_max) throw std::invalid_argument("l_comb orders must be non-decreasing");
        std::set<std::vector<std::pair<int, int>>> uniq;
        for (const auto& tpc : tp_combs[order]) {
            std::vector<std::pair<int, int>> ms;
            for (int j = 0; j < order; ++j) ms.push_back({lc[j], tpc[j]});
            std::sort(ms.begin(), ms.end());
            uniq.insert(ms);
        }
        for (const auto& ms : uniq) {
            std::vector<int> tpc;
            for (const auto& pr : ms) tpc.push_back(pr.second);
            std::vector<int> n_list;
            for (int n = 0; n < n_fn; ++n) {
                bool ok = true;
                for (int tp : tpc) ok = ok && tp_nid[tp][n] >= 0;
                if (ok) n_list.push_back(n);
            }
            std::vector<int> t1;
            for (int t = 0; t < nt; ++t)
                if (comb_in_type(tpc, tps_of_type[t])) t1.push_back(t);
            for (int n : n_list) lin_by_n[n].push_back({n, lcid, order, tpc, t1});
        }
    }
    linear.clear();
    if (pair_features) {
        // pair models enumerate (tp, n) tp-major, n in the order of the type pair's active list
        // (polymlp_mapping.cpp:75-88 set_ntp_global_attrs)
        for (int tp = 0; tp < n_tp; ++tp)
            for (int n : fp.cond[tp])
                for (auto& x : lin_by_n[n])
                    if (x.tp_comb[0] == tp) linear.push_back(x);
    } else {
        for (auto& v : lin_by_n)
            for (auto& x : v) linear.push_back(x);
    }
    n_linear = (int)linear.size();

    // ---- per-type features: term lists ---------------------------------------------------
    for (int t = 0; t < nt; ++t) {
        TypeTables& T = types[t];
        T.max_order = order_max;
        T.term_off.push_back(0);
    }
    // local feature order of a type = global order sorted by radial index (stable): gtinv models are n-major
    // already; pair models (tp-major columns) are regrouped so that a feature tile still has ONE radial index
    std::vector<int> fid_by_n(n_linear);
    for (int k = 0; k < n_linear; ++k) fid_by_n[k] = k;
    std::stable_sort(fid_by_n.begin(), fid_by_n.end(), [&](int a, int b) { return linear[a].n < linear[b].n; });
    for (int fid : fid_by_n) {
        const LinearTerm& lt = linear[fid];
        const auto& lmlist = fp.lm_seq[lt.lcid];
        const auto& cf = fp.lm_coeffs[lt.lcid];
        for (int t : lt.types) {
            TypeTables& T = types[t];
            T.feat_gid.push_back(fid);
            for (size_t i = 0; i < lmlist.size(); ++i) {
                std::vector<int> gids;
                for (int k = 0; k < lt.order; ++k) {
                    const int lm = lmlist[i][k];
                    if (lm < 0 || lm >= n_lm) throw std::invalid_argument("lm index exceeds max_l");
                    const int gid = gindex[((size_t)lt.n * n_lm + lm) * n_tp + lt.tp_comb[k]];
                    if (gid < 0) throw std::runtime_error("inconsistent (n, lm, tp) key");
                    gids.push_back(gid);
                }
                std::sort(gids.begin(), gids.end());
                T.term_coeff.push_back(cf[i]);
                T.term_order.push_back(lt.order);
                for (int k = 0; k < order_max; ++k)
                    T.term_ids.push_back(k < lt.order ? g2l[t][gids[k]] : -1);
            }
            T.term_off.push_back((int)T.term_coeff.size());
        }
    }

    // ---- padded feature tiles (8 features, one radial index per tile) -------------------
    for (int t = 0; t < nt; ++t) {
        TypeTables& T = types[t];
        T.n_feat = (int)T.feat_gid.size();
        T.feat_pad.assign(T.n_feat, -1);
        int cur_n = -1;
        for (int f = 0; f < T.n_feat; ++f) {
            const int n = linear[T.feat_gid[f]].n;
            if (n != cur_n) {
                while (T.pad_feat.size() % 8) { T.pad_feat.push_back(-1); }
                cur_n = n;
            }
            if (T.pad_feat.size() % 8 == 0) T.tile_n.push_back(n);
            T.feat_pad[f] = (int)T.pad_feat.size();
            T.pad_feat.push_back(f);
        }
        while (T.pad_feat.size() % 8) T.pad_feat.push_back(-1);
        T.n_fpad = (int)T.pad_feat.size();
        T.pad_gid.assign(T.n_fpad, -1);
        for (int p = 0; p < T.n_fpad; ++p)
            if (T.pad_feat[p] >= 0) T.pad_gid[p] = T.feat_gid[T.pad_feat[p]];
    }

    // ---- G entries and block-sparse pattern ---------------------------------------------------
    for (int t = 0; t < nt; ++t) {
        TypeTables& T = types[t];
        // head -> (segment, position)
        std::vector<int> head_seg(T.n_head, -1), head_pos(T.n_head, -1);
        for (int u = 0; u < nt; ++u) {
            if (u > 0 && T.seg_tp[u] == T.seg_tp[u - 1]) continue;
            for (int p = 0; p < (int)T.seg_heads[u].size(); ++p) {
                const int h = T.seg_heads[u][p];
                if (h >= 0 && head_seg[h] < 0) { head_seg[h] = u; head_pos[h] = p; }
            }
        }
        // gather contributions keyed by (feature, head)
        struct Key { int f, h; bool operator<(const Key& o) const { return f != o.f ? f < o.f : h < o.h; } };
        struct CKey {
            int conj; std::vector<int> ids;
            bool operator<(const CKey& o) const { return conj != o.conj ? conj < o.conj : ids < o.ids; }
        };
        std::map<Key, std::map<CKey, double>> ent;
        const int mo = T.max_order;
        for (int f = 0; f < T.n_feat; ++f)
            for (int ti = T.term_off[f]; ti < T.term_off[f + 1]; ++ti) {
                const int o = T.term_order[ti];
                const int* ids = &T.term_ids[(size_t)ti * mo];
                for (int k = 0; k < o; ++k) {
                    const int full = ids[k];
                    const int h = T.full_head[full];
                    CKey ck;
                    ck.conj = T.full_conj[full];
                    for (int q = 0; q < o; ++q)
                        if (q != k) ck.ids.push_back(ids[q]);
                    std::sort(ck.ids.begin(), ck.ids.end());
                    const double c = T.term_coeff[ti] * (ck.conj ? T.full_cc[full] : 1.0);
                    ent[{f, h}][ck] += c;
                    ++T.n_deriv_pairs;
                }
            }
        // block pattern
        std::set<std::array<int, 3>> blkset;  // (seg, tile, kchunk)
        for (const auto& e : ent) {
            const int fp_ = T.feat_pad[e.first.f];
            blkset.insert({head_seg[e.first.h], fp_ / 8, head_pos[e.first.h] / 2});
        }
        {   // small radial groups (<= 8 k-chunks, <= 4 feature tiles): store EVERY block of a (segment, tile) so
            // that the K4a fast path addresses B fragments as base + constant offset
            const int n_tiles_ = T.n_fpad / 8;
            int max_kc = 0, max_tpn = 0;
            std::vector<int> cnt(fp.n_fn, 0);
            for (int n_ : T.tile_n) cnt[n_]++;
            for (int v : cnt) max_tpn = std::max(max_tpn, v);
            for (int u = 0; u < nt; ++u)
                for (int n = 0; n < n_fn; ++n) max_kc = std::max(max_kc, (T.seg_n_off[u][n + 1] - T.seg_n_off[u][n]) / 2);
            T.dense_blocks = max_kc <= 8 && max_tpn <= 4;
            if (T.dense_blocks)
                for (int u = 0; u < nt; ++u) {
                    if (u > 0 && T.seg_tp[u] == T.seg_tp[u - 1]) continue;
                    for (int tile = 0; tile < n_tiles_; ++tile) {
                        const int n = T.tile_n[tile];
                        for (int kc = T.seg_n_off[u][n] / 2; kc < T.seg_n_off[u][n + 1] / 2; ++kc) blkset.insert({u, tile, kc});
                    }
                }
        }
        std::map<std::array<int, 3>, int> blkid;
        const int n_tiles = T.n_fpad / 8;
        T.tile_blk_off.assign(nt, std::vector<int>(n_tiles + 1, 0));
        for (const auto& b : blkset) {
            blkid[b] = (int)T.blocks.size();
            T.blocks.push_back({b[0], b[1], b[2]});
        }
        for (int u = 0; u < nt; ++u) {
            // blocks are sorted by (seg, tile, kchunk): build tile ranges for this segment
            int idx = 0;
            while (idx < (int)T.blocks.size() && T.blocks[idx].seg < u) ++idx;
            for (int tile = 0; tile < n_tiles; ++tile) {
                T.tile_blk_off[u][tile] = idx;
                while (idx < (int)T.blocks.size() && T.blocks[idx].seg == u && T.blocks[idx].tile == tile) ++idx;
            }
            T.tile_blk_off[u][n_tiles] = idx;
        }
        T.g_size = 32L * (long)T.blocks.size();
        T.ent_off.push_back(0);
        for (const auto& e : ent) {
            const int fp_ = T.feat_pad[e.first.f];
            const int pos = head_pos[e.first.h];
            const int b = blkid[{head_seg[e.first.h], fp_ / 8, pos / 2}];
            const int nn = fp_ % 8, kre = 2 * (pos % 2);
            T.ent_pos_re.push_back(32 * b + nn * 4 + kre);
            T.ent_pos_im.push_back(32 * b + nn * 4 + kre + 1);
            for (const auto& c : e.second) {
                if (c.second == 0.0) continue;
                Contribution cb;
                cb.coeff = c.second;
                cb.conj = c.first.conj;
                cb.n_ids = (int)c.first.ids.size();
                if (cb.n_ids > 5) throw std::runtime_error("gtinv order > 6 is not supported");
                for (int q = 0; q < 5; ++q) cb.ids[q] = q < cb.n_ids ? c.first.ids[q] : 0;
                T.contribs.push_back(cb);
            }
            T.ent_off.push_back((int)T.contribs.size());
        }
    }

    // ---- polynomial combinations -------------------------------------------------------------------
    std::vector<int> pidx;
    if (fp.model_type == 2) {
        for (int k = 0; k < n_linear; ++k) pidx.push_back(k);
    } else if (fp.model_type > 2) {
        const int mo = fp.model_type == 3 ? 1 : 2;
        for (int k = 0; k < n_linear; ++k)
            if (linear[k].order <= mo) pidx.push_back(k);
    }
    auto inter = [&](const std::vector<int>& a, const std::vector<int>& b) {
        std::vector<int> r;
        std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r));
        return r;
    };
    comb2.clear();
    comb3.clear();
    std::vector<std::vector<int>> comb2_t(nt), comb3_t(nt);
    if (fp.model_type > 1 && fp.maxp > 1) {
        for (size_t i1 = 0; i1 < pidx.size(); ++i1)
            for (size_t i2 = 0; i2 <= i1; ++i2) {
                auto is = inter(linear[pidx[i1]].types, linear[pidx[i2]].types);
                if (is.empty()) continue;
                for (int t : is) comb2_t[t].push_back((int)comb2.size());
                comb2.push_back({pidx[i2], pidx[i1]});
            }
    }
    if (fp.model_type > 1 && fp.maxp > 2) {
        for (size_t i1 = 0; i1 < pidx.size(); ++i1)
            for (size_t i2 = 0; i2 <= i1; ++i2)
                for (size_t i3 = 0; i3 <= i2; ++i3) {
                    auto is = inter(inter(linear[pidx[i1]].types, linear[pidx[i2]].types), linear[pidx[i3]].types);
                    if (is.empty()) continue;
                    for (int t : is) comb3_t[t].push_back((int)comb3.size());
                    comb3.push_back({pidx[i3], pidx[i2], pidx[i1]});
                }
    }
    has_order3 = !comb3.empty();
    n_variables = n_linear + (int)comb2.size() + (int)comb3.size();

    colterm.assign(nt, std::vector<PolyTerm>(n_variables, PolyTerm{-1, 0, {-1, -1, -1}}));
    for (int t = 0; t < nt; ++t) {
        TypeTables& T = types[t];
        std::vector<int> gid2pad(n_linear, -1);
        for (int f = 0; f < T.n_feat; ++f) gid2pad[T.feat_gid[f]] = T.feat_pad[f];
        for (int f = 0; f < T.n_feat; ++f) T.poly.push_back({T.feat_gid[f], 1, {T.feat_pad[f], -1, -1}});
        for (int i : comb2_t[t])
            T.poly.push_back({n_linear + i, 2, {gid2pad[comb2[i][0]], gid2pad[comb2[i][1]], -1}});
        const int b3 = n_linear + (int)comb2.size();
        for (int i : comb3_t[t])
            T.poly.push_back({b3 + i, 3, {gid2pad[comb3[i][0]], gid2pad[comb3[i][1]], gid2pad[comb3[i][2]]}});
        for (const auto& p : T.poly) colterm[t][p.col] = p;
    }

    // dense polynomial-variable space (order-2 terms)
    std::set<int> pvset;
    for (const auto& c : comb2) { pvset.insert(c[0]); pvset.insert(c[1]); }
    pv_gid.assign(pvset.begin(), pvset.end());
    std::vector<int> gid2pv(n_linear, -1);
    for (int a = 0; a < (int)pv_gid.size(); ++a) gid2pv[pv_gid[a]] = a;
    pv_fp.assign(nt, std::vector<int>(pv_gid.size(), -1));
    for (int t = 0; t < nt; ++t) {
        const TypeTables& T = types[t];
        for (int f = 0; f < T.n_feat; ++f) {
            const int a = gid2pv[T.feat_gid[f]];
            if (a >= 0) pv_fp[t][a] = T.feat_pad[f];
        }
    }
    pair_terms.clear();
    for (int i = 0; i < (int)comb2.size(); ++i)
        pair_terms.push_back({n_linear + i, gid2pv[comb2[i][0]], gid2pv[comb2[i][1]]});
}

// ================================================================================================
// gtinv reader
// ================================================================================================

namespace {

struct Cursor {
    const unsigned char* p;
    const unsigned char* end;
    int32_t i32() {
        if (p + 4 > end) throw std::runtime_error("Binary file is broken.");
        uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        p += 4;
        return (int32_t)v;
    }
    double f64() {
        if (p + 8 > end) throw std::runtime_error("Binary file is broken.");
        uint64_t u = 0;
        for (int k = 0; k < 8; ++k) u |= (uint64_t)p[k] << (8 * k);
        p += 8;
        double x;
        std::memcpy(&x, &u, 8);
        return x;
    }
};

// One table in the reference's little-endian layout ("DATA", n_blocks, then typed ragged blocks;
// reference writer: src/pypolymlp/polyinv/binary/convert_binary.py:24-67).
void parse_table(const unsigned char* b, size_t n, std::vector<std::vector<int>>& l_all,
                 std::vector<std::vector<double>>& c_all, std::vector<std::vector<std::vector<int>>>& m_all) {
    Cursor c{b, b + n};
    if (n < 8 || std::memcmp(b, "DATA", 4) != 0) throw std::runtime_error("Binary file is broken.");
    c.p += 4;
    c.i32();  // n_blocks
    c.i32();
    l_all.resize(c.i32());
    for (auto& v : l_all) { v.resize(c.i32()); for (auto& x : v) x = c.i32(); }
    c.i32();
    c_all.resize(c.i32());
    for (auto& v : c_all) { v.resize(c.i32()); for (auto& x : v) x = c.f64(); }
    c.i32();
    m_all.resize(c.i32());
    for (auto& v : m_all) { v.resize(c.i32()); for (auto& w : v) { w.resize(c.i32()); for (auto& x : w) x = c.i32(); } }
}

std::vector<unsigned char> slurp(const std::string& path) {
    std::ifstream ifs(path, std::ios::binary);
    if (!ifs.is_open()) return {};
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
}

std::string default_datadir() {
    if (const char* e = std::getenv("POLYMLP_B200_GTINV_DIR")) return e;
    Dl_info info;
    if (dladdr((void*)&default_datadir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t s = p.find_last_of('/');
        const std::string dir = s == std::string::npos ? "." : p.substr(0, s);
        return dir + "/../data";
    }
    return ".";
}

// gtinv.pack: "PMGT", int32 n_tables, then per table: int32 version, int32 order, int32 n_bytes, bytes.
std::vector<unsigned char> load_table(const std::string& dir, int version, int order) {
    auto raw = slurp(dir + "/polymlp_gtinv_data_v" + std::to_string(version) + "_order" + std::to_string(order) + ".bin");
    if (!raw.empty()) return raw;
    auto pack = slurp(dir + "/gtinv.pack");
    if (pack.size() >= 8 && std::memcmp(pack.data(), "PMGT", 4) == 0) {
        Cursor c{pack.data() + 4, pack.data() + pack.size()};
        const int n = c.i32();
        for (int i = 0; i < n; ++i) {
            const int v = c.i32(), o = c.i32(), nb = c.i32();
            if (c.p + nb > c.end) throw std::runtime_error("Binary file is broken.");
            if (v == version && o == order) return std::vector<unsigned char>(c.p, c.p + nb);
            c.p += nb;
        }
    }
    throw std::runtime_error("Binary file not found.");
}

}  // namespace

GtinvTables read_gtinv(const std::string& datadir, int order, const std::vector<int>& maxl, int version) {
    if (order < 1 || order > 6) throw std::invalid_argument("Invalid order");
    if (version < 1 || version > 2) throw std::invalid_argument("Invalid version");
    if ((int)maxl.size() < order - 1) throw std::invalid_argument("gtinv_maxl is shorter than gtinv_order - 1");
    const std::string dir = datadir.empty() ? default_datadir() : datadir;
    GtinvTables out;
    for (int o = 1; o <= order; ++o) {
        auto raw = load_table(dir, version, o);
        std::vector<std::vector<int>> l_all;
        std::vector<std::vector<double>> c_all;
        std::vector<std::vector<std::vector<int>>> m_all;
        parse_table(raw.data(), raw.size(), l_all, c_all, m_all);
        const int ml = o > 1 ? maxl[o - 2] : 0;
        for (size_t i = 0; i < l_all.size(); ++i) {
            const auto& lc = l_all[i];
            if (ml < lc.back()) continue;
            std::vector<std::vector<int>> seq(m_all[i].size(), std::vector<int>(o));
            for (size_t j = 0; j < m_all[i].size(); ++j)
                for (int k = 0; k < o; ++k) seq[j][k] = lc[k] * lc[k] + lc[k] + m_all[i][j][k];
            out.l_comb.push_back(lc);
            out.lm_seq.push_back(seq);
            out.lm_coeffs.push_back(c_all[i]);
        }
    }
    return out;
}

// ================================================================================================
// lattice translations (host; O(n_trans) per structure)
// ================================================================================================

namespace {

struct Cell {
    double a[3][3];
    double d00, d11, d22, d01, d02, d12;
    bool r01, r02, r12;
    int ref_sum;
    void metric() {
        auto dot = [&](int c1, int c2) { return a[0][c1] * a[0][c2] + a[1][c1] * a[1][c2] + a[2][c1] * a[2][c2]; };
        d00 = dot(0, 0); d11 = dot(1, 1); d22 = dot(2, 2);
        d01 = dot(0, 1); d02 = dot(0, 2); d12 = dot(1, 2);
        // The reference writes `abs(dot01)` with only ::abs(int) visible in that translation unit
        // (neighbor_cell.cpp:36-38), i.e. the dot product is truncated to int first.  Reproduced
        // on purpose: the choice of cell decides the translation list and hence the pair order.
        auto iabs = [](double x) { return (double)std::abs((int)x); };
        r01 = iabs(d01) > 0.5 * d00 || iabs(d01) > 0.5 * d11;
        r02 = iabs(d02) > 0.5 * d00 || iabs(d02) > 0.5 * d22;
        r12 = iabs(d12) > 0.5 * d11 || iabs(d12) > 0.5 * d22;
        ref_sum = (int)r01 + (int)r02 + (int)r12;
    }
    template <typename T> void cart(T i, T j, T k, double* v) const {
        v[0] = a[0][0] * i + a[0][1] * j + a[0][2] * k;
        v[1] = a[1][0] * i + a[1][1] * j + a[1][2] * k;
        v[2] = a[2][0] * i + a[2][1] * j + a[2][2] * k;
    }
    double dist(int i, int j, int k) const {
        double v[3];
        cart(i, j, k, v);
        return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    void replace(int i, int j, int k, int col) {
        double v[3];
        cart(i, j, k, v);
        for (int r = 0; r < 3; ++r) a[r][col] = v[r];
        metric();
    }
};

}  // namespace

void find_translations(const double* axis9, double* pos, int n_atom, double cutoff, CellTranslations& out) {
    Cell c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.a[i][j] = axis9[3 * i + j];
    c.metric();
    out.refined = c.ref_sum > 0;
    if (c.ref_sum > 0) {
        int iter = 0;
        while (c.ref_sum > 0 && iter < 100) {
            if (c.r01) {
                if (c.d00 > c.d11) c.replace(1, (int)std::round(-(c.d01 * 1 + c.d12 * 0) / c.d11), 0, 0);
                else c.replace((int)std::round(-(c.d01 * 1 + c.d02 * 0) / c.d00), 1, 0, 1);
            }
            if (c.r02) {
                if (c.d00 > c.d22) c.replace(1, 0, (int)std::round(-(c.d02 * 1 + c.d12 * 0) / c.d22), 0);
                else c.replace((int)std::round(-(c.d01 * 0 + c.d02 * 1) / c.d00), 0, 1, 2);
            }
            if (c.r12) {
                if (c.d11 > c.d22) c.replace(0, 1, (int)std::round(-(c.d02 * 0 + c.d12 * 1) / c.d22), 1);
                else c.replace(0, (int)std::round(-(c.d01 * 0 + c.d12 * 1) / c.d11), 1, 2);
            }
            ++iter;
        }
        const auto& a = c.a;
        const double det = a[0][0] * a[1][1] * a[2][2] + a[0][1] * a[1][2] * a[2][0] + a[0][2] * a[1][0] * a[2][1]
                         - a[0][2] * a[1][1] * a[2][0] - a[0][1] * a[1][0] * a[2][2] - a[0][0] * a[1][2] * a[2][1];
        double inv[3][3];
        inv[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
        inv[0][1] = -(a[0][1] * a[2][2] - a[0][2] * a[2][1]);
        inv[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
        inv[1][0] = -(a[1][0] * a[2][2] - a[1][2] * a[2][0]);
        inv[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
        inv[1][2] = -(a[0][0] * a[1][2] - a[0][2] * a[1][0]);
        inv[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
        inv[2][1] = -(a[0][0] * a[2][1] - a[0][1] * a[2][0]);
        inv[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) inv[i][j] /= det;
        for (int j = 0; j < n_atom; ++j) {
            const double pc[3] = {pos[j], pos[n_atom + j], pos[2 * n_atom + j]};
            double f[3];
            for (int r = 0; r < 3; ++r) {
                f[r] = inv[r][0] * pc[0] + inv[r][1] * pc[1] + inv[r][2] * pc[2];
                f[r] -= std::floor(f[r]);
            }
            double v[3];
            c.cart(f[0], f[1], f[2], v);
            pos[j] = v[0]; pos[n_atom + j] = v[1]; pos[2 * n_atom + j] = v[2];
        }
    }
    const int mx[3] = {(int)(std::ceil(cutoff / c.dist(1, 0, 0)) + 1), (int)(std::ceil(cutoff / c.dist(0, 1, 0)) + 1),
                       (int)(std::ceil(cutoff / c.dist(0, 0, 1)) + 1)};
    double max_len = 0.0;
    for (int i = -1; i < 2; ++i)
        for (int j = -1; j < 2; ++j)
            for (int k = -1; k < 2; ++k) max_len = std::max(max_len, c.dist(i, j, k));
    out.trans.clear();
    out.tmap.clear();
    for (int d = 0; d < 3; ++d) out.mx[d] = mx[d];
    for (int i = -mx[0]; i <= mx[0]; ++i)
        for (int j = -mx[1]; j <= mx[1]; ++j)
            for (int k = -mx[2]; k <= mx[2]; ++k) {
                double v[3];
                c.cart(i, j, k, v);
                if (std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) < max_len + cutoff) {
                    out.tmap.push_back((int)(out.trans.size() / 3));
                    out.trans.push_back(v[0]); out.trans.push_back(v[1]); out.trans.push_back(v[2]);
                } else {
                    out.tmap.push_back(-1);
                }
            }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out.axis[3 * i + j] = c.a[i][j];
}

}  // namespace pm

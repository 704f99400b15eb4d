// pm_tables.hpp -- host-side model tables of a gtinv polymlp (new code, B200 layout).
//
// Semantics follow the reference's table construction (file:line relative to
// /root/reference/src/pypolymlp/cxx/src):
//   type pairs, (n,lm,tp) ordering, local ids      polymlp/polymlp_mapping.cpp:36-279
//   linear terms (n, l-comb, tp-comb, type set)    polymlp/polymlp_model_params_gtinv.cpp:53-151
//   polynomial combinations by model_type / max_p  polymlp/polymlp_model_params_polynomial.cpp:58-110,185-251
//   per-type term lists                            polymlp/polymlp_features_utils.cpp:48-76
//   per-type polynomial terms, n_variables         polymlp/polymlp_features_polynomial.cpp:22-58
// The *layout* is ours: instead of the reference's product / mapped-feature /
// potential-term maps we build, per centre type,
//   * a head list (m <= 0 order parameters) grouped by neighbour type pair and radial index,
//   * padded feature tiles (8 features, one radial index per tile),
//   * a block-sparse (4 x 8 blocks = one DMMA B fragment) pattern of
//     G[f, head] = d(feature f)/d(a_head) with the list of monomials feeding each entry,
//   * polynomial tables in a dense "polynomial variable" index space for the gather GEMM.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace pm {

struct FeatureParams {
    int n_type = 0;
    int n_fn = 0;
    std::vector<std::array<double, 2>> params;  // (beta, mu) per radial function
    std::vector<std::vector<int>> cond;         // per type pair (i<=j row-major): active radial ids
    double cutoff = 0.0;
    int model_type = 1, maxp = 1, maxl = 0;
    int feature_type = 0;                       // 0 = gtinv, 1 = pair (radial sums only; compute/local_pair.cpp)
    std::vector<std::vector<int>> l_comb;               // [n_lcomb][order]
    std::vector<std::vector<std::vector<int>>> lm_seq;  // [n_lcomb][n_terms][order], lm = l*l+l+m
    std::vector<std::vector<double>> lm_coeffs;         // [n_lcomb][n_terms]
};

struct LinearTerm {
    int n, lcid, order;
    std::vector<int> tp_comb;
    std::vector<int> types;  // centre types that carry this feature
};

struct PolyTerm {
    int col;      // global column of X
    int order;    // 1..3
    int fp[3];    // padded local feature ids of the centre type (unused = -1)
};

// One monomial feeding a G entry: coeff * prod_{k<n_ids} a_full[ids[k]], conjugated if conj.
struct Contribution {
    double coeff;
    int conj;
    int n_ids;
    int ids[5];
};

struct TypeTables {
    int type = 0;
    // ---- order parameters -------------------------------------------------
    int n_full = 0;                    // (n, lm, tp) with all m
    int n_head = 0;                    // m <= 0 only
    std::vector<int> full_head;        // full id -> head id (of its m<=0 partner)
    std::vector<int8_t> full_conj;     // 1 if m > 0
    std::vector<double> full_cc;       // (-1)^m
    std::vector<int> head_full;        // head id -> full id
    std::vector<int> head_n, head_nid, head_key, head_l, head_m, head_tp;
    // heads regrouped per neighbour type (type2 -> tp): k-space of the G / V matrices.
    // For type2 = u: heads with tp(type, u) ordered by (n, lm); per radial index a segment
    // padded to a multiple of 2 heads (= 4 reals = one DMMA k-step).
    std::vector<int> seg_tp;                    // [n_type] tp of (type, u)
    std::vector<std::vector<int>> seg_heads;    // [n_type] padded head list (-1 = padding)
    std::vector<std::vector<int>> seg_n_off;    // [n_type][n_fn+1] offsets into seg_heads (in heads)
    // ---- linear features ---------------------------------------------------
    int n_feat = 0;                    // real local features
    int n_fpad = 0;                    // padded (tiles of 8, one radial index per tile)
    std::vector<int> feat_gid;         // local -> global linear feature id
    std::vector<int> feat_pad;         // local -> padded id
    std::vector<int> pad_feat;         // padded -> local (-1 padding)
    std::vector<int> pad_gid;          // padded -> global linear id (-1 padding)
    std::vector<int> tile_n;           // [n_fpad/8] radial index of the tile
    int max_order = 1;
    std::vector<int> term_off;         // [n_feat+1]
    std::vector<double> term_coeff;
    std::vector<int> term_order;
    std::vector<int> term_ids;         // [n_terms][max_order] full ids
    // ---- G = d feature / d head, block sparse --------------------------------
    // Blocks are grouped per (type2 segment, feature tile); block = 4 reals (2 heads) x 8 features.
    struct Block { int seg; int tile; int kchunk; };  // kchunk: index of 4-real step inside the segment
    std::vector<Block> blocks;
    std::vector<std::vector<int>> tile_blk_off;  // [n_type seg][n_tiles+1] -> range in blocks (blocks sorted by seg, tile, kchunk)
    // entries: one per non-zero complex G[f, head]
    std::vector<int> ent_pos_re, ent_pos_im;     // position (in doubles) inside the atom's G buffer
    std::vector<int> ent_off;                    // [n_ent+1] -> contributions
    std::vector<Contribution> contribs;
    long g_size = 0;                             // doubles per atom (= 32 * n_blocks)
    bool dense_blocks = false;                   // every (segment, tile, k-chunk of the tile's radial group) block exists
    // ---- polynomial ----------------------------------------------------------
    std::vector<PolyTerm> poly;
    long n_deriv_pairs = 0;  // (f, full head) pairs before folding, for reporting
};

struct HostModel {
    FeatureParams fp;
    int n_tp = 0;
    std::vector<std::vector<int>> type_pairs;   // [t1][t2] -> tp
    std::vector<std::array<int, 2>> tp_types;
    std::vector<std::vector<int>> tp_nid;       // [tp][n] -> index inside cond[tp] or -1
    int n_lm_half = 0;
    std::vector<LinearTerm> linear;
    std::vector<std::array<int, 2>> comb2;
    std::vector<std::array<int, 3>> comb3;
    int n_linear = 0, n_variables = 0;
    std::vector<TypeTables> types;
    // dense polynomial-variable space for order-2 terms (gather GEMM)
    std::vector<int> pv_gid;                    // PV index -> global linear id
    std::vector<std::vector<int>> pv_fp;        // [type][PV] -> padded local id or -1
    std::vector<std::array<int, 3>> pair_terms; // (col, a, b) in PV indices, a <= b position-wise as in comb2
    // per type, per column: order and padded ids (order 0 = column absent for this type)
    std::vector<std::vector<PolyTerm>> colterm; // [type][n_variables]
    bool has_order3 = false;

    void build(const FeatureParams& fp);
};

// gtinv coupling coefficients ------------------------------------------------
struct GtinvTables {
    std::vector<std::vector<int>> l_comb;
    std::vector<std::vector<std::vector<int>>> lm_seq;
    std::vector<std::vector<double>> lm_coeffs;
};
// Reads orders 1..order of `version` and screens by maxl (reference: polymlp_read_gtinv.cpp:23-58).
// `datadir` may hold the reference's polymlp_gtinv_data_v{V}_order{O}.bin files or our gtinv.pack;
// empty -> $POLYMLP_B200_GTINV_DIR, else <dir of this shared library>/../data.
GtinvTables read_gtinv(const std::string& datadir, int order, const std::vector<int>& maxl, int version);

// Lattice translations + (possibly reduced) cell, reference: compute/neighbor_cell.cpp:11-19,126-168,203-256.
struct CellTranslations {
    double axis[9];                  // row-major 3x3, columns are a, b, c (possibly refined)
    bool refined = false;
    std::vector<double> trans;       // [n_trans][3]
    int mx[3] = {0, 0, 0};           // the list enumerates integer triples (i, j, k), |i| <= mx[0] ..., in loop order
    std::vector<int> tmap;           // [(2 mx0 + 1)(2 mx1 + 1)(2 mx2 + 1)] index into trans of a triple, or -1 (filtered out)
};
// positions_c (3 x n_atom, row-major) is rewritten in place when the cell is refined.
void find_translations(const double* axis9, double* positions_c, int n_atom, double cutoff, CellTranslations& out);

}  // namespace pm

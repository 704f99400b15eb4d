// pm_device.cuh -- device-side views of the model tables and of one chunk of structures.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pm_tables.hpp"

namespace pm {

constexpr int MAXT = 4;       // max number of atom types on the device path
constexpr int MAX_NH = 231;   // (L+1)(L+2)/2 for L = 20

struct DevContribution {
    double coeff;
    int conj;
    int n_ids;
    int ids[5];
    int pad;
};

// lane = atom eval kernel (k_eval_features_la): one 16-byte item per gtinv term / per contribution, read warp-uniformly
struct __align__(16) LaItem {
    double coeff;
    unsigned w0, w1;   // terms: full ids, 16 bits each (w0 = id0 | id1 << 16, w1 = id2 | id3 << 16); contributions: see la_hitems
};

struct DevPolyTerm {
    int order;  // 0 = column absent for this centre type
    int fp0, fp1, fp2;
};

struct DevType {
    int n_full, n_head, n_feat, n_fpad, n_tiles, max_order, n_ent, n_blocks;
    long g_size;
    const int* full_head;
    const signed char* full_conj;
    const double* full_cc;
    const int* head_nid;
    const int* head_key;
    const int* head_seg;      // neighbour type u whose pairs feed this head
    const int* seg_heads[MAXT];
    int seg_len[MAXT];        // padded number of heads in the segment (even)
    const int* seg_key[MAXT];   // [seg_len] ylm key of the head at each position (-1 padding)
    const int* seg_n_off[MAXT]; // [n_fn + 1] head offsets of each radial group inside the segment
    const int* seg_nid[MAXT];   // [n_fn] radial id inside the pair record for each radial index (-1 inactive)
    const int* tile_n_off;      // [n_fn + 1] feature tiles of each radial index
    const int* blkmap[MAXT];    // [n_tiles][kpn] block index of (tile, k-chunk inside the tile's radial group) or -1
    const int* term_off;
    const double* term_coeff;
    const int* term_order;
    const int* term_ids;
    const int* feat_pad;
    const int* ent_pos_re;
    const int* ent_pos_im;
    const int* ent_off;
    const DevContribution* contribs;
    // sliced (ELL-like) copies of the term / contribution tables for k_features_v3: a slice is a group of rows
    // (8 features x 4 k-lanes, or 32 G entries) whose slots are stored slot-major so that a warp reads 32
    // consecutive slots per iteration; padded slots carry coefficient 0.  Slices have a uniform product order.
    int n_fsl, n_esl, sl_words;   // sl_words = 32-bit id words per slot (two 16-bit full ids each)
    long n_slots;
    const int4* fsl_meta;         // [n_fsl] (first slot, iterations, order, 0)
    const int* fsl_out;           // [n_fsl][8] padded feature id or -1
    const int4* esl_meta;         // [n_esl] (first slot, iterations, number of ids, 0)
    const int2* esl_out;          // [n_esl][32] (pos_re, pos_im) in the atom's G buffer, or (-1, -1)
    const int2* esl_fh;           // [n_esl][32] (padded feature id, segment << 20 | real position inside the segment's
                                  //  head list) of the entry, or (-1, -1): the eval path folds G into the head adjoint
    // plain (warp-uniform) tables of k_eval_features_la; la_ok = 0: not built (product order > 4, mixed orders, ids >= 32768)
    int la_ok, n_la_heads;
    const LaItem* la_terms;       // [n_terms] in the order of term_off
    const int* la_forder;         // [n_feat] product order of the feature's terms
    const LaItem* la_hitems;      // contributions grouped by head, then by entry: w1 = id2 | padded feature id << 15 |
                                  // ids of the contribution << 27 | last contribution of its entry << 30
    const int* la_hoff;           // [n_la_heads + 1] item range of each head (most expensive head first)
    const int* la_hpos;           // [n_la_heads] segment << 20 | real position inside the segment's head list
    // radial-batched tables of k_eval_features_lb (lb_nr = 0: the model is not a radial replication)
    int lb_nr, lb_S, lb_Fs, lb_Ps;   // replicas; strides of the full id, the padded feature id and the head position per radial index
    int lb_G, lb_nwork, lb_nfwork;
    const LaItem* lb_terms;       // terms of the radial-0 features, most terms first
    const int* lb_foff;           // [lb_G + 1]
    const int* lb_forder;         // [lb_G]
    const int* lb_fpad;           // [lb_G] padded feature id of the radial-0 feature
    const LaItem* lb_hitems;      // contributions of the radial-0 heads (format of la_hitems)
    const int4* lb_fwork;         // [lb_nfwork] (first term, end, feature index g, 0): chunks of one feature's terms
    const int4* lb_work;          // [lb_nwork] (first item, end, head position key, 0): chunks ending on entry boundaries
    // radial-0 slices of k_features_v4r, large models (r_nr = 0: not a radial replication): the sliced tables restricted to the
    // features / entries of radial index 0; radial index n adds n * r_S to every full id, n * r_Fs to the padded feature id
    // and n * (stride of its segment) to the G positions
    int r_nr, r_S, r_Fs, r_Gs, r_n_fsl, r_n_esl;
    long r_n_slots;
    const int4* r_fsl_meta;
    const int* r_fsl_out;
    const int4* r_esl_meta;
    const int4* r_esl_out;        // (pos_re, pos_im, G stride per radial index of the entry's segment, 0)
    const double* r_sl_coeff;
    const unsigned* r_sl_ids;
    const double* sl_coeff;       // [n_slots]
    const unsigned* sl_ids;       // [sl_words][n_slots]; bit 31 of word 0 = conjugate flag (entries)
    const int* blk_kchunk;
    const int* tile_blk_off[MAXT];
    const int* tile_order[MAXT];  // [n_tiles] tiles of each radial index sorted by decreasing block count in segment u
    const int* pad_gid;     // [n_fpad] global linear column of a padded feature id or -1
    const int* pad_pv;      // [n_fpad] polynomial-variable index of a padded feature id or -1
    const DevPolyTerm* colterm;
};

struct DevModel {
    int n_type, n_fn, n_tp, maxl, nh, n_variables, n_linear;
    int fpad;      // padded width of X-tilde = [X | y | 0...], multiple of 128
    int fl;        // stride of per-centre linear-feature rows (max n_fpad over types)
    int hmax;      // max n_head over types
    int pbstride;  // doubles per pair-basis record
    int dense;     // 1 if every type stores dense blocks (B fragments addressed as base + (tile*kc_n + kc)*32)
    int kpn;       // k-chunks (2 heads) per radial group, padded to the fast-path template value (0 = no fast path)
    int tpn;       // max feature tiles per radial index, padded likewise
    int front2;    // 1: single-type model whose radial groups list all m <= 0 heads in Y_lm key order (k_lrows_v4 / k_xrows_v6)
    long gstride;  // doubles per atom in the G buffer
    double cutoff;
    const double* tp_params;  // [n_tp][n_fn][2], compacted by radial id of the pair
    const int* tp_nfn;        // [n_tp]
    const int* type_pairs;    // [n_type * n_type]
    int npv;                  // polynomial variables (order-2 terms)
    int npv_pad;              // padded to a multiple of 8
    const int* pv_fp;         // [n_type][npv_pad] padded local id or -1
    int n_pair_terms;
    const int* pair_terms;    // [n_pair_terms][3] = (col, a, b)
    const int* pair_colof;    // [64][64] column of the unordered PV pair (min, max), -1 elsewhere (npv_pad <= 64 only)
    const int* lin_fp;        // [n_type][n_linear] padded local id of each global linear feature or -1
    const int* pv_of_lin;     // [n_linear] polynomial-variable index of a linear feature or -1
    DevType types[MAXT];
};

// pair-basis record items (doubles): dx dy dz 1/r | fn[n_fn] | fn'[n_fn] | Y (re,im)[nh] | Yx | Yy | Yz
// Storage is blocked: 32 consecutive pairs form a block [item][32], so a warp that writes item k of 32
// pairs touches 256 contiguous bytes, and the records of consecutive pairs are contiguous per item.
constexpr int CL_NI = 12;      // ints per structure in DevBatch::cl_int
constexpr int CL_CAP = 160;    // neighbours of one atom the cell-list fill pass can order (else: masked-sweep fill)
constexpr int PB_BLK = 32;
struct PBRec {
    const double* base;
    __device__ __forceinline__ double operator[](int item) const { return base[(size_t)item * PB_BLK]; }
    __device__ __forceinline__ PBRec operator+(int items) const { return PBRec{base + (size_t)items * PB_BLK}; }
};
struct PBRecW {
    double* base;
    __device__ __forceinline__ double& operator[](int item) const { return base[(size_t)item * PB_BLK]; }
    __device__ __forceinline__ PBRecW operator+(int items) const { return PBRecW{base + (size_t)items * PB_BLK}; }
};
__device__ __forceinline__ PBRec pb_rec(const double* PB, int p, int stride) {
    return PBRec{PB + (size_t)(p >> 5) * stride * PB_BLK + (p & 31)};
}
__device__ __forceinline__ PBRecW pb_rec_w(double* PB, int p, int stride) {
    return PBRecW{PB + (size_t)(p >> 5) * stride * PB_BLK + (p & 31)};
}
__host__ __device__ inline size_t pb_doubles(size_t n_pairs, int stride) {
    return (n_pairs + PB_BLK - 1) / PB_BLK * PB_BLK * (size_t)stride;
}
__host__ __device__ inline int pb_fn(const DevModel& m) { return 4; }
__host__ __device__ inline int pb_fnd(const DevModel& m) { return 4 + m.n_fn; }
__host__ __device__ inline int pb_y(const DevModel& m, int comp) { return 4 + 2 * m.n_fn + comp * 2 * m.nh; }

struct DevBatch {
    int n_st, n_atoms, n_pairs, n_rows;
    int max_trans;          // largest number of lattice translations of a structure in the chunk
    int mask_stride;        // > 0: the count pass stores its hit masks at masks[i * mask_stride + j_local] for the fill pass
    ulonglong2* masks;
    int need_agg;           // 1: K2b also builds the aggregated derivative rows (fit), 0: a_nlm only (eval)
    const int* atom_off;    // [n_st + 1]
    const int* st_of_atom;  // [n_atoms]
    const int* types;       // [n_atoms]
    const double* x;        // [n_atoms] Cartesian (possibly wrapped into the refined cell)
    const double* y;
    const double* z;
    const int* trans_off;   // [n_st + 1]
    const double* trans;    // [n_trans][3]
    const int* force;       // [n_st]
    const int* erow;        // [n_st] row of the energy entry in the chunk
    const int* srow;        // [n_st] first stress row or -1
    const int* frow;        // [n_st] first force row or -1
    const double* w;        // [n_rows]
    const double* yv;       // [n_rows] weighted targets
    // device cell list (K1, when it prunes): per structure CL_NI ints {nb[3], R[3], mx[3], tmap offset, first bin} and the
    // inverse axis; per atom its bin and lattice image; bins as CSR over the chunk
    int use_cl;             // 1: cell-list kernels, 0: masked sweep over (atom, translation)
    int cl_nmax, cl_tmax;   // key = (type * cl_nmax + local atom) * cl_tmax + translation
    const int* cl_int;      // [n_st][CL_NI]
    const double* cl_ainv;  // [n_st][9]
    const int* cl_tmap;     // translation index of an integer triple or -1
    int* cl_bin_start;      // [n_bins + 1]
    int* cl_bin_atoms;      // [n_atoms]
    int* cl_atom_bin;       // [n_atoms] chunk-global bin
    int* cl_atom_img;       // [n_atoms][3] floor of the fractional coordinates
    int* cl_hkey;           // [n_atoms][CL_CAP] hits of the count pass (sort key), reused by the fill pass
    double* cl_hd;          // [n_atoms][CL_CAP][3] their displacements
    int* seg_off;           // [n_atoms * n_type + 1] pair offsets by (atom, neighbour type)
    int* nbr;               // [n_pairs] neighbour atom (chunk-global index)
    int* centre;            // [n_pairs]
    int* rev;               // [n_pairs] index of the reverse pair
};

}  // namespace pm

/* polymlp_b200.h -- C ABI of the B200-native pypolymlp hot path.
 *
 * This is the drop-in boundary: every entry point below replaces a piece of the
 * reference's pybind11 module `libmlpcpp` (reference paths are relative to
 * /root/reference/src/pypolymlp/cxx/src).  Plain C types only, caller-owned
 * buffers, no torch / pybind types.  Every function returns 0 on success and a
 * non-zero status otherwise; pm_last_error() then holds the message (thread local).
 * There is NO CPU fallback: entry points that compute need a CUDA device and fail
 * with PM_ERR_CUDA if none is usable.
 *
 * Conventions shared with the reference boundary (python/pybind11_mlp.cpp:12-67):
 *   axis         3x3 row-major, COLUMNS are the lattice vectors a, b, c
 *   positions_c  Cartesian, (3, N) row-major per structure; structures concatenated
 *   types        N ints per structure, concatenated
 *   X row order  energies | stress (6 per force structure: xx,yy,zz,xy,yz,zx) |
 *                forces (3N per force structure, row 3*atom+alpha)   (compute/py_model.cpp:58-106)
 */
#ifndef POLYMLP_B200_H
#define POLYMLP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK 0
#define PM_ERR_INVALID 1   /* std::invalid_argument in the reference -> ValueError   */
#define PM_ERR_RUNTIME 2   /* std::runtime_error in the reference   -> RuntimeError */
#define PM_ERR_CUDA 3      /* CUDA driver/runtime failure, or no device              */

typedef struct pm_model pm_model;     /* host tables of one polymlp (no GPU needed)   */
typedef struct pm_context pm_context; /* one GPU: device tables, workspaces, streams  */

const char* pm_last_error(void);
const char* pm_version(void);

/* ---- Readgtinv (polymlp/polymlp_read_gtinv.cpp:10-63; pybind11_mlp.cpp:84-94) ------------
 * Two-pass: call with all output pointers NULL to obtain sizes, then with buffers.
 * sizes[0] = n_lcomb, [1] = sum(order), [2] = sum(n_terms), [3] = sum(n_terms * order).
 * datadir NULL -> $POLYMLP_B200_GTINV_DIR or <library dir>/../data. */
int pm_gtinv_read(const char* datadir, int order, const int* maxl, int n_maxl, int version,
                  int64_t sizes[4], int* lcomb_order, int* l_comb, int* n_terms, int* lm_seq,
                  double* lm_coeffs);

/* ---- feature_params (polymlp/polymlp_mlpcpp.h:54-68, compute/py_params.cpp:11-60) ------------ */
typedef struct {
    int n_type;
    int n_fn;                  /* number of radial functions                                  */
    const double* pair_params; /* [n_fn][2] = (beta, mu)                                       */
    const int* cond_offsets;   /* [n_tp + 1], type pairs enumerated (i <= j) row-major          */
    const int* cond_values;    /* active radial ids per type pair                               */
    double cutoff;
    int model_type;            /* 1..4 */
    int max_p;                 /* 1..3 */
    int max_l;
    int n_lcomb;
    const int* lcomb_order;    /* [n_lcomb]                                                     */
    const int* l_comb;         /* concatenated, sum(order)                                      */
    const int* n_terms;        /* [n_lcomb]                                                     */
    const int* lm_seq;         /* concatenated [n_terms][order] per l-comb, lm = l*l + l + m    */
    const double* lm_coeffs;   /* concatenated [n_terms] per l-comb                             */
    int feature_type;          /* PM_FEATURE_GTINV (0) or PM_FEATURE_PAIR (1: radial sums only,
                                  compute/local_pair.cpp:57-122; the gtinv arrays are ignored, n_lcomb may be 0,
                                  model_type must be 1 or 2, polymlp_model_params_polynomial.cpp:40-56)   */
} pm_feature_params;
#define PM_FEATURE_GTINV 0
#define PM_FEATURE_PAIR 1

/* Builds all index tables on the host (replaces Features::Features, polymlp_features.cpp:27-58). */
int pm_model_create(const pm_feature_params* fp, pm_model** out);
void pm_model_destroy(pm_model* m);
/* FeaturesAttr / get_num_features replacement (compute/py_features_attr.cpp): n_variables. */
int pm_model_n_features(const pm_model* m);
/* info[0]=n_type, [1]=n_linear, [2]=n_comb2, [3]=n_comb3, [4]=n_variables, [5]=n_polyvars */
int pm_model_info(const pm_model* m, int64_t info[8]);
/* per centre type: out[0]=n_full, [1]=n_head, [2]=n_feat, [3]=n_fpad, [4]=n_terms,
 * [5]=n_G_entries, [6]=n_contributions, [7]=n_blocks, [8]=n_poly_terms, [9]=n_deriv_pairs */
int pm_model_type_info(const pm_model* m, int type, int64_t out[12]);
/* polynomial terms of a centre type: global column, order, local feature ids (-1 padded).
 * Returns the count through *n; buffers may be NULL. */
int pm_model_polynomial(const pm_model* m, int type, int* n, int* col, int* order, int* local_ids3);
/* FeaturesAttr getters (compute/py_features_attr.cpp:11-63, pybind11_mlp.cpp:70-82; consumed by get_features_attr /
 * get_num_features, PY/mlp_dev/core/features_attr.py:11-49).  Per linear feature, in column order: the radial index,
 * the l-combination id (gtinv models only: n_gtinv_ids = 0 for pair models) and its type-pair combination (CSR
 * tcomb_off[n_linear+1] / tcomb_ids); the polynomial columns (comb2 then comb3) as global linear ids (CSR poly_off /
 * poly_ids); type_pairs[n_type * n_type].  sizes[0]=n_linear, [1]=n_gtinv_ids, [2]=len(tcomb_ids), [3]=n_poly,
 * [4]=len(poly_ids), [5]=n_type.  Every buffer may be NULL (size query). */
int pm_model_feature_attrs(const pm_model* m, int64_t sizes[6], int* radial_ids, int* gtinv_ids, int* tcomb_off,
                           int* tcomb_ids, int* poly_off, int* poly_ids, int* type_pairs);
/* Algorithmic FLOPs of one structure given its neighbour statistics:
 * pairs_tt[t*n_type+u] = number of ordered pairs with centre type t, neighbour type u;
 * atoms_t[t] = atoms of type t.  out[0]=W_syrk, [1]=W_xty, [2]=W_poly, [3]=W_deriv, [4]=W_anlm. */
int pm_model_count_flops(const pm_model* m, const int64_t* atoms_t, const int64_t* pairs_tt,
                         int force, double out[5]);

/* ---- device context --------------------------------------------------------------------------- */
int pm_device_count(int* n);
/* workspace_bytes: cap for per-chunk scratch (0 -> default 12 GiB). flags: PM_FLAG_* */
#define PM_FLAG_SIMPLE_KERNELS 1 /* use the straightforward (non tensor-core) kernels: debug/parity */
#define PM_FLAG_SCATTER 2        /* K4a adds the linear columns into X with RED.F64 and keeps only the polynomial-variable
                                    derivative rows (4x less scratch per structure, same speed, run-to-run summation order) */
int pm_context_create(const pm_model* m, int device, size_t workspace_bytes, int flags, pm_context** out);
void pm_context_destroy(pm_context* c);

/* a batch of structures in HOST memory */
typedef struct {
    int n_st;
    const double* axis;        /* [n_st][9]                    */
    const double* positions_c; /* concat of (3, N_s) row-major */
    const int* types;          /* concat                       */
    const int* n_atoms;        /* [n_st]                       */
    const int* force;          /* [n_st] 0/1: build force and stress rows */
} pm_structures;

/* number of X rows of a batch in the PyModel layout */
int64_t pm_batch_rows(const pm_structures* st);

/* ---- neighbour list (compute/neighbor_full.cpp:10-76) ------------------------------------------
 * Full list of ONE structure, CSR.  Entries of an atom are grouped by neighbour type, then ordered
 * by (j, translation) as in the reference; for single-type cells this IS the reference order.
 * Two-pass: offsets[n_atom+1] is always filled; the rest only if neigh != NULL. */
int pm_neighbor_full(pm_context* c, const double* axis, const double* positions_c, const int* types,
                     int n_atom, int* offsets, int* neigh, double* dx, double* dy, double* dz);

/* ---- lattice translations / cell reduction (NeighborCell, compute/neighbor_cell.cpp:11-19,126-168,203-256) ----
 * Host only (no device needed).  positions_c: (3, n_atom) row-major.  axis_out[9] / positions_out (same shape) receive
 * the possibly refined cell and the atoms wrapped into it; *n_trans is always set, trans[cap][3] is filled when
 * cap >= *n_trans.  Same translation order as the reference. */
int pm_cell_translations(const double* axis, const double* positions_c, int n_atom, double cutoff, double* axis_out,
                         double* positions_out, int* n_trans, double* trans, int cap);

/* ---- design matrix (PotentialModel::get_x, compute/py_model.cpp:10-54) -----------------------------
 * x: row-major (pm_batch_rows, n_features) in HOST memory, unweighted. */
int pm_features_x(pm_context* c, const pm_structures* st, double* x);

/* ---- fused feature + X^T X / X^T y accumulation ------------------------------------------------------
 * Replaces compute_features -> apply_weights -> x.T @ x of
 * src/pypolymlp/mlp_dev/core/data_sequential.py:97-156 without materialising X on the host.
 * w, y: per row of the batch (PyModel layout): row weight and WEIGHTED target (y = w * target),
 * exactly the arrays the reference's apply_weights (utils_weights.py:50-98) produces. */
int pm_fit_reset(pm_context* c);
int pm_fit_accumulate(pm_context* c, const pm_structures* st, const double* w, const double* y);
/* Staged variant for benchmarking with inputs resident in HBM: stage copies the batch to the device
 * once; accumulate_staged runs the whole device pipeline on it. */
int pm_fit_stage(pm_context* c, const pm_structures* st, const double* w, const double* y);
int pm_fit_accumulate_staged(pm_context* c);
/* Device pointer and length (doubles) of the packed accumulator
 *   [ C (fpad x fpad, row-major, upper tiles valid) | xe_sum (fpad) | xe_sq_sum (fpad) | n_data (1) ]
 * where C = [X|y]^T [X|y]; exposed so that the caller can all-reduce it across GPUs (NCCL) before
 * pm_fit_finalize.  fpad = pm_fit_fpad(). */
int pm_fit_accumulator(pm_context* c, void** dev_ptr, size_t* n_doubles);
int pm_fit_fpad(pm_context* c);
/* Copies results to HOST: xtx (F x F, symmetric, row-major), xty (F), xe_sum (F), xe_sq_sum (F). */
int pm_fit_finalize(pm_context* c, double* xtx, double* xty, double* xe_sum, double* xe_sq_sum,
                    double* y_sq_norm, int64_t* n_data);
/* Same result without the final host copy: *packed points at the context's pinned host buffer
 *   [ xtx (F*F, symmetric, row-major) | xty (F) | xe_sum (F) | xe_sq_sum (F) | y_sq_norm | n_data ]
 * valid until the next pm_fit_finalize* call on this context or pm_context_destroy. */
int pm_fit_finalize_view(pm_context* c, const double** packed, size_t* n_doubles);
/* ---- multi-GPU reduce (replaces the OpenMP-over-structures loop of compute/py_model.cpp:39-53 and the batch
 * sum of src/pypolymlp/mlp_dev/core/data_sequential.py:49-70 across devices) -----------------------------------------------
 * Structures shard over GPUs, every GPU accumulates its own partial sums, and ONE ncclReduce (fp64 sum over NVLink) of the
 * packed upper 128 x 128 tiles of C plus xe_sum / xe_sq_sum / n_data combines them on the root.  NCCL is resolved at run
 * time (libnccl.so.2, override with $POLYMLP_B200_NCCL_LIB); nothing else of the library needs it.
 * (a) one process per GPU (torchrun / mpirun): rank 0 calls pm_comm_unique_id, the 128 bytes travel out of band
 *     (file, socket, MPI), every rank calls pm_comm_init_rank on its context; pm_fit_reduce is collective. */
#define PM_UNIQUE_ID_BYTES 128
int pm_comm_unique_id(char id[PM_UNIQUE_ID_BYTES]);
int pm_comm_init_rank(pm_context* c, int n_ranks, int rank, const char id[PM_UNIQUE_ID_BYTES]);
int pm_comm_size(pm_context* c);
int pm_comm_rank(pm_context* c);
/* sums the ranks' accumulators into the root's (in place on the root; the other ranks keep their partial sums).
 * A context without a communicator is a single rank: no-op. */
int pm_fit_reduce(pm_context* c, int root);
int64_t pm_fit_reduce_bytes(pm_context* c);   /* bytes each rank contributes to the reduce */
/* small host-value collectives for drivers (timing: max over ranks): op 0 = sum, 1 = max, 2 = min; n <= 64 */
int pm_comm_allreduce(pm_context* c, double* values, int n, int op);
int pm_comm_barrier(pm_context* c);
/* (b) several GPUs driven by one process (the pybind11 module's PotentialXtX(params, devices=[...])): one context per
 *     device, a single-process ncclCommInitAll communicator, one host thread per device during accumulate. */
typedef struct pm_multi pm_multi;
int pm_multi_create(const pm_model* m, const int* devices, int n_dev, size_t workspace_bytes, int flags, pm_multi** out);
void pm_multi_destroy(pm_multi* mg);
int pm_multi_size(const pm_multi* mg);
pm_context* pm_multi_context(pm_multi* mg, int k);
int pm_multi_fit_reset(pm_multi* mg);
/* shards the batch over the devices (contiguous slices balanced by atom count); w / y as in pm_fit_accumulate */
int pm_multi_fit_accumulate(pm_multi* mg, const pm_structures* st, const double* w, const double* y);
int pm_multi_fit_reduce(pm_multi* mg, int root);
/* reduce onto device 0, clear the other accumulators, then pm_fit_finalize of device 0 */
int pm_multi_fit_finalize(pm_multi* mg, double* xtx, double* xty, double* xe_sum, double* xe_sq_sum, double* y_sq_norm,
                          int64_t* n_data);

/* ---- ridge solve on the device (reported separately from the hot path) ---------------------------------
 * Replaces the tail of calc_xtx_xty (scales, zeroing, normalisation; data_sequential.py:72-92), solver_ridge
 * (Cholesky per alpha with incremental diagonal update; src/pypolymlp/mlp_dev/standard/solvers.py:48-84) and
 * compute_rmse from X^T X (utils_model_selection.py:37-72) on the accumulator that is already resident:
 * cuSOLVER potrf/potrs + cuBLAS symv.  scales_in may be NULL (then they are computed from xe_sum / xe_sq_sum
 * with n_energy = number of energy rows accumulated, as compute_scales does).  Outputs (host): scales_out[F],
 * coefs[n_alpha][F] (in the scaled basis, like the reference's coefs_array columns), rmse[n_alpha].
 * A failed factorisation yields coefficients 1e30 for that alpha (solvers.py:40-45). */
int pm_fit_solve_ridge(pm_context* c, const double* alphas, int n_alpha, const double* scales_in, int64_t n_energy,
                       int include_force, double scale_threshold, double* scales_out, double* coefs, double* rmse);

/* Blocks until all queued device work of the context is finished. */
int pm_synchronize(pm_context* c);
/* CUDA stream (cudaStream_t) the context launches on, for event timing by the caller. */
void* pm_stream(pm_context* c);
/* Device-side stopwatch: CUDA events on the context's stream (every kernel, copy and NCCL call of the context is ordered
 * on it).  start synchronises the stream first; stop records, waits and returns the elapsed milliseconds. */
int pm_timer_start(pm_context* c);
int pm_timer_stop(pm_context* c, double* ms);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
int64_t pm_launch_count(pm_context* c);
/* Accumulated device time (ms, CUDA events) per pipeline stage since the last reset; names are
 * returned through a static table: pm_stage_name(i).  on = 1: every stage, with a host synchronisation after each
 * (single stream; the idle gaps inflate the stage times by a few percent).  on = 2: only the SYRK launches, bracketed
 * by event pairs on the stream they run on, no synchronisation (the live figure bench.py's roofline uses). */
int pm_profile_enable(pm_context* c, int on);
int pm_profile_get(pm_context* c, int* n_stages, double* ms, int64_t* launches);
const char* pm_stage_name(int i);

/* ---- energy / force / stress (PotentialPropertiesFast, compute/py_properties_fast.cpp:10-76) --------
 * coeffs: n_features regression coefficients (already divided by scales, as in polymlp.yaml).
 * energies [n_st] (eV/cell), forces concat [N_s][3] (eV/A), stresses [n_st][6] (eV/cell). */
int pm_eval_set_coeffs(pm_context* c, const double* coeffs, int n);
int pm_eval(pm_context* c, const pm_structures* st, double* energies, double* forces, double* stresses);

/* ---- test hooks: intermediates of the LAST processed chunk (device -> host) --------------------------
 * what: 0 = a_nlm heads (complex, [atom][n_head_max]), 1 = linear features d ([atom][fl]),
 *       2 = pair basis records, 3 = X-tilde chunk ([rows][fpad]).  *n returns the number of doubles. */
int pm_debug_fetch(pm_context* c, int what, double* out, size_t cap, size_t* n);
/* micro-benchmarks: which = 0 DFMA, 1 DMMA (mma.sync m8n8k4 f64), 2 both interleaved, 3 cuBLAS DGEMM n^3 (TFLOP/s);
 * 4: fp64 RED throughput (giga atomics / s); 5: DMMA dependent-issue probe, n = 100 * warps + independent chains per
 * warp on one SM, returns SM cycles per DMMA of one warp. */
int pm_microbench(pm_context* c, int which, int n, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* POLYMLP_B200_H */

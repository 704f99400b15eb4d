// pm_kernels.cuh -- launch wrappers of the sm_100a kernels (definitions in pm_kernels*.cu).
#pragma once
#include "pm_device.cuh"

namespace pm {

struct Workspace {
    // per chunk scratch (device)
    int* counts = nullptr;      // [n_atoms * n_type]
    double* PB = nullptr;       // [n_pairs][pbstride]
    double2* anc = nullptr;     // [n_atoms][hmax]
    double2* agg = nullptr;     // [n_atoms][hmax][9]
    double* dfeat = nullptr;    // [n_atoms][fl]
    double* dpv = nullptr;      // [n_atoms][64] polynomial variables of each atom, compact (single-type models; else null)
    double* Gbuf = nullptr;     // [n_atoms][gstride]
    double* Lbuf = nullptr;     // [n_pairs * 3][fl]   (not used in scatter mode)
    bool lt = false;            // Lbuf is target-major ("Lt", k_lrows_v4 -> k_xrows_v6): [(n_pairs + n_atoms) * 3][fl], block of
                                // atom k starts at row 3 * (seg_off[k] + k): own row, then the negated rows of its neighbours
    double* Lpv = nullptr;      // [n_pairs * 3][npv_pad] derivative rows of the polynomial variables (scatter mode)
    bool scatter = false;       // K4a adds the linear columns straight into X with RED.F64, K4b only does the GEMM part
    double* Xown = nullptr;     // [n_atoms][3][fl]
    double* Sbuf = nullptr;     // [n_atoms][6][fl]
    double* X = nullptr;        // [n_rows][fpad]
    double* Ah = nullptr;       // [n_atoms][ah_stride] head adjoints (eval)
    bool pairs_rc = false;      // eval: PB holds only the displacements, the pair pass recomputes the records (k_eval_pairs_rc)
    bool pairs_rads = false;    // ... and 1 / r, f_n, f_n' (left there by k_anlm_eval or the storing K2 kernels): only the angular part is recomputed
    const double* cmat = nullptr;  // [n_type][64][64] order-2 coefficient matrix over the polynomial variables (eval, max_p = 2) or null
    const double* clin = nullptr;  // [n_type][fl] linear-column coefficient of each padded feature (eval, with cmat) or null
    int* errflag = nullptr;     // device error flag
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: remembers, per (device, kernel), the
// largest size configured so far (several contexts on several devices may live in one process)
void ensure_smem(const void* kernel, size_t smem_bytes, bool max_carveout = false);

// K1: neighbour list
void launch_neighbor_count(const DevModel& m, const DevBatch& b, int* counts, cudaStream_t s);
void launch_neighbor_fill(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s);
void launch_neighbor_rev(const DevModel& m, const DevBatch& b, const double* PB, int* errflag, cudaStream_t s);
// K1 with a device cell list: bins of >= r_c / 2 in fractional space, +-2 bins searched, every candidate image decided by
// the same exact arithmetic as the sweep, hits ordered by (neighbour type, j, translation) -> the identical list
void launch_cl_bins(const DevModel& m, const DevBatch& b, int n_bins, int* bin_count, cudaStream_t s);
void launch_neighbor_cl_count(const DevModel& m, const DevBatch& b, int* counts, int* max_count, cudaStream_t s);
void launch_neighbor_cl_fill(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s);
// K2: pair basis + order parameters
void launch_pair_basis(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s);
void launch_anlm(const DevModel& m, const DevBatch& b, const double* PB, double2* anc, double2* agg, cudaStream_t s,
                 bool small_footprint = false);
// K2a + K2b fused for small angular expansions; false -> call launch_pair_basis + launch_anlm instead
bool launch_pair_anlm(const DevModel& m, const DevBatch& b, double* PB, double2* anc, double2* agg, cudaStream_t s,
                      bool store_records = true);
// eval: the pair pass recomputes the basis records from the displacements (k_eval_pairs_rc), so the fused K2 kernel need
// not store them
bool eval_pairs_rc_supported(const DevModel& m);
// K2 of the eval path when the pair pass recomputes: a_nlm only, several atoms per CTA, nothing stored per pair; false if
// the model is not served (then launch_pair_anlm / launch_anlm)
bool launch_anlm_eval(const DevModel& m, const DevBatch& b, double* PB, double2* anc, cudaStream_t s, bool store_radial);
// K3: invariants and G = d feature / d head
void launch_features(const DevModel& m, const DevBatch& b, const double2* anc, double* dfeat, double* Gbuf,
                     size_t smem_bytes, cudaStream_t s, bool zero_g = true,
                     double* dpv = nullptr);
// K4a: per-centre derivative rows L = V . G
void launch_lrows(const DevModel& m, const DevBatch& b, const Workspace& ws, bool simple, bool apply_weights, cudaStream_t s);
bool scatter_mode_supported(const DevModel& m);
// second-generation front end (pm_kernels_front.cu): K4a writes target-major rows, K4b streams them
bool front_v2_supported(const DevModel& m);
bool launch_lrows_v4(const DevModel& m, const DevBatch& b, const Workspace& ws, cudaStream_t s);
void launch_xrows_v6(const DevModel& m, const DevBatch& b, const Workspace& ws, bool apply_weights, cudaStream_t s);
// K4b: gather + polynomial expansion -> weighted X-tilde rows
void launch_xrows(const DevModel& m, const DevBatch& b, const Workspace& ws, double* xe_sum, double* xe_sq,
                  bool simple, bool apply_weights, cudaStream_t s);
// K4b (tensor-core flavour) writes every column of every row for this model: no memset of the X chunk needed
bool xrows_fills_rows(const DevModel& m, bool scatter);
// exclusive prefix sum (segment offsets of the neighbour list)
void launch_scan_exclusive(const int* in, int* out, int n, cudaStream_t s);
// K5: C += Xt^T Xt (upper tiles).  The scratch holds the parked partial tiles of the deterministic stream-K fix-up
// (2 slots of 128 x 128 doubles per CTA) and one arrival counter per tile (zero between launches).
struct SyrkScratch {
    double* partials = nullptr;
    int* counters = nullptr;
    long n_slots = 0;
    long n_counters = 0;
};
void launch_syrk(const double* X, int n_rows, int fpad, double* C, bool simple, cudaStream_t s,
                 const SyrkScratch* scratch = nullptr);
int syrk_launches(int n_rows, int fpad, bool simple);
// eval: E/F/S through contraction with coefficients
// feat_smem > 0: fused feature + polynomial-adjoint kernel (no G buffer, no separate K3 launch); it must be the
// bytes of one atom's full a_nlm array and eval_fused_supported() must hold
void launch_eval_adjoint(const DevModel& m, const DevBatch& b, const Workspace& ws, const double* coeffs,
                         double* energies, double* forces, double* stresses, cudaStream_t s, size_t feat_smem = 0);
bool eval_fused_supported(const DevModel& m, size_t feat_smem);
// K3 for large radial-replication models (DevType::r_nr > 0): each decoded table slot serves a block of radial indices;
// false if the model is not served (then launch_features runs k_features_v3 & co.)
bool launch_features_radial(const DevModel& m, const DevBatch& b, const double2* anc, double* dfeat, double* Gbuf,
                            size_t smem_bytes, cudaStream_t s, bool zero_g, double* dpv);
// micro-benchmarks (TFLOP/s)
double microbench_fp64(int which, cudaStream_t s);

}  // namespace pm

// pm_kernels_front.cu -- second generation of the derivative front end (K4a + K4b) for single-type models whose
// radial groups are small (the BASELINE config-1/2/5 model family).  Replaces, behind the same launch wrappers,
//   k_lrows_v3 -> k_lrows_v4    (Local::compute_anlmtp_d derivative part + Features::compute_features_deriv,
//                                compute/local.cpp:138-196, polymlp/polymlp_features.cpp:208-248)
//   k_xrows_v5 -> k_xrows_v6    (Model::model_order1..2, compute/model.cpp:159-236, apply_weights,
//                                src/pypolymlp/mlp_dev/core/utils_weights.py:50-98)
//
// What changed against the first generation:
//  * no block-level barrier in either main loop.  K4a is a persistent kernel whose warps run free over
//    (centre atom, 16-row job) work items; the only shared state is the centre's G matrix (already stored as DMMA
//    B fragments), which a producer thread streams into a two-deep shared-memory ring with one cp.async.bulk per
//    centre (mbarrier full / empty).  A operands are built in registers straight from the pair-basis records.
//  * the derivative rows leave K4a in TARGET-MAJOR order ("Lt"): the rows that X row atom k needs -- its own row
//    first, then one (x, y, z) triple per neighbour, already negated -- are contiguous, so K4b streams
//    [3 (M_k + 1)] x fl doubles with a plain TMA ring instead of chasing 55 pointers.
//  * K4b is warp specialised: one producer warp feeds a 4-stage ring (4 centres = one DMMA k-step per stage), eight
//    MMA warps take their A / B fragments straight from the staged rows (no extraction pass, no Lambda tile) and
//    fold the linear-column sums into the same pass.
#include "pm_kernels.cuh"
#include "pm_mma.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace pm {

// ================================================================================================
// K4a v4
// ================================================================================================
constexpr int L4_W = 12;                        // consumer warps (3 per SM sub-partition)
constexpr int L4_THREADS = (L4_W + 1) * 32;     // + the producer warp
constexpr int L4_JOB_ROWS = 16;                 // two DMMA row tiles per job

// DMMAs of one (16-row job, radial group): A fragments formed on the fly from the lm factors and the two radial scalars
// of each row; B fragments = coalesced 256-byte reads of the centre's G blocks in shared memory.  FULL: the group has
// exactly TPN x KPN dense blocks (constant offsets); AGG: some rows of the job are aggregated rows (own x/y/z, six
// virial rows) whose A operand is the K2b sum of the head itself.
template <int TPN, int KPN, bool FULL, bool AGG>
__device__ __forceinline__ void l4_group(double (&acc)[TPN][2][2], const double (&a1)[2][KPN], const double (&a2)[2][KPN],
                                         const double (&cd)[2], const double (&cf)[2], const double* __restrict__ Gn,
                                         const int* __restrict__ bm, int ntile, int kcn,
                                         const double2* __restrict__ aggi, const int* __restrict__ hd,
                                         const int (&ragg)[2], int q) {
#pragma unroll
    for (int kc = 0; kc < KPN; ++kc) {
        double af[2];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) af[rt] = cd[rt] * a1[rt][kc] + cf[rt] * a2[rt][kc];
        if (AGG) {
            const int k = 4 * kc + q;
            const int h = hd[k >> 1];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
                if (ragg[rt] >= 0) {
                    double2 v = make_double2(0.0, 0.0);
                    if (h >= 0) v = aggi[(size_t)h * 9 + ragg[rt]];
                    af[rt] = (k & 1) ? v.y : v.x;
                }
        }
#pragma unroll
        for (int tt = 0; tt < TPN; ++tt) {
            double bf;
            if (FULL) {
                bf = Gn[(tt * KPN + kc) * 32];
            } else {
                const int bi = (tt < ntile && kc < kcn) ? bm[tt * KPN + kc] : -1;
                bf = bi >= 0 ? Gn[32 * (size_t)bi] : 0.0;
            }
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) dmma(acc[tt][rt][0], acc[tt][rt][1], af[rt], bf);
        }
    }
}

template <int TPN, int KPN>
__global__ void __maxnreg__(152)
k_lrows_v4(DevModel m, DevBatch b, const double* __restrict__ PB, const double2* __restrict__ agg,
           const double* __restrict__ Gbuf, double* __restrict__ Lt, double* __restrict__ Sbuf) {
    extern __shared__ __align__(128) double smem[];
    const DevType& T = m.types[0];
    const int gsz = (int)T.g_size;
    double* Gs = smem;                                                          // [2][gsz] B fragments of two centres
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Gs + 2 * (size_t)gsz);   // full[2], empty[2]
    int* s_base = reinterpret_cast<int*>(bars + 4);      // [n_fn] first G block of the radial index (dense groups)
    int* s_tile0 = s_base + m.n_fn;                      // [n_fn + 1]
    int* s_kcn = s_tile0 + m.n_fn + 1;                   // [n_fn] k-chunks of the radial group
    int* s_nid = s_kcn + m.n_fn;                         // [n_fn] radial id inside the pair record or -1
    int* s_head = s_nid + m.n_fn;                        // [n_fn][2 * KPN] head id per position or -1
    int* s_bmap = s_head + m.n_fn * 2 * KPN;             // [n_tiles][KPN]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int fl = m.fl;

    for (int e = tid; e <= m.n_fn; e += L4_THREADS) s_tile0[e] = T.tile_n_off[e];
    for (int e = tid; e < m.n_fn; e += L4_THREADS) {
        const int o0 = T.seg_n_off[0][e], o1 = T.seg_n_off[0][e + 1];
        s_kcn[e] = (o1 - o0) >> 1;
        s_nid[e] = T.seg_nid[0][e];
        s_base[e] = T.blkmap[0][(size_t)T.tile_n_off[e] * KPN];
    }
    for (int e = tid; e < m.n_fn * 2 * KPN; e += L4_THREADS) {
        const int n = e / (2 * KPN), hq = e - n * 2 * KPN;
        const int o0 = T.seg_n_off[0][n], o1 = T.seg_n_off[0][n + 1];
        s_head[e] = hq < o1 - o0 ? T.seg_heads[0][o0 + hq] : -1;
    }
    for (int e = tid; e < T.n_tiles * KPN; e += L4_THREADS) s_bmap[e] = T.blkmap[0][e];
    if (tid == 0) {
        mbar_init(smem_u32(bars + 0), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_init(smem_u32(bars + 2), L4_W);
        mbar_init(smem_u32(bars + 3), L4_W);
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 2);
    const unsigned gbytes = (unsigned)(gsz * sizeof(double));

    if (warp == L4_W) {   // ---- producer: one cp.async.bulk per centre into the ring -----------------------------
        if (lane == 0) {
            int c_idx = 0;
            for (int i = blockIdx.x; i < b.n_atoms; i += gridDim.x) {
                if (!b.force[b.st_of_atom[i]]) continue;
                const int buf = c_idx & 1;
                if (c_idx >= 2) mbar_wait(bar_empty + 8 * buf, (unsigned)((c_idx >> 1) - 1) & 1u);
                mbar_expect_tx(bar_full + 8 * buf, gbytes);
                bulk_g2s(smem_u32(Gs + (size_t)buf * gsz), Gbuf + (size_t)i * m.gstride, gbytes, bar_full + 8 * buf);
                ++c_idx;
            }
        }
        return;
    }

    // ---- consumers ----------------------------------------------------------------------------------------------
    const int oy = pb_y(m, 0);
    const int k_real = 2 * m.nh;
    const size_t pbblk = (size_t)m.pbstride * PB_BLK;
    int c_idx = 0, jobmod = 0;   // jobmod = (jobs handed out so far) mod L4_W
    for (int i = blockIdx.x; i < b.n_atoms; i += gridDim.x) {
        if (!b.force[b.st_of_atom[i]]) continue;
        const int buf = c_idx & 1;
        const int p0 = b.seg_off[i], np = b.seg_off[i + 1] - p0;
        const int nrow = 3 * np, nrow_all = nrow + 9;
        const int njobs = (nrow_all + L4_JOB_ROWS - 1) / L4_JOB_ROWS;
        mbar_wait(bar_full + 8 * buf, (unsigned)(c_idx >> 1) & 1u);
        const double* G = Gs + (size_t)buf * gsz;

        int tj = warp - jobmod;
        if (tj < 0) tj += L4_W;
        for (; tj < njobs; tj += L4_W) {
            const int row0 = tj * L4_JOB_ROWS;
            double a1[2][KPN], a2[2][KPN];
            const double* recp[2];
            double* rowp[2];
            int ragg[2];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                const int r = row0 + rt * 8 + g;
                ragg[rt] = (r >= nrow && r < nrow_all) ? r - nrow : -1;
                recp[rt] = nullptr;
                rowp[rt] = nullptr;
                if (r < nrow) {
                    const int pl = r / 3, al = r - 3 * pl, p = p0 + pl;
                    const double* rec = PB + (size_t)(p >> 5) * pbblk + (p & 31);
                    recp[rt] = rec;
                    const double dal = rec[al * PB_BLK] * rec[3 * PB_BLK];
                    const double* ry = rec + (size_t)oy * PB_BLK;
                    const double* rya = rec + (size_t)pb_y(m, 1 + al) * PB_BLK;
#pragma unroll
                    for (int kc = 0; kc < KPN; ++kc) {
                        const int k = 4 * kc + q;
                        const bool ok = k < k_real;
                        a1[rt][kc] = ok ? ry[(size_t)k * PB_BLK] * dal : 0.0;
                        a2[rt][kc] = ok ? rya[(size_t)k * PB_BLK] : 0.0;
                    }
                    // force on atom j is -d/dr_j: the row of pair (i -> j) lands, negated, in the slot of the reverse
                    // pair in j's block of Lt (slot 0 of a block is the atom's own row)
                    const int j = b.nbr[p];
                    rowp[rt] = Lt + ((size_t)(b.rev[p] + j + 1) * 3 + al) * fl + 2 * q;
                } else {
#pragma unroll
                    for (int kc = 0; kc < KPN; ++kc) { a1[rt][kc] = 0.0; a2[rt][kc] = 0.0; }
                    if (ragg[rt] >= 0)
                        rowp[rt] = (ragg[rt] < 3 ? Lt + ((size_t)(p0 + i) * 3 + ragg[rt]) * fl
                                                 : Sbuf + ((size_t)i * 6 + (ragg[rt] - 3)) * fl) + 2 * q;
                }
            }
            const bool has_agg = row0 + L4_JOB_ROWS > nrow;   // warp uniform
            double cdn[2] = {0.0, 0.0}, cfn[2] = {0.0, 0.0};
            {
                const int nid = s_nid[0];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt)
                    if (recp[rt] && nid >= 0) {
                        cdn[rt] = recp[rt][(size_t)(4 + m.n_fn + nid) * PB_BLK];
                        cfn[rt] = recp[rt][(size_t)(4 + nid) * PB_BLK];
                    }
            }
            for (int n = 0; n < m.n_fn; ++n) {
                const double cd[2] = {cdn[0], cdn[1]}, cf[2] = {cfn[0], cfn[1]};
                if (n + 1 < m.n_fn) {   // radial scalars of the next group: in flight during this group's DMMAs
                    const int nid = s_nid[n + 1];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        cdn[rt] = 0.0; cfn[rt] = 0.0;
                        if (recp[rt] && nid >= 0) {
                            cdn[rt] = recp[rt][(size_t)(4 + m.n_fn + nid) * PB_BLK];
                            cfn[rt] = recp[rt][(size_t)(4 + nid) * PB_BLK];
                        }
                    }
                }
                const int tile0 = s_tile0[n];
                const int ntile = s_tile0[n + 1] - tile0;
                const int kcn = s_kcn[n];
                double acc[TPN][2][2];
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt)
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) { acc[tt][rt][0] = 0.0; acc[tt][rt][1] = 0.0; }
                if (kcn > 0) {
                    const bool full = m.dense && kcn == KPN && ntile == TPN;
                    const double* Gn = G + 32 * (size_t)(full ? s_base[n] : 0) + lane;
                    const int* bm = s_bmap + tile0 * KPN;
                    const double2* aggi = agg + (size_t)i * m.hmax * 9;
                    const int* hd = s_head + n * 2 * KPN;
                    if (full) {
                        if (!has_agg) l4_group<TPN, KPN, true, false>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                        else l4_group<TPN, KPN, true, true>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                    } else {
                        if (!has_agg) l4_group<TPN, KPN, false, false>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                        else l4_group<TPN, KPN, false, true>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                    }
                }
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt) {
                    if (tt >= ntile) break;
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        if (!rowp[rt]) continue;
                        const double sg = ragg[rt] < 0 ? -1.0 : 1.0;
                        *reinterpret_cast<double2*>(rowp[rt] + (tile0 + tt) * 8) =
                            make_double2(sg * acc[tt][rt][0], sg * acc[tt][rt][1]);
                    }
                }
            }
        }
        jobmod = (jobmod + njobs) % L4_W;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
        ++c_idx;
    }
}

static size_t lrows_v4_smem(const DevModel& m, int kpn) {
    const DevType& T = m.types[0];
    const size_t ints = 4ull * m.n_fn + 2 + (size_t)m.n_fn * 2 * kpn + (size_t)T.n_tiles * kpn;
    return 2ull * (size_t)T.g_size * sizeof(double) + 4 * sizeof(unsigned long long) + ints * sizeof(int) + 128;
}

template <int TPN, int KPN>
static void launch_lrows_v4_t(const DevModel& m, const DevBatch& b, const Workspace& ws, cudaStream_t s) {
    static int n_sm = 0;
    static size_t set_for = 0;
    const size_t smem = lrows_v4_smem(m, KPN);
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    if (set_for < smem) {
        cudaFuncSetAttribute(k_lrows_v4<TPN, KPN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        set_for = smem;
    }
    const int grid = std::min(n_sm, b.n_atoms);
    k_lrows_v4<TPN, KPN><<<grid, L4_THREADS, smem, s>>>(m, b, ws.PB, ws.agg, ws.Gbuf, ws.Lbuf, ws.Sbuf);
}

bool launch_lrows_v4(const DevModel& m, const DevBatch& b, const Workspace& ws, cudaStream_t s) {
#define PM_L4_CASE(TP, KP) if (m.tpn == TP && m.kpn == KP) { launch_lrows_v4_t<TP, KP>(m, b, ws, s); return true; }
    PM_L4_CASE(1, 2) PM_L4_CASE(2, 2) PM_L4_CASE(3, 2) PM_L4_CASE(4, 2)
    PM_L4_CASE(1, 4) PM_L4_CASE(2, 4) PM_L4_CASE(3, 4) PM_L4_CASE(4, 4)
    PM_L4_CASE(1, 8) PM_L4_CASE(2, 8) PM_L4_CASE(3, 8) PM_L4_CASE(4, 8)
#undef PM_L4_CASE
    return false;
}

// ================================================================================================
// K4b v6
// ================================================================================================
constexpr int X6_MMA_WARPS = 8;
constexpr int X6_THREADS = (X6_MMA_WARPS + 1) * 32;
constexpr int X6_STAGES = 4;
constexpr int X6_KC = 4;           // centres per stage = one DMMA k-step
constexpr int X6_LDD = 68;         // row stride of the staged polynomial-variable rows (== 4 mod 16)

// upper-triangle tiles (ta <= tb) of the 8 x 8 grid of 8 x 8 blocks over the <= 64 polynomial variables, dealt to
// the eight MMA warps as row strips so that the A-side fragments of a strip are loaded once: warp 2w owns
// (w, w .. w+4); warp 2w+1 owns the rest of row w and row 7-w (rows w and 7-w hold 9 tiles together).
__host__ __device__ constexpr int x6_ntiles(int W) { return (W & 1) ? 4 : 5; }
__host__ __device__ constexpr int x6_ta(int W, int i) {
    const int w = W >> 1;
    return (W & 1) ? (i < 3 - w ? w : 7 - w) : w;
}
__host__ __device__ constexpr int x6_tb(int W, int i) {
    const int w = W >> 1;
    return (W & 1) ? (i < 3 - w ? w + 5 + i : 7 - w + (i - (3 - w))) : w + i;
}

struct X6Args {
    const double* stage;   // [X6_STAGES][X6_KC][sl]
    const double* sD;      // [X6_STAGES][X6_KC][X6_LDD]
    unsigned bar_full, bar_empty;
    int sl, fl, n_cent, nst;
};

template <int W>
__device__ __forceinline__ void x6_consume(const DevModel& m, const X6Args& a, int lane, int ctid, double (&lin)[3],
                                           double (&acc)[5][3][2]) {
    constexpr int NT = x6_ntiles(W);
    const int g = lane >> 2, q = lane & 3;
    // padded feature id of polynomial variable t * 8 + g for the eight tile rows / columns (-1: no such variable)
    int fpv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) fpv[t] = (t * 8 + g < m.npv_pad) ? m.pv_fp[t * 8 + g] : -1;
    const int nlin = 3 * a.fl;
    for (int st = 0; st < a.nst; ++st) {
        const int slot = st % X6_STAGES;
        mbar_wait(a.bar_full + 8 * slot, (unsigned)(st / X6_STAGES) & 1u);
        const double* sl = a.stage + (size_t)slot * X6_KC * a.sl;
        const double* sd = a.sD + (size_t)slot * X6_KC * X6_LDD;
        const int nval = min(X6_KC, a.n_cent - st * X6_KC);
        // linear columns: plain sums over the centres (own row +, neighbours' rows arrive negated)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int idx = ctid + j * X6_MMA_WARPS * 32;
            if (idx < nlin) {
                double s_ = lin[j];
                for (int c = 0; c < nval; ++c) s_ += sl[(size_t)c * a.sl + idx];
                lin[j] = s_;
            }
        }
        if (m.n_pair_terms > 0) {
            const bool okq = q < nval;
            const double* lq = sl + (size_t)q * a.sl;
            const double* dq = sd + q * X6_LDD + g;
            auto ldL = [&](int t, int r) -> double {
                return (okq && fpv[t] >= 0) ? lq[r * a.fl + fpv[t]] : 0.0;
            };
            auto ldD = [&](int t) -> double { return okq ? dq[t * 8] : 0.0; };
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int ta = x6_ta(W, i), tb = x6_tb(W, i);
                const double fDa = ldD(ta), fDb = ldD(tb);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double fLa = ldL(ta, r), fLb = ldL(tb, r);
                    dmma(acc[i][r][0], acc[i][r][1], fDa, fLb);
                    dmma(acc[i][r][0], acc[i][r][1], fLa, fDb);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(a.bar_empty + 8 * slot);
    }
}

template <int W>
__device__ __forceinline__ void x6_store_pairs(const DevModel& m, double* __restrict__ X, const int (&rows)[3],
                                               const double (&wrow)[3], int lane, const double (&acc)[5][3][2]) {
    constexpr int NT = x6_ntiles(W);
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int a = x6_ta(W, i) * 8 + g, bq = x6_tb(W, i) * 8 + 2 * q;
        const int col0 = m.pair_colof[a * 64 + bq], col1 = m.pair_colof[a * 64 + bq + 1];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double* xr = X + (size_t)rows[r] * m.fpad;
            if (col0 >= 0) xr[col0] = wrow[r] * acc[i][r][0];
            if (col1 >= 0) xr[col1] = wrow[r] * acc[i][r][1];
        }
    }
}

__global__ void __maxnreg__(112)
k_xrows_v6(DevModel m, DevBatch b, const double* __restrict__ dpv, const double* __restrict__ Lt,
           double* __restrict__ X, int apply_w, int sl) {
    extern __shared__ __align__(128) double smem[];
    double* stage = smem;                                             // [X6_STAGES][X6_KC][sl]
    double* sD = stage + (size_t)X6_STAGES * X6_KC * sl;             // [X6_STAGES][X6_KC][X6_LDD]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sD + X6_STAGES * X6_KC * X6_LDD);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int k_atom = blockIdx.x;
    const int s = b.st_of_atom[k_atom];
    if (!b.force[s]) return;
    const int p0 = b.seg_off[k_atom];
    const int n_cent = 1 + b.seg_off[k_atom + 1] - p0;
    const int nst = (n_cent + X6_KC - 1) / X6_KC;
    if (tid == 0) {
        for (int k = 0; k < X6_STAGES; ++k) {
            mbar_init(smem_u32(bars + k), X6_KC);
            mbar_init(smem_u32(bars + X6_STAGES + k), X6_MMA_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned bar_full = smem_u32(bars), bar_empty = smem_u32(bars + X6_STAGES);

    if (warp == X6_MMA_WARPS) {   // ---- producer: lane c brings in centre c of each stage -------------------------
        if (lane < X6_KC) {
            const unsigned bytes_l = (unsigned)(3 * m.fl * sizeof(double)), bytes_d = 64 * sizeof(double);
            const double* src0 = Lt + (size_t)(p0 + k_atom) * 3 * m.fl;   // the atom's block: own row, then neighbours
            for (int st = 0; st < nst; ++st) {
                const int slot = st % X6_STAGES;
                if (st >= X6_STAGES) mbar_wait(bar_empty + 8 * slot, (unsigned)(st / X6_STAGES - 1) & 1u);
                const int c = st * X6_KC + lane;
                const unsigned bar = bar_full + 8 * slot;
                if (c < n_cent) {
                    const int atom = c == 0 ? k_atom : b.nbr[p0 + c - 1];
                    mbar_expect_tx(bar, bytes_l + bytes_d);
                    bulk_g2s(smem_u32(stage + ((size_t)slot * X6_KC + lane) * sl), src0 + (size_t)c * 3 * m.fl, bytes_l, bar);
                    bulk_g2s(smem_u32(sD + ((size_t)slot * X6_KC + lane) * X6_LDD), dpv + (size_t)atom * 64, bytes_d, bar);
                } else {
                    mbar_arrive(bar);
                }
            }
        }
        return;
    }

    // ---- MMA warps ----------------------------------------------------------------------------------------------
    int rows[3];
    double wrow[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rows[r] = b.frow[s] + 3 * (k_atom - b.atom_off[s]) + r;
        wrow[r] = apply_w ? b.w[rows[r]] : 1.0;
    }
    double lin[3] = {0.0, 0.0, 0.0};
    double acc[5][3][2];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int r = 0; r < 3; ++r) { acc[i][r][0] = 0.0; acc[i][r][1] = 0.0; }
    X6Args a;
    a.stage = stage; a.sD = sD; a.bar_full = bar_full; a.bar_empty = bar_empty;
    a.sl = sl; a.fl = m.fl; a.n_cent = n_cent; a.nst = nst;
    switch (warp) {
        case 0: x6_consume<0>(m, a, lane, tid, lin, acc); break;
        case 1: x6_consume<1>(m, a, lane, tid, lin, acc); break;
        case 2: x6_consume<2>(m, a, lane, tid, lin, acc); break;
        case 3: x6_consume<3>(m, a, lane, tid, lin, acc); break;
        case 4: x6_consume<4>(m, a, lane, tid, lin, acc); break;
        case 5: x6_consume<5>(m, a, lane, tid, lin, acc); break;
        case 6: x6_consume<6>(m, a, lane, tid, lin, acc); break;
        default: x6_consume<7>(m, a, lane, tid, lin, acc); break;
    }
    // ---- epilogue: linear columns, y column and padding, order-2 columns -------------------------------------------
    {
        const int* pad_gid = m.types[0].pad_gid;
        const int nlin = 3 * m.fl;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int idx = tid + j * X6_MMA_WARPS * 32;
            if (idx >= nlin) continue;
            const int r = idx / m.fl, fp = idx - r * m.fl;
            const int gcol = pad_gid[fp];
            if (gcol < 0) continue;
            const int row = r == 0 ? rows[0] : (r == 1 ? rows[1] : rows[2]);
            const double wv = r == 0 ? wrow[0] : (r == 1 ? wrow[1] : wrow[2]);
            X[(size_t)row * m.fpad + gcol] = wv * lin[j];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        if (tid == r) X[(size_t)rows[r] * m.fpad + m.n_variables] = apply_w ? b.yv[rows[r]] : 0.0;
        for (int cz = m.n_variables + 1 + tid; cz < m.fpad; cz += X6_MMA_WARPS * 32) X[(size_t)rows[r] * m.fpad + cz] = 0.0;
    }
    if (m.n_pair_terms == 0) return;
    switch (warp) {
        case 0: x6_store_pairs<0>(m, X, rows, wrow, lane, acc); break;
        case 1: x6_store_pairs<1>(m, X, rows, wrow, lane, acc); break;
        case 2: x6_store_pairs<2>(m, X, rows, wrow, lane, acc); break;
        case 3: x6_store_pairs<3>(m, X, rows, wrow, lane, acc); break;
        case 4: x6_store_pairs<4>(m, X, rows, wrow, lane, acc); break;
        case 5: x6_store_pairs<5>(m, X, rows, wrow, lane, acc); break;
        case 6: x6_store_pairs<6>(m, X, rows, wrow, lane, acc); break;
        default: x6_store_pairs<7>(m, X, rows, wrow, lane, acc); break;
    }
}

// staged row stride of one centre (3 rows of fl doubles + padding so that the stride is == 4 mod 16)
static int xrows_v6_sl(const DevModel& m) { return 3 * m.fl + ((3 * m.fl) % 16 == 0 ? 4 : 12); }
static size_t xrows_v6_smem(const DevModel& m) {
    return ((size_t)X6_STAGES * X6_KC * (xrows_v6_sl(m) + X6_LDD)) * sizeof(double) + 2 * X6_STAGES * sizeof(unsigned long long) + 128;
}

void launch_xrows_v6(const DevModel& m, const DevBatch& b, const Workspace& ws, bool apply_weights, cudaStream_t s) {
    static size_t set_for = 0;
    const size_t smem = xrows_v6_smem(m);
    if (set_for < smem) {
        cudaFuncSetAttribute(k_xrows_v6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        set_for = smem;
    }
    k_xrows_v6<<<b.n_atoms, X6_THREADS, smem, s>>>(m, b, ws.dpv, ws.Lbuf, ws.X, apply_weights ? 1 : 0, xrows_v6_sl(m));
}

// Models served by the second-generation front end: one atom type, dense small radial groups whose head order is the
// Y_lm key order (DevModel::front2, checked on the host tables), <= 64 polynomial variables, 3 * fl <= 768 (three
// linear-column sums per MMA thread), and shared-memory footprints that fit.
bool front_v2_supported(const DevModel& m) {
    const bool off = getenv("PM_FRONT_V1") != nullptr;   // read per call: tests flip it to compare the two generations
    if (off || !m.front2 || m.n_type != 1 || m.kpn == 0 || m.tpn == 0) return false;
    if (m.npv_pad > 64 || m.npv <= 0 || m.pair_colof == nullptr || m.n_pair_terms <= 0) return false;
    if ((m.fl & 7) != 0 || 3 * m.fl > 3 * X6_MMA_WARPS * 32) return false;
    if (lrows_v4_smem(m, m.kpn) > 200 * 1024 || xrows_v6_smem(m) > 110 * 1024) return false;
    return true;
}

}  // namespace pm

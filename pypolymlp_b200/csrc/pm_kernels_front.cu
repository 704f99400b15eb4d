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
//  * K4b: a 4-stage TMA ring (4 centres = one DMMA k-step per stage) refilled by whichever warp releases a slot last;
//    the eight warps take their A / B fragments straight from the staged rows (no extraction pass, no Lambda tile) and
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
// Two shapes: RT = 2 row tiles (16 rows) per job with 11 consumer warps at 160 registers (each B fragment read from shared
// memory feeds two DMMAs), or RT = 1 (8 rows) with 15 consumer warps at 128 registers (more warps to hide the operand
// latencies, one DMMA per B-fragment read).  NW consumer warps + 1 producer warp per CTA, one CTA per SM.  The register
// file hands out warps in groups of four: 12 x 160 and 16 x 128 registers x 32 lanes both fill it exactly.

// DMMAs of one (16-row job, radial group): A fragments formed on the fly from the lm factors and the two radial scalars
// of each row; B fragments = coalesced 256-byte reads of the centre's G blocks in shared memory.  FULL: the group has
// exactly TPN x KPN dense blocks (constant offsets); AGG: some rows of the job are aggregated rows (own x/y/z, six
// virial rows) whose A operand is the K2b sum of the head itself.
template <int TPN, int KPN, int RT, bool FULL, bool AGG>
__device__ __forceinline__ void l4_group(double (&acc)[TPN][RT][2], const double (&a1)[RT][KPN], const double (&a2)[RT][KPN],
                                         const double (&cd)[RT], const double (&cf)[RT], const double* __restrict__ Gn,
                                         const int* __restrict__ bm, int ntile, int kcn,
                                         const double2* __restrict__ aggi, const int* __restrict__ hd,
                                         const int (&ragg)[RT], int q) {
    // aggregated rows: all KPN operand loads of the radial group are issued together (one global latency per group
    // instead of one per k-step; the a1 / a2 factors of those rows are unused zeros)
    double ag[RT][KPN];
    if (AGG) {
        const double* aggd = reinterpret_cast<const double*>(aggi);
#pragma unroll
        for (int kc = 0; kc < KPN; ++kc) {
            const int k = 4 * kc + q;
            const int h = hd[k >> 1];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                ag[rt][kc] = 0.0;
                if (ragg[rt] >= 0 && h >= 0) ag[rt][kc] = aggd[((size_t)h * 9 + ragg[rt]) * 2 + (k & 1)];
            }
        }
    }
#pragma unroll
    for (int kc = 0; kc < KPN; ++kc) {
        double af[RT];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) af[rt] = cd[rt] * a1[rt][kc] + cf[rt] * a2[rt][kc];
        if (AGG) {
#pragma unroll
            for (int rt = 0; rt < RT; ++rt)
                if (ragg[rt] >= 0) af[rt] = ag[rt][kc];
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
            for (int rt = 0; rt < RT; ++rt) dmma(acc[tt][rt][0], acc[tt][rt][1], af[rt], bf);
        }
    }
}

template <int TPN, int KPN, int RT, int L4_W>
__device__ __forceinline__ void l4_body(const DevModel& m, const DevBatch& b, const double* __restrict__ PB,
                                        const double2* __restrict__ agg, const double* __restrict__ Gbuf,
                                        double* __restrict__ Lt, double* __restrict__ Sbuf, double* smem) {
    constexpr int L4_THREADS = (L4_W + 1) * 32;
    constexpr int L4_JOB_ROWS = 8 * RT;
    const DevType& T = m.types[0];
    const int gsz = (int)T.g_size;
    const int asz = T.n_head * 18;                                              // doubles of one atom's K2b sums
    double* Gs = smem;                                                          // [2][gsz] B fragments of two centres
    double* As = Gs + 2 * (size_t)gsz;                                          // [2][asz] aggregated-row operands
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(As + 2 * (size_t)asz);   // full[2], empty[2]
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
    const unsigned gbytes = (unsigned)(gsz * sizeof(double)), abytes = (unsigned)(asz * sizeof(double));

    if (warp == L4_W) {   // ---- producer: one cp.async.bulk per centre into the ring -----------------------------
        if (lane == 0) {
            int c_idx = 0;
            for (int i = blockIdx.x; i < b.n_atoms; i += gridDim.x) {
                if (!b.force[b.st_of_atom[i]]) continue;
                const int buf = c_idx & 1;
                if (c_idx >= 2) mbar_wait(bar_empty + 8 * buf, (unsigned)((c_idx >> 1) - 1) & 1u);
                mbar_expect_tx(bar_full + 8 * buf, gbytes + abytes);
                bulk_g2s(smem_u32(Gs + (size_t)buf * gsz), Gbuf + (size_t)i * m.gstride, gbytes, bar_full + 8 * buf);
                bulk_g2s(smem_u32(As + (size_t)buf * asz), agg + (size_t)i * m.hmax * 9, abytes, bar_full + 8 * buf);
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
            double a1[RT][KPN], a2[RT][KPN];
            const double* recp[RT];
            double* rowp[RT];
            int ragg[RT];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
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
            double cdn[RT], cfn[RT];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) { cdn[rt] = 0.0; cfn[rt] = 0.0; }
            {
                const int nid = s_nid[0];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt)
                    if (recp[rt] && nid >= 0) {
                        cdn[rt] = recp[rt][(size_t)(4 + m.n_fn + nid) * PB_BLK];
                        cfn[rt] = recp[rt][(size_t)(4 + nid) * PB_BLK];
                    }
            }
            for (int n = 0; n < m.n_fn; ++n) {
                double cd[RT], cf[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) { cd[rt] = cdn[rt]; cf[rt] = cfn[rt]; }
                if (n + 1 < m.n_fn) {   // radial scalars of the next group: in flight during this group's DMMAs
                    const int nid = s_nid[n + 1];
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
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
                double acc[TPN][RT][2];
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt)
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) { acc[tt][rt][0] = 0.0; acc[tt][rt][1] = 0.0; }
                if (kcn > 0) {
                    const bool full = m.dense && kcn == KPN && ntile == TPN;
                    const double* Gn = G + 32 * (size_t)(full ? s_base[n] : 0) + lane;
                    const int* bm = s_bmap + tile0 * KPN;
                    const double2* aggi = reinterpret_cast<const double2*>(As + (size_t)buf * asz);
                    const int* hd = s_head + n * 2 * KPN;
                    if (full) {
                        if (!has_agg) l4_group<TPN, KPN, RT, true, false>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                        else l4_group<TPN, KPN, RT, true, true>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                    } else {
                        if (!has_agg) l4_group<TPN, KPN, RT, false, false>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                        else l4_group<TPN, KPN, RT, false, true>(acc, a1, a2, cd, cf, Gn, bm, ntile, kcn, aggi, hd, ragg, q);
                    }
                }
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt) {
                    if (tt >= ntile) break;
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) {
                        if (!rowp[rt]) continue;
                        const double sg = ragg[rt] < 0 ? -1.0 : 1.0;
                        *reinterpret_cast<double2*>(rowp[rt] + (tile0 + tt) * 8) =
                            make_double2(sg * acc[tt][rt][0], sg * acc[tt][rt][1]);
                    }
                }
            }
        }
        jobmod = (jobmod + njobs + 1) % L4_W;   // + 1: the last job (aggregated rows, extra latency) rotates over the warps
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
        ++c_idx;
    }
}

// ------------------------------------------------------------------------------------------------
// Alpha-major job shape: a job is 8 pairs x 3 Cartesian components = three DMMA row tiles whose row g is the SAME pair
// (tile rt = component alpha).  The lm factor Y_lm of a pair is then loaded once and shared by its three rows
// (A = f_n' (D_alpha / r) Y + f_n dY/dalpha), the two radial scalars are per pair instead of per row, and every B
// fragment read from shared memory feeds three DMMAs.  The 9 aggregated rows (own x/y/z, six virial rows) are one
// extra job of two row tiles.
// ------------------------------------------------------------------------------------------------
template <int TPN, int KPN, bool FULL>
__device__ __forceinline__ void l4a_group(double (&acc)[TPN][3][2], const double (&yk)[KPN], const double (&a2)[3][KPN],
                                          const double (&c)[3], double cf, const double* __restrict__ Gn,
                                          const int* __restrict__ bm, int ntile, int kcn) {
#pragma unroll
    for (int kc = 0; kc < KPN; ++kc) {
        double af[3];
#pragma unroll
        for (int rt = 0; rt < 3; ++rt) af[rt] = c[rt] * yk[kc] + cf * a2[rt][kc];
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
            for (int rt = 0; rt < 3; ++rt) dmma(acc[tt][rt][0], acc[tt][rt][1], af[rt], bf);
        }
    }
}

template <int TPN, int KPN, int L4_W>
__device__ __forceinline__ void l4a_body(const DevModel& m, const DevBatch& b, const double* __restrict__ PB,
                                         const double2* __restrict__ agg, const double* __restrict__ Gbuf,
                                         double* __restrict__ Lt, double* __restrict__ Sbuf, double* smem) {
    constexpr int L4_THREADS = (L4_W + 1) * 32;
    const DevType& T = m.types[0];
    const int gsz = (int)T.g_size;
    const int asz = T.n_head * 18;                                              // doubles of one atom's K2b sums
    double* Gs = smem;                                                          // [2][gsz] B fragments of two centres
    double* As = Gs + 2 * (size_t)gsz;                                          // [2][asz] aggregated-row operands
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(As + 2 * (size_t)asz);   // full[2], empty[2]
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
    const unsigned gbytes = (unsigned)(gsz * sizeof(double)), abytes = (unsigned)(asz * sizeof(double));

    if (warp == L4_W) {   // ---- producer: one cp.async.bulk per centre into the ring -----------------------------
        if (lane == 0) {
            int c_idx = 0;
            for (int i = blockIdx.x; i < b.n_atoms; i += gridDim.x) {
                if (!b.force[b.st_of_atom[i]]) continue;
                const int buf = c_idx & 1;
                if (c_idx >= 2) mbar_wait(bar_empty + 8 * buf, (unsigned)((c_idx >> 1) - 1) & 1u);
                mbar_expect_tx(bar_full + 8 * buf, gbytes + abytes);
                bulk_g2s(smem_u32(Gs + (size_t)buf * gsz), Gbuf + (size_t)i * m.gstride, gbytes, bar_full + 8 * buf);
                bulk_g2s(smem_u32(As + (size_t)buf * asz), agg + (size_t)i * m.hmax * 9, abytes, bar_full + 8 * buf);
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
        const int npj = (np + 7) >> 3;     // pair jobs; job npj holds the aggregated rows
        const int njobs = npj + 1;
        mbar_wait(bar_full + 8 * buf, (unsigned)(c_idx >> 1) & 1u);
        const double* G = Gs + (size_t)buf * gsz;

        int tj = warp - jobmod;
        if (tj < 0) tj += L4_W;
        for (; tj < njobs; tj += L4_W) {
            const bool is_agg = tj == npj;   // warp uniform
            double yk[KPN], a2[3][KPN], dal[3];
            const double* rec = nullptr;
            double* rowp[3] = {nullptr, nullptr, nullptr};
            if (!is_agg) {
                const int pl = tj * 8 + g;
                if (pl < np) {
                    const int p = p0 + pl;
                    rec = PB + (size_t)(p >> 5) * pbblk + (p & 31);
                    const double rinv = rec[3 * PB_BLK];
                    const double* ry = rec + (size_t)oy * PB_BLK;
#pragma unroll
                    for (int kc = 0; kc < KPN; ++kc) {
                        const int k = 4 * kc + q;
                        yk[kc] = k < k_real ? ry[(size_t)k * PB_BLK] : 0.0;
                    }
#pragma unroll
                    for (int al = 0; al < 3; ++al) {
                        dal[al] = rec[al * PB_BLK] * rinv;
                        const double* rya = rec + (size_t)pb_y(m, 1 + al) * PB_BLK;
#pragma unroll
                        for (int kc = 0; kc < KPN; ++kc) {
                            const int k = 4 * kc + q;
                            a2[al][kc] = k < k_real ? rya[(size_t)k * PB_BLK] : 0.0;
                        }
                    }
                    // force on atom j is -d/dr_j: the rows of pair (i -> j) land, negated, in the slot of the reverse
                    // pair in j's block of Lt (slot 0 of a block is the atom's own row)
                    const int j = b.nbr[p];
                    double* r0 = Lt + (size_t)(b.rev[p] + j + 1) * 3 * fl + 2 * q;
                    rowp[0] = r0; rowp[1] = r0 + fl; rowp[2] = r0 + 2 * fl;
                }
            } else {
                // aggregated rows 0..8: tile 0 row g, tile 1 row 0 -> row 8
                rowp[0] = (g < 3 ? Lt + ((size_t)(p0 + i) * 3 + g) * fl : Sbuf + ((size_t)i * 6 + (g - 3)) * fl) + 2 * q;
                if (g == 0) rowp[1] = Sbuf + ((size_t)i * 6 + 5) * fl + 2 * q;
            }
            if (!rec) {
#pragma unroll
                for (int kc = 0; kc < KPN; ++kc) { yk[kc] = 0.0; a2[0][kc] = 0.0; a2[1][kc] = 0.0; a2[2][kc] = 0.0; }
                dal[0] = dal[1] = dal[2] = 0.0;
            }
            // radial scalars of the pair, prefetched two radial groups ahead
            double cdq[2] = {0.0, 0.0}, cfq[2] = {0.0, 0.0};
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int nid = d < m.n_fn ? s_nid[d] : -1;
                if (rec && nid >= 0) {
                    cdq[d] = rec[(size_t)(4 + m.n_fn + nid) * PB_BLK];
                    cfq[d] = rec[(size_t)(4 + nid) * PB_BLK];
                }
            }
            const double* aggd = As + (size_t)buf * asz;
            for (int n = 0; n < m.n_fn; ++n) {
                const double cd = cdq[0], cf = cfq[0];
                cdq[0] = cdq[1]; cfq[0] = cfq[1];
                cdq[1] = 0.0; cfq[1] = 0.0;
                if (n + 2 < m.n_fn) {
                    const int nid = s_nid[n + 2];
                    if (rec && nid >= 0) {
                        cdq[1] = rec[(size_t)(4 + m.n_fn + nid) * PB_BLK];
                        cfq[1] = rec[(size_t)(4 + nid) * PB_BLK];
                    }
                }
                const int tile0 = s_tile0[n];
                const int ntile = s_tile0[n + 1] - tile0;
                const int kcn = s_kcn[n];
                double acc[TPN][3][2];
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt)
#pragma unroll
                    for (int rt = 0; rt < 3; ++rt) { acc[tt][rt][0] = 0.0; acc[tt][rt][1] = 0.0; }
                if (kcn > 0) {
                    double c[3] = {cd * dal[0], cd * dal[1], cd * dal[2]};
                    double cfv = cf;
                    if (is_agg) {
                        // the A operand of an aggregated row is the K2b sum of the head itself: parked in the (otherwise
                        // unused) a2 registers of tiles 0 / 1 and passed through with c = 0, cf = 1
                        const int* hd = s_head + n * 2 * KPN;
#pragma unroll
                        for (int kc = 0; kc < KPN; ++kc) {
                            const int k = 4 * kc + q;
                            const int h = hd[k >> 1];
                            a2[0][kc] = h >= 0 ? aggd[((size_t)h * 9 + g) * 2 + (k & 1)] : 0.0;
                            a2[1][kc] = (h >= 0 && g == 0) ? aggd[((size_t)h * 9 + 8) * 2 + (k & 1)] : 0.0;
                        }
                        c[0] = c[1] = c[2] = 0.0;
                        cfv = 1.0;
                    }
                    const bool full = m.dense && kcn == KPN && ntile == TPN;
                    const double* Gn = G + 32 * (size_t)(full ? s_base[n] : 0) + lane;
                    const int* bm = s_bmap + tile0 * KPN;
                    if (full) l4a_group<TPN, KPN, true>(acc, yk, a2, c, cfv, Gn, bm, ntile, kcn);
                    else l4a_group<TPN, KPN, false>(acc, yk, a2, c, cfv, Gn, bm, ntile, kcn);
                }
                const double sg = is_agg ? 1.0 : -1.0;
#pragma unroll
                for (int tt = 0; tt < TPN; ++tt) {
                    if (tt >= ntile) break;
#pragma unroll
                    for (int rt = 0; rt < 3; ++rt) {
                        if (!rowp[rt]) continue;
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

template <int TPN, int KPN>
__global__ void __maxnreg__(160)
k_lrows_v4a(DevModel m, DevBatch b, const double* __restrict__ PB, const double2* __restrict__ agg,
            const double* __restrict__ Gbuf, double* __restrict__ Lt, double* __restrict__ Sbuf) {
    extern __shared__ __align__(128) double smem[];
    l4a_body<TPN, KPN, 11>(m, b, PB, agg, Gbuf, Lt, Sbuf, smem);
}

template <int TPN, int KPN>
__global__ void __maxnreg__(160)
k_lrows_v4(DevModel m, DevBatch b, const double* __restrict__ PB, const double2* __restrict__ agg,
           const double* __restrict__ Gbuf, double* __restrict__ Lt, double* __restrict__ Sbuf) {
    extern __shared__ __align__(128) double smem[];
    l4_body<TPN, KPN, 2, 11>(m, b, PB, agg, Gbuf, Lt, Sbuf, smem);
}

template <int TPN, int KPN>
__global__ void __maxnreg__(128)
k_lrows_v4n(DevModel m, DevBatch b, const double* __restrict__ PB, const double2* __restrict__ agg,
            const double* __restrict__ Gbuf, double* __restrict__ Lt, double* __restrict__ Sbuf) {
    extern __shared__ __align__(128) double smem[];
    l4_body<TPN, KPN, 1, 15>(m, b, PB, agg, Gbuf, Lt, Sbuf, smem);
}

static size_t lrows_v4_smem(const DevModel& m, int kpn) {
    const DevType& T = m.types[0];
    const size_t ints = 4ull * m.n_fn + 2 + (size_t)m.n_fn * 2 * kpn + (size_t)T.n_tiles * kpn;
    return 2ull * ((size_t)T.g_size + (size_t)T.n_head * 18) * sizeof(double) + 4 * sizeof(unsigned long long) +
           ints * sizeof(int) + 128;
}

template <int TPN, int KPN>
static void launch_lrows_v4_t(const DevModel& m, const DevBatch& b, const Workspace& ws, cudaStream_t s) {
    static int n_sm = 0;
    const size_t smem = lrows_v4_smem(m, KPN);
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    ensure_smem((const void*)k_lrows_v4<TPN, KPN>, smem);
    ensure_smem((const void*)k_lrows_v4n<TPN, KPN>, smem);
    ensure_smem((const void*)k_lrows_v4a<TPN, KPN>, smem);
    const int grid = std::min(n_sm, b.n_atoms);
    // measured on config 2 (us / structure): 8-row jobs x 15 warps 31.5, 16-row jobs x 11 warps 36.9, alpha-major 24-row jobs
    // x 11 warps 36.7 -- the warp count (latency hiding) beats the B-fragment reuse
    const char* e = getenv("PM_L4_RT");
    const int shape = e ? atoi(e) : 1;
    if (shape == 3)
        k_lrows_v4a<TPN, KPN><<<grid, 12 * 32, smem, s>>>(m, b, ws.PB, ws.agg, ws.Gbuf, ws.Lbuf, ws.Sbuf);
    else if (shape == 2)
        k_lrows_v4<TPN, KPN><<<grid, 12 * 32, smem, s>>>(m, b, ws.PB, ws.agg, ws.Gbuf, ws.Lbuf, ws.Sbuf);
    else
        k_lrows_v4n<TPN, KPN><<<grid, 16 * 32, smem, s>>>(m, b, ws.PB, ws.agg, ws.Gbuf, ws.Lbuf, ws.Sbuf);
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
constexpr int X6_THREADS = X6_MMA_WARPS * 32;   // 8 warps x 112 registers: two CTAs per SM (warps are allocated in fours)
constexpr int X6_STAGES = 4;
constexpr int X6_KC = 4;           // centres per stage = one DMMA k-step
constexpr int X6_LDD = 68;         // row stride of the staged polynomial-variable rows (== 4 mod 16)
constexpr int X6_MAXC = 160;       // centre atoms kept in shared memory (beyond that: read from global)

// upper-triangle tiles (ta <= tb) of the 8 x 8 grid of 8 x 8 blocks over the <= 64 polynomial variables, dealt to the
// eight MMA warps as one row strip of <= 3 tiles plus one of <= 2 tiles each (the A-side fragments of a strip are loaded
// once).  {row, first column, tiles} x 2; warps w and w + 4 share an SM sub-partition and hold 9 tiles together.
// One code path for all warps: per-warp template instances of this loop (8 x 14 KB of SASS) thrashed the instruction
// cache (ncu: 64 % of the stall samples were no_instruction).
__constant__ signed char c_x6_strips[X6_MMA_WARPS][6] = {
    {0, 0, 3, 0, 6, 2}, {0, 3, 3, 1, 4, 2}, {1, 1, 3, 1, 6, 2}, {2, 2, 3, 3, 6, 2},
    {2, 5, 3, 4, 7, 1}, {3, 3, 3, 7, 7, 1}, {4, 4, 3, 5, 7, 1}, {5, 5, 2, 6, 6, 2}};

__global__ void __maxnreg__(112)
k_xrows_v6(DevModel m, DevBatch b, const double* __restrict__ dpv, const double* __restrict__ Lt,
           double* __restrict__ X, int apply_w, int sl) {
    extern __shared__ __align__(128) double smem[];
    double* stage = smem;                                             // [X6_STAGES][X6_KC][sl]
    double* sD = stage + (size_t)X6_STAGES * X6_KC * sl;             // [X6_STAGES][X6_KC][X6_LDD]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sD + X6_STAGES * X6_KC * X6_LDD);   // full[X6_STAGES]
    int* s_cnt = reinterpret_cast<int*>(bars + X6_STAGES);           // [X6_STAGES] warps done with the slot
    int* s_atom = s_cnt + X6_STAGES;                                  // [X6_MAXC] atom of each centre
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int k_atom = blockIdx.x;
    const int s = b.st_of_atom[k_atom];
    if (!b.force[s]) return;
    const int p0 = b.seg_off[k_atom];
    const int n_cent = 1 + b.seg_off[k_atom + 1] - p0;
    const int nst = (n_cent + X6_KC - 1) / X6_KC;
    for (int c = tid; c < min(n_cent, X6_MAXC); c += X6_THREADS) s_atom[c] = c == 0 ? k_atom : b.nbr[p0 + c - 1];
    if (tid == 0) {
        for (int k = 0; k < X6_STAGES; ++k) {
            mbar_init(smem_u32(bars + k), X6_KC);
            s_cnt[k] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned bar_full = smem_u32(bars);
    const unsigned bytes_l = (unsigned)(3 * m.fl * sizeof(double)), bytes_d = 64 * sizeof(double);
    const double* src0 = Lt + (size_t)(p0 + k_atom) * 3 * m.fl;   // the atom's block: own row, then its neighbours
    // lane c (< X6_KC) of the calling warp brings in centre c of stage st: 3 derivative rows + the centre's
    // polynomial-variable row, one cp.async.bulk each, completion counted on the slot's mbarrier
    auto issue_stage = [&](int st) {
        if (lane < X6_KC) {
            const int slot = st % X6_STAGES;
            const int c = st * X6_KC + lane;
            const unsigned bar = bar_full + 8 * slot;
            if (c < n_cent) {
                const int atom = c < X6_MAXC ? s_atom[c] : b.nbr[p0 + c - 1];
                mbar_expect_tx(bar, bytes_l + bytes_d);
                bulk_g2s(smem_u32(stage + ((size_t)slot * X6_KC + lane) * sl), src0 + (size_t)c * 3 * m.fl, bytes_l, bar);
                bulk_g2s(smem_u32(sD + ((size_t)slot * X6_KC + lane) * X6_LDD), dpv + (size_t)atom * 64, bytes_d, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    if (warp == 0)
        for (int st = 0; st < min(nst, X6_STAGES); ++st) issue_stage(st);

    // ---- MMA warps ----------------------------------------------------------------------------------------------
    const int g = lane >> 2, q = lane & 3;
    const int rA = c_x6_strips[warp][0], cA = c_x6_strips[warp][1], nA = c_x6_strips[warp][2];
    const int rB = c_x6_strips[warp][3], cB = c_x6_strips[warp][4], nB = c_x6_strips[warp][5];
    // Fragment addresses = (warp-uniform slot base) + (per-lane byte offset fixed for the whole kernel).  A polynomial
    // variable slot beyond npv reads column 0 instead of a zero: it only feeds accumulator entries that are never stored
    // (pair_colof = -1 there) and every entry depends on its own (a, b) pair alone.
    auto fp_of = [&](int t) {
        const int v = (t * 8 + g < m.npv_pad) ? m.pv_fp[t * 8 + g] : -1;
        return v >= 0 ? v : 0;
    };
    const int fl = m.fl;
    const unsigned lane_l = (unsigned)(q * sl) * 8u, lane_d = (unsigned)(q * X6_LDD + g) * 8u;
    unsigned oLa[2], oDa[2], oLb[5], oDb[5];
    oLa[0] = lane_l + 8u * fp_of(rA); oLa[1] = lane_l + 8u * fp_of(rB);
    oDa[0] = lane_d + 64u * rA; oDa[1] = lane_d + 64u * rB;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int t = i < 3 ? cA + min(i, nA - 1) : cB + min(i - 3, nB - 1);
        oLb[i] = lane_l + 8u * fp_of(t);
        oDb[i] = lane_d + 64u * t;
    }
    const unsigned flb = (unsigned)fl * 8u, slb = (unsigned)sl * 8u;
    const unsigned stage_u = smem_u32(stage), sd_u = smem_u32(sD);
    auto lds = [](unsigned addr) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
        return v;
    };
    double lin[3] = {0.0, 0.0, 0.0};
    double acc[5][3][2];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int r = 0; r < 3; ++r) { acc[i][r][0] = 0.0; acc[i][r][1] = 0.0; }
    const int nlin = 3 * fl;
    for (int st = 0; st < nst; ++st) {
        const int slot = st % X6_STAGES;
        mbar_wait(bar_full + 8 * slot, (unsigned)(st / X6_STAGES) & 1u);
        const unsigned sl_u = stage_u + (unsigned)slot * X6_KC * slb;
        const unsigned sdu = sd_u + (unsigned)slot * (X6_KC * X6_LDD * 8);
        const int nval = min(X6_KC, n_cent - st * X6_KC);
        if (nval == X6_KC) {
            // ---- full stage: no predication -------------------------------------------------------------------
            // linear columns: plain sums over the centres (own row +, neighbours' rows arrive negated)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int idx = tid + j * X6_THREADS;
                if (idx < nlin) {
                    const unsigned a0 = sl_u + 8u * idx;
                    const double v0 = lds(a0), v1 = lds(a0 + slb), v2 = lds(a0 + 2 * slb), v3 = lds(a0 + 3 * slb);
                    lin[j] += (v0 + v1) + (v2 + v3);
                }
            }
            if (m.n_pair_terms > 0) {
#pragma unroll
                for (int strip = 0; strip < 2; ++strip) {
                    const int nn = strip == 0 ? nA : nB;
                    const double fDa = lds(sdu + oDa[strip]);
                    double fLa[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) fLa[r] = lds(sl_u + oLa[strip] + r * flb);
#pragma unroll
                    for (int i = 0; i < (strip == 0 ? 3 : 2); ++i) {
                        if (i >= nn) break;
                        const int si = strip == 0 ? i : 3 + i;
                        const double fDb = lds(sdu + oDb[si]);
                        double fLb[3];
#pragma unroll
                        for (int r = 0; r < 3; ++r) fLb[r] = lds(sl_u + oLb[si] + r * flb);
#pragma unroll
                        for (int r = 0; r < 3; ++r) dmma(acc[si][r][0], acc[si][r][1], fDa, fLb[r]);
#pragma unroll
                        for (int r = 0; r < 3; ++r) dmma(acc[si][r][0], acc[si][r][1], fLa[r], fDb);
                    }
                }
            }
        } else {
            // ---- last, partly filled stage: the unused centre slots hold stale or uninitialised data -------------
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int idx = tid + j * X6_THREADS;
                if (idx < nlin) {
                    double s_ = lin[j];
                    for (int c = 0; c < nval; ++c) s_ += lds(sl_u + 8u * idx + c * slb);
                    lin[j] = s_;
                }
            }
            if (m.n_pair_terms > 0) {
                const bool okq = q < nval;
#pragma unroll
                for (int strip = 0; strip < 2; ++strip) {
                    const int nn = strip == 0 ? nA : nB;
                    const double fDa = okq ? lds(sdu + oDa[strip]) : 0.0;
                    double fLa[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) fLa[r] = okq ? lds(sl_u + oLa[strip] + r * flb) : 0.0;
#pragma unroll
                    for (int i = 0; i < (strip == 0 ? 3 : 2); ++i) {
                        if (i >= nn) break;
                        const int si = strip == 0 ? i : 3 + i;
                        const double fDb = okq ? lds(sdu + oDb[si]) : 0.0;
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const double fLb = okq ? lds(sl_u + oLb[si] + r * flb) : 0.0;
                            dmma(acc[si][r][0], acc[si][r][1], fDa, fLb);
                            dmma(acc[si][r][0], acc[si][r][1], fLa[r], fDb);
                        }
                    }
                }
            }
        }
        // the warp that finishes a slot last refills it (no producer warp: 8 warps x 112 registers let two CTAs share an SM)
        // (release: this warp's reads of the slot are ordered before its count; acquire + proxy fence: the refill, an
        // async-proxy write, is ordered after every warp's count)
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = atomicAdd(s_cnt + slot, 1) == X6_MMA_WARPS - 1;
            if (last) {
                s_cnt[slot] = 0;
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && st + X6_STAGES < nst) issue_stage(st + X6_STAGES);
    }
    // ---- epilogue: linear columns, y column and padding, order-2 columns -------------------------------------------
    int rows[3];
    double wrow[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rows[r] = b.frow[s] + 3 * (k_atom - b.atom_off[s]) + r;
        wrow[r] = apply_w ? b.w[rows[r]] : 1.0;
    }
    {
        const int* pad_gid = m.types[0].pad_gid;
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
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int ta = i < 3 ? rA : rB, tb = i < 3 ? cA + i : cB + (i - 3);
        if ((i < 3 ? i : i - 3) >= (i < 3 ? nA : nB)) continue;
        const int a = ta * 8 + g, bq = tb * 8 + 2 * q;
        const int col0 = m.pair_colof[a * 64 + bq], col1 = m.pair_colof[a * 64 + bq + 1];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double* xr = X + (size_t)rows[r] * m.fpad;
            if (col0 >= 0) xr[col0] = wrow[r] * acc[i][r][0];
            if (col1 >= 0) xr[col1] = wrow[r] * acc[i][r][1];
        }
    }
}

// staged row stride of one centre (3 rows of fl doubles + padding so that the stride is == 4 mod 16)
static int xrows_v6_sl(const DevModel& m) { return 3 * m.fl + ((3 * m.fl) % 16 == 0 ? 4 : 12); }
static size_t xrows_v6_smem(const DevModel& m) {
    return ((size_t)X6_STAGES * X6_KC * (xrows_v6_sl(m) + X6_LDD)) * sizeof(double) + X6_STAGES * sizeof(unsigned long long) +
           (X6_STAGES + X6_MAXC) * sizeof(int) + 128;
}

void launch_xrows_v6(const DevModel& m, const DevBatch& b, const Workspace& ws, bool apply_weights, cudaStream_t s) {
    const size_t smem = xrows_v6_smem(m);
    // two CTAs per SM need ~204 KB of the 256 KB L1 / shared array: ask for the largest carve-out
    ensure_smem((const void*)k_xrows_v6, smem, true);
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

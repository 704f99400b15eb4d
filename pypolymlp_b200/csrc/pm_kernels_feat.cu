// pm_kernels_feat.cu -- K3 for LARGE radial-replication models (DevType::r_nr > 0): k_features_v4r.
//
// K3 computes, per atom, the linear gtinv invariants d_f (sums over products of order parameters, SURVEY.md section 8;
// reference: compute/polymlp_features.cpp + model.cpp through Features::compute_features{,_deriv}) and, for force
// structures, the entries G[f][h] = d d_f / d a_h that K4a multiplies with the pair-basis derivative rows.
// k_features_v3 (pm_kernels.cu) walks sliced (ELL-like) term tables, one slot per lane.  For max_l ~ 12 those tables are
// ~15 MB per atom type and every CTA streams them from L2: the kernel is bound by that stream.  In the usual gtinv model
// every product shares one radial index and the lists of radial index n are those of radial index 0 with every
// order-parameter id shifted by n * S (verified entry by entry on the host, pm_capi.cu).  This kernel walks the slices of
// radial index 0 only -- tables 1 / n_radial of the size -- and applies each decoded slot to NB radial indices (NB
// independent gather / multiply / accumulate chains).  Work item = (slice, block of NB radial indices).
// (For small models -- config 2 -- the same kernel measured slower than v3, 11.3 - 14.3 vs 10.7 us per structure: too few
// radial-0 slices per CTA; it is used for the big a_nlm arrays only.)
#include "pm_kernels.cuh"

#include <cstdlib>

namespace pm {

namespace {

__device__ __forceinline__ double2 cmulf(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int AT, int NT, int NB>
__global__ void __launch_bounds__(NT, 2) k_features_v4r(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                      double* __restrict__ dfeat, double* __restrict__ Gbuf,
                                                      int nfull_max, int zero_g, double* __restrict__ dpv) {
    extern __shared__ double2 afull[];   // [AT][nfull_max]
    const int i0 = blockIdx.x * AT;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = NT / 32;
    int ty[AT];
    bool fo[AT];
#pragma unroll
    for (int a = 0; a < AT; ++a) {
        const int i = i0 + a;
        ty[a] = i < b.n_atoms ? b.types[i] : -1;
        fo[a] = i < b.n_atoms ? b.force[b.st_of_atom[i]] != 0 : false;
    }
#pragma unroll
    for (int a = 0; a < AT; ++a) {
        if (ty[a] < 0) continue;
        const int i = i0 + a;
        const DevType& T = m.types[ty[a]];
        for (int k = tid; k < T.n_full; k += NT) {
            double2 v = anc[(size_t)i * m.hmax + T.full_head[k]];
            if (T.full_conj[k]) {
                const double cc = T.full_cc[k];
                v = make_double2(cc * v.x, -cc * v.y);
            }
            afull[(size_t)a * nfull_max + k] = v;
        }
        double* drow = dfeat + (size_t)i * m.fl;
        for (int k = tid; k < m.fl; k += NT) drow[k] = 0.0;
        if (fo[a] && zero_g) {   // single-type models: the structural zeros of G are written once per allocation (host)
            double2* G = reinterpret_cast<double2*>(Gbuf + (size_t)i * m.gstride);
            for (long k = tid; k < T.g_size / 2; k += NT) G[k] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false, anyf = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) { any = any || ty[a] == tt; anyf = anyf || (ty[a] == tt && fo[a]); }
        if (!any) continue;
        const DevType& T = m.types[tt];
        const double* __restrict__ coef = T.r_sl_coeff;
        const unsigned* __restrict__ ids = T.r_sl_ids;
        const long ns = T.r_n_slots;
        const int nblk = T.r_nr / NB;
        const int S = T.r_S, Fs = T.r_Fs;
        const int4* __restrict__ fmeta = T.r_fsl_meta;
        const int4* __restrict__ emeta = T.r_esl_meta;
        const int* __restrict__ fout = T.r_fsl_out;
        const int4* __restrict__ eout = T.r_esl_out;
        const int* __restrict__ pad_pv = T.pad_pv;
        const int n_fw = T.r_n_fsl * nblk, n_ew = T.r_n_esl * nblk;
        // ---- linear invariants: work item = (feature slice, block of NB radial indices) -------------------------------
        for (int w = warp; w < n_fw; w += NWARP) {
            const int s = w / nblk, n0 = (w - s * nblk) * NB;
            const int4 meta = fmeta[s];
            const int o = meta.z;
            double sum[AT][NB];
#pragma unroll
            for (int a = 0; a < AT; ++a)
#pragma unroll
                for (int n = 0; n < NB; ++n) sum[a][n] = 0.0;
            const double2* af0 = afull + (size_t)n0 * S;
#pragma unroll 2
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                const unsigned v0 = ids[slot];
                const unsigned v1 = o > 2 ? ids[ns + slot] : 0u;
                const int id0 = v0 & 0xffffu, id1 = v0 >> 16, id2 = v1 & 0xffffu, id3 = v1 >> 16;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double2* af = af0 + (size_t)a * nfull_max;
                    if (o == 2) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = af[id0 + n * S], q = af[id1 + n * S];
                            sum[a][n] += cf * (p.x * q.x - p.y * q.y);
                        }
                    } else if (o == 3) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = cmulf(af[id0 + n * S], af[id1 + n * S]), q = af[id2 + n * S];
                            sum[a][n] += cf * (p.x * q.x - p.y * q.y);
                        }
                    } else if (o == 1) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) sum[a][n] += cf * af[id0 + n * S].x;
                    } else {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = cmulf(af[id0 + n * S], af[id1 + n * S]);
                            const double2 q = cmulf(af[id2 + n * S], af[id3 + n * S]);
                            sum[a][n] += cf * (p.x * q.x - p.y * q.y);
                        }
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < AT; ++a)
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    sum[a][n] += __shfl_xor_sync(0xffffffffu, sum[a][n], 8);
                    sum[a][n] += __shfl_xor_sync(0xffffffffu, sum[a][n], 16);
                }
            if (lane < 8) {
                const int fp0 = fout[s * 8 + lane];
                if (fp0 >= 0) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const int fp = fp0 + (n0 + n) * Fs;
                        const int pv = dpv ? pad_pv[fp] : -1;
#pragma unroll
                        for (int a = 0; a < AT; ++a)
                            if (ty[a] == tt) {
                                dfeat[(size_t)(i0 + a) * m.fl + fp] = sum[a][n];
                                if (pv >= 0) dpv[(size_t)(i0 + a) * 64 + pv] = sum[a][n];
                            }
                    }
                }
            }
        }
        if (!anyf) continue;
        // ---- G entries: work item = (entry slice, block of NB radial indices) ---------------------------------------
        for (int w = warp; w < n_ew; w += NWARP) {
            const int s = w / nblk, n0 = (w - s * nblk) * NB;
            const int4 meta = emeta[s];
            const int cn = meta.z;
            double gr[AT][NB], gi[AT][NB];
#pragma unroll
            for (int a = 0; a < AT; ++a)
#pragma unroll
                for (int n = 0; n < NB; ++n) { gr[a][n] = 0.0; gi[a][n] = 0.0; }
            const double2* af0 = afull + (size_t)n0 * S;
#pragma unroll 2
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                const unsigned v0 = ids[slot];
                const unsigned v1 = cn > 2 ? ids[ns + slot] : 0u;
                const int id0 = v0 & 0xffffu, id1 = (v0 >> 16) & 0x7fffu, id2 = v1 & 0xffffu;
                const double cfi = (v0 & 0x80000000u) ? -cf : cf;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt || !fo[a]) continue;
                    const double2* af = af0 + (size_t)a * nfull_max;
                    if (cn == 2) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = af[id0 + n * S], q = af[id1 + n * S];
                            gr[a][n] += cf * (p.x * q.x - p.y * q.y);
                            gi[a][n] += cfi * (p.x * q.y + p.y * q.x);
                        }
                    } else if (cn == 1) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = af[id0 + n * S];
                            gr[a][n] += cf * p.x;
                            gi[a][n] += cfi * p.y;
                        }
                    } else if (cn == 0) {
#pragma unroll
                        for (int n = 0; n < NB; ++n) gr[a][n] += cf;
                    } else {
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const double2 p = cmulf(cmulf(af[id0 + n * S], af[id1 + n * S]), af[id2 + n * S]);
                            gr[a][n] += cf * p.x;
                            gi[a][n] += cfi * p.y;
                        }
                    }
                }
            }
            const int4 pos = eout[s * 32 + lane];   // pos_re, pos_im, G stride per radial index
            if (pos.x >= 0) {
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt || !fo[a]) continue;
                    double* G = Gbuf + (size_t)(i0 + a) * m.gstride + (size_t)n0 * pos.z;
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        G[pos.x + n * pos.z] = gr[a][n];
                        G[pos.y + n * pos.z] = -gi[a][n];
                    }
                }
            }
        }
    }
}

}  // namespace

// true if the radial-batched kernel served the batch: big a_nlm arrays (the 512-thread regime of launch_features), all types
// radial replications with the same number of radial indices, product order <= 4; else the caller runs k_features_v3 & co.
bool launch_features_radial(const DevModel& m, const DevBatch& b, const double2* anc, double* dfeat, double* Gbuf,
                            size_t smem_bytes, cudaStream_t s, bool zero_g, double* dpv) {
    if (b.n_atoms == 0) return true;
    if (getenv("PM_FEAT_V3") != nullptr) return false;
    if (4 * smem_bytes <= 96 * 1024) return false;   // small models: k_features_v3 is faster (measured)
    const int nr = m.types[0].r_nr;
    for (int t = 0; t < m.n_type; ++t)
        if (m.types[t].r_nr <= 0 || m.types[t].r_nr != nr || m.types[t].max_order > 4) return false;
    int nb = nr % 5 == 0 ? 5 : (nr % 4 == 0 ? 4 : (nr % 3 == 0 ? 3 : (nr % 2 == 0 ? 2 : 1)));
    if (getenv("PM_FEAT_NB")) { const int want = atoi(getenv("PM_FEAT_NB")); if (want >= 1 && want <= 5 && nr % want == 0) nb = want; }
    const int nfull_max = (int)(smem_bytes / sizeof(double2));
    // one atom per CTA, two 512-thread CTAs per SM at 64 registers (measured on configs 3 / 4: 12.0 / 16.6 ms against 13.1 /
    // 20.4 ms with two atoms per CTA and 14.7 / 19.4 ms with one 107-register CTA per SM)
    int at = 1;
    if (getenv("PM_FEAT_AT")) { const int want = atoi(getenv("PM_FEAT_AT")); if (want == 1 || want == 2) at = want; }
    if ((size_t)at * smem_bytes > 226 * 1024) at = 1;
    if (smem_bytes > 226 * 1024) return false;
    const size_t smem = smem_bytes * at;
    const int grid = (b.n_atoms + at - 1) / at;
#define PM_F4_LAUNCH(AT_, NB_)                                                                                      \
    {                                                                                                               \
        ensure_smem((const void*)k_features_v4r<AT_, 512, NB_>, smem);                                              \
        k_features_v4r<AT_, 512, NB_><<<grid, 512, smem, s>>>(m, b, anc, dfeat, Gbuf, nfull_max, zero_g ? 1 : 0, dpv); \
    }
#define PM_F4_NB(NB_)                                                                                               \
    case NB_:                                                                                                       \
        if (at == 2) PM_F4_LAUNCH(2, NB_) else PM_F4_LAUNCH(1, NB_)                                                 \
        break;
    switch (nb) { PM_F4_NB(5) PM_F4_NB(4) PM_F4_NB(3) PM_F4_NB(2) PM_F4_NB(1) }
#undef PM_F4_NB
#undef PM_F4_LAUNCH
    return true;
}

}  // namespace pm

// pm_kernels.cu -- hand-written sm_100a kernels of the pypolymlp hot path (fp64 throughout).
//
// Pipeline per chunk of structures (all buffers stay in HBM/L2, X is only ever a chunk):
//   K1  neighbour list            NeighborFull            compute/neighbor_full.cpp:10-76
//   K2  pair basis + a_nlm        Local::compute_anlmtp_d compute/local.cpp:116-217,
//                                 get_fn_/get_ylm_        polymlp/polymlp_functions_interface.cpp:41-142
//   K3  invariants d_f and G      Features::compute_features{,_deriv}  polymlp/polymlp_features.cpp:173-248
//   K4a L = V.G per centre        (neighbour-sparse replacement of the dense (n_nlmtp x N) arrays)
//   K4b gather + polynomial       Model::model_order1..3  compute/model.cpp:159-267, apply_weights
//   K5  C += Xt^T Xt              numpy x.T @ x           src/pypolymlp/mlp_dev/core/utils_sequential.py:94-109
// This file holds the straightforward kernels (one thread per output); the tensor-core (DMMA)
// variants of K4a/K4b/K5 live in pm_kernels_mma.cu and are validated against these.
#include "pm_kernels.cuh"
#include "pm_mma.cuh"

#include <algorithm>
#include <cstdio>
#include <mutex>
#include <vector>

namespace pm {

void ensure_smem(const void* kernel, size_t smem_bytes, bool max_carveout) {
    if (smem_bytes <= 48 * 1024 && !max_carveout) return;
    struct Ent { const void* fn; size_t sz; };
    static std::vector<Ent> tab[64];
    static std::mutex mu;
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : tab[d])
        if (e.fn == kernel) {
            if (e.sz >= smem_bytes) return;
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            e.sz = smem_bytes;
            return;
        }
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (max_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    tab[d].push_back({kernel, smem_bytes});
}

// ================================================================================================
// K1: neighbour list.  One warp per centre atom; lanes sweep the (j, translation) candidates in the
// reference's order, ballot + popc keeps that order in the output.  The arithmetic uses explicit
// round-to-nearest intrinsics so that no FMA contraction can change a bit w.r.t. the reference:
//   dx = (x_j - x_i) + t_x ;  r2 = dx*dx + dy*dy + dz*dz ;  keep if r2 < rc^2 and r2 > 1e-20.
// ================================================================================================
template <bool FILL>
__global__ void __launch_bounds__(256) k_neighbor(DevModel m, DevBatch b, int* __restrict__ counts,
                                                   double* __restrict__ PB, double cutoff_sq, double tol_sq) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= b.n_atoms) return;
    const int i = warp;
    const int s = b.st_of_atom[i];
    const int a0 = b.atom_off[s];
    const int N = b.atom_off[s + 1] - a0;
    const int t0 = b.trans_off[s];
    const int T = b.trans_off[s + 1] - t0;
    const double xi = b.x[i], yi = b.y[i], zi = b.z[i];
    const int nt = m.n_type;
    const int total = N * T;
    for (int u = 0; u < nt; ++u) {
        int cnt = 0;
        const int base = FILL ? b.seg_off[i * nt + u] : 0;
        // (j, t) of this lane's candidate, advanced by 32 per iteration without an integer division
        int j = lane / T, t = lane - (lane / T) * T;
        const int dj = 32 / T, dt = 32 - (32 / T) * T;
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int c = c0 + lane;
            bool hit = false;
            double dx = 0.0, dy = 0.0, dz = 0.0;
            const int jc = j;
            if (c < total) {
                if (nt == 1 || b.types[a0 + jc] == u) {
                    const double* tr = b.trans + 3 * (size_t)(t0 + t);
                    dx = __dadd_rn(__dsub_rn(b.x[a0 + jc], xi), tr[0]);
                    dy = __dadd_rn(__dsub_rn(b.y[a0 + jc], yi), tr[1]);
                    dz = __dadd_rn(__dsub_rn(b.z[a0 + jc], zi), tr[2]);
                    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    hit = (r2 < cutoff_sq) && (r2 > tol_sq);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (FILL && hit) {
                const int pos = base + cnt + __popc(mask & ((1u << lane) - 1u));
                b.nbr[pos] = a0 + jc;
                b.centre[pos] = i;
                const PBRecW rec = pb_rec_w(PB, pos, m.pbstride);
                rec[0] = dx; rec[1] = dy; rec[2] = dz;
            }
            cnt += __popc(mask);
            j += dj; t += dt;
            if (t >= T) { t -= T; ++j; }
        }
        if (!FILL && lane == 0) counts[i * nt + u] = cnt;
    }
}

// Faster sweep for the common case (<= 128 lattice translations): lane = neighbour atom j (32 at a time), the
// translation loop is sequential per lane and records hits in a bit mask, so there is no per-candidate ballot and
// no integer division; a warp prefix sum over the lanes' hit counts restores the reference's (j, translation)
// output order.  Same exact arithmetic as k_neighbor.
template <bool FILL>
__global__ void __launch_bounds__(256) k_neighbor_mask(DevModel m, DevBatch b, int* __restrict__ counts,
                                                        double* __restrict__ PB, double cutoff_sq, double tol_sq) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= b.n_atoms) return;
    const int i = warp;
    const int s = b.st_of_atom[i];
    const int a0 = b.atom_off[s];
    const int N = b.atom_off[s + 1] - a0;
    const int t0 = b.trans_off[s];
    const int T = b.trans_off[s + 1] - t0;
    const double xi = b.x[i], yi = b.y[i], zi = b.z[i];
    const int nt = m.n_type;
    const double* __restrict__ tr = b.trans + 3 * (size_t)t0;
    for (int u = 0; u < nt; ++u) {
        int cnt = 0;
        const int base = FILL ? b.seg_off[i * nt + u] : 0;
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int j = j0 + lane;
            unsigned long long m0 = 0ull, m1 = 0ull;
            double dxij = 0.0, dyij = 0.0, dzij = 0.0;
            const bool act = j < N && (nt == 1 || b.types[a0 + j] == u);
            if (act) {
                if (FILL && b.mask_stride > 0) {   // hit masks were stored by the count pass: no second distance sweep
                    const ulonglong2 mm = b.masks[(size_t)i * b.mask_stride + j];
                    m0 = mm.x; m1 = mm.y;
                    if (m0 | m1) {
                        dxij = __dsub_rn(b.x[a0 + j], xi);
                        dyij = __dsub_rn(b.y[a0 + j], yi);
                        dzij = __dsub_rn(b.z[a0 + j], zi);
                    }
                } else {
                    dxij = __dsub_rn(b.x[a0 + j], xi);
                    dyij = __dsub_rn(b.y[a0 + j], yi);
                    dzij = __dsub_rn(b.z[a0 + j], zi);
                    for (int t = 0; t < T; ++t) {
                        const double dx = __dadd_rn(dxij, tr[3 * t]);
                        const double dy = __dadd_rn(dyij, tr[3 * t + 1]);
                        const double dz = __dadd_rn(dzij, tr[3 * t + 2]);
                        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (r2 < cutoff_sq && r2 > tol_sq) {
                            if (t < 64) m0 |= 1ull << t; else m1 |= 1ull << (t - 64);
                        }
                    }
                    if (!FILL && b.mask_stride > 0) b.masks[(size_t)i * b.mask_stride + j] = make_ulonglong2(m0, m1);
                }
            }
            const int mine = __popcll(m0) + __popcll(m1);
            int incl = mine;   // inclusive warp scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (FILL && mine) {
                int pos = base + cnt + incl - mine;
                for (int w = 0; w < 2; ++w) {
                    unsigned long long mm = w == 0 ? m0 : m1;
                    while (mm) {
                        const int t = __ffsll((long long)mm) - 1 + 64 * w;
                        mm &= mm - 1;
                        b.nbr[pos] = a0 + j;
                        b.centre[pos] = i;
                        const PBRecW rec = pb_rec_w(PB, pos, m.pbstride);
                        rec[0] = __dadd_rn(dxij, tr[3 * t]);
                        rec[1] = __dadd_rn(dyij, tr[3 * t + 1]);
                        rec[2] = __dadd_rn(dzij, tr[3 * t + 2]);
                        ++pos;
                    }
                }
            }
            cnt += total;
        }
        if (!FILL && lane == 0) counts[i * nt + u] = cnt;
    }
}

// Exclusive prefix sum of the per-(atom, neighbour type) counts: one CTA, 256 threads, each thread owns a
// contiguous slice (n is ~12k per chunk).
__global__ void __launch_bounds__(256) k_scan_exclusive(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int part[256];
    const int tid = threadIdx.x;
    const int per = (n + 255) / 256;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int k = lo; k < hi; ++k) sum += in[k];
    part[tid] = sum;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const int v = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - sum;
    for (int k = lo; k < hi; ++k) { const int v = in[k]; out[k] = run; run += v; }
}

void launch_scan_exclusive(const int* in, int* out, int n, cudaStream_t s) {
    k_scan_exclusive<<<1, 256, 0, s>>>(in, out, n);
}

void launch_neighbor_count(const DevModel& m, const DevBatch& b, int* counts, cudaStream_t s) {

    const int threads = 256;
    const int blocks = (b.n_atoms * 32 + threads - 1) / threads;
    const double tol = 1e-10;
    if (b.max_trans <= 128) {
        k_neighbor_mask<false><<<blocks, threads, 0, s>>>(m, b, counts, nullptr, m.cutoff * m.cutoff, tol * tol);
        return;
    }
    k_neighbor<false><<<blocks, threads, 0, s>>>(m, b, counts, nullptr, m.cutoff * m.cutoff, tol * tol);
}

void launch_neighbor_fill(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s) {
    const int threads = 256;
    const int blocks = (b.n_atoms * 32 + threads - 1) / threads;
    const double tol = 1e-10;
    if (b.max_trans <= 128) {
        k_neighbor_mask<true><<<blocks, threads, 0, s>>>(m, b, nullptr, PB, m.cutoff * m.cutoff, tol * tol);
        return;
    }
    k_neighbor<true><<<blocks, threads, 0, s>>>(m, b, nullptr, PB, m.cutoff * m.cutoff, tol * tol);
}

// ------------------------------------------------------------------------------------------------
// K1 with a cell list.  The masked sweep above tests every (centre, atom, translation) triple: 2.2 M tests for 13 824
// neighbours per 256-atom config-2 structure, 15.7 M for a 512-atom config-5 cell.  Here the atoms are binned in
// fractional space (bin width >= r_c / 2 along every lattice direction), a centre looks at the +-R bins around its own
// (R = 2 except for very small cells) and every candidate IMAGE is one (atom, integer lattice triple): ~200 candidates
// per centre.  The decision and the stored displacement use the reference's arithmetic on the reference's own
// translation vector (triple -> index through the host's lookup table), so the list is bit-identical; the hits are
// ranked by (neighbour type, j, translation index) to restore the reference order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cl_bin(DevBatch b, int* __restrict__ bin_count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= b.n_atoms) return;
    const int s = b.st_of_atom[a];
    const int* ci = b.cl_int + CL_NI * s;
    const double* ai = b.cl_ainv + 9 * s;
    const double x = b.x[a], y = b.y[a], z = b.z[a];
    int bin = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double f = ai[3 * d] * x + ai[3 * d + 1] * y + ai[3 * d + 2] * z;
        const double fl_ = floor(f);
        const int nb = ci[d];
        const int bd = min(nb - 1, max(0, (int)((f - fl_) * nb)));
        b.cl_atom_img[3 * a + d] = (int)fl_;
        bin = bin * nb + bd;
    }
    bin += ci[10];
    b.cl_atom_bin[a] = bin;
    atomicAdd(bin_count + bin, 1);
}

__global__ void __launch_bounds__(256) k_cl_fill(DevBatch b, int* __restrict__ cursor) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= b.n_atoms) return;
    const int bin = b.cl_atom_bin[a];
    b.cl_bin_atoms[b.cl_bin_start[bin] + atomicAdd(cursor + bin, 1)] = a;
}

__device__ __forceinline__ int cl_floordiv(int a, int n) { return a >= 0 ? a / n : -((-a + n - 1) / n); }

// Count pass: the search.  Every hit is also parked in the per-atom scratch (key, displacement) so that the fill pass only
// has to rank the hits (no second search).
__global__ void __launch_bounds__(256) k_neighbor_cl_count(DevModel m, DevBatch b, int* __restrict__ counts,
                                                            int* __restrict__ max_count, double cutoff_sq, double tol_sq) {
    __shared__ int s_nh[8];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wib;
    if (i >= b.n_atoms) return;
    const int s = b.st_of_atom[i];
    const int a0 = b.atom_off[s];
    const int t0 = b.trans_off[s];
    const int* ci = b.cl_int + CL_NI * s;
    const int nb0 = ci[0], nb1 = ci[1], nb2 = ci[2];
    const int R0 = ci[3], R1 = ci[4], R2 = ci[5];
    const int mx0 = ci[6], mx1 = ci[7], mx2 = ci[8];
    const int* tmap = b.cl_tmap + ci[9];
    const int bin0 = ci[10];
    const int nt = m.n_type;
    const double xi = b.x[i], yi = b.y[i], zi = b.z[i];
    const int ni0 = b.cl_atom_img[3 * i], ni1 = b.cl_atom_img[3 * i + 1], ni2 = b.cl_atom_img[3 * i + 2];
    int bi = b.cl_atom_bin[i] - bin0;
    const int b2 = bi % nb2; bi /= nb2;
    const int b1 = bi % nb1;
    const int b0 = bi / nb1;
    if (lane == 0) s_nh[wib] = 0;
    __syncwarp();
    int* hkey = b.cl_hkey + (size_t)i * CL_CAP;
    double* hd = b.cl_hd + (size_t)i * CL_CAP * 3;
    int cnt[MAXT] = {0, 0, 0, 0};
    const int s1 = 2 * R1 + 1, s2 = 2 * R2 + 1;
    const int nsearch = (2 * R0 + 1) * s1 * s2;
    for (int bidx = lane; bidx < nsearch; bidx += 32) {   // lanes walk the searched bins
        const int d2 = bidx % s2 - R2, d1 = (bidx / s2) % s1 - R1, d0 = bidx / (s1 * s2) - R0;
        const int u0 = ni0 * nb0 + b0 + d0, u1 = ni1 * nb1 + b1 + d1, u2 = ni2 * nb2 + b2 + d2;   // unwrapped bin
        const int q0 = cl_floordiv(u0, nb0), q1 = cl_floordiv(u1, nb1), q2 = cl_floordiv(u2, nb2);
        const int bin = ((u0 - q0 * nb0) * nb1 + (u1 - q1 * nb1)) * nb2 + (u2 - q2 * nb2) + bin0;
        for (int k = b.cl_bin_start[bin]; k < b.cl_bin_start[bin + 1]; ++k) {
            const int j = b.cl_bin_atoms[k];
            // the image of j in this bin is j shifted by the lattice triple L = q - floor(frac_j)
            const int L0 = q0 - b.cl_atom_img[3 * j], L1 = q1 - b.cl_atom_img[3 * j + 1], L2 = q2 - b.cl_atom_img[3 * j + 2];
            if (abs(L0) > mx0 || abs(L1) > mx1 || abs(L2) > mx2) continue;   // not in the reference's translation list
            const int t = tmap[((L0 + mx0) * (2 * mx1 + 1) + (L1 + mx1)) * (2 * mx2 + 1) + (L2 + mx2)];
            if (t < 0) continue;
            const double* tr = b.trans + 3 * (size_t)(t0 + t);
            const double dx = __dadd_rn(__dsub_rn(b.x[j], xi), tr[0]);
            const double dy = __dadd_rn(__dsub_rn(b.y[j], yi), tr[1]);
            const double dz = __dadd_rn(__dsub_rn(b.z[j], zi), tr[2]);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (!(r2 < cutoff_sq && r2 > tol_sq)) continue;
            const int u = nt == 1 ? 0 : b.types[j];
#pragma unroll
            for (int v = 0; v < MAXT; ++v) cnt[v] += (u == v) ? 1 : 0;
            const int pos = atomicAdd(&s_nh[wib], 1);
            if (pos < CL_CAP) {
                hkey[pos] = (u * b.cl_nmax + (j - a0)) * b.cl_tmax + t;
                hd[3 * pos] = dx; hd[3 * pos + 1] = dy; hd[3 * pos + 2] = dz;
            }
        }
    }
    int tot = 0;
#pragma unroll
    for (int v = 0; v < MAXT; ++v) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt[v] += __shfl_xor_sync(0xffffffffu, cnt[v], d);
        if (v < nt) { if (lane == 0) counts[i * nt + v] = cnt[v]; tot += cnt[v]; }
    }
    if (lane == 0) atomicMax(max_count, tot);
}

// Fill pass: rank the parked hits of each atom by (neighbour type, j, translation) and write them in that order.
__global__ void __launch_bounds__(256) k_neighbor_cl_fill(DevModel m, DevBatch b, double* __restrict__ PB) {
    __shared__ int s_key[8][CL_CAP];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wib;
    if (i >= b.n_atoms) return;
    const int nt = m.n_type;
    const int a0 = b.atom_off[b.st_of_atom[i]];
    const int nh = min(CL_CAP, b.seg_off[i * nt + nt] - b.seg_off[i * nt]);
    const int* hkey = b.cl_hkey + (size_t)i * CL_CAP;
    const double* hd = b.cl_hd + (size_t)i * CL_CAP * 3;
    for (int hh = lane; hh < nh; hh += 32) s_key[wib][hh] = hkey[hh];
    __syncwarp();
    const int per_type = b.cl_nmax * b.cl_tmax;
    for (int hh = lane; hh < nh; hh += 32) {
        const int key = s_key[wib][hh];
        const int u = key / per_type;
        int rank = 0;
        for (int o = 0; o < nh; ++o) {
            const int ko = s_key[wib][o];
            rank += (ko < key && ko / per_type == u) ? 1 : 0;
        }
        const int pos = b.seg_off[i * nt + u] + rank;
        b.nbr[pos] = a0 + (key - u * per_type) / b.cl_tmax;
        b.centre[pos] = i;
        const PBRecW rec = pb_rec_w(PB, pos, m.pbstride);
        rec[0] = hd[3 * hh]; rec[1] = hd[3 * hh + 1]; rec[2] = hd[3 * hh + 2];
    }
}

void launch_cl_bins(const DevModel& m, const DevBatch& b, int n_bins, int* bin_count, cudaStream_t s) {
    // bin_count: [n_bins + 1] scratch (counts, then fill cursors); b.cl_bin_start receives the exclusive scan
    cudaMemsetAsync(bin_count, 0, (size_t)(n_bins + 1) * sizeof(int), s);
    k_cl_bin<<<(b.n_atoms + 255) / 256, 256, 0, s>>>(b, bin_count);
    launch_scan_exclusive(bin_count, b.cl_bin_start, n_bins + 1, s);
    cudaMemsetAsync(bin_count, 0, (size_t)(n_bins + 1) * sizeof(int), s);
    k_cl_fill<<<(b.n_atoms + 255) / 256, 256, 0, s>>>(b, bin_count);
}

void launch_neighbor_cl_count(const DevModel& m, const DevBatch& b, int* counts, int* max_count, cudaStream_t s) {
    const double tol = 1e-10;
    k_neighbor_cl_count<<<(b.n_atoms + 7) / 8, 256, 0, s>>>(m, b, counts, max_count, m.cutoff * m.cutoff, tol * tol);
}

void launch_neighbor_cl_fill(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s) {
    k_neighbor_cl_fill<<<(b.n_atoms + 7) / 8, 256, 0, s>>>(m, b, PB);
}

// reverse pair: (i -> j, D) <-> (j -> i, -D); exact because negation is exact in every step above.
__global__ void __launch_bounds__(256) k_neighbor_rev(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                       int* __restrict__ errflag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n_pairs) return;
    const int i = b.centre[p], j = b.nbr[p];
    const PBRec rec = pb_rec(PB, p, m.pbstride);
    const double dx = -rec[0], dy = -rec[1], dz = -rec[2];
    const int u = b.types[i];
    const int q0 = b.seg_off[j * m.n_type + u], q1 = b.seg_off[j * m.n_type + u + 1];
    int found = -1;
    for (int q = q0; q < q1; ++q) {
        if (b.nbr[q] != i) continue;
        const PBRec rq = pb_rec(PB, q, m.pbstride);
        if (rq[0] == dx && rq[1] == dy && rq[2] == dz) { found = q; break; }
    }
    b.rev[p] = found;
    if (found < 0) atomicExch(errflag, 1);
}

void launch_neighbor_rev(const DevModel& m, const DevBatch& b, const double* PB, int* errflag, cudaStream_t s) {
    if (b.n_pairs == 0) return;
    k_neighbor_rev<<<(b.n_pairs + 255) / 256, 256, 0, s>>>(m, b, PB, errflag);
}

// ================================================================================================
// K2a: per-pair basis record: radial functions (Gaussian x cosine cutoff) with d/dr, complex Y_lm
// (m <= 0) and Cartesian gradients by the normalised associated-Legendre recurrences.
// ================================================================================================
// pair-independent coefficients of the Legendre recurrences (filled once on the host with the same
// IEEE expressions the reference evaluates per pair)
constexpr int CT_L = 21;                       // l = 0..20
__constant__ double c_c1[CT_L];                // -sqrt(1 + 0.5 / l)
__constant__ double c_c2[CT_L];                // sqrt(2 (l - 1) + 3)
__constant__ double c_alm[CT_L * CT_L];        // sqrt((4 l^2 - 1) / (l^2 - m^2))
__constant__ double c_blm[CT_L * CT_L];        // -sqrt(((l-1)^2 - m^2) / (4 (l-1)^2 - 1))
__constant__ double c_sq0[CT_L];               // sqrt(0.5 l (l + 1))
__constant__ double c_sqd[CT_L * CT_L];        // sqrt((l - m)(l + m + 1))

void init_pair_basis_tables() {
    static bool done[64] = {false};
    static std::mutex mu;   // (several host threads may start their first chunk at the same time: pm_multi, the eval lanes)
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return;
    double c1[CT_L] = {0}, c2[CT_L] = {0}, sq0[CT_L] = {0};
    static double alm[CT_L * CT_L], blm[CT_L * CT_L], sqd[CT_L * CT_L];
    for (int l = 0; l < CT_L; ++l) {
        if (l >= 1) { c1[l] = -sqrt(1.0 + 0.5 / l); c2[l] = sqrt(2.0 * (l - 1.0) + 3.0); }
        sq0[l] = sqrt(0.5 * l * (l + 1));
        const double ls = (double)(l * l), lm1s = (double)((l - 1) * (l - 1));
        for (int mm = 0; mm < CT_L; ++mm) {
            const double ms = (double)(mm * mm);
            alm[l * CT_L + mm] = (l >= 2 && mm <= l - 2) ? sqrt((4.0 * ls - 1.0) / (ls - ms)) : 0.0;
            blm[l * CT_L + mm] = (l >= 2 && mm <= l - 2) ? -sqrt((lm1s - ms) / (4.0 * lm1s - 1.0)) : 0.0;
            sqd[l * CT_L + mm] = (mm <= l) ? sqrt((double)((l - mm) * (l + mm + 1))) : 0.0;
        }
    }
    cudaMemcpyToSymbol(c_c1, c1, sizeof(c1));
    cudaMemcpyToSymbol(c_c2, c2, sizeof(c2));
    cudaMemcpyToSymbol(c_sq0, sq0, sizeof(sq0));
    cudaMemcpyToSymbol(c_alm, alm, sizeof(alm));
    cudaMemcpyToSymbol(c_blm, blm, sizeof(blm));
    cudaMemcpyToSymbol(c_sqd, sqd, sizeof(sqd));
    done[dev] = true;
}

// Radial part of a pair: cosine cutoff x Gaussians and d/dr, handed to rad(n, f_n, f_n') for n < m.n_fn.
__device__ __forceinline__ void radial_cutoff(const DevModel& m, double r, double& fc, double& fcd) {
    const double pi = 3.1415926535897932384626433832795;
    fc = 0.0; fcd = 0.0;
    if (r < m.cutoff) {
        const double v1 = pi / m.cutoff, v2 = v1 * r;
        double sv, cv;
        sincos(v2, &sv, &cv);
        fc = 0.5 * (cv + 1.0);
        fcd = -0.5 * v1 * sv;
    }
}

__device__ __forceinline__ void radial_one(const double* __restrict__ prm, int nfn, int n, double r, double fc, double fcd,
                                           double& fn, double& fnd) {
    fn = 0.0; fnd = 0.0;
    if (n < nfn) {
        const double beta = prm[2 * n], mu = prm[2 * n + 1];
        const double d = r - mu;
        const double bf = exp(-beta * d * d);
        const double bfd = -2.0 * beta * d * bf;
        fn = bf * fc;
        fnd = bfd * fc + bf * fcd;
        if (fn < 1e-20) { fn = 0.0; fnd = 0.0; }  // reference skip rule (local.cpp:164)
    }
}

template <class Rad>
__device__ __forceinline__ void pair_radial(const DevModel& m, double r, int tp, Rad rad) {
    double fc, fcd;
    radial_cutoff(m, r, fc, fcd);
    const int nfn = m.tp_nfn[tp];
    const double* prm = m.tp_params + (size_t)tp * m.n_fn * 2;
    for (int n = 0; n < m.n_fn; ++n) {
        double fn, fnd;
        radial_one(prm, nfn, n, r, fc, fcd, fn, fnd);
        rad(n, fn, fnd);
    }
}

// Angular part of a pair: complex Y_lm (m <= 0) and its Cartesian gradient, handed key by key to
// ang(key, Re Y, Im Y, Re dY/dx, Im dY/dx, Re dY/dy, Im dY/dy, Re dY/dz, Im dY/dz); key = l (l + 1) / 2 + l - |m| is the
// index of (l, m) in the record.  Order of the calls: m = 0 for every l, then |m| = 1, 2, ... for l >= |m|.
template <int LT, class Ang>
__device__ __forceinline__ void pair_angular(const DevModel& m, double dx, double dy, double dz, double r, double rinv, Ang ang) {
    const int L = LT >= 0 ? LT : m.maxl;
    constexpr int NHT = LT >= 0 ? (LT + 1) * (LT + 2) / 2 : MAX_NH;
    // a true division, as the reference does (polymlp_functions_interface.cpp:124 cos_theta = z / r): dz * rinv can be
    // 1 - 1 ulp for a neighbour exactly on the z axis, and st = sqrt(1 - ct^2) then comes out as 1.5e-8 instead of 0
    // (seen as 6e-8 eV/A residual forces in the ideal perovskite cell, where the reference gives 1e-16)
    const double ct = dz / r;
    const double rho = hypot(dx, dy);
    double cp = 1.0, sp = 0.0;
    if (rho > 0.0) { cp = dx / rho; sp = dy / rho; }
    double pl[NHT], ql[NHT];
    const double s2pi = 0.39894228040143267794;
    const double st = sqrt(1.0 - ct * ct);
#define LM2I(l, mm) ((l) * ((l) + 1) / 2 + (mm))
    pl[0] = s2pi; ql[0] = 0.0;
    if (L >= 1) {
        pl[LM2I(1, 0)] = ct * 1.7320508075688772935 * s2pi; ql[LM2I(1, 0)] = 0.0;
        pl[LM2I(1, 1)] = -st * 1.2247448713915890491 * s2pi; ql[LM2I(1, 1)] = -1.2247448713915890491 * s2pi;
#pragma unroll
        for (int l = 2; l <= L; ++l) {
            const double c1 = c_c1[l] * st;
            pl[LM2I(l, l)] = c1 * pl[LM2I(l - 1, l - 1)];
            ql[LM2I(l, l)] = c1 * ql[LM2I(l - 1, l - 1)];
            const double c2 = c_c2[l] * ct;
            pl[LM2I(l, l - 1)] = c2 * pl[LM2I(l - 1, l - 1)];
            ql[LM2I(l, l - 1)] = c2 * ql[LM2I(l - 1, l - 1)];
        }
#pragma unroll
        for (int l = 2; l <= L; ++l) {
#pragma unroll
            for (int mm = 0; mm <= l - 2; ++mm) {
                const double alm = c_alm[l * CT_L + mm];
                const double blm = c_blm[l * CT_L + mm];
                pl[LM2I(l, mm)] = alm * (ct * pl[LM2I(l - 1, mm)] + blm * pl[LM2I(l - 2, mm)]);
                ql[LM2I(l, mm)] = alm * (ct * ql[LM2I(l - 1, mm)] + blm * ql[LM2I(l - 2, mm)]);
            }
        }
    }
    const double hs2 = 0.70710678118654752440;
#pragma unroll
    for (int l = 0; l <= L; ++l) {
        const int idx = LM2I(l, 0) + l;
        double common = 0.0;
        if (l >= 1) common = ql[LM2I(l, 1)] * st * rinv * c_sq0[l];
        ang(idx, pl[LM2I(l, 0)] * hs2, 0.0, common * ct * cp, 0.0, common * ct * sp, 0.0, -common * st, 0.0);
    }
    double c1 = 1.0, c2 = cp, s1 = 0.0, s2 = -sp;
    const double tc = 2.0 * c2;
    double sign = -1.0;
#pragma unroll
    for (int mp = 1; mp <= L; ++mp) {
        const double sn = tc * s1 - s2;
        const double cs = tc * c1 - c2;
        c2 = c1; c1 = cs; s2 = s1; s1 = sn;
#pragma unroll
        for (int l = mp; l <= L; ++l) {
            const int idx = LM2I(l, -mp) + l;
            const double tmp = sign * pl[LM2I(l, mp)] * hs2;
            // common = e^{i m phi} / sqrt(2) / r
            const double cr = cs * hs2 * rinv, ci = sn * hs2 * rinv;
            double dth = mp * ct * ql[LM2I(l, mp)];
            if (mp != l) dth += c_sqd[l * CT_L + mp] * ql[LM2I(l, mp + 1)] * st;
            const double dph = mp * ql[LM2I(l, mp)];  // dphi = i * dph
            // x: common * (dth*ct*cp - i*dph*sp) ; y: common * (dth*ct*sp + i*dph*cp) ; z: -common*dth*st
            const double ax = dth * ct * cp, bx = -dph * sp;
            const double ay = dth * ct * sp, by = dph * cp;
            const double az = -dth * st;
            // (cr + i ci)(a + i b) = (cr a - ci b) + i (cr b + ci a); result = sign * conj(.)
            ang(idx, tmp * cs, -tmp * sn, sign * (cr * ax - ci * bx), -sign * (cr * bx + ci * ax),
                sign * (cr * ay - ci * by), -sign * (cr * by + ci * ay), sign * (cr * az), -sign * (ci * az));
        }
        sign = -sign;
    }
#undef LM2I
}

// One pair-basis record: every item is handed to put(item, value) (k_pair_basis stores it in the blocked global layout,
// the fused k_pair_anlm also keeps it in its shared-memory tile).
template <int LT, class Put>
__device__ __forceinline__ void pair_basis_items(const DevModel& m, double dx, double dy, double dz, int tp, Put put) {
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const double rinv = 1.0 / r;
    put(3, rinv);
    pair_radial(m, r, tp, [&](int n, double fn, double fnd) {
        put(4 + n, fn);
        put(4 + m.n_fn + n, fnd);
    });
    const int oY = pb_y(m, 0), oYx = pb_y(m, 1), oYy = pb_y(m, 2), oYz = pb_y(m, 3);
    pair_angular<LT>(m, dx, dy, dz, r, rinv, [&](int idx, double yr, double yi, double xr, double xi, double wr, double wi,
                                                  double zr, double zi) {
        put(oY + 2 * idx, yr); put(oY + 2 * idx + 1, yi);
        put(oYx + 2 * idx, xr); put(oYx + 2 * idx + 1, xi);
        put(oYy + 2 * idx, wr); put(oYy + 2 * idx + 1, wi);
        put(oYz + 2 * idx, zr); put(oYz + 2 * idx + 1, zi);
    });
}

template <int LT>
__global__ void __launch_bounds__(128) k_pair_basis(DevModel m, DevBatch b, double* __restrict__ PB) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n_pairs) return;
    const PBRecW rec = pb_rec_w(PB, p, m.pbstride);
    const double dx = rec[0], dy = rec[1], dz = rec[2];
    const int ti = b.types[b.centre[p]], tj = b.types[b.nbr[p]];
    const int tp = m.type_pairs[ti * m.n_type + tj];
    pair_basis_items<LT>(m, dx, dy, dz, tp, [&](int item, double v) { rec[item] = v; });
}

void launch_pair_basis(const DevModel& m, const DevBatch& b, double* PB, cudaStream_t s) {
    if (b.n_pairs == 0) return;
    init_pair_basis_tables();
    const int blocks = (b.n_pairs + 127) / 128;
    switch (m.maxl) {
        case 0: k_pair_basis<0><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 1: k_pair_basis<1><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 2: k_pair_basis<2><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 3: k_pair_basis<3><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 4: k_pair_basis<4><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 5: k_pair_basis<5><<<blocks, 128, 0, s>>>(m, b, PB); break;
        case 6: k_pair_basis<6><<<blocks, 128, 0, s>>>(m, b, PB); break;
        default: k_pair_basis<-1><<<blocks, 128, 0, s>>>(m, b, PB); break;
    }
}

// ================================================================================================
// K2b: a_nlm(i) = sum_j f_n Y_lm for the m <= 0 heads, plus the aggregated derivative rows
//   own[alpha]  = sum_j v_alpha(ij)            (row of atom i itself)
//   str[ab]     = -sum_j v_alpha(ij) D_beta    (six virial rows xx,yy,zz,xy,yz,zx)
// with v_alpha = f_n' Y D_alpha / r + f_n dY/dalpha (local.cpp:175-196).  One CTA per atom, one
// thread per head; accumulation is thread-private (deterministic).
// ================================================================================================
__global__ void __launch_bounds__(128) k_anlm(DevModel m, DevBatch b, const double* __restrict__ PB,
                                               double2* __restrict__ anc, double2* __restrict__ agg) {
    const int i = blockIdx.x;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const bool force = b.force[b.st_of_atom[i]] != 0 && b.need_agg != 0;
    const int nt = m.n_type;
    const int oy = pb_y(m, 0), oyx = pb_y(m, 1), oyy = pb_y(m, 2), oyz = pb_y(m, 3);
    for (int h = threadIdx.x; h < T.n_head; h += blockDim.x) {
        const int u = T.head_seg[h], nid = T.head_nid[h], key = T.head_key[h];
        const int p0 = b.seg_off[i * nt + u], p1 = b.seg_off[i * nt + u + 1];
        double ar = 0.0, ai = 0.0;
        double gr[9], gi[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { gr[k] = 0.0; gi[k] = 0.0; }
        for (int p = p0; p < p1; ++p) {
            const PBRec rec = pb_rec(PB, p, m.pbstride);
            const double fn = rec[4 + nid];
            if (fn == 0.0) continue;
            const double yr = rec[oy + 2 * key], yi = rec[oy + 2 * key + 1];
            ar += fn * yr; ai += fn * yi;
            if (force) {
                const double dx = rec[0], dy = rec[1], dz = rec[2];
                const double d1 = rec[4 + m.n_fn + nid] * rec[3];
                const double d1r = d1 * yr, d1i = d1 * yi;
                const double vxr = d1r * dx + fn * rec[oyx + 2 * key], vxi = d1i * dx + fn * rec[oyx + 2 * key + 1];
                const double vyr = d1r * dy + fn * rec[oyy + 2 * key], vyi = d1i * dy + fn * rec[oyy + 2 * key + 1];
                const double vzr = d1r * dz + fn * rec[oyz + 2 * key], vzi = d1i * dz + fn * rec[oyz + 2 * key + 1];
                gr[0] += vxr; gi[0] += vxi; gr[1] += vyr; gi[1] += vyi; gr[2] += vzr; gi[2] += vzi;
                gr[3] -= vxr * dx; gi[3] -= vxi * dx;
                gr[4] -= vyr * dy; gi[4] -= vyi * dy;
                gr[5] -= vzr * dz; gi[5] -= vzi * dz;
                gr[6] -= vxr * dy; gi[6] -= vxi * dy;
                gr[7] -= vyr * dz; gi[7] -= vyi * dz;
                gr[8] -= vzr * dx; gi[8] -= vzi * dx;
            }
        }
        anc[(size_t)i * m.hmax + h] = make_double2(ar, ai);
        if (force) {
            double2* g = agg + ((size_t)i * m.hmax + h) * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) g[k] = make_double2(gr[k], gi[k]);
        }
    }
}

// Same sums with the pair records staged through shared memory (coalesced 16-byte copies of 8 records at a
// time, double buffered with cp.async); one thread per head, so accumulation stays thread-private.
// AN_PT = pairs per staged tile: 16 standalone; 8 keeps the CTA under 24 KB of shared memory so that it fits on an SM
// next to the persistent SYRK CTA when K5 of the previous chunk runs concurrently (pm_capi.cu, stream2).
template <int AN_PT>
__global__ void __launch_bounds__(256) k_anlm_v2(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                  double2* __restrict__ anc, double2* __restrict__ agg) {
    extern __shared__ __align__(16) double sm_an[];
    const int i = blockIdx.x;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const bool force = b.force[b.st_of_atom[i]] != 0 && b.need_agg != 0;
    const int nt = m.n_type;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int stride = m.pbstride;
    int pa = b.seg_off[i * nt], pe = b.seg_off[i * nt + nt];
    const int oy = pb_y(m, 0), oyx = pb_y(m, 1), oyy = pb_y(m, 2), oyz = pb_y(m, 3);
    // models with more than 256 heads per atom (max_l ~ 12): blockIdx.y selects a pass of blockDim.x heads; the heads are
    // a pass stages only the pair range that covers the neighbour-type segments of its heads
    const int h = blockIdx.y * nthr + tid;
    const bool active = h < T.n_head;
    int nid = 0, key = 0, ps0 = 0, ps1 = 0;
    if (active) {
        const int u = T.head_seg[h];
        nid = T.head_nid[h];
        key = T.head_key[h];
        ps0 = b.seg_off[i * nt + u];
        ps1 = b.seg_off[i * nt + u + 1];
    }
    if (gridDim.y > 1) {
        __shared__ int s_pa, s_pe;
        if ((int)(blockIdx.y * nthr) >= T.n_head) return;   // (uniform: no head in this pass)
        if (tid == 0) { s_pa = pe; s_pe = pa; }
        __syncthreads();
        if (active) { atomicMin(&s_pa, ps0); atomicMax(&s_pe, ps1); }
        __syncthreads();
        pa = s_pa;
        pe = max(s_pe, s_pa);
    }
    double ar = 0.0, ai = 0.0, gr[9], gi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { gr[k] = 0.0; gi[k] = 0.0; }
    auto issue = [&](int pfirst, double* dst) {
        const int np = min(AN_PT, pe - pfirst);
        const int tot = np * stride;
        for (int e = tid; e < tot; e += nthr) {
            const int it = e / np, pp = e - it * np;
            const int p = pfirst + pp;
            const double* src = PB + ((size_t)(p >> 5) * stride + it) * PB_BLK + (p & 31);
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + it * (AN_PT + 1) + pp);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(src));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    if (pe > pa) issue(pa, sm_an);
    int buf = 0;
    for (int pf = pa; pf < pe; pf += AN_PT, buf ^= 1) {
        __syncthreads();
        if (pf + AN_PT < pe) {
            issue(pf + AN_PT, sm_an + (size_t)(buf ^ 1) * (AN_PT + 1) * stride);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        if (!active) continue;
        const double* tile = sm_an + (size_t)buf * (AN_PT + 1) * stride;
        const int q0 = max(pf, ps0), q1 = min(min(pf + AN_PT, pe), ps1);
        for (int p = q0; p < q1; ++p) {
            const double* rb = tile + (p - pf);
#define REC(item) rb[(item) * (AN_PT + 1)]
            const double fn = REC(4 + nid);
            if (fn == 0.0) continue;
            const double2 y = make_double2(REC(oy + 2 * key), REC(oy + 2 * key + 1));
            ar += fn * y.x; ai += fn * y.y;
            if (force) {
                const double dx = REC(0), dy = REC(1), dz = REC(2);
                const double d1 = REC(4 + m.n_fn + nid) * REC(3);
                const double d1r = d1 * y.x, d1i = d1 * y.y;
                const double2 yx = make_double2(REC(oyx + 2 * key), REC(oyx + 2 * key + 1));
                const double2 yy = make_double2(REC(oyy + 2 * key), REC(oyy + 2 * key + 1));
                const double2 yz = make_double2(REC(oyz + 2 * key), REC(oyz + 2 * key + 1));
                const double vxr = d1r * dx + fn * yx.x, vxi = d1i * dx + fn * yx.y;
                const double vyr = d1r * dy + fn * yy.x, vyi = d1i * dy + fn * yy.y;
                const double vzr = d1r * dz + fn * yz.x, vzi = d1i * dz + fn * yz.y;
                gr[0] += vxr; gi[0] += vxi; gr[1] += vyr; gi[1] += vyi; gr[2] += vzr; gi[2] += vzi;
                gr[3] -= vxr * dx; gi[3] -= vxi * dx;
                gr[4] -= vyr * dy; gi[4] -= vyi * dy;
                gr[5] -= vzr * dz; gi[5] -= vzi * dz;
                gr[6] -= vxr * dy; gi[6] -= vxi * dy;
                gr[7] -= vyr * dz; gi[7] -= vyi * dz;
                gr[8] -= vzr * dx; gi[8] -= vzi * dx;
            }
#undef REC
        }
    }
    if (!active) return;
    anc[(size_t)i * m.hmax + h] = make_double2(ar, ai);
    if (force) {
        double2* gp = agg + ((size_t)i * m.hmax + h) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) gp[k] = make_double2(gr[k], gi[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// K2a + K2b fused (models whose pair record fits a 64-pair shared-memory tile: max_l <= ~6): one CTA per atom.
// Phase 1: thread = pair, the whole basis record is computed in registers and written BOTH to the global pair-basis
// buffer (K4a / the eval pair pass read it) and to the CTA's shared-memory tile [item][pair]; phase 2: thread = head,
// thread-private sums over the tile as in k_anlm_v2.  Compared with the two separate kernels this removes the 16 MB /
// structure re-read of the records and the ~7800 eight-byte cp.async staging copies per atom that paced k_anlm_v2.
// ------------------------------------------------------------------------------------------------
constexpr int PA_PT = 64;          // pairs per tile
constexpr int PA_LD = PA_PT + 1;   // odd row stride: the head threads read different items of one pair

template <int LT, bool STORE>
__global__ void __maxnreg__(80) k_pair_anlm(DevModel m, DevBatch b, double* __restrict__ PB,
                                                    double2* __restrict__ anc, double2* __restrict__ agg) {
    extern __shared__ __align__(16) double sm_pa[];   // [pbstride][PA_LD]
    const int i = blockIdx.x;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const bool force = b.force[b.st_of_atom[i]] != 0 && b.need_agg != 0;
    const int nt = m.n_type;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int pa = b.seg_off[i * nt], pe = b.seg_off[i * nt + nt];
    const int oy = pb_y(m, 0), oyx = pb_y(m, 1), oyy = pb_y(m, 2), oyz = pb_y(m, 3);
    const int h = tid;
    const bool active = h < T.n_head;
    int nid = 0, key = 0, ps0 = 0, ps1 = 0;
    if (active) {
        const int u = T.head_seg[h];
        nid = T.head_nid[h];
        key = T.head_key[h];
        ps0 = b.seg_off[i * nt + u];
        ps1 = b.seg_off[i * nt + u + 1];
    }
    // Tile by tile (one tile covers <= 64 neighbours, i.e. nearly every atom): the sums are tile-local so that no
    // accumulator is live across the register-hungry record computation; later tiles add to the stored sums (same
    // thread, fixed order: deterministic).  An atom without neighbours still stores its zeros.
    for (int pf = pa; pf < pe || pf == pa; pf += PA_PT) {
        if (pf > pa) __syncthreads();   // the previous tile has been consumed
        for (int pp = tid; pp < min(PA_PT, pe - pf); pp += nthr) {
            const int p = pf + pp;
            const PBRecW rec = pb_rec_w(PB, p, m.pbstride);
            const double dx = rec[0], dy = rec[1], dz = rec[2];
            const int tp = m.type_pairs[t * nt + b.types[b.nbr[p]]];
            double* col = sm_pa + pp;
            col[0] = dx; col[PA_LD] = dy; col[2 * PA_LD] = dz;
            pair_basis_items<LT>(m, dx, dy, dz, tp, [&](int item, double v) {
                if (STORE) rec[item] = v;   // (the eval pair pass recomputes the record from D: no 1.2 KB / pair store)
                col[item * PA_LD] = v;
            });
        }
        __syncthreads();
        if (!active) continue;
        double ar = 0.0, ai = 0.0, gr[9], gi[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { gr[k] = 0.0; gi[k] = 0.0; }
        const int q0 = max(pf, ps0), q1 = min(min(pf + PA_PT, pe), ps1);
        for (int p = q0; p < q1; ++p) {
            const double* rb = sm_pa + (p - pf);
#define REC(item) rb[(item) * PA_LD]
            const double fn = REC(4 + nid);
            if (fn == 0.0) continue;   // (a branch-free, unrolled body measured slower: 14.4 vs 13.2 us / structure)
            const double2 y = make_double2(REC(oy + 2 * key), REC(oy + 2 * key + 1));
            ar += fn * y.x; ai += fn * y.y;
            if (force) {
                const double dx = REC(0), dy = REC(1), dz = REC(2);
                const double d1 = REC(4 + m.n_fn + nid) * REC(3);
                const double d1r = d1 * y.x, d1i = d1 * y.y;
                const double2 yx = make_double2(REC(oyx + 2 * key), REC(oyx + 2 * key + 1));
                const double2 yy = make_double2(REC(oyy + 2 * key), REC(oyy + 2 * key + 1));
                const double2 yz = make_double2(REC(oyz + 2 * key), REC(oyz + 2 * key + 1));
                const double vxr = d1r * dx + fn * yx.x, vxi = d1i * dx + fn * yx.y;
                const double vyr = d1r * dy + fn * yy.x, vyi = d1i * dy + fn * yy.y;
                const double vzr = d1r * dz + fn * yz.x, vzi = d1i * dz + fn * yz.y;
                gr[0] += vxr; gi[0] += vxi; gr[1] += vyr; gi[1] += vyi; gr[2] += vzr; gi[2] += vzi;
                gr[3] -= vxr * dx; gi[3] -= vxi * dx;
                gr[4] -= vyr * dy; gi[4] -= vyi * dy;
                gr[5] -= vzr * dz; gi[5] -= vzi * dz;
                gr[6] -= vxr * dy; gi[6] -= vxi * dy;
                gr[7] -= vyr * dz; gi[7] -= vyi * dz;
                gr[8] -= vzr * dx; gi[8] -= vzi * dx;
            }
#undef REC
        }
        double2* ap = anc + (size_t)i * m.hmax + h;
        double2* gp = agg + ((size_t)i * m.hmax + h) * 9;
        if (pf == pa) {
            *ap = make_double2(ar, ai);
            if (force) {
#pragma unroll
                for (int k = 0; k < 9; ++k) gp[k] = make_double2(gr[k], gi[k]);
            }
        } else {
            const double2 o = *ap;
            *ap = make_double2(o.x + ar, o.y + ai);
            if (force) {
#pragma unroll
                for (int k = 0; k < 9; ++k) { const double2 g0 = gp[k]; gp[k] = make_double2(g0.x + gr[k], g0.y + gi[k]); }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 for evaluation (the pair pass recomputes the records, nothing is stored and no derivative item is needed here): one CTA
// per AE_AT consecutive atoms.  Phase 1: thread = pair of any of these atoms, f_n and Y_lm only (the compiler drops the
// gradient arithmetic of pair_angular) into a shared-memory tile [item][pair]; phase 2: thread = (atom, head), private sum
// over the atom's pairs of the head's neighbour-type segment, in pair order (deterministic).  Against k_pair_anlm (one atom
// per CTA, a third of the threads busy in phase 1, 1.2 KB stored per pair) this took the eval K2 stage from 0.45 to ... ms
// per 21 000 atoms.
// ------------------------------------------------------------------------------------------------
constexpr int AE_AT = 4;      // atoms per CTA
constexpr int AE_PT = 256;    // pairs per tile
constexpr int AE_NI = 4;      // (atom, head) items per thread: AE_AT * hmax <= AE_NI * 256

// RECT (DevModel::front2: one atom type, every radial group lists all m <= 0 heads in Y_lm key order): phase 2 is the small
// matrix product A[n][y] = sum_p F[n][p] Y[y][p] (n: radial index, y: Re / Im of a Y_lm key, p: the atom's pairs) on the
// FP64 tensor cores, two warps per atom (each takes half of the y tiles): 14 k-steps of 8 DMMA per atom instead of 150 heads
// x 54 pairs x (3 LDS + 2 DFMA).  The tile row stride is 4 (mod 16) so that the fragment loads are conflict-free.
template <int LT, bool RECT>
__global__ void __launch_bounds__(256) k_anlm_eval(DevModel m, DevBatch b, double* __restrict__ PBw,
                                                    double2* __restrict__ anc, int store_radial) {
    constexpr int AE_LD = RECT ? AE_PT + 4 : AE_PT + 1;
    constexpr int NH = LT >= 0 ? (LT + 1) * (LT + 2) / 2 : MAX_NH;
    constexpr int NTW = ((2 * NH + 7) / 8 + 1) / 2;   // 8-column tiles of y per warp (RECT)
    const double* PB = PBw;
    extern __shared__ __align__(16) double sm_ae[];   // [n_fn + 2 nh][AE_LD]
    const int i0 = blockIdx.x * AE_AT;
    const int na = min(AE_AT, b.n_atoms - i0);
    const int nt = m.n_type;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pa = b.seg_off[i0 * nt], pe = b.seg_off[(i0 + na) * nt];
    const int oy = m.n_fn;   // tile rows: f_n (n_fn), then Re / Im Y per key
    // generic path: the (atom, head) items of this thread
    int it_nid[AE_NI], it_key[AE_NI], it_p0[AE_NI], it_p1[AE_NI];
    double ar[AE_NI], ai[AE_NI];
    // RECT path: fragment rows of this lane and the accumulator tiles [n tile][y tile]
    int arow[2] = {-1, -1}, brow[NTW], ahead[2] = {-1, -1};
    double acc[2][NTW][2];
    if (RECT) {
        const DevType& T = m.types[0];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int n = mt * 8 + (lane >> 2);
            if (n < m.n_fn && T.seg_n_off[0][n + 1] > T.seg_n_off[0][n]) {
                ahead[mt] = T.seg_n_off[0][n];                       // position of (n, key 0) in the segment's head list
                arow[mt] = T.head_nid[T.seg_heads[0][ahead[mt]]];    // tile row of f_n
            }
        }
#pragma unroll
        for (int q = 0; q < NTW; ++q) {
            const int y = (((warp & 1) * NTW + q) * 8) + (lane >> 2);
            brow[q] = y < 2 * m.nh ? oy + y : -1;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { acc[mt][q][0] = 0.0; acc[mt][q][1] = 0.0; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < AE_NI; ++k) {
            const int w = tid + k * 256;
            const int a = w / m.hmax, h = w - a * m.hmax;
            it_p0[k] = 0; it_p1[k] = 0; it_nid[k] = 0; it_key[k] = 0; ar[k] = 0.0; ai[k] = 0.0;
            if (a < na) {
                const int i = i0 + a;
                const DevType& T = m.types[b.types[i]];
                if (h < T.n_head) {
                    const int u = T.head_seg[h];
                    it_nid[k] = T.head_nid[h];
                    it_key[k] = T.head_key[h];
                    it_p0[k] = b.seg_off[i * nt + u];
                    it_p1[k] = b.seg_off[i * nt + u + 1];
                }
            }
        }
    }
    for (int pf = pa; pf < pe; pf += AE_PT) {
        if (pf > pa) __syncthreads();   // the previous tile has been consumed
        const int np = min(AE_PT, pe - pf);
        for (int pp = tid; pp < np; pp += 256) {
            const int p = pf + pp;
            const PBRec rec = pb_rec(PB, p, m.pbstride);
            const double dx = rec[0], dy = rec[1], dz = rec[2];
            const int tp = m.type_pairs[b.types[b.centre[p]] * nt + b.types[b.nbr[p]]];
            double* col = sm_ae + pp;
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            const double rinv = 1.0 / r;
            // f_n goes to the tile.  store_radial (A/B switch PM_EVAL_RC_RADS): 1 / r, f_n and f_n' also go to the pair
            // record and the pair pass reads these 21 doubles instead of evaluating the Gaussians a second time --
            // measured slower than recomputing them (4.40 vs 4.26 ms per 131 072 atoms), so it is off by default
            const PBRecW recw = pb_rec_w(PBw, p, m.pbstride);
            if (store_radial) recw[3] = rinv;
            pair_radial(m, r, tp, [&](int n, double fn, double fnd) {
                col[n * AE_LD] = fn;
                if (store_radial) { recw[4 + n] = fn; recw[4 + m.n_fn + n] = fnd; }
            });
            pair_angular<LT>(m, dx, dy, dz, r, rinv, [&](int key, double yr, double yi, double, double, double, double,
                                                         double, double) {
                col[(oy + 2 * key) * AE_LD] = yr;
                col[(oy + 2 * key + 1) * AE_LD] = yi;
            });
        }
        __syncthreads();
        if (RECT) {
            const int a = warp >> 1;
            if (a < na) {
                const int q0 = max(pf, b.seg_off[i0 + a]), q1 = min(pf + np, b.seg_off[i0 + a + 1]);
                for (int k0 = q0; k0 < q1; k0 += 4) {
                    const int p = k0 + (lane & 3);
                    const bool pv = p < q1;
                    const double* colp = sm_ae + (p - pf);
                    double av[2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) av[mt] = (pv && arow[mt] >= 0) ? colp[arow[mt] * AE_LD] : 0.0;
#pragma unroll
                    for (int q = 0; q < NTW; ++q) {
                        const double bv = (pv && brow[q] >= 0) ? colp[brow[q] * AE_LD] : 0.0;
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) dmma(acc[mt][q][0], acc[mt][q][1], av[mt], bv);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < AE_NI; ++k) {
                const int q0 = max(pf, it_p0[k]), q1 = min(pf + np, it_p1[k]);
                const double* rf = sm_ae + it_nid[k] * AE_LD - pf;
                const double* ry = sm_ae + (oy + 2 * it_key[k]) * AE_LD - pf;
                for (int p = q0; p < q1; ++p) {
                    const double fn = rf[p];
                    if (fn == 0.0) continue;
                    ar[k] += fn * ry[p];
                    ai[k] += fn * ry[p + AE_LD];
                }
            }
        }
    }
    if (RECT) {
        // accumulator (row, columns 2 c, 2 c + 1) = (Re, Im) of head (n = tile row, key = 4 * y tile + c)
        const int a = warp >> 1;
        if (a < na) {
            const DevType& T = m.types[0];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (ahead[mt] < 0) continue;
#pragma unroll
                for (int q = 0; q < NTW; ++q) {
                    const int key = ((warp & 1) * NTW + q) * 4 + (lane & 3);
                    if (key < m.nh) {
                        const int h = T.seg_heads[0][ahead[mt] + key];
                        anc[(size_t)(i0 + a) * m.hmax + h] = make_double2(acc[mt][q][0], acc[mt][q][1]);
                    }
                }
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < AE_NI; ++k) {
        const int w = tid + k * 256;
        const int a = w / m.hmax, h = w - a * m.hmax;
        if (a < na && h < m.types[b.types[i0 + a]].n_head) anc[(size_t)(i0 + a) * m.hmax + h] = make_double2(ar[k], ai[k]);
    }
}

bool launch_anlm_eval(const DevModel& m, const DevBatch& b, double* PB, double2* anc, cudaStream_t s, bool store_radial) {
    if (b.n_atoms == 0) return true;
    const bool rect = m.front2 && m.n_fn <= 16 && getenv("PM_EVAL_K2_NO_DMMA") == nullptr;
    const size_t smem = (size_t)(m.n_fn + 2 * m.nh) * (AE_PT + (rect ? 4 : 1)) * sizeof(double);
    if (getenv("PM_EVAL_K2_FIT") != nullptr || m.maxl > 6 || AE_AT * m.hmax > AE_NI * 256 || smem > 200 * 1024) return false;
    init_pair_basis_tables();
    const int grid = (b.n_atoms + AE_AT - 1) / AE_AT;
#define PM_AE_CASE(L_)                                                          \
    case L_:                                                                    \
        if (rect) {                                                             \
            ensure_smem((const void*)k_anlm_eval<L_, true>, smem);              \
            k_anlm_eval<L_, true><<<grid, 256, smem, s>>>(m, b, PB, anc, store_radial ? 1 : 0);  \
        } else {                                                                \
            ensure_smem((const void*)k_anlm_eval<L_, false>, smem);             \
            k_anlm_eval<L_, false><<<grid, 256, smem, s>>>(m, b, PB, anc, store_radial ? 1 : 0); \
        }                                                                       \
        break;
    switch (m.maxl) { PM_AE_CASE(0) PM_AE_CASE(1) PM_AE_CASE(2) PM_AE_CASE(3) PM_AE_CASE(4) PM_AE_CASE(5) PM_AE_CASE(6) }
#undef PM_AE_CASE
    return true;
}

// true if the fused kernel served the model (then neither launch_pair_basis nor launch_anlm must be called)
bool launch_pair_anlm(const DevModel& m, const DevBatch& b, double* PB, double2* anc, double2* agg, cudaStream_t s,
                      bool store_records) {
    if (b.n_atoms == 0) return true;
    const bool off = getenv("PM_K2_SPLIT") != nullptr;   // A/B switch: the two separate kernels
    const int threads = std::max(64, (m.hmax + 31) / 32 * 32);
    const size_t smem = (size_t)m.pbstride * PA_LD * sizeof(double);
    if (off || threads > 256 || smem > 100 * 1024 || m.maxl > 6) return false;
    init_pair_basis_tables();
#define PM_PA_CASE(L_)                                                                 \
    case L_:                                                                           \
        if (store_records) {                                                           \
            ensure_smem((const void*)k_pair_anlm<L_, true>, smem);                     \
            k_pair_anlm<L_, true><<<b.n_atoms, threads, smem, s>>>(m, b, PB, anc, agg);  \
        } else {                                                                       \
            ensure_smem((const void*)k_pair_anlm<L_, false>, smem);                    \
            k_pair_anlm<L_, false><<<b.n_atoms, threads, smem, s>>>(m, b, PB, anc, agg); \
        }                                                                              \
        break;
    switch (m.maxl) {
        PM_PA_CASE(0) PM_PA_CASE(1) PM_PA_CASE(2) PM_PA_CASE(3) PM_PA_CASE(4) PM_PA_CASE(5) PM_PA_CASE(6)
    }
#undef PM_PA_CASE
    return true;
}

void launch_anlm(const DevModel& m, const DevBatch& b, const double* PB, double2* anc, double2* agg, cudaStream_t s,
                 bool small_footprint) {
    if (b.n_atoms == 0) return;
    const int threads = (m.hmax + 31) / 32 * 32;
    const size_t smem16 = 2ull * (16 + 1) * m.pbstride * sizeof(double);
    const size_t smem8 = 2ull * (8 + 1) * m.pbstride * sizeof(double);
    if (threads <= 256 && small_footprint && smem8 <= 23 * 1024) {
        k_anlm_v2<8><<<b.n_atoms, threads, smem8, s>>>(m, b, PB, anc, agg);
        return;
    }
    if (threads <= 256 && smem16 <= 48 * 1024) {
        k_anlm_v2<16><<<b.n_atoms, threads, smem16, s>>>(m, b, PB, anc, agg);
        return;
    }
    if (threads > 256 && getenv("PM_ANLM_V1") == nullptr) {
        // more than 256 heads per atom: passes of 256 heads (blockIdx.y), records staged through shared memory as above
        // instead of the thread-per-head kernel that reads every record item of every pair from L2 (8.4 -> 7.1 ms per
        // 24 config-3 structures, 220 -> 232 structures/s)
        const size_t smem4 = 2ull * (4 + 1) * m.pbstride * sizeof(double);
        const dim3 grid(b.n_atoms, (m.hmax + 255) / 256);
        if (smem8 <= 110 * 1024) {
            ensure_smem((const void*)k_anlm_v2<8>, smem8);
            k_anlm_v2<8><<<grid, 256, smem8, s>>>(m, b, PB, anc, agg);
            return;
        }
        if (smem4 <= 110 * 1024) {
            ensure_smem((const void*)k_anlm_v2<4>, smem4);
            k_anlm_v2<4><<<grid, 256, smem4, s>>>(m, b, PB, anc, agg);
            return;
        }
    }
    k_anlm<<<b.n_atoms, 128, 0, s>>>(m, b, PB, anc, agg);
}

// ================================================================================================
// K3: linear invariants d_f = sum_t c_t Re(prod a) and G[f, head] = d d_f / d a_head written straight
// into the block-sparse layout consumed by K4a (one 4x8 block = one DMMA B fragment).
// ================================================================================================
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) k_features(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                   double* __restrict__ dfeat, double* __restrict__ Gbuf) {
    extern __shared__ double2 afull[];
    const int i = blockIdx.x;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const bool force = b.force[b.st_of_atom[i]] != 0;
    for (int k = threadIdx.x; k < T.n_full; k += blockDim.x) {
        double2 v = anc[(size_t)i * m.hmax + T.full_head[k]];
        if (T.full_conj[k]) {
            const double cc = T.full_cc[k];
            v = make_double2(cc * v.x, -cc * v.y);
        }
        afull[k] = v;
    }
    double* drow = dfeat + (size_t)i * m.fl;
    for (int k = threadIdx.x; k < m.fl; k += blockDim.x) drow[k] = 0.0;
    double* G = Gbuf + (size_t)i * m.gstride;
    if (force)
        for (long k = threadIdx.x; k < T.g_size; k += blockDim.x) G[k] = 0.0;
    __syncthreads();
    const int mo = T.max_order;
    for (int f = threadIdx.x; f < T.n_feat; f += blockDim.x) {
        double sum = 0.0;
        for (int ti = T.term_off[f]; ti < T.term_off[f + 1]; ++ti) {
            const int o = T.term_order[ti];
            const int* ids = T.term_ids + (size_t)ti * mo;
            double2 pr = afull[ids[0]];
            for (int k = 1; k < o; ++k) pr = cmul(pr, afull[ids[k]]);
            sum += T.term_coeff[ti] * pr.x;
        }
        drow[T.feat_pad[f]] = sum;
    }
    if (!force) return;
    for (int e = threadIdx.x; e < T.n_ent; e += blockDim.x) {
        double gr = 0.0, gi = 0.0;
        for (int c = T.ent_off[e]; c < T.ent_off[e + 1]; ++c) {
            const DevContribution& cb = T.contribs[c];
            double2 pr = make_double2(1.0, 0.0);
            for (int q = 0; q < cb.n_ids; ++q) pr = cmul(pr, afull[cb.ids[q]]);
            if (cb.conj) pr.y = -pr.y;
            gr += cb.coeff * pr.x;
            gi += cb.coeff * pr.y;
        }
        G[T.ent_pos_re[e]] = gr;
        G[T.ent_pos_im[e]] = -gi;
    }
}

// Several atoms per CTA: every table entry (term / contribution) is read once and applied to all atoms of
// the group that have the matching centre type, which divides the table traffic by the group size.
template <int AT, int MO>
__global__ void __launch_bounds__(256) k_features_v2(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                      double* __restrict__ dfeat, double* __restrict__ Gbuf,
                                                      int nfull_max) {
    extern __shared__ double2 afull[];   // [AT][nfull_max]
    const int i0 = blockIdx.x * AT;
    const int tid = threadIdx.x, nthr = blockDim.x;
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
        for (int k = tid; k < T.n_full; k += nthr) {
            double2 v = anc[(size_t)i * m.hmax + T.full_head[k]];
            if (T.full_conj[k]) {
                const double cc = T.full_cc[k];
                v = make_double2(cc * v.x, -cc * v.y);
            }
            afull[(size_t)a * nfull_max + k] = v;
        }
        double* drow = dfeat + (size_t)i * m.fl;
        for (int k = tid; k < m.fl; k += nthr) drow[k] = 0.0;
        if (fo[a]) {
            double* G = Gbuf + (size_t)i * m.gstride;
            for (long k = tid; k < T.g_size; k += nthr) G[k] = 0.0;
        }
    }
    __syncthreads();
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false, anyf = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) { any = any || ty[a] == tt; anyf = anyf || (ty[a] == tt && fo[a]); }
        if (!any) continue;
        const DevType& T = m.types[tt];
        const int mo = T.max_order;
        for (int f = tid; f < T.n_feat; f += nthr) {
            double sum[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) sum[a] = 0.0;
            for (int ti = T.term_off[f]; ti < T.term_off[f + 1]; ++ti) {
                const int o = T.term_order[ti];
                const int* ids = T.term_ids + (size_t)ti * mo;
                const double cf = T.term_coeff[ti];
                int id[MO];
#pragma unroll
                for (int k = 0; k < MO; ++k) id[k] = k < o ? ids[k] : 0;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = af[id[0]];
#pragma unroll
                    for (int k = 1; k < MO; ++k)
                        if (k < o) pr = cmul(pr, af[id[k]]);
                    sum[a] += cf * pr.x;
                }
            }
            const int fp = T.feat_pad[f];
#pragma unroll
            for (int a = 0; a < AT; ++a)
                if (ty[a] == tt) dfeat[(size_t)(i0 + a) * m.fl + fp] = sum[a];
        }
        if (!anyf) continue;
        for (int e = tid; e < T.n_ent; e += nthr) {
            double gr[AT], gi[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) { gr[a] = 0.0; gi[a] = 0.0; }
            for (int c = T.ent_off[e]; c < T.ent_off[e + 1]; ++c) {
                const DevContribution* cbp = T.contribs + c;
                const double ccf = cbp->coeff;
                const int cconj = cbp->conj, cn = cbp->n_ids;
                int cid[MO > 1 ? MO - 1 : 1];
#pragma unroll
                for (int qq = 0; qq < MO - 1; ++qq) cid[qq] = qq < cn ? cbp->ids[qq] : 0;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt || !fo[a]) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = make_double2(1.0, 0.0);
                    if (MO > 1 && cn > 0) pr = af[cid[0]];
#pragma unroll
                    for (int qq = 1; qq < MO - 1; ++qq)
                        if (qq < cn) pr = cmul(pr, af[cid[qq]]);
                    if (cconj) pr.y = -pr.y;
                    gr[a] += ccf * pr.x;
                    gi[a] += ccf * pr.y;
                }
            }
            const int pr_ = T.ent_pos_re[e], pi_ = T.ent_pos_im[e];
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                if (ty[a] != tt || !fo[a]) continue;
                double* G = Gbuf + (size_t)(i0 + a) * m.gstride;
                G[pr_] = gr[a];
                G[pi_] = -gi[a];
            }
        }
    }
}

// v3: same work split as v2 (AT atoms per CTA share every table read) but the term / contribution tables are
// read through the sliced copies (DevType::fsl_* / esl_*): every warp iteration loads 32 consecutive slots
// (coefficient + packed 16-bit ids), so table reads are coalesced and independent of the accumulation chain.
template <int AT, int MO, int NT>
__global__ void __launch_bounds__(NT) k_features_v3(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                      double* __restrict__ dfeat, double* __restrict__ Gbuf,
                                                      int nfull_max, int zero_g, double* __restrict__ dpv) {
    extern __shared__ double2 afull[];   // [AT][nfull_max]
    constexpr int NW = (MO + 1) / 2;
    constexpr int UN = NT > 256 ? 4 : 2;   // table slots in flight per lane (4 for the small-model CTAs too: measured no gain)
    const int i0 = blockIdx.x * AT;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
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
        for (int k = tid; k < T.n_full; k += nthr) {
            double2 v = anc[(size_t)i * m.hmax + T.full_head[k]];
            if (T.full_conj[k]) {
                const double cc = T.full_cc[k];
                v = make_double2(cc * v.x, -cc * v.y);
            }
            afull[(size_t)a * nfull_max + k] = v;
        }
        double* drow = dfeat + (size_t)i * m.fl;
        for (int k = tid; k < m.fl; k += nthr) drow[k] = 0.0;
        if (fo[a] && zero_g) {   // single-type models: the structural zeros of G are written once per allocation (host)
            double2* G = reinterpret_cast<double2*>(Gbuf + (size_t)i * m.gstride);
            for (long k = tid; k < T.g_size / 2; k += nthr) G[k] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false, anyf = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) { any = any || ty[a] == tt; anyf = anyf || (ty[a] == tt && fo[a]); }
        if (!any) continue;
        const DevType& T = m.types[tt];
        const double* __restrict__ coef = T.sl_coeff;
        const unsigned* __restrict__ ids = T.sl_ids;
        const long ns = T.n_slots;
        for (int s = warp; s < T.n_fsl; s += nwarp) {
            const int4 meta = T.fsl_meta[s];
            const int o = meta.z;
            double sum[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) sum[a] = 0.0;
#pragma unroll UN
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                int id[2 * NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const unsigned v = ids[w * ns + slot];
                    id[2 * w] = v & 0xffffu;
                    id[2 * w + 1] = v >> 16;
                }
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = af[id[0]];
#pragma unroll
                    for (int k = 1; k < MO; ++k)
                        if (k < o) pr = cmul(pr, af[id[k]]);
                    sum[a] += cf * pr.x;
                }
            }
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], 8);
                sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], 16);
            }
            if (lane < 8) {
                const int fp = T.fsl_out[s * 8 + lane];
                if (fp >= 0) {
                    const int pv = dpv ? T.pad_pv[fp] : -1;
#pragma unroll
                    for (int a = 0; a < AT; ++a)
                        if (ty[a] == tt) {
                            dfeat[(size_t)(i0 + a) * m.fl + fp] = sum[a];
                            if (pv >= 0) dpv[(size_t)(i0 + a) * 64 + pv] = sum[a];
                        }
                }
            }
        }
        if (!anyf) continue;
        for (int s = warp; s < T.n_esl; s += nwarp) {
            const int4 meta = T.esl_meta[s];
            const int cn = meta.z;
            double gr[AT], gi[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) { gr[a] = 0.0; gi[a] = 0.0; }
#pragma unroll UN
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                int id[2 * NW];
                unsigned w0 = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const unsigned v = ids[w * ns + slot];
                    if (w == 0) w0 = v;
                    id[2 * w] = v & 0xffffu;
                    id[2 * w + 1] = (v >> 16) & 0x7fffu;
                }
                const double cfi = (w0 & 0x80000000u) ? -cf : cf;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt || !fo[a]) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = make_double2(1.0, 0.0);
                    if (MO > 1 && cn > 0) pr = af[id[0]];
#pragma unroll
                    for (int qq = 1; qq < MO - 1; ++qq)
                        if (qq < cn) pr = cmul(pr, af[id[qq]]);
                    gr[a] += cf * pr.x;
                    gi[a] += cfi * pr.y;
                }
            }
            const int2 pos = T.esl_out[s * 32 + lane];
            if (pos.x >= 0) {
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt || !fo[a]) continue;
                    double* G = Gbuf + (size_t)(i0 + a) * m.gstride;
                    G[pos.x] = gr[a];
                    G[pos.y] = -gi[a];
                }
            }
        }
    }
}


// compact polynomial-variable rows from dfeat (only after the fallback feature kernels; v3 writes them itself)
__global__ void __launch_bounds__(256) k_dpv_gather(DevModel m, int n_atoms, const double* __restrict__ dfeat,
                                                     double* __restrict__ dpv) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_atoms * 64) return;
    const int atom = (int)(idx >> 6), a = (int)(idx & 63);
    const int fa = a < m.npv_pad ? m.pv_fp[a] : -1;
    dpv[idx] = fa >= 0 ? dfeat[(size_t)atom * m.fl + fa] : 0.0;
}

void launch_features(const DevModel& m, const DevBatch& b, const double2* anc, double* dfeat, double* Gbuf,
                     size_t smem_bytes, cudaStream_t s, bool zero_g, double* dpv) {
    if (b.n_atoms == 0) return;
    struct DpvFallback {   // runs when a fallback kernel (which does not know dpv) was used
        const DevModel& m; const DevBatch& b; const double* dfeat; double* dpv; cudaStream_t s; bool armed = true;
        ~DpvFallback() {
            if (armed && dpv) k_dpv_gather<<<(int)(((long)b.n_atoms * 64 + 255) / 256), 256, 0, s>>>(m, b.n_atoms, dfeat, dpv);
        }
    } dpv_fallback{m, b, dfeat, dpv, s};
    // large radial-replication models: the radial-batched kernel (pm_kernels_feat.cu).  (On config 2 it measured slower,
    // 11.3 - 14.3 vs 10.7 us per structure: the 128 radial-0 entries pack badly into 32-row slices and there are too few
    // work items per CTA -- it declines small models.)
    if (launch_features_radial(m, b, anc, dfeat, Gbuf, smem_bytes, s, zero_g, dpv)) {
        dpv_fallback.armed = false;
        return;
    }
    // smem_bytes = bytes of one atom's full a_nlm array
    const int nfull_max = (int)(smem_bytes / sizeof(double2));
    constexpr int AT = 4;
    const size_t smem4 = smem_bytes * AT;
    int mo = 1;
    for (int t = 0; t < m.n_type; ++t) mo = max(mo, m.types[t].max_order);
    bool sliced = mo <= 6;
    for (int t = 0; t < m.n_type; ++t) sliced = sliced && m.types[t].n_fsl > 0;
    if (sliced) {
        // atoms per CTA (every table read is shared by the CTA's atoms).  Small models: 4 atoms, 256 threads, several
        // CTAs per SM; big a_nlm arrays (max_l ~ 12): 512-thread CTAs with 4, 2 or 1 atoms, two per SM when they fit
        const size_t cap = 226 * 1024;
        // measured (config 3 / 4): two 512-thread CTAs per SM beat one CTA with twice the atoms (latency-bound gathers)
        const size_t half = 113 * 1024;
        int at = smem4 <= half ? 4 : (2 * smem_bytes <= half ? 2 : (smem_bytes <= half ? 1 : (2 * smem_bytes <= cap ? 2 : 1)));
        if (smem4 > 96 * 1024 && getenv("PM_FEAT_AT")) {
            const int want = atoi(getenv("PM_FEAT_AT"));
            if ((want == 1 || want == 2 || want == 4) && want * smem_bytes <= cap) at = want;
        }
        const bool big = smem4 > 96 * 1024;
        const size_t smem = smem_bytes * at;
        const int grid = (b.n_atoms + at - 1) / at;
#define PM_FEAT3_LAUNCH(AT_, MO_, NT_)                                                                              \
    {                                                                                                              \
        ensure_smem((const void*)k_features_v3<AT_, MO_, NT_>, smem);                                              \
        k_features_v3<AT_, MO_, NT_><<<grid, NT_, smem, s>>>(m, b, anc, dfeat, Gbuf, nfull_max, zero_g ? 1 : 0, dpv); \
    }
#define PM_FEAT3_CASE(MO_)                                                                                          \
    case MO_:                                                                                                      \
        if (!big) PM_FEAT3_LAUNCH(4, MO_, 256)                                                                     \
        else if (at == 4) PM_FEAT3_LAUNCH(4, MO_, 512)                                                             \
        else if (at == 2) PM_FEAT3_LAUNCH(2, MO_, 512)                                                             \
        else PM_FEAT3_LAUNCH(1, MO_, 512)                                                                          \
        break;
        switch (mo) {
            PM_FEAT3_CASE(1) PM_FEAT3_CASE(2) PM_FEAT3_CASE(3) PM_FEAT3_CASE(4) PM_FEAT3_CASE(5) PM_FEAT3_CASE(6)
        }
#undef PM_FEAT3_CASE
#undef PM_FEAT3_LAUNCH
        dpv_fallback.armed = false;
        return;
    }
    if (smem4 <= 96 * 1024 && mo <= 6) {
        const int grid = (b.n_atoms + AT - 1) / AT;
#define PM_FEAT_CASE(MO_)                                                                                        \
    case MO_:                                                                                                    \
        ensure_smem((const void*)k_features_v2<AT, MO_>, smem4);                                                 \
        k_features_v2<AT, MO_><<<grid, 256, smem4, s>>>(m, b, anc, dfeat, Gbuf, nfull_max);                      \
        break;
        switch (mo) {
            PM_FEAT_CASE(1) PM_FEAT_CASE(2) PM_FEAT_CASE(3) PM_FEAT_CASE(4) PM_FEAT_CASE(5) PM_FEAT_CASE(6)
        }
#undef PM_FEAT_CASE
        return;
    }
    ensure_smem((const void*)k_features, smem_bytes);
    k_features<<<b.n_atoms, 256, smem_bytes, s>>>(m, b, anc, dfeat, Gbuf);
}

void set_features_smem(size_t smem_bytes) { ensure_smem((const void*)k_features, smem_bytes); }

// ================================================================================================
// K4a (straightforward): L[(pair, alpha), f] = sum_heads Re(G[f, head] v_alpha,head(pair)), plus the
// aggregated own/virial rows from K2b.  One CTA per centre atom, one thread per output element.
// ================================================================================================
__global__ void __launch_bounds__(256) k_lrows_simple(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                       const double2* __restrict__ agg, const double* __restrict__ Gbuf,
                                                       double* __restrict__ Lbuf, double* __restrict__ Xown,
                                                       double* __restrict__ Sbuf) {
    const int i = blockIdx.x;
    if (!b.force[b.st_of_atom[i]]) return;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const int nt = m.n_type;
    const double* G = Gbuf + (size_t)i * m.gstride;
    const int pb = b.seg_off[i * nt], pe = b.seg_off[i * nt + nt];
    const int np = pe - pb;
    const int nrows = 3 * np + 9;
    const int nf = T.n_fpad;
    const int oy = pb_y(m, 0);
    for (int idx = threadIdx.x; idx < nrows * nf; idx += blockDim.x) {
        const int row = idx / nf, fp = idx - row * nf;
        const int tile = fp >> 3, nn = fp & 7;
        double acc = 0.0;
        if (row < 3 * np) {
            const int pl = row / 3, al = row - 3 * pl;
            const int p = pb + pl;
            const int u = b.types[b.nbr[p]];
            const PBRec rec = pb_rec(PB, p, m.pbstride);
            const double dal = rec[al] * rec[3];
            const int oya = pb_y(m, 1 + al);
            const int* sh = T.seg_heads[u];
            for (int bk = T.tile_blk_off[u][tile]; bk < T.tile_blk_off[u][tile + 1]; ++bk) {
                const int kc = T.blk_kchunk[bk];
                const double* B = G + 32 * (size_t)bk + nn * 4;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int h = sh[2 * kc + hh];
                    if (h < 0) continue;
                    const int nid = T.head_nid[h], key = T.head_key[h];
                    const double fn = rec[4 + nid];
                    if (fn == 0.0) continue;
                    const double d1 = rec[4 + m.n_fn + nid] * dal;
                    const double vr = d1 * rec[oy + 2 * key] + fn * rec[oya + 2 * key];
                    const double vi = d1 * rec[oy + 2 * key + 1] + fn * rec[oya + 2 * key + 1];
                    acc += vr * B[2 * hh] + vi * B[2 * hh + 1];
                }
            }
            Lbuf[((size_t)p * 3 + al) * m.fl + fp] = acc;
        } else {
            const int r = row - 3 * np;
            for (int u = 0; u < nt; ++u) {
                const int* sh = T.seg_heads[u];
                for (int bk = T.tile_blk_off[u][tile]; bk < T.tile_blk_off[u][tile + 1]; ++bk) {
                    const int kc = T.blk_kchunk[bk];
                    const double* B = G + 32 * (size_t)bk + nn * 4;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int h = sh[2 * kc + hh];
                        if (h < 0) continue;
                        const double2 v = agg[((size_t)i * m.hmax + h) * 9 + r];
                        acc += v.x * B[2 * hh] + v.y * B[2 * hh + 1];
                    }
                }
            }
            if (r < 3) Xown[((size_t)i * 3 + r) * m.fl + fp] = acc;
            else Sbuf[((size_t)i * 6 + (r - 3)) * m.fl + fp] = acc;
        }
    }
}

// ================================================================================================
// K4b (straightforward): gather over the centres that touch a row, expand the polynomial, apply the
// row weight, write X-tilde = [w X | w y | 0].
//   force row (k, alpha): centres = k itself (Lambda = own row) and every neighbour c of k
//   (Lambda = -L_c[(c -> k), alpha], found through the reverse-pair index).
// ================================================================================================
__device__ __forceinline__ double poly_contrib(const DevPolyTerm& tm, const double* __restrict__ d,
                                               const double* __restrict__ L) {
    if (tm.order == 1) return L[tm.fp0];
    if (tm.order == 2) return d[tm.fp1] * L[tm.fp0] + d[tm.fp0] * L[tm.fp1];
    const double d0 = d[tm.fp0], d1 = d[tm.fp1], d2 = d[tm.fp2];
    return d1 * d2 * L[tm.fp0] + d0 * d2 * L[tm.fp1] + d0 * d1 * L[tm.fp2];
}

__device__ __forceinline__ double poly_value(const DevPolyTerm& tm, const double* __restrict__ d) {
    if (tm.order == 1) return d[tm.fp0];
    if (tm.order == 2) return d[tm.fp0] * d[tm.fp1];
    return d[tm.fp0] * d[tm.fp1] * d[tm.fp2];
}

__global__ void __launch_bounds__(256) k_xrows_force_simple(DevModel m, DevBatch b, const double* __restrict__ dfeat,
                                                             const double* __restrict__ Lbuf,
                                                             const double* __restrict__ Xown, double* __restrict__ X,
                                                             int apply_w) {
    const int k = blockIdx.x, al = blockIdx.y;
    const int s = b.st_of_atom[k];
    if (!b.force[s]) return;
    const int row = b.frow[s] + 3 * (k - b.atom_off[s]) + al;
    const double w = apply_w ? b.w[row] : 1.0;
    const int nt = m.n_type;
    const int tk = b.types[k];
    const int p0 = b.seg_off[k * nt], p1 = b.seg_off[k * nt + nt];
    double* xr = X + (size_t)row * m.fpad;
    for (int col = threadIdx.x; col < m.n_variables; col += blockDim.x) {
        double val = 0.0;
        {
            const DevPolyTerm tm = m.types[tk].colterm[col];
            if (tm.order) val += poly_contrib(tm, dfeat + (size_t)k * m.fl, Xown + ((size_t)k * 3 + al) * m.fl);
        }
        for (int p = p0; p < p1; ++p) {
            const int c = b.nbr[p];
            const DevPolyTerm tm = m.types[b.types[c]].colterm[col];
            if (tm.order) val -= poly_contrib(tm, dfeat + (size_t)c * m.fl, Lbuf + ((size_t)b.rev[p] * 3 + al) * m.fl);
        }
        xr[col] = w * val;
    }
    if (threadIdx.x == 0) xr[m.n_variables] = apply_w ? b.yv[row] : 0.0;
}

// energy row (r = 0) and the six virial rows (r = 1..6) of each structure
__global__ void __launch_bounds__(256) k_xrows_struct_simple(DevModel m, DevBatch b, const double* __restrict__ dfeat,
                                                              const double* __restrict__ Sbuf, double* __restrict__ X,
                                                              double* __restrict__ xe_sum, double* __restrict__ xe_sq,
                                                              int apply_w) {
    const int s = blockIdx.x, r = blockIdx.y;
    if (r > 0 && !b.force[s]) return;
    const int row = r == 0 ? b.erow[s] : b.srow[s] + r - 1;
    const double w = apply_w ? b.w[row] : 1.0;
    const int a0 = b.atom_off[s], a1 = b.atom_off[s + 1];
    double* xr = X + (size_t)row * m.fpad;
    for (int col = threadIdx.x; col < m.n_variables; col += blockDim.x) {
        double val = 0.0;
        for (int a = a0; a < a1; ++a) {
            const DevPolyTerm tm = m.types[b.types[a]].colterm[col];
            if (!tm.order) continue;
            if (r == 0) val += poly_value(tm, dfeat + (size_t)a * m.fl);
            else val += poly_contrib(tm, dfeat + (size_t)a * m.fl, Sbuf + ((size_t)a * 6 + (r - 1)) * m.fl);
        }
        if (r == 0 && xe_sum) {
            atomicAdd(xe_sum + col, val);
            atomicAdd(xe_sq + col, val * val);
        }
        xr[col] = w * val;
    }
    if (threadIdx.x == 0) xr[m.n_variables] = apply_w ? b.yv[row] : 0.0;
}

// ================================================================================================
// K5 (straightforward): C[i,j] += sum_r Xt[r,i] Xt[r,j] for upper 64x64 tiles, DFMA, 4x4 per thread.
// ================================================================================================
__global__ void __launch_bounds__(256) k_syrk_simple(const double* __restrict__ X, int n_rows, int fpad,
                                                      double* __restrict__ C) {
    const int nt = fpad / 64;
    // map blockIdx.x -> (ti <= tj)
    int ti = 0, rem = blockIdx.x;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;
    __shared__ double sa[16][64 + 4];
    __shared__ double sb[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
    for (int r0 = 0; r0 < n_rows; r0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int rr = e >> 6, cc = e & 63;
            const int r = r0 + rr;
            sa[rr][cc] = r < n_rows ? X[(size_t)r * fpad + ti * 64 + cc] : 0.0;
            sb[rr][cc] = r < n_rows ? X[(size_t)r * fpad + tj * 64 + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { av[a] = sa[rr][ty * 4 + a]; bv[a] = sb[rr][tx * 4 + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] += av[a] * bv[c];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            C[(size_t)(ti * 64 + ty * 4 + a) * fpad + tj * 64 + tx * 4 + c] += acc[a][c];
}

// ================================================================================================
// eval: E/F/S with trained coefficients (PolymlpEval::eval_gtinv, compute/polymlp_eval.cpp:182-335).
// Adjoint form: per atom E_i and w_f = dE_i/dd_f from the polynomial, then the head adjoint
// Ah[k] = sum_f G[k, f] w_f, then one pass over the pairs: f_alpha(pair) = sum_k V[(pair,alpha),k] Ah[k].
// ================================================================================================
__global__ void __launch_bounds__(256) k_eval_atom(DevModel m, DevBatch b, const double* __restrict__ dfeat,
                                                    const double* __restrict__ Gbuf, const double* __restrict__ coeffs,
                                                    double* __restrict__ wbuf, double* __restrict__ Ah,
                                                    int ah_stride, double* __restrict__ energies) {
    // wbuf: [n_atoms][fl] scratch (reuses Xown); Ah: [n_atoms][ah_stride] with per-segment offsets u * (ah_stride / nt)
    const int i = blockIdx.x;
    const int t = b.types[i];
    const DevType& T = m.types[t];
    const double* d = dfeat + (size_t)i * m.fl;
    double* w = wbuf + (size_t)i * m.fl;
    for (int k = threadIdx.x; k < m.fl; k += blockDim.x) w[k] = 0.0;
    __syncthreads();
    double e = 0.0;
    for (int col = threadIdx.x; col < m.n_variables; col += blockDim.x) {
        const DevPolyTerm tm = T.colterm[col];
        if (!tm.order) continue;
        const double c = coeffs[col];
        if (tm.order == 1) {
            e += c * d[tm.fp0];
            atomicAdd(w + tm.fp0, c);
        } else if (tm.order == 2) {
            e += c * d[tm.fp0] * d[tm.fp1];
            atomicAdd(w + tm.fp0, c * d[tm.fp1]);
            atomicAdd(w + tm.fp1, c * d[tm.fp0]);
        } else {
            const double d0 = d[tm.fp0], d1 = d[tm.fp1], d2 = d[tm.fp2];
            e += c * d0 * d1 * d2;
            atomicAdd(w + tm.fp0, c * d1 * d2);
            atomicAdd(w + tm.fp1, c * d0 * d2);
            atomicAdd(w + tm.fp2, c * d0 * d1);
        }
    }
    // block reduce e
    __shared__ double red[256];
    red[threadIdx.x] = e;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if (threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(energies + b.st_of_atom[i], red[0]);
    // head adjoint
    const int segstride = ah_stride / m.n_type;
    double* ah = Ah + (size_t)i * ah_stride;
    for (int k = threadIdx.x; k < ah_stride; k += blockDim.x) ah[k] = 0.0;
    __syncthreads();
    const double* G = Gbuf + (size_t)i * m.gstride;
    // one thread per (block, k in 0..3)
    for (int idx = threadIdx.x; idx < T.n_blocks * 4; idx += blockDim.x) {
        const int bk = idx >> 2, k = idx & 3;
        // find segment and tile of the block: stored implicitly through tile_blk_off; search segments
        int u = 0;
        while (u + 1 < m.n_type && bk >= T.tile_blk_off[u + 1][0]) ++u;
        // binary search tile
        int lo = 0, hi = T.n_tiles;
        const int* off = T.tile_blk_off[u];
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= bk) lo = mid; else hi = mid; }
        const int tile = lo;
        double sum = 0.0;
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) sum += G[32 * (size_t)bk + nn * 4 + k] * w[tile * 8 + nn];
        atomicAdd(ah + u * segstride + 4 * T.blk_kchunk[bk] + k, sum);
    }
}

__global__ void __launch_bounds__(128) k_eval_pairs(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                     const double* __restrict__ Ah, int ah_stride,
                                                     double* __restrict__ forces, double* __restrict__ stresses) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b.n_pairs * 3) return;
    const int p = idx / 3, al = idx - 3 * p;
    const int i = b.centre[p], j = b.nbr[p];
    const DevType& T = m.types[b.types[i]];
    const int u = b.types[j];
    const int segstride = ah_stride / m.n_type;
    const double* ah = Ah + (size_t)i * ah_stride + u * segstride;
    const PBRec rec = pb_rec(PB, p, m.pbstride);
    const double dal = rec[al] * rec[3];
    const int oy = pb_y(m, 0), oya = pb_y(m, 1 + al);
    const int* sh = T.seg_heads[u];
    double g = 0.0;
    for (int q = 0; q < T.seg_len[u]; ++q) {
        const int h = sh[q];
        if (h < 0) continue;
        const int nid = T.head_nid[h], key = T.head_key[h];
        const double fn = rec[4 + nid];
        if (fn == 0.0) continue;
        const double d1 = rec[4 + m.n_fn + nid] * dal;
        const double vr = d1 * rec[oy + 2 * key] + fn * rec[oya + 2 * key];
        const double vi = d1 * rec[oy + 2 * key + 1] + fn * rec[oya + 2 * key + 1];
        g += vr * ah[2 * q] + vi * ah[2 * q + 1];
    }
    // X force rows: +g on the centre, -g on the neighbour; virial rows: -g_alpha * D_beta
    atomicAdd(forces + (size_t)i * 3 + al, g);
    atomicAdd(forces + (size_t)j * 3 + al, -g);
    const int s = b.st_of_atom[i];
    double* S = stresses + (size_t)s * 6;
    // (alpha,beta) pairs: xx(0) yy(1) zz(2) xy(3) yz(4) zx(5)
    atomicAdd(S + al, -g * rec[al]);
    const int be = (al + 1) % 3;
    atomicAdd(S + 3 + al, -g * rec[be]);
}

// One thread per pair (consecutive pairs are contiguous per item in the blocked pair-basis layout, so every
// load is coalesced); Y_lm and its gradient are read once per lm and reused for all radial indices; the three
// Cartesian components are accumulated together.  Virial sums are reduced over the warp before the atomics.
constexpr int EV_MAXFN = 16;

constexpr int EP_NC = 6;   // centre atoms whose head adjoints a CTA of k_eval_pairs_v2 keeps in shared memory

// The head adjoints Ah of the centres this CTA's 128 consecutive pairs belong to (2-4 centres at ~54 neighbours each) are
// staged in shared memory first: ncu showed 72 % long-scoreboard stalls on the 300 dependent-latency global loads of
// Ah per thread; a pair whose centre lies beyond the staged window reads global memory as before.
__global__ void __launch_bounds__(128) k_eval_pairs_v2(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                        const double* __restrict__ Ah, int ah_stride,
                                                        double* __restrict__ forces, double* __restrict__ stresses, int nc_max) {
    extern __shared__ __align__(16) double s_ah[];   // [nc_max][ah_stride]
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = p < b.n_pairs;
    const int c_first = b.centre[min(blockIdx.x * blockDim.x, b.n_pairs - 1)];
    const int c_last = b.centre[min(blockIdx.x * blockDim.x + (int)blockDim.x - 1, b.n_pairs - 1)];
    const int n_stage = min(nc_max, c_last - c_first + 1);
    for (int e = threadIdx.x; e < n_stage * ah_stride; e += blockDim.x)
        s_ah[e] = Ah[(size_t)c_first * ah_stride + e];
    __syncthreads();
    double g[3] = {0.0, 0.0, 0.0};
    double dl[3] = {0.0, 0.0, 0.0};
    int i = 0, j = 0, s = -1;
    if (act) {
        i = b.centre[p];
        j = b.nbr[p];
        s = b.st_of_atom[i];
        const DevType& T = m.types[b.types[i]];
        const int u = b.types[j];
        const int segstride = ah_stride / m.n_type;
        const double* ah = (i - c_first < n_stage ? s_ah + (size_t)(i - c_first) * ah_stride : Ah + (size_t)i * ah_stride) + u * segstride;
        const PBRec rec = pb_rec(PB, p, m.pbstride);
        const double rinv = rec[3];
        dl[0] = rec[0]; dl[1] = rec[1]; dl[2] = rec[2];
        const int* snid = T.seg_nid[u];
        const int* snoff = T.seg_n_off[u];
        double fn[EV_MAXFN], fd[EV_MAXFN];
#pragma unroll
        for (int n = 0; n < EV_MAXFN; ++n) {
            fn[n] = 0.0; fd[n] = 0.0;
            if (n < m.n_fn) {
                const int nid = snid[n];
                if (nid >= 0) { fn[n] = rec[4 + nid]; fd[n] = rec[4 + m.n_fn + nid] * rinv; }
            }
        }
        const int oy = pb_y(m, 0), oyx = pb_y(m, 1), oyy = pb_y(m, 2), oyz = pb_y(m, 3);
        for (int key = 0; key < m.nh; ++key) {
            const double yr = rec[oy + 2 * key], yi = rec[oy + 2 * key + 1];
            const double xr = rec[oyx + 2 * key], xi = rec[oyx + 2 * key + 1];
            const double wr = rec[oyy + 2 * key], wi = rec[oyy + 2 * key + 1];
            const double zr = rec[oyz + 2 * key], zi = rec[oyz + 2 * key + 1];
            // sum over radial indices of Ah[(n, lm)] weighted by f_n'/r and f_n
            double s1r = 0.0, s1i = 0.0, s2r = 0.0, s2i = 0.0;
#pragma unroll
            for (int n = 0; n < EV_MAXFN; ++n) {
                if (n < m.n_fn && snoff[n + 1] > snoff[n]) {
                    const int q = snoff[n] + key;   // every radial group lists all lm keys in order
                    const double ar = ah[2 * q], ai = ah[2 * q + 1];
                    s1r += fd[n] * ar; s1i += fd[n] * ai;
                    s2r += fn[n] * ar; s2i += fn[n] * ai;
                }
            }
            // g_alpha += Re-part contraction: (f' Y D_alpha / r + f dY_alpha) . Ah
            const double t1 = yr * s1r + yi * s1i;
            g[0] += t1 * dl[0] + xr * s2r + xi * s2i;
            g[1] += t1 * dl[1] + wr * s2r + wi * s2i;
            g[2] += t1 * dl[2] + zr * s2r + zi * s2i;
        }
#pragma unroll
        for (int al = 0; al < 3; ++al) {
            atomicAdd(forces + (size_t)i * 3 + al, g[al]);
            atomicAdd(forces + (size_t)j * 3 + al, -g[al]);
        }
    }
    // virial: xx yy zz xy yz zx = -g_alpha * D_beta, reduced over the warp when all lanes share the structure
    double sv[6] = {-g[0] * dl[0], -g[1] * dl[1], -g[2] * dl[2], -g[0] * dl[1], -g[1] * dl[2], -g[2] * dl[0]};
    const int s0 = __shfl_sync(0xffffffffu, s, 0);
    const bool uniform = __all_sync(0xffffffffu, s == s0 || s < 0);
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sv[k] += __shfl_xor_sync(0xffffffffu, sv[k], d);
        }
        const int sl = __reduce_max_sync(0xffffffffu, s);
        if ((threadIdx.x & 31) == 0 && sl >= 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(stresses + (size_t)sl * 6 + k, sv[k]);
        }
    } else if (act) {
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(stresses + (size_t)s * 6 + k, sv[k]);
    }
}


// Pair pass that RECOMPUTES the basis record of its pair from the displacement (3 doubles) instead of reading the 1.2 KB
// record k_pair_anlm would have stored: per chunk of 41 x 512 atoms that removes a 1.4 GB write and a 1.4 GB read.  Same
// contraction as k_eval_pairs_v2, key by key as pair_angular produces them.  The head adjoints of the CTA's centres arrive
// by one TMA bulk copy that overlaps the radial part; when every centre of the CTA is staged (the normal case) they are
// read with ld.shared at compile-time offsets, otherwise from global memory.  NF = radial functions per segment, padded
// (absent ones carry f = f' = 0 and point at group 0).  Used when the fused K2 kernel ran without storing
// (single decision in pm_capi.cu: eval_pairs_rc_supported).
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int LT, int NF, bool RADS>
__global__ void __launch_bounds__(128, 4) k_eval_pairs_rc(DevModel m, DevBatch b, const double* __restrict__ PB,
                                                           const double* __restrict__ Ah, int ah_stride,
                                                           double* __restrict__ forces, double* __restrict__ stresses, int nc_max) {
    extern __shared__ __align__(16) double s_ah[];   // [nc_max][ah_stride] | mbarrier
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = p < b.n_pairs;
    const int c_first = b.centre[min(blockIdx.x * blockDim.x, b.n_pairs - 1)];
    const int c_last = b.centre[min(blockIdx.x * blockDim.x + (int)blockDim.x - 1, b.n_pairs - 1)];
    const bool staged = c_last - c_first + 1 <= nc_max;   // CTA-uniform
    const unsigned bar = smem_u32(s_ah + (size_t)nc_max * ah_stride);
    if (staged && threadIdx.x == 0) {
        const unsigned bytes = (unsigned)((c_last - c_first + 1) * ah_stride * sizeof(double));
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(s_ah), Ah + (size_t)c_first * ah_stride, bytes, bar);
    }
    __syncthreads();   // the barrier is initialised before anyone polls it
    double g[3] = {0.0, 0.0, 0.0};
    double dl[3] = {0.0, 0.0, 0.0};
    int i = 0, j = 0, s = -1;
    if (act) {
        i = b.centre[p];
        j = b.nbr[p];
        s = b.st_of_atom[i];
        const int ti = b.types[i], u = b.types[j];
        const DevType& T = m.types[ti];
        const int segstride = ah_stride / m.n_type;
        const PBRec rec = pb_rec(PB, p, m.pbstride);
        dl[0] = rec[0]; dl[1] = rec[1]; dl[2] = rec[2];
        const double r = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
        const double rinv = 1.0 / r;
        const int* snid = T.seg_nid[u];
        const int* snoff = T.seg_n_off[u];
        // f_n and f_n' / r of the radial functions of this (centre type, neighbour type) segment, in segment order, and the
        // element offset of each radial group in the adjoint row.  RADS: K2 left f_n, f_n' in the pair record (k_anlm_eval);
        // else they are evaluated here.
        double fr[NF], fd[NF];
        int qo[NF];
        {
            double fc = 0.0, fcd = 0.0;
            if (!RADS) radial_cutoff(m, r, fc, fcd);
            const int tp = m.type_pairs[ti * m.n_type + u];
            const int nfn = m.tp_nfn[tp];
            const double* prm = m.tp_params + (size_t)tp * m.n_fn * 2;
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                fr[n] = 0.0; fd[n] = 0.0; qo[n] = 0;
                if (n < m.n_fn) {
                    const int nid = snid[n];
                    const int q0 = snoff[n];
                    if (nid >= 0 && snoff[n + 1] > q0) {
                        if (RADS) {
                            fr[n] = rec[4 + nid];
                            fd[n] = rec[4 + m.n_fn + nid] * rinv;
                        } else {
                            double fnd;
                            radial_one(prm, nfn, nid, r, fc, fcd, fr[n], fnd);
                            fd[n] = fnd * rinv;
                        }
                        qo[n] = 2 * q0;
                    }
                }
            }
        }
        auto contract = [&](auto load) {
            pair_angular<LT>(m, dl[0], dl[1], dl[2], r, rinv, [&](int key, double yr, double yi, double xr, double xi,
                                                                  double wr, double wi, double zr, double zi) {
                double s1r = 0.0, s1i = 0.0, s2r = 0.0, s2i = 0.0;
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const double2 a = load(n, key);   // every radial group lists all lm keys in order
                    s1r += fd[n] * a.x; s1i += fd[n] * a.y;
                    s2r += fr[n] * a.x; s2i += fr[n] * a.y;
                }
                // g_alpha += Re-part contraction: (f' Y D_alpha / r + f dY_alpha) . Ah
                const double t1 = yr * s1r + yi * s1i;
                g[0] += t1 * dl[0] + xr * s2r + xi * s2i;
                g[1] += t1 * dl[1] + wr * s2r + wi * s2i;
                g[2] += t1 * dl[2] + zr * s2r + zi * s2i;
            });
        };
        if (staged) {
            const unsigned base = smem_u32(s_ah) + (unsigned)(((i - c_first) * ah_stride + u * segstride) * sizeof(double));
            unsigned qa[NF];
#pragma unroll
            for (int n = 0; n < NF; ++n) qa[n] = base + qo[n] * (unsigned)sizeof(double);
            mbar_wait(bar, 0);
            contract([&](int n, int key) { return lds_f64x2(qa[n] + 16u * key); });
        } else {
            const double* ah = Ah + (size_t)i * ah_stride + u * segstride;
            contract([&](int n, int key) { return *reinterpret_cast<const double2*>(ah + qo[n] + 2 * key); });
        }
#pragma unroll
        for (int al = 0; al < 3; ++al) {
            atomicAdd(forces + (size_t)i * 3 + al, g[al]);
            atomicAdd(forces + (size_t)j * 3 + al, -g[al]);
        }
    }
    // virial: xx yy zz xy yz zx = -g_alpha * D_beta, reduced over the warp when all lanes share the structure
    double sv[6] = {-g[0] * dl[0], -g[1] * dl[1], -g[2] * dl[2], -g[0] * dl[1], -g[1] * dl[2], -g[2] * dl[0]};
    const int s0 = __shfl_sync(0xffffffffu, s, 0);
    const bool uniform = __all_sync(0xffffffffu, s == s0 || s < 0);
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sv[k] += __shfl_xor_sync(0xffffffffu, sv[k], d);
        }
        const int sl = __reduce_max_sync(0xffffffffu, s);
        if ((threadIdx.x & 31) == 0 && sl >= 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(stresses + (size_t)sl * 6 + k, sv[k]);
        }
    } else if (act) {
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(stresses + (size_t)s * 6 + k, sv[k]);
    }
}

static int eval_ah_stride(const DevModel& m);
bool eval_pairs_rc_supported(const DevModel& m) {
    // (ah_stride even: the adjoints are read as double2; it is n_type * 2 * maxseg by construction)
    return getenv("PM_EVAL_STORED_PB") == nullptr && m.maxl <= 6 && m.n_fn <= EV_MAXFN &&
           (size_t)eval_ah_stride(m) * sizeof(double) + 16 <= 160 * 1024;
}


// Fused eval front end (single launch instead of K3 + k_eval_atom, no G buffer): for AT atoms per CTA
//   (1) linear invariants d_f from the sliced term tables (as k_features_v3),
//   (2) E_i and w_f = dE_i/dd_f from the polynomial terms (shared-memory atomics),
//   (3) the head adjoint Ah[h] = sum_f w_f dd_f/da_h accumulated straight from the sliced contribution tables --
//       the 48 KB / atom G matrix of the fit path is never written (polymlp_eval.cpp:182-290 restated in adjoint form).
template <int AT, int MO>
__global__ void __launch_bounds__(256) k_eval_features(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                        const double* __restrict__ coeffs, double* __restrict__ Ah,
                                                        int ah_stride, double* __restrict__ energies, int nfull_max,
                                                        const double* __restrict__ cmat) {
    extern __shared__ double2 afull[];   // [AT][nfull_max] | sd [AT][fl] | sw [AT][fl] | sah [AT][ah_stride]
    __shared__ double s_dpv[AT * 64];
    double* sd = reinterpret_cast<double*>(afull + (size_t)AT * nfull_max);
    double* sw = sd + (size_t)AT * m.fl;
    double* sah = sw + (size_t)AT * m.fl;
    __shared__ double s_e[AT];
    constexpr int NW = (MO + 1) / 2;
    const int i0 = blockIdx.x * AT;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int segstride = ah_stride / m.n_type;
    int ty[AT];
#pragma unroll
    for (int a = 0; a < AT; ++a) ty[a] = i0 + a < b.n_atoms ? b.types[i0 + a] : -1;
#pragma unroll
    for (int a = 0; a < AT; ++a) {
        if (ty[a] < 0) continue;
        const DevType& T = m.types[ty[a]];
        for (int k = tid; k < T.n_full; k += nthr) {
            double2 v = anc[(size_t)(i0 + a) * m.hmax + T.full_head[k]];
            if (T.full_conj[k]) {
                const double cc = T.full_cc[k];
                v = make_double2(cc * v.x, -cc * v.y);
            }
            afull[(size_t)a * nfull_max + k] = v;
        }
    }
    for (int k = tid; k < AT * m.fl; k += nthr) { sd[k] = 0.0; sw[k] = 0.0; }
    for (int k = tid; k < AT * ah_stride; k += nthr) sah[k] = 0.0;
    if (tid < AT) s_e[tid] = 0.0;
    __syncthreads();
    // (1) linear invariants
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) any = any || ty[a] == tt;
        if (!any) continue;
        const DevType& T = m.types[tt];
        const double* __restrict__ coef = T.sl_coeff;
        const unsigned* __restrict__ ids = T.sl_ids;
        const long ns = T.n_slots;
        for (int s = warp; s < T.n_fsl; s += nwarp) {
            const int4 meta = T.fsl_meta[s];
            const int o = meta.z;
            double sum[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) sum[a] = 0.0;
#pragma unroll 2
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                int id[2 * NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const unsigned v = ids[w * ns + slot];
                    id[2 * w] = v & 0xffffu;
                    id[2 * w + 1] = v >> 16;
                }
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = af[id[0]];
#pragma unroll
                    for (int k = 1; k < MO; ++k)
                        if (k < o) pr = cmul(pr, af[id[k]]);
                    sum[a] += cf * pr.x;
                }
            }
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], 8);
                sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], 16);
            }
            if (lane < 8) {
                const int fp = T.fsl_out[s * 8 + lane];
                if (fp >= 0) {
#pragma unroll
                    for (int a = 0; a < AT; ++a)
                        if (ty[a] == tt) sd[a * m.fl + fp] = sum[a];
                }
            }
        }
    }
    __syncthreads();
    // (2) polynomial: energy and w_f
    double e[AT];
#pragma unroll
    for (int a = 0; a < AT; ++a) e[a] = 0.0;
    if (cmat) {
        // max_p = 2 with <= 64 polynomial variables: w = c_lin + C' d over the polynomial variables, C' the symmetric
        // matrix of the order-2 coefficients with a doubled diagonal (pm_eval_set_coeffs); E_2 = d . C' d / 2.
        // No two threads touch the same w entry, so no atomics (the CAS loops of the generic branch below took more
        // than half of this kernel: ~10 retries per update on the 60 hot entries).
        for (int k = tid; k < AT * 64; k += nthr) {
            const int a = k >> 6, pv = k & 63;
            double v = 0.0;
            if (ty[a] >= 0 && pv < m.npv_pad) {
                const int fp = m.pv_fp[(size_t)ty[a] * m.npv_pad + pv];
                if (fp >= 0) v = sd[a * m.fl + fp];
            }
            s_dpv[k] = v;
        }
        for (int tt = 0; tt < m.n_type; ++tt) {
            bool any = false;
#pragma unroll
            for (int a = 0; a < AT; ++a) any = any || ty[a] == tt;
            if (!any) continue;
            const DevPolyTerm* __restrict__ ct = m.types[tt].colterm;
            for (int col = tid; col < m.n_linear; col += nthr) {
                const DevPolyTerm tm = ct[col];
                if (tm.order != 1) continue;
                const double c = coeffs[col];
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    e[a] += c * sd[a * m.fl + tm.fp0];
                    sw[a * m.fl + tm.fp0] = c;
                }
            }
        }
        __syncthreads();
        for (int k = tid; k < AT * 64; k += nthr) {
            const int a = k >> 6, pv = k & 63;
            if (ty[a] < 0 || pv >= m.npv_pad) continue;
            const int fp = m.pv_fp[(size_t)ty[a] * m.npv_pad + pv];
            if (fp < 0) continue;
            const double* __restrict__ cm = cmat + (size_t)ty[a] * 4096 + pv;   // column pv = row pv (symmetric): coalesced
            const double* dv = s_dpv + a * 64;
            double w2 = 0.0;
#pragma unroll 8
            for (int q = 0; q < 64; ++q) w2 += cm[q * 64] * dv[q];
            sw[a * m.fl + fp] += w2;
#pragma unroll
            for (int a2 = 0; a2 < AT; ++a2)
                if (a2 == a) e[a2] += 0.5 * dv[pv] * w2;
        }
    } else
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) any = any || ty[a] == tt;
        if (!any) continue;
        const DevPolyTerm* __restrict__ ct = m.types[tt].colterm;
        for (int col = tid; col < m.n_variables; col += nthr) {
            const DevPolyTerm tm = ct[col];
            if (!tm.order) continue;
            const double c = coeffs[col];
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                if (ty[a] != tt) continue;
                const double* d = sd + a * m.fl;
                double* w = sw + a * m.fl;
                if (tm.order == 1) {
                    e[a] += c * d[tm.fp0];
                    atomicAdd(w + tm.fp0, c);
                } else if (tm.order == 2) {
                    const double d0 = d[tm.fp0], d1 = d[tm.fp1];
                    e[a] += c * d0 * d1;
                    atomicAdd(w + tm.fp0, c * d1);
                    atomicAdd(w + tm.fp1, c * d0);
                } else {
                    const double d0 = d[tm.fp0], d1 = d[tm.fp1], d2 = d[tm.fp2];
                    e[a] += c * d0 * d1 * d2;
                    atomicAdd(w + tm.fp0, c * d1 * d2);
                    atomicAdd(w + tm.fp1, c * d0 * d2);
                    atomicAdd(w + tm.fp2, c * d0 * d1);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < AT; ++a) {
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) e[a] += __shfl_xor_sync(0xffffffffu, e[a], dlt);
        if (lane == 0 && ty[a] >= 0) atomicAdd(&s_e[a], e[a]);
    }
    __syncthreads();
    if (tid < AT && ty[tid] >= 0) atomicAdd(energies + b.st_of_atom[i0 + tid], s_e[tid]);
    // (3) head adjoint from the contribution slices
    for (int tt = 0; tt < m.n_type; ++tt) {
        bool any = false;
#pragma unroll
        for (int a = 0; a < AT; ++a) any = any || ty[a] == tt;
        if (!any) continue;
        const DevType& T = m.types[tt];
        const double* __restrict__ coef = T.sl_coeff;
        const unsigned* __restrict__ ids = T.sl_ids;
        const long ns = T.n_slots;
        for (int s = warp; s < T.n_esl; s += nwarp) {
            const int4 meta = T.esl_meta[s];
            const int cn = meta.z;
            double gr[AT], gi[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) { gr[a] = 0.0; gi[a] = 0.0; }
#pragma unroll 2
            for (int it = 0; it < meta.y; ++it) {
                const long slot = meta.x + it * 32 + lane;
                const double cf = coef[slot];
                int id[2 * NW];
                unsigned w0 = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const unsigned v = ids[w * ns + slot];
                    if (w == 0) w0 = v;
                    id[2 * w] = v & 0xffffu;
                    id[2 * w + 1] = (v >> 16) & 0x7fffu;
                }
                const double cfi = (w0 & 0x80000000u) ? -cf : cf;
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double2* af = afull + (size_t)a * nfull_max;
                    double2 pr = make_double2(1.0, 0.0);
                    if (MO > 1 && cn > 0) pr = af[id[0]];
#pragma unroll
                    for (int qq = 1; qq < MO - 1; ++qq)
                        if (qq < cn) pr = cmul(pr, af[id[qq]]);
                    gr[a] += cf * pr.x;
                    gi[a] += cfi * pr.y;
                }
            }
            const int2 fh = T.esl_fh[s * 32 + lane];
            if (fh.x >= 0) {
                const int pos = (fh.y >> 20) * segstride + (fh.y & 0xfffff);
#pragma unroll
                for (int a = 0; a < AT; ++a) {
                    if (ty[a] != tt) continue;
                    const double wf = sw[a * m.fl + fh.x];
                    atomicAdd(sah + a * ah_stride + pos, wf * gr[a]);
                    atomicAdd(sah + a * ah_stride + pos + 1, -wf * gi[a]);
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < AT; ++a) {
        if (ty[a] < 0) continue;
        double* ah = Ah + (size_t)(i0 + a) * ah_stride;
        for (int k = tid; k < ah_stride; k += nthr) ah[k] = sah[a * ah_stride + k];
    }
}

// ------------------------------------------------------------------------------------------------
// Fused eval front end, lane = atom: one CTA takes 32 consecutive atoms, lane a of EVERY warp works on atom a, and a warp
// walks one gtinv term list (stage 1: one feature; stage 3: all entries of one head) whose items it reads warp-uniformly.
// The a_nlm arrays live in shared memory as [full id][atom], so every gather is one conflict-free 512-byte row (4 wavefronts
// for 32 useful 16-byte values; the slice kernel above averages ~11 wavefronts for the same 32 values at random addresses,
// and it re-streams the tables for every 4 atoms instead of every 32).  No shared-memory atomics: a head's adjoint is
// summed in registers by the one warp that owns the head.  Needs the max_p = 2 coefficient matrix path (cmat / clin),
// product order <= 4 and (16 n_full + 8 fl + 512) * 32 bytes of shared memory; large batches only (grid = atoms / 32).
// ------------------------------------------------------------------------------------------------
constexpr int LA_AT = 32;
constexpr int LA_NT = 1024;
constexpr int LA_PF = 4;      // table items a warp keeps in flight

template <int MO>
__global__ void __launch_bounds__(LA_NT) k_eval_features_la(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                             double* __restrict__ Ah, int ah_stride,
                                                             double* __restrict__ energies, int nfull_max,
                                                             const double* __restrict__ cmat, const double* __restrict__ clin) {
    extern __shared__ double2 la_af[];   // [nfull_max][32] | sdw [fl][32] | sdpv [64][32] | s_e [32] | s_next [2 MAXT]
    double* sdw = reinterpret_cast<double*>(la_af + (size_t)nfull_max * LA_AT);
    double* sdpv = sdw + (size_t)m.fl * LA_AT;
    double* s_e = sdpv + 64 * LA_AT;
    int* s_next = reinterpret_cast<int*>(s_e + LA_AT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LA_NT / 32;
    const int i = blockIdx.x * LA_AT + lane;
    const bool valid = i < b.n_atoms;
    const int ty = valid ? b.types[i] : -1;
    const int segstride = ah_stride / m.n_type;
    for (int k = warp; k < nfull_max; k += NW) {
        double2 v = make_double2(0.0, 0.0);
        if (valid) {
            const DevType& T = m.types[ty];
            if (k < T.n_full) {
                v = anc[(size_t)i * m.hmax + T.full_head[k]];
                if (T.full_conj[k]) {
                    const double cc = T.full_cc[k];
                    v = make_double2(cc * v.x, -cc * v.y);
                }
            }
        }
        la_af[k * LA_AT + lane] = v;
    }
    for (int k = tid; k < m.fl * LA_AT; k += LA_NT) sdw[k] = 0.0;
    if (tid < LA_AT) s_e[tid] = 0.0;
    if (tid < 2 * MAXT) s_next[tid] = 0;
    __syncthreads();
    // (1) linear invariants: a warp takes the next feature, lane = atom
    for (int tt = 0; tt < m.n_type; ++tt) {
        if (!__any_sync(0xffffffffu, ty == tt)) continue;   // (the same answer in every warp: lanes map to the same atoms)
        const DevType& T = m.types[tt];
        const LaItem* __restrict__ terms = T.la_terms;
        for (;;) {
            int f = 0;
            if (lane == 0) f = atomicAdd(&s_next[tt], 1);
            f = __shfl_sync(0xffffffffu, f, 0);
            if (f >= T.n_feat) break;
            const int o = T.la_forder[f];
            const int t0 = T.term_off[f], t1 = T.term_off[f + 1];
            double sum = 0.0;
            LaItem buf[LA_PF];
#pragma unroll
            for (int k = 0; k < LA_PF; ++k) buf[k] = terms[max(min(t0 + k, t1 - 1), 0)];
            for (int tb = t0; tb < t1; tb += LA_PF) {   // (a feature without terms: t1 == t0, the clamped loads above read a neighbour's item and nothing is used)
                LaItem cur[LA_PF];
#pragma unroll
                for (int k = 0; k < LA_PF; ++k) { cur[k] = buf[k]; buf[k] = terms[min(tb + LA_PF + k, t1 - 1)]; }
#pragma unroll
                for (int k = 0; k < LA_PF; ++k) {
                    if (tb + k < t1) {
                        const LaItem it = cur[k];
                        double2 pr = la_af[(it.w0 & 0xffffu) * LA_AT + lane];
                        if (MO > 1 && o > 1) pr = cmul(pr, la_af[(it.w0 >> 16) * LA_AT + lane]);
                        if (MO > 2 && o > 2) pr = cmul(pr, la_af[(it.w1 & 0xffffu) * LA_AT + lane]);
                        if (MO > 3 && o > 3) pr = cmul(pr, la_af[(it.w1 >> 16) * LA_AT + lane]);
                        sum += it.coeff * pr.x;
                    }
                }
            }
            if (ty == tt) sdw[T.feat_pad[f] * LA_AT + lane] = sum;
        }
    }
    __syncthreads();
    // (2) polynomial (max_p = 2 over <= 64 polynomial variables): E and w = dE/dd, w replaces d in sdw
    for (int k = tid; k < 64 * LA_AT; k += LA_NT) {
        const int pv = k >> 5;
        double v = 0.0;
        if (valid && pv < m.npv_pad) {
            const int fp = m.pv_fp[(size_t)ty * m.npv_pad + pv];
            if (fp >= 0) v = sdw[fp * LA_AT + lane];
        }
        sdpv[k] = v;
    }
    __syncthreads();
    double e = 0.0;
    if (valid)
        for (int k = tid; k < m.fl * LA_AT; k += LA_NT) {
            const double cl = clin[(size_t)ty * m.fl + (k >> 5)];
            e += cl * sdw[k];
            sdw[k] = cl;
        }
    __syncthreads();
    if (valid)
        for (int pv = warp; pv < m.npv_pad; pv += NW) {
            const int fp = m.pv_fp[(size_t)ty * m.npv_pad + pv];
            if (fp < 0) continue;
            const double* __restrict__ cm = cmat + (size_t)ty * 4096 + pv;   // column pv = row pv (symmetric)
            double w2 = 0.0;
#pragma unroll 8
            for (int q = 0; q < 64; ++q) w2 += cm[q * 64] * sdpv[q * LA_AT + lane];
            sdw[fp * LA_AT + lane] += w2;
            e += 0.5 * sdpv[pv * LA_AT + lane] * w2;
        }
    atomicAdd(&s_e[lane], e);
    __syncthreads();
    if (warp == 0 && valid) atomicAdd(energies + b.st_of_atom[i], s_e[lane]);
    // (3) head adjoints: a warp takes the next head and sums all its (feature, contribution) entries
    for (int tt = 0; tt < m.n_type; ++tt) {
        if (!__any_sync(0xffffffffu, ty == tt)) continue;
        const DevType& T = m.types[tt];
        const LaItem* __restrict__ items = T.la_hitems;
        for (;;) {
            int j = 0;
            if (lane == 0) j = atomicAdd(&s_next[MAXT + tt], 1);
            j = __shfl_sync(0xffffffffu, j, 0);
            if (j >= T.n_la_heads) break;
            const int q0 = T.la_hoff[j], q1 = T.la_hoff[j + 1];
            const int hp = T.la_hpos[j];
            double accr = 0.0, acci = 0.0, gr = 0.0, gi = 0.0;
            // the item list is streamed LA_PF items ahead of its use (warp-uniform 16-byte loads, L2 latency)
            LaItem buf[LA_PF];
#pragma unroll
            for (int k = 0; k < LA_PF; ++k) buf[k] = items[max(min(q0 + k, q1 - 1), 0)];
            for (int qb = q0; qb < q1; qb += LA_PF) {
                LaItem cur[LA_PF];
#pragma unroll
                for (int k = 0; k < LA_PF; ++k) { cur[k] = buf[k]; buf[k] = items[min(qb + LA_PF + k, q1 - 1)]; }
#pragma unroll
                for (int k = 0; k < LA_PF; ++k) {
                    if (qb + k < q1) {
                        const LaItem it = cur[k];
                        const int cn = (it.w1 >> 27) & 7;
                        double2 pr = make_double2(1.0, 0.0);
                        if (MO > 1 && cn > 0) pr = la_af[(it.w0 & 0xffffu) * LA_AT + lane];
                        if (MO > 2 && cn > 1) pr = cmul(pr, la_af[((it.w0 >> 16) & 0x7fffu) * LA_AT + lane]);
                        if (MO > 3 && cn > 2) pr = cmul(pr, la_af[(it.w1 & 0x7fffu) * LA_AT + lane]);
                        gr += it.coeff * pr.x;
                        gi += ((it.w0 >> 31) ? -it.coeff : it.coeff) * pr.y;
                        if (it.w1 & (1u << 30)) {   // last contribution of its (feature, head) entry
                            const double w = sdw[((it.w1 >> 15) & 0xfffu) * LA_AT + lane];
                            accr += w * gr;
                            acci -= w * gi;
                            gr = 0.0; gi = 0.0;
                        }
                    }
                }
            }
            if (ty == tt) {
                double* ah = Ah + (size_t)i * ah_stride + (hp >> 20) * segstride + (hp & 0xfffff);
                ah[0] = accr;
                ah[1] = acci;
            }
        }
    }
}

// Radial-batched variant of the lane = atom kernel for models that are a radial replication (DevType::lb_nr > 0): the term
// lists of radial index n are those of radial index 0 with shifted ids, so a warp decodes each table item ONCE and applies
// it to NB radial indices (NB independent gather / multiply / accumulate chains): ~2.5x fewer instructions per useful
// gather than k_eval_features_la.  Work items: stage 1 (radial-0 feature, block of NB radial indices); stage 3 (chunk of a
// radial-0 head's contribution list, block of NB radial indices) -- the chunks of one head add into Ah with RED.F64.
constexpr int LB_NT = 768;   // 24 warps at <= 80 registers

template <int NB>
__global__ void __launch_bounds__(LB_NT) k_eval_features_lb(DevModel m, DevBatch b, const double2* __restrict__ anc,
                                                             double* __restrict__ Ah, int ah_stride,
                                                             double* __restrict__ energies, int nfull_max,
                                                             const double* __restrict__ cmat, const double* __restrict__ clin) {
    extern __shared__ double2 la_af[];   // [nfull_max][32] | sdw [fl][32] | sdpv [64][32] | s_e [32] | s_next [2 MAXT]
    double* sdw = reinterpret_cast<double*>(la_af + (size_t)nfull_max * LA_AT);
    double* sdpv = sdw + (size_t)m.fl * LA_AT;
    double* s_e = sdpv + 64 * LA_AT;
    int* s_next = reinterpret_cast<int*>(s_e + LA_AT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LB_NT / 32;
    const int i = blockIdx.x * LA_AT + lane;
    const bool valid = i < b.n_atoms;
    const int ty = valid ? b.types[i] : -1;
    const int segstride = ah_stride / m.n_type;
#pragma unroll 4
    for (int k = warp; k < nfull_max; k += NW) {
        double2 v = make_double2(0.0, 0.0);
        if (valid) {
            const DevType& T = m.types[ty];
            if (k < T.n_full) {
                v = anc[(size_t)i * m.hmax + T.full_head[k]];
                if (T.full_conj[k]) {
                    const double cc = T.full_cc[k];
                    v = make_double2(cc * v.x, -cc * v.y);
                }
            }
        }
        la_af[k * LA_AT + lane] = v;
    }
    for (int k = tid; k < m.fl * LA_AT; k += LB_NT) sdw[k] = 0.0;
    if (tid < LA_AT) s_e[tid] = 0.0;
    if (tid < 2 * MAXT) s_next[tid] = 0;
    __syncthreads();
    // shared-memory byte addresses: a_nlm row of full id k for this lane = af_l + k * 512; radial stride S * 512
    const unsigned af_l = smem_u32(la_af) + (unsigned)lane * 16u;
    const unsigned sdw_l = smem_u32(sdw) + (unsigned)lane * 8u;
    // (1) linear invariants.  (NB divides the number of radial indices: no tail; the product order is uniform per feature,
    // so the branch on it sits outside the radial loop and the loop bodies are branch-free.)
    for (int tt = 0; tt < m.n_type; ++tt) {
        if (!__any_sync(0xffffffffu, ty == tt)) continue;
        const DevType& T = m.types[tt];
        const LaItem* __restrict__ terms = T.lb_terms;
        const int4* __restrict__ fwork = T.lb_fwork;
        const int nblk = T.lb_nr / NB, nwork = T.lb_nfwork * nblk;
        const unsigned Sb = (unsigned)T.lb_S * 512u, Fb = (unsigned)T.lb_Fs * 256u;
        for (;;) {
            int wk = 0;
            if (lane == 0) wk = atomicAdd(&s_next[tt], 1);
            wk = __shfl_sync(0xffffffffu, wk, 0);
            if (wk >= nwork) break;
            const int cidx = wk / nblk, n0 = (wk - cidx * nblk) * NB;
            const int4 wi = fwork[cidx];   // first term, end, feature index
            const int o = T.lb_forder[wi.z];
            const unsigned afn = af_l + (unsigned)n0 * Sb;
            double sum[NB];
#pragma unroll
            for (int n = 0; n < NB; ++n) sum[n] = 0.0;
            LaItem nxt = terms[wi.x], nx2 = terms[min(wi.x + 1, wi.y - 1)];
            for (int ti = wi.x; ti < wi.y; ++ti) {
                const LaItem it = nxt;
                nxt = nx2;
                nx2 = terms[min(ti + 2, wi.y - 1)];
                const unsigned a0 = afn + (it.w0 & 0xffffu) * 512u, a1 = afn + (it.w0 >> 16) * 512u;
                const unsigned a2 = afn + (it.w1 & 0xffffu) * 512u, a3 = afn + (it.w1 >> 16) * 512u;
                if (o == 2) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = lds_f64x2(a0 + n * Sb), q = lds_f64x2(a1 + n * Sb);
                        sum[n] += it.coeff * (p.x * q.x - p.y * q.y);
                    }
                } else if (o == 3) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = cmul(lds_f64x2(a0 + n * Sb), lds_f64x2(a1 + n * Sb)), q = lds_f64x2(a2 + n * Sb);
                        sum[n] += it.coeff * (p.x * q.x - p.y * q.y);
                    }
                } else if (o == 1) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) sum[n] += it.coeff * lds_f64x2(a0 + n * Sb).x;
                } else {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = cmul(lds_f64x2(a0 + n * Sb), lds_f64x2(a1 + n * Sb));
                        const double2 q = cmul(lds_f64x2(a2 + n * Sb), lds_f64x2(a3 + n * Sb));
                        sum[n] += it.coeff * (p.x * q.x - p.y * q.y);
                    }
                }
            }
            if (ty == tt) {
                double* dst = sdw + (size_t)(T.lb_fpad[wi.z] + n0 * T.lb_Fs) * LA_AT + lane;
#pragma unroll
                for (int n = 0; n < NB; ++n) atomicAdd(dst + (size_t)n * T.lb_Fs * LA_AT, sum[n]);   // (chunks of one feature)
            }
        }
    }
    __syncthreads();
    // (2) polynomial (max_p = 2 over <= 64 polynomial variables): E and w = dE/dd, w replaces d in sdw
    for (int k = tid; k < 64 * LA_AT; k += LB_NT) {
        const int pv = k >> 5;
        double v = 0.0;
        if (valid && pv < m.npv_pad) {
            const int fp = m.pv_fp[(size_t)ty * m.npv_pad + pv];
            if (fp >= 0) v = sdw[fp * LA_AT + lane];
        }
        sdpv[k] = v;
    }
    __syncthreads();
    double e = 0.0;
    if (valid)
        for (int k = tid; k < m.fl * LA_AT; k += LB_NT) {
            const double cl = clin[(size_t)ty * m.fl + (k >> 5)];
            e += cl * sdw[k];
            sdw[k] = cl;
        }
    __syncthreads();
    if (valid) {
        // four polynomial variables per warp pass: four independent sums share each d_q (and the coefficient loads are
        // warp-uniform 32-byte rows)
        const double* __restrict__ cmt = cmat + (size_t)ty * 4096;
        for (int pv0 = warp * 4; pv0 < m.npv_pad; pv0 += NW * 4) {
            double w2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
            for (int q = 0; q < 64; ++q) {
                const double dq = sdpv[q * LA_AT + lane];
                const double4 cq = *reinterpret_cast<const double4*>(cmt + q * 64 + pv0);   // row q = column q (symmetric)
                w2[0] += cq.x * dq; w2[1] += cq.y * dq; w2[2] += cq.z * dq; w2[3] += cq.w * dq;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int pv = pv0 + r;
                const int fp = pv < m.npv_pad ? m.pv_fp[(size_t)ty * m.npv_pad + pv] : -1;
                if (fp < 0) continue;
                sdw[fp * LA_AT + lane] += w2[r];
                e += 0.5 * sdpv[pv * LA_AT + lane] * w2[r];
            }
        }
    }
    atomicAdd(&s_e[lane], e);
    __syncthreads();
    if (warp == 0 && valid) atomicAdd(energies + b.st_of_atom[i], s_e[lane]);
    // (3) head adjoints: the branch on the number of factors of a contribution is warp-uniform and outside the radial loop
    for (int tt = 0; tt < m.n_type; ++tt) {
        if (!__any_sync(0xffffffffu, ty == tt)) continue;
        const DevType& T = m.types[tt];
        const LaItem* __restrict__ items = T.lb_hitems;
        const int4* __restrict__ work = T.lb_work;
        const int nblk = T.lb_nr / NB, nwork = T.lb_nwork * nblk;
        const unsigned Sb = (unsigned)T.lb_S * 512u, Fb = (unsigned)T.lb_Fs * 256u;
        const int Ps = T.lb_Ps;
        for (;;) {
            int wk = 0;
            if (lane == 0) wk = atomicAdd(&s_next[MAXT + tt], 1);
            wk = __shfl_sync(0xffffffffu, wk, 0);
            if (wk >= nwork) break;
            const int c = wk / nblk, n0 = (wk - c * nblk) * NB;
            const int4 wi = work[c];   // first item, end, head position key
            const unsigned afn = af_l + (unsigned)n0 * Sb, swn = sdw_l + (unsigned)n0 * Fb;
            double accr[NB], acci[NB], gr[NB], gi[NB];
#pragma unroll
            for (int n = 0; n < NB; ++n) { accr[n] = 0.0; acci[n] = 0.0; gr[n] = 0.0; gi[n] = 0.0; }
            LaItem nxt = items[wi.x], nx2 = items[min(wi.x + 1, wi.y - 1)];
            for (int q = wi.x; q < wi.y; ++q) {
                const LaItem it = nxt;
                nxt = nx2;
                nx2 = items[min(q + 2, wi.y - 1)];
                const int cn = (it.w1 >> 27) & 7;
                const unsigned a0 = afn + (it.w0 & 0xffffu) * 512u, a1 = afn + ((it.w0 >> 16) & 0x7fffu) * 512u;
                const unsigned a2 = afn + (it.w1 & 0x7fffu) * 512u;
                const double cfi = (it.w0 >> 31) ? -it.coeff : it.coeff;
                if (cn == 2) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = lds_f64x2(a0 + n * Sb), r = lds_f64x2(a1 + n * Sb);
                        gr[n] += it.coeff * (p.x * r.x - p.y * r.y);
                        gi[n] += cfi * (p.x * r.y + p.y * r.x);
                    }
                } else if (cn == 1) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = lds_f64x2(a0 + n * Sb);
                        gr[n] += it.coeff * p.x;
                        gi[n] += cfi * p.y;
                    }
                } else if (cn == 0) {
#pragma unroll
                    for (int n = 0; n < NB; ++n) gr[n] += it.coeff;
                } else {
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const double2 p = cmul(cmul(lds_f64x2(a0 + n * Sb), lds_f64x2(a1 + n * Sb)), lds_f64x2(a2 + n * Sb));
                        gr[n] += it.coeff * p.x;
                        gi[n] += cfi * p.y;
                    }
                }
                if (it.w1 & (1u << 30)) {   // last contribution of its (feature, head) entry
                    const unsigned fo = swn + ((it.w1 >> 15) & 0xfffu) * 256u;
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        double w;
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w) : "r"(fo + n * Fb));
                        accr[n] += w * gr[n];
                        acci[n] -= w * gi[n];
                        gr[n] = 0.0; gi[n] = 0.0;
                    }
                }
            }
            if (ty == tt) {
                double* ah = Ah + (size_t)i * ah_stride + (wi.z >> 20) * segstride + (wi.z & 0xfffff) + (size_t)n0 * Ps;
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    atomicAdd(ah + n * Ps, accr[n]);
                    atomicAdd(ah + n * Ps + 1, acci[n]);
                }
            }
        }
    }
}

static size_t eval_la_smem(const DevModel& m, size_t feat_smem) {
    return (size_t)LA_AT * (feat_smem + (size_t)m.fl * sizeof(double) + 64 * sizeof(double) + sizeof(double)) + 2 * MAXT * sizeof(int);
}
static bool eval_la_supported(const DevModel& m, const DevBatch& b, const Workspace& ws, size_t feat_smem) {
    if (getenv("PM_EVAL_LA") && atoi(getenv("PM_EVAL_LA")) == 0) return false;
    const int min_atoms = getenv("PM_EVAL_LA_MIN") ? atoi(getenv("PM_EVAL_LA_MIN")) : 148 * LA_AT * 2;
    if (!ws.cmat || !ws.clin || m.npv_pad > 64 || b.n_atoms < min_atoms) return false;
    for (int t = 0; t < m.n_type; ++t)
        if (!m.types[t].la_ok || m.types[t].max_order > 4) return false;
    return eval_la_smem(m, feat_smem) <= 227 * 1024;
}

constexpr int EVF_AT = 4;
static int eval_ah_stride(const DevModel& m) {
    int maxseg = 0;
    for (int t = 0; t < m.n_type; ++t)
        for (int u = 0; u < m.n_type; ++u) maxseg = max(maxseg, m.types[t].seg_len[u]);
    return m.n_type * 2 * maxseg;
}
static size_t eval_fused_smem(const DevModel& m, size_t feat_smem) {
    return EVF_AT * (feat_smem + (2ull * m.fl + eval_ah_stride(m)) * sizeof(double));
}
bool eval_fused_supported(const DevModel& m, size_t feat_smem) {
    int mo = 1;
    for (int t = 0; t < m.n_type; ++t) {
        mo = max(mo, m.types[t].max_order);
        if (m.types[t].n_fsl <= 0 || m.types[t].n_esl <= 0) return false;
    }
    return mo <= 6 && eval_fused_smem(m, feat_smem) <= 160 * 1024;
}

void launch_eval_adjoint(const DevModel& m, const DevBatch& b, const Workspace& ws, const double* coeffs,
                         double* energies, double* forces, double* stresses, cudaStream_t s, size_t feat_smem) {
    if (b.n_atoms == 0) return;
    const int ah_stride = eval_ah_stride(m);
    if (feat_smem > 0 && eval_la_supported(m, b, ws, feat_smem)) {
        const size_t smem = eval_la_smem(m, feat_smem);
        const int nfull_max = (int)(feat_smem / sizeof(double2));
        const int grid = (b.n_atoms + LA_AT - 1) / LA_AT;
        int mo = 1;
        for (int t = 0; t < m.n_type; ++t) mo = max(mo, m.types[t].max_order);
        cudaMemsetAsync(ws.Ah, 0, (size_t)b.n_atoms * ah_stride * sizeof(double), s);   // positions no head entry writes
        bool radial = getenv("PM_EVAL_LB") == nullptr || atoi(getenv("PM_EVAL_LB")) != 0;
        int nr = m.types[0].lb_nr;
        for (int t = 0; t < m.n_type; ++t) radial = radial && m.types[t].lb_nr > 0 && m.types[t].lb_nr == nr;
        radial = radial && (nr % 2 == 0 || nr % 3 == 0 || nr % 5 == 0);   // NB must divide the number of radial indices
        if (radial) {
#define PM_LB_CASE(NB_)                                                                                               \
    {                                                                                                                 \
        ensure_smem((const void*)k_eval_features_lb<NB_>, smem);                                                      \
        k_eval_features_lb<NB_><<<grid, LB_NT, smem, s>>>(m, b, ws.anc, ws.Ah, ah_stride, energies, nfull_max, ws.cmat, ws.clin); \
    }
            if (nr % 5 == 0) PM_LB_CASE(5) else if (nr % 4 == 0) PM_LB_CASE(4) else if (nr % 6 == 0) PM_LB_CASE(6)
            else if (nr % 3 == 0) PM_LB_CASE(3) else PM_LB_CASE(2)
#undef PM_LB_CASE
        } else
#define PM_LA_CASE(MO_)                                                                                               \
    case MO_:                                                                                                         \
        ensure_smem((const void*)k_eval_features_la<MO_>, smem);                                                      \
        k_eval_features_la<MO_><<<grid, LA_NT, smem, s>>>(m, b, ws.anc, ws.Ah, ah_stride, energies, nfull_max, ws.cmat, ws.clin); \
        break;
        switch (mo) { PM_LA_CASE(1) PM_LA_CASE(2) PM_LA_CASE(3) PM_LA_CASE(4) }
#undef PM_LA_CASE
    } else if (feat_smem > 0) {
        const size_t smem = eval_fused_smem(m, feat_smem);
        const int nfull_max = (int)(feat_smem / sizeof(double2));
        const int grid = (b.n_atoms + EVF_AT - 1) / EVF_AT;
        int mo = 1;
        for (int t = 0; t < m.n_type; ++t) mo = max(mo, m.types[t].max_order);
#define PM_EVF_CASE(MO_)                                                                                              \
    case MO_:                                                                                                         \
        ensure_smem((const void*)k_eval_features<EVF_AT, MO_>, smem);                                                 \
        k_eval_features<EVF_AT, MO_><<<grid, 256, smem, s>>>(m, b, ws.anc, coeffs, ws.Ah, ah_stride, energies, nfull_max, ws.cmat); \
        break;
        switch (mo) {
            PM_EVF_CASE(1) PM_EVF_CASE(2) PM_EVF_CASE(3) PM_EVF_CASE(4) PM_EVF_CASE(5) PM_EVF_CASE(6)
        }
#undef PM_EVF_CASE
    } else {
        k_eval_atom<<<b.n_atoms, 256, 0, s>>>(m, b, ws.dfeat, ws.Gbuf, coeffs, ws.Xown, ws.Ah, ah_stride, energies);
    }
    if (b.n_pairs > 0 && ws.pairs_rc) {
        const int nc = (int)std::max<size_t>(1, std::min<size_t>(EP_NC, (48 * 1024 - 16) / ((size_t)ah_stride * sizeof(double))));
        const size_t smem = (size_t)nc * ah_stride * sizeof(double) + 16;
        const int grid = (b.n_pairs + 127) / 128;
        init_pair_basis_tables();
#define PM_RC_CASE(L_, NF_)                                                                                           \
    {                                                                                                                 \
        if (ws.pairs_rads) {                                                                                          \
            ensure_smem((const void*)k_eval_pairs_rc<L_, NF_, true>, smem);                                           \
            k_eval_pairs_rc<L_, NF_, true><<<grid, 128, smem, s>>>(m, b, ws.PB, ws.Ah, ah_stride, forces, stresses, nc);  \
        } else {                                                                                                      \
            ensure_smem((const void*)k_eval_pairs_rc<L_, NF_, false>, smem);                                          \
            k_eval_pairs_rc<L_, NF_, false><<<grid, 128, smem, s>>>(m, b, ws.PB, ws.Ah, ah_stride, forces, stresses, nc); \
        }                                                                                                             \
    }
#define PM_RC_L(L_)                                                                                                   \
    case L_:                                                                                                          \
        if (m.n_fn <= 8) PM_RC_CASE(L_, 8) else if (m.n_fn <= 12) PM_RC_CASE(L_, 12) else PM_RC_CASE(L_, 16)          \
        break;
        switch (m.maxl) { PM_RC_L(0) PM_RC_L(1) PM_RC_L(2) PM_RC_L(3) PM_RC_L(4) PM_RC_L(5) PM_RC_L(6) }
#undef PM_RC_L
#undef PM_RC_CASE
    } else if (b.n_pairs > 0) {
        if (m.n_fn <= EV_MAXFN)
            // (a half-pair variant -- each unordered pair once with Ah_i + (-1)^l Ah_j, the reference's trick -- measured
            // slower, 11.2 vs 8.7 ms per 131072 atoms: the idle lanes of a warp still pull the same 32-byte sectors of the
            // blocked pair records, and the second adjoint is an extra L2 gather)
        {
            const int nc = (int)std::min<size_t>(EP_NC, (48 * 1024) / ((size_t)ah_stride * sizeof(double)));   // 0: adjoints from global
            k_eval_pairs_v2<<<(b.n_pairs + 127) / 128, 128, (size_t)nc * ah_stride * sizeof(double), s>>>(m, b, ws.PB, ws.Ah, ah_stride,
                                                                                                    forces, stresses, nc);
        }
        else
            k_eval_pairs<<<(b.n_pairs * 3 + 127) / 128, 128, 0, s>>>(m, b, ws.PB, ws.Ah, ah_stride, forces, stresses);
    }
}

// ================================================================================================
// launch wrappers with simple / tensor-core dispatch
// ================================================================================================
void launch_lrows_mma(const DevModel& m, const DevBatch& b, const Workspace& ws, bool apply_weights, cudaStream_t s);
void launch_xrows_mma(const DevModel& m, const DevBatch& b, const Workspace& ws, double* xe_sum, double* xe_sq,
                      bool apply_weights, cudaStream_t s);
void launch_syrk_mma(const double* X, int n_rows, int fpad, double* C, cudaStream_t s, const SyrkScratch* scr);

void launch_lrows(const DevModel& m, const DevBatch& b, const Workspace& ws, bool simple, bool apply_weights, cudaStream_t s) {
    if (b.n_atoms == 0) return;
    if (simple) k_lrows_simple<<<b.n_atoms, 256, 0, s>>>(m, b, ws.PB, ws.agg, ws.Gbuf, ws.Lbuf, ws.Xown, ws.Sbuf);
    else launch_lrows_mma(m, b, ws, apply_weights, s);
}

void launch_xrows(const DevModel& m, const DevBatch& b, const Workspace& ws, double* xe_sum, double* xe_sq,
                  bool simple, bool apply_weights, cudaStream_t s) {
    if (b.n_atoms == 0) return;
    if (simple) {
        k_xrows_force_simple<<<dim3(b.n_atoms, 3), 256, 0, s>>>(m, b, ws.dfeat, ws.Lbuf, ws.Xown, ws.X, apply_weights ? 1 : 0);
        k_xrows_struct_simple<<<dim3(b.n_st, 7), 256, 0, s>>>(m, b, ws.dfeat, ws.Sbuf, ws.X, xe_sum, xe_sq, apply_weights ? 1 : 0);
    } else {
        launch_xrows_mma(m, b, ws, xe_sum, xe_sq, apply_weights, s);
    }
}

void launch_syrk(const double* X, int n_rows, int fpad, double* C, bool simple, cudaStream_t s, const SyrkScratch* scr) {
    if (n_rows == 0) return;
    if (simple) {
        const int nt = fpad / 64;
        k_syrk_simple<<<nt * (nt + 1) / 2, 256, 0, s>>>(X, n_rows, fpad, C);
    } else {
        launch_syrk_mma(X, n_rows, fpad, C, s, scr);
    }
}

}  // namespace pm

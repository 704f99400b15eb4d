// pm_solver.cu -- ridge regression on the device-resident accumulator (cuSOLVER Cholesky + cuBLAS).
// Reference: src/pypolymlp/mlp_dev/core/utils_scales.py:6-40, data_sequential.py:72-92,
// src/pypolymlp/mlp_dev/standard/solvers.py:9-84, src/pypolymlp/mlp_dev/core/utils_model_selection.py:37-72.
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include <cublas_v2.h>
#include <cusolverDn.h>

namespace pm {

// A[i][j] = C[i][j] / (s_i s_j) on the upper triangle (row-major upper == column-major lower for LAPACK),
// zero rows / columns flagged in `zero`; rhs[i] = C[i][F] / s_i.
__global__ void k_scale_system(const double* __restrict__ C, int fpad, int F, const double* __restrict__ sinv,
                               const int* __restrict__ zero, double* __restrict__ A, double* __restrict__ rhs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= F) return;
    for (int i = blockIdx.y; i < F; i += gridDim.y) {   // rows strided over gridDim.y (<= 65535)
        const bool z = zero[i] || zero[j];
        if (j >= i) A[(size_t)i * F + j] = z ? 0.0 : C[(size_t)i * fpad + j] * sinv[i] * sinv[j];
        if (j == 0) rhs[i] = zero[i] ? 0.0 : C[(size_t)i * fpad + F] * sinv[i];
    }
}

__global__ void k_add_diag(double* __restrict__ A, int F, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) A[(size_t)i * F + i] += v;
}

#define SOLVER_CK(call, what)                                                          \
    do {                                                                                \
        if ((call) != 0) throw std::runtime_error(std::string("ridge solve: ") + what); \
    } while (0)

// C: device accumulator (fpad x fpad, upper triangle valid, column F = X^T y, C[F][F] = y^T y);
// xe_sum / xe_sq: device vectors behind C.  Outputs on the host.
void solve_ridge_device(const double* C, int fpad, int F, const double* xe_sum_h, const double* xe_sq_h,
                        double y_sq_norm, long n_data, const double* alphas, int n_alpha, const double* scales_in,
                        long n_energy, bool include_force, double threshold, double* scales_out, double* coefs,
                        double* rmse, cudaStream_t stream) {
    // ---- scales (host, O(F)) -----------------------------------------------------------------------
    std::vector<double> scales(F), sinv(F);
    std::vector<int> zero(F, 0);
    for (int i = 0; i < F; ++i) {
        double sc;
        if (scales_in) sc = scales_in[i];
        else {
            const double mean = xe_sum_h[i] / (double)n_energy;
            double var = xe_sq_h[i] / (double)n_energy - mean * mean;
            if (var < 0.0) var = 1.0;
            sc = std::sqrt(var);
        }
        const double thr = include_force ? threshold : threshold * threshold;
        if (std::fabs(sc) < thr) { zero[i] = 1; sc = 1.0; }
        scales[i] = sc;
        sinv[i] = 1.0 / sc;
        if (scales_out) scales_out[i] = sc;
    }
    double *d_sinv = nullptr, *A = nullptr, *A0 = nullptr, *rhs = nullptr, *x = nullptr, *work = nullptr, *tmp = nullptr;
    int *d_zero = nullptr, *d_info = nullptr;
    cusolverDnHandle_t sol = nullptr;
    cublasHandle_t blas = nullptr;
    auto cleanup = [&] {
        cudaFree(d_sinv); cudaFree(A); cudaFree(A0); cudaFree(rhs); cudaFree(x); cudaFree(work); cudaFree(tmp);
        cudaFree(d_zero); cudaFree(d_info);
        if (sol) cusolverDnDestroy(sol);
        if (blas) cublasDestroy(blas);
    };
    try {
        const size_t nA = (size_t)F * F;
        SOLVER_CK(cudaMalloc(&d_sinv, F * sizeof(double)), "alloc");
        SOLVER_CK(cudaMalloc(&d_zero, F * sizeof(int)), "alloc");
        SOLVER_CK(cudaMalloc(&A, nA * sizeof(double)), "alloc A");
        SOLVER_CK(cudaMalloc(&A0, nA * sizeof(double)), "alloc A0");
        SOLVER_CK(cudaMalloc(&rhs, F * sizeof(double)), "alloc");
        SOLVER_CK(cudaMalloc(&x, F * sizeof(double)), "alloc");
        SOLVER_CK(cudaMalloc(&tmp, F * sizeof(double)), "alloc");
        SOLVER_CK(cudaMalloc(&d_info, sizeof(int)), "alloc");
        SOLVER_CK(cudaMemcpyAsync(d_sinv, sinv.data(), F * sizeof(double), cudaMemcpyHostToDevice, stream), "h2d");
        SOLVER_CK(cudaMemcpyAsync(d_zero, zero.data(), F * sizeof(int), cudaMemcpyHostToDevice, stream), "h2d");
        SOLVER_CK(cudaMemsetAsync(A0, 0, nA * sizeof(double), stream), "memset");
        k_scale_system<<<dim3((F + 255) / 256, std::min(F, 32768)), 256, 0, stream>>>(C, fpad, F, d_sinv, d_zero, A0, rhs);
        SOLVER_CK(cudaGetLastError(), "k_scale_system launch");
        SOLVER_CK(cusolverDnCreate(&sol), "cusolverDnCreate");
        SOLVER_CK(cusolverDnSetStream(sol, stream), "cusolverDnSetStream");
        SOLVER_CK(cublasCreate(&blas), "cublasCreate");
        SOLVER_CK(cublasSetStream(blas, stream), "cublasSetStream");
        int lwork = 0;
        // row-major upper triangle == column-major lower triangle
        SOLVER_CK(cusolverDnDpotrf_bufferSize(sol, CUBLAS_FILL_MODE_LOWER, F, A, F, &lwork), "potrf_bufferSize");
        SOLVER_CK(cudaMalloc(&work, (size_t)std::max(lwork, 1) * sizeof(double)), "alloc work");
        const double one = 1.0, zero_d = 0.0;
        double alpha_prev = 0.0;
        for (int k = 0; k < n_alpha; ++k) {
            // incremental diagonal update on the pristine matrix (solvers.py:76-83), then factorise a copy
            k_add_diag<<<(F + 255) / 256, 256, 0, stream>>>(A0, F, alphas[k] - alpha_prev);
            SOLVER_CK(cudaGetLastError(), "k_add_diag launch");
            alpha_prev = alphas[k];
            SOLVER_CK(cudaMemcpyAsync(A, A0, nA * sizeof(double), cudaMemcpyDeviceToDevice, stream), "copy");
            SOLVER_CK(cudaMemcpyAsync(x, rhs, F * sizeof(double), cudaMemcpyDeviceToDevice, stream), "copy");
            SOLVER_CK(cusolverDnDpotrf(sol, CUBLAS_FILL_MODE_LOWER, F, A, F, work, lwork, d_info), "potrf");
            int info = 0;
            SOLVER_CK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, stream), "d2h");
            SOLVER_CK(cudaStreamSynchronize(stream), "sync");
            double* ck = coefs + (size_t)k * F;
            if (info != 0) {
                for (int i = 0; i < F; ++i) ck[i] = 1e30;
                rmse[k] = 1e10;
                continue;
            }
            SOLVER_CK(cusolverDnDpotrs(sol, CUBLAS_FILL_MODE_LOWER, F, 1, A, F, x, F, d_info), "potrs");
            // mse = (c^T (XtX) c - 2 c^T Xty + y^T y) / n  with the alpha-free matrix: A0 carries +alpha on the diagonal
            SOLVER_CK(cublasDsymv(blas, CUBLAS_FILL_MODE_LOWER, F, &one, A0, F, x, 1, &zero_d, tmp, 1), "symv");
            double cAc = 0.0, cb = 0.0, cc = 0.0;
            SOLVER_CK(cublasDdot(blas, F, x, 1, tmp, 1, &cAc), "dot");
            SOLVER_CK(cublasDdot(blas, F, x, 1, rhs, 1, &cb), "dot");
            SOLVER_CK(cublasDdot(blas, F, x, 1, x, 1, &cc), "dot");
            SOLVER_CK(cudaMemcpyAsync(ck, x, F * sizeof(double), cudaMemcpyDeviceToHost, stream), "d2h");
            SOLVER_CK(cudaStreamSynchronize(stream), "sync");
            const double mse = (cAc - alphas[k] * cc - 2.0 * cb + y_sq_norm) / (double)n_data;
            rmse[k] = mse >= 0.0 ? std::sqrt(mse) : 1e10;
        }
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace pm

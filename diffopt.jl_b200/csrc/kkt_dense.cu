// Direct branch of `solve_system` for ONE KKT system given as Julia's SparseMatrixCSC (or its Adjoint):
// `LHS \ RHS` of QuadraticProgram.jl:486-492, called from reverse_differentiate! (:335, LHS) and
// forward_differentiate! (:438, LHS').  The matrix is scattered into a dense column-major array on the device
// (augmented with the right-hand sides) and factorised by a partially pivoted LU in one persistent CTA; zero pivot
// -> info > 0 (the reference throws SingularException there).  Intended for the reference's problem sizes
// (N up to a few thousand, the whole matrix stays L2 resident); many right-hand sides share the factorisation,
// which the reference does not do (it refactorises per direction).
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int LU_THREADS = 1024;

__global__ void csc_scatter_kernel(int64_t N, int64_t ld, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                   const double* __restrict__ nzval, int trans, double* __restrict__ M) {
    // one warp per column of the CSC matrix
    const int64_t col = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= N) return;
    for (int64_t k = colptr[col] - 1 + lane; k < colptr[col + 1] - 1; k += 32) {
        const int64_t row = rowval[k] - 1;
        // duplicates are summed like SparseArrays does on construction
        if (trans) atomicAdd(&M[row * ld + col], nzval[k]);
        else atomicAdd(&M[col * ld + row], nzval[k]);
    }
}

// M: N x (N + nrhs) column-major with leading dimension ld = N; on exit columns N.. hold the solutions.
__global__ void __launch_bounds__(LU_THREADS, 1) dense_lu_solve_kernel(int N, int nrhs, double* __restrict__ M, int* info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double s_rinv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = N + nrhs;
    const size_t ld = (size_t)N;
    for (int k = 0; k < N; ++k) {
        // pivot search in column k, rows k..N-1
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < N; i += LU_THREADS) {
            const double v = fabs(M[k * ld + i]);
            if (v > best) {
                best = v;
                bi = i;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            s_val[warp] = best;
            s_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane];
            bi = s_idx[lane];
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) {
                    best = ov;
                    bi = oi;
                }
            }
            if (lane == 0) {
                s_piv = bi;
                s_rinv = best > 0.0 ? 1.0 / M[k * ld + bi] : 0.0;
            }
        }
        __syncthreads();
        const int p = s_piv;
        const double rinv = s_rinv;
        if (rinv == 0.0) {  // exactly zero column below the diagonal: singular
            if (tid == 0) *info = k + 1;
            return;
        }
        // row interchange k <-> p on all columns, then multipliers
        if (p != k) {
            for (int c = tid; c < W; c += LU_THREADS) {
                const double a = M[c * ld + k], b = M[c * ld + p];
                M[c * ld + k] = b;
                M[c * ld + p] = a;
            }
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < N; i += LU_THREADS) M[k * ld + i] *= rinv;
        __syncthreads();
        // rank-1 update of the trailing block (and of the right-hand sides): thread grid over (row, column)
        const int rows = N - k - 1, cols = W - k - 1;
        if (rows > 0) {
            // a warp walks down a column segment so that accesses are coalesced
            const int rchunks = (rows + 31) >> 5;
            const long long items = (long long)rchunks * cols;
            for (long long it = warp; it < items; it += LU_THREADS / 32) {
                const int c = k + 1 + (int)(it / rchunks);
                const int i = k + 1 + (int)(it % rchunks) * 32 + lane;
                if (i < N) M[c * ld + i] = fma(-M[k * ld + i], M[c * ld + k], M[c * ld + i]);
            }
        }
        __syncthreads();
    }
    // back substitution U x = y on the nrhs columns (column oriented)
    for (int k = N - 1; k >= 0; --k) {
        const double rinv = 1.0 / M[k * ld + k];
        for (int r = tid; r < nrhs; r += LU_THREADS) M[(N + r) * ld + k] *= rinv;
        __syncthreads();
        const long long items = (long long)k * nrhs;
        for (long long it = tid; it < items; it += LU_THREADS) {
            const int r = (int)(it / k), i = (int)(it % k);
            M[(N + r) * ld + i] = fma(-M[k * ld + i], M[(N + r) * ld + k], M[(N + r) * ld + i]);
        }
        __syncthreads();
    }
    if (tid == 0) *info = 0;
}

}  // namespace

extern "C" int32_t diffopt_b200_kkt_solve_csc(diffopt_b200_ctx* ctx, int64_t N, const int64_t* colptr, const int64_t* rowval,
                                              const double* nzval, int32_t trans, int64_t nrhs, const double* rhs, double* x_out,
                                              int32_t memspace) {
    if (!ctx) return -1;
    if (N <= 0 || nrhs <= 0 || !colptr || !rowval || !nzval || !rhs || !x_out) BAD_ARG(ctx, "kkt_solve_csc: bad argument");
    DeviceGuard guard_(ctx->device);
    // The dense kernel is one CTA: fine for the reference's own problem sizes, hopeless beyond ~1000 unknowns.  Larger
    // systems go through the sparse factorisation (multifrontal LU, sparse_mf.cu), which is what the reference's sparse
    // `\` does at any size.  DIFFOPT_B200_DENSE_MAX moves the switch-over.
    int64_t dense_max = 1024;
    if (const char* dm = getenv("DIFFOPT_B200_DENSE_MAX")) dense_max = atoll(dm);
    if (N > dense_max) {
        std::vector<int64_t> hc, hr;
        std::vector<double> hv;
        const int64_t *pc = colptr, *pr = rowval;
        const double* pvv = nzval;
        if (memspace == DIFFOPT_B200_DEVICE) {  // the analysis runs on the host: fetch the matrix
            hc.resize((size_t)N + 1);
            DO_CUDA(ctx, cudaMemcpy(hc.data(), colptr, sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyDeviceToHost));
            const int64_t nz = hc[(size_t)N] - 1;
            if (nz < 0) BAD_ARG(ctx, "kkt_solve_csc: colptr must be 1-based");
            hr.resize((size_t)nz);
            hv.resize((size_t)nz);
            DO_CUDA(ctx, cudaMemcpy(hr.data(), rowval, sizeof(int64_t) * (size_t)nz, cudaMemcpyDeviceToHost));
            DO_CUDA(ctx, cudaMemcpy(hv.data(), nzval, sizeof(double) * (size_t)nz, cudaMemcpyDeviceToHost));
            pc = hc.data(); pr = hr.data(); pvv = hv.data();
        }
        int32_t rc = diffopt_b200_sparse_setup(ctx, N, pc, pr, pvv, trans, nullptr);
        if (rc != 0) return rc;
        const double factor_ms = ctx->last_ms;
        rc = diffopt_b200_sparse_solve(ctx, nrhs, rhs, x_out, memspace);
        ctx->last_ms += factor_ms;
        return rc;
    }
    int64_t nnz = 0;
    std::vector<int64_t> hcol;
    const void *dcol = nullptr, *drow = nullptr, *dval = nullptr;
    if (memspace == DIFFOPT_B200_HOST) {
        nnz = colptr[N] - 1;
    } else {
        int64_t last = 0;
        DO_CUDA(ctx, cudaMemcpyAsync(&last, colptr + N, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        nnz = last - 1;
    }
    if (nnz < 0) BAD_ARG(ctx, "kkt_solve_csc: colptr must be 1-based");
    DO_CUDA(ctx, stage_in(ctx, ctx->in[0], colptr, sizeof(int64_t) * (size_t)(N + 1), memspace, &dcol));
    DO_CUDA(ctx, stage_in(ctx, ctx->in[1], rowval, sizeof(int64_t) * (size_t)nnz, memspace, &drow));
    DO_CUDA(ctx, stage_in(ctx, ctx->in[2], nzval, sizeof(double) * (size_t)nnz, memspace, &dval));
    const size_t mbytes = sizeof(double) * (size_t)N * (size_t)(N + nrhs);
    DO_CUDA(ctx, ctx->in[3].reserve(mbytes));
    double* M = ctx->in[3].as<double>();
    DO_CUDA(ctx, cudaMemsetAsync(M, 0, sizeof(double) * (size_t)N * (size_t)N, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(M + (size_t)N * N, rhs, sizeof(double) * (size_t)N * (size_t)nrhs,
                                 memspace == DIFFOPT_B200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    DO_CUDA(ctx, ctx->info.reserve(sizeof(int)));
    int* dinfo = ctx->info.as<int>();
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (nnz > 0) {
        const int64_t blocks = (N * 32 + 255) / 256;
        csc_scatter_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(N, N, (const int64_t*)dcol, (const int64_t*)drow,
                                                                      (const double*)dval, trans, M);
        ctx->launches++;
    }
    dense_lu_solve_kernel<<<1, LU_THREADS, 0, ctx->stream>>>((int)N, (int)nrhs, M, dinfo);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    int hinfo = 0;
    DO_CUDA(ctx, cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(x_out, M + (size_t)N * N, sizeof(double) * (size_t)N * (size_t)nrhs,
                                 memspace == DIFFOPT_B200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    return hinfo;
}

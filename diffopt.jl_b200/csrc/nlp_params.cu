// The two callers next to the hot path, on top of the kernels of sparse_mf.cu / qp_ex.cu:
//
//  * diffopt_b200_sparse_setup_inertia -- the NonLinearProgram backend's factorisation
//    (`_lu_with_inertia_correction` / `_inertia_correction`, NonLinearProgram.jl:356-435): K = lu(M); when that reports
//    a singular matrix, J = M + c st D with D = diag(+1 on the first num_w rows, -1 on the next num_cons rows, +1 on
//    the rest) for c = 1, 2, ... until the factorisation succeeds (at most max_corrections corrected factorisations).
//    `K \ N` for all parameter columns at once (nlp_utilities.jl:436-444, `ldiv!(ds, K, N)`) is diffopt_b200_sparse_solve.
//
//  * diffopt_b200_param_pullback -- reverse-mode accumulation into PARAMETERS (src/parameters.jl:341-534): every
//    parametric term contributes  coefficient x (gradient entry)  to its parameter, where the gradient entry is a
//    constant (affine p terms), a coefficient (p v terms) or a constant times the other parameter's value (p p terms,
//    folded into the coefficient by the caller).  In array form: out[param] += sum over terms coef * flat[index] with
//    `flat` the (batch-summed, all-reduced) gradient block of diffopt_b200_qp_batch_shared_grads -- one deterministic
//    sparse product on the device, so the layer's parameter gradient never visits the host in between.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

// terms sorted by parameter (host side): thread p sums its terms in storage order
__global__ void param_pullback_kernel(const int64_t nparams, const int64_t* __restrict__ ptr, const int64_t* __restrict__ idx,
                                      const double* __restrict__ coef, const double* __restrict__ flat, double* __restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nparams; p += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int64_t e = ptr[p]; e < ptr[p + 1]; ++e) acc = fma(coef[e], flat[idx[e]], acc);
        out[p] = acc;
    }
}

}  // namespace

extern "C" {

int32_t diffopt_b200_sparse_setup_inertia(diffopt_b200_ctx* ctx, int64_t N, const int64_t* colptr, const int64_t* rowval,
                                          const double* nzval, int64_t num_w, int64_t num_cons, double st, int32_t max_corrections,
                                          int32_t* corrections_out) {
    if (!ctx) return -1;
    if (corrections_out) *corrections_out = 0;
    if (N <= 0 || !colptr || !rowval || !nzval) BAD_ARG(ctx, "sparse_setup_inertia: bad argument");
    if (num_w < 0 || num_cons < 0 || num_w + num_cons > N) BAD_ARG(ctx, "sparse_setup_inertia: num_w + num_cons exceeds N");
    int32_t rc = diffopt_b200_sparse_setup(ctx, N, colptr, rowval, nzval, 0, nullptr);   // K = lu(M; check = false)
    if (rc <= 0) return rc;                                                                // fine, or an error
    // status == 1 in the reference: singular.  J = M + c st D needs every diagonal entry in the pattern.
    const int64_t nnz = colptr[N] - 1;
    std::vector<int64_t> cp((size_t)N + 1), rv;
    std::vector<double> nz;
    std::vector<int64_t> dpos((size_t)N);
    rv.reserve((size_t)(nnz + N));
    nz.reserve((size_t)(nnz + N));
    cp[0] = 1;
    for (int64_t j = 0; j < N; ++j) {
        bool placed = false;
        for (int64_t e = colptr[j] - 1; e < colptr[j + 1] - 1; ++e) {
            const int64_t r = rowval[e] - 1;
            if (!placed && r >= j) {
                if (r != j) {   // the diagonal entry is not stored: insert an explicit zero in front of row r
                    dpos[(size_t)j] = (int64_t)rv.size();
                    rv.push_back(j + 1);
                    nz.push_back(0.0);
                } else dpos[(size_t)j] = (int64_t)rv.size();
                placed = true;
            }
            rv.push_back(r + 1);
            nz.push_back(nzval[e]);
        }
        if (!placed) {
            dpos[(size_t)j] = (int64_t)rv.size();
            rv.push_back(j + 1);
            nz.push_back(0.0);
        }
        cp[(size_t)j + 1] = (int64_t)rv.size() + 1;
    }
    int32_t num_c = 0;
    while (rc > 0 && num_c < max_corrections) {
        for (int64_t j = 0; j < N; ++j) nz[(size_t)dpos[(size_t)j]] += (j >= num_w && j < num_w + num_cons) ? -st : st;
        ++num_c;
        rc = diffopt_b200_sparse_setup(ctx, N, cp.data(), rv.data(), nz.data(), 0, nullptr);
    }
    if (corrections_out) *corrections_out = num_c;
    if (rc > 0) ctx->err = "sparse_setup_inertia: still singular after the allowed corrections (reference: \"Inertia correction failed.\")";
    return rc;
}

int32_t diffopt_b200_param_pullback(diffopt_b200_ctx* ctx, int64_t nterms, const int64_t* term_param, const int64_t* term_index,
                                    const double* term_coef, int64_t nflat, const double* flat, int64_t nparams, double* out,
                                    int32_t memspace) {
    if (!ctx) return -1;
    DeviceGuard guard_(ctx->device);
    if (nterms < 0 || nparams <= 0 || nflat <= 0 || !flat || !out || (nterms > 0 && (!term_param || !term_index || !term_coef)))
        BAD_ARG(ctx, "param_pullback: bad argument");
    // term lists are host arrays (they describe the model, not the data): bucket them by parameter, keeping their order
    std::vector<int64_t> ptr((size_t)nparams + 1, 0), idx((size_t)nterms);
    std::vector<double> coef((size_t)nterms);
    for (int64_t e = 0; e < nterms; ++e) {
        const int64_t p = term_param[e] - 1, i = term_index[e] - 1;   // 1-based like every index at this boundary
        if (p < 0 || p >= nparams || i < 0 || i >= nflat) BAD_ARG(ctx, "param_pullback: term index out of range");
        ++ptr[(size_t)p + 1];
    }
    for (int64_t p = 0; p < nparams; ++p) ptr[(size_t)p + 1] += ptr[(size_t)p];
    {
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t e = 0; e < nterms; ++e) {
            const int64_t p = term_param[e] - 1, at = fill[(size_t)p]++;
            idx[(size_t)at] = term_index[e] - 1;
            coef[(size_t)at] = term_coef[e];
        }
    }
    const size_t d8 = sizeof(double), i8 = sizeof(int64_t);
    DO_CUDA(ctx, ctx->in[12].reserve(i8 * ((size_t)nparams + 1)));
    DO_CUDA(ctx, ctx->in[13].reserve(i8 * std::max<size_t>((size_t)nterms, 1)));
    DO_CUDA(ctx, ctx->in[14].reserve(d8 * std::max<size_t>((size_t)nterms, 1)));
    DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[12].ptr, ptr.data(), i8 * ((size_t)nparams + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nterms > 0) {
        DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[13].ptr, idx.data(), i8 * (size_t)nterms, cudaMemcpyHostToDevice, ctx->stream));
        DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[14].ptr, coef.data(), d8 * (size_t)nterms, cudaMemcpyHostToDevice, ctx->stream));
    }
    const void* dflat;
    void* dout;
    DO_CUDA(ctx, stage_in(ctx, ctx->in[11], flat, d8 * (size_t)nflat, memspace, &dflat));
    DO_CUDA(ctx, stage_out_prepare(ctx->out[4], out, d8 * (size_t)nparams, memspace, &dout));
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    const int64_t blocks = std::min<int64_t>((nparams + 127) / 128, (int64_t)ctx->sm_count * 8);
    param_pullback_kernel<<<(unsigned)blocks, 128, 0, ctx->stream>>>(nparams, ctx->in[12].as<int64_t>(), ctx->in[13].as<int64_t>(),
                                                                     ctx->in[14].as<double>(), (const double*)dflat, (double*)dout);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    DO_CUDA(ctx, stage_out_finish(ctx, dout, out, d8 * (size_t)nparams, memspace));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host term vectors die at return
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    return 0;
}

}  // extern "C"

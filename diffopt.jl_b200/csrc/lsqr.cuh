// Persistent cooperative LSQR (Paige-Saunders) over two kinds of linear operator:
//   OP_CSR   : explicit sparse matrix (the QP backend's `lsqr(LHS, RHS)` branch, QuadraticProgram.jl:488)
//   OP_CONIC : matrix-free M = [0 A'Dpi c; -A I-Dpi b; -c' -b'Dpi 0]  (ConicProgram.jl:243-247), used by
//              `lsqr(M, g)` at ConicProgram.jl:323,372 -- the dense Dpi blocks are never formed.
// One kernel runs the whole iteration: vectors stay in L2/HBM, Golub-Kahan scalars live replicated in
// registers, norms are reduced deterministically (per-CTA partials + fixed-order sum), grid.sync()
// separates the phases; there is no host round-trip per iteration.
#pragma once
#include "common.cuh"

struct CsrView {
    int nrows, ncols;
    const int* rowptr;
    const int* colind;
    const double* val;
    const int* blk;  // row blocks of the streaming SpMV (nblk + 1 entries), see CsrDev
    int nblk;
    // column-ordered copy of every streaming block and each nonzero's CSR slot inside its block (or null)
    const double* sval;
    const int* scol;
    const unsigned short* spos;
};

struct ConicOpView {
    int n, m;
    CsrView A;   // m x n
    CsrView At;  // n x m
    const double *b, *c;
    const double* diag;       // per row: 1 (zero cone -> dual is free), (sign(v)+1)/2 (nonneg); unused on SOC/PSD rows
    const signed char* kind;  // per row: 0 diag row, 2 SOC row, 3 PSD row
    int nsoc;
    const int* soc_off;   // first row of each SOC cone
    const int* soc_dim;
    const int* soc_case;  // 0: identity (|x| <= t), 1: zero (|x| <= -t), 2: general
    const double* soc_nx; // |x|
    const double* v;      // y - s (length m)
    // PSD cones (operator form: U (B o (U' X U)) U')
    int npsd;
    const int* psd_off;        // first row
    const int* psd_d;          // side
    const long long* psd_uoff; // offset (in doubles) into U / Bm / work
    const double* psd_U;       // eigenvectors, column-major d x d per cone
    const double* psd_Bm;      // B matrix, d x d per cone
    const int* psd_ident;      // 1 if all eigenvalues >= 0 (Dpi = I)
    const int* psd_toff;       // first 32 x 32 output tile of each cone in the flattened tile list (npsd + 1 entries)
    int psd_ntiles;
    double* psd_w0;            // 3 scratch matrices per cone (d x d each)
    double* psd_w1;
    double* psd_w2;
    double* wc;   // scratch, length m : Dpi * t2   (forward)  /  r = A u1 - u2 - b u3 (transpose)
};

struct LsqrParams {
    double atol, btol, conlim;
    long long maxiter;
};

struct LsqrVectors {
    double *u, *v, *w, *x;  // lengths nrows, ncols, ncols, ncols
    double* partials;       // [NSLOTS][gridDim.x]
    double* stats;          // out: istop, itn, rnorm, arnorm, anorm, acond, xnorm
};

int32_t lsqr_run_csr(diffopt_b200_ctx* ctx, const CsrDev& M, bool trans, const double* rhs_dev, LsqrParams prm,
                     double* x_dev, double* stats_host7);
int32_t lsqr_run_conic(diffopt_b200_ctx* ctx, const double* rhs_dev, LsqrParams prm, double* x_dev,
                       double* stats_host7);
int32_t conic_apply_M(diffopt_b200_ctx* ctx, const double* t_dev, bool transpose, double* out_dev);
int32_t conic_apply_dpi(diffopt_b200_ctx* ctx, const double* t_dev, bool transpose, double* out_dev);
int32_t lsqr_run_conic_batch(diffopt_b200_ctx* ctx, std::vector<ConicState>& states, int C, bool stage, double* work, size_t stride,
                             LsqrParams prm, double zero_below, DevBuf& ops_buf, DevBuf& vecs_buf, DevBuf& rhs_buf);
constexpr int LSQR_NSLOTS = 8;  // == NSLOTS of lsqr.cu (partial-sum slots per CTA)
int32_t csr_from_csc_host(diffopt_b200_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* colptr,
                          const int64_t* rowval, const double* nzval, CsrDev& out);

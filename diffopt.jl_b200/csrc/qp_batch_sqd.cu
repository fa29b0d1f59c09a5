// Fast path of the batched KKT sensitivity solve for the headline shape n=64, m=64, p=16:
// pivot-free blocked LDL' of the *symmetric quasi-definite* form of the reduced KKT system.
//
// The reference solves with LHS = [Q G'diag(lam) A'; G diag(Gz-h) 0; A 0 0] (QuadraticProgram.jl:256-282):
// reverse mode LHS x = [dl_dz;0;0] (:324-335), forward mode LHS' x = rhs (:429-438), outputs -x.
//   1. Column singletons (what the reference's sparse `\` removes in its preprocessing): lam_i == 0 makes
//      column n+i of LHS a singleton (only D_i = (Gz-h)_i), so that unknown decouples exactly:
//        LHS  x = r : x_lam_i = (r_i - G_i x_z) / D_i      LHS' x = r : x_lam_i = r_i / D_i = 0
//   2. On the remaining unknowns (z, ACTIVE inequalities a, equalities) both systems are the SYMMETRIC matrix
//        Ks = [Q Ga' A'; Ga diag(D_a/lam_a) 0; A 0 0]
//      reverse: Ks [x_z; lam_a.x_lam_a; x_nu] = [dl_dz;0;0]       (column scaling by lam_a)
//      forward: Ks [x_z; x_lam_a; x_nu] = [r_z; (dG z - dh)_a; dA z - db]   (row scaling by 1/lam_a)
//      With Q > 0, D_a/lam_a <= 0 and [Ga; A] of full row rank, Ks is quasi-definite: LDL' exists for the
//      natural order with d_k > 0 on the z block and d_k < 0 on the rest -- no pivoting, half the flops of LU.
//   3. Every pivot is checked (expected sign, |d_k| > 1e-12 x its starting diagonal); an instance that fails
//      (Q only semidefinite, rank-deficient active set, wrong-sign duals, lam_i = D_i = 0) is appended to a
//      device list and re-solved by the partially pivoted LU kernel (qp_batch_n144.cu), so results never
//      depend on this path's assumptions.  Q is read from its lower triangle (the reference's Q is symmetric,
//      utils.jl:46-69).
//
// Execution: 128-thread CTAs, four resident per SM, each streaming over instances.  The reduced matrix lives in
// shared memory as the lower triangle of an nt x nt grid of 8 x 8 row-major tiles (40 KB at 16 active rows); the
// blocked right-looking LDL' (block 8) does the diagonal block in registers of one warp, the panel
// W = A21 L11^-T and the trailing update C -= W D^-1 W' on the FP64 tensor pipe (mma.sync m8n8k4 DMMA), with
// both right-hand sides riding along.  Inputs are read once from HBM (tile-shaped 16-byte loads, next instance
// prefetched to L2); the KKT matrix and its factors never touch HBM.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

int32_t qp_lu_launch_list(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, const int* list, const int* count);

namespace {

constexpr int NV = 64, MI = 64, PE = 16, N = NV + MI + PE, NTZ = NV / 8;
constexpr int THREADS = 128, NWARP = THREADS / 32;
constexpr unsigned FULL = 0xffffffffu;
constexpr double PIV_RTOL = 1e-12;

struct __align__(16) Hdr {
    double yf[N + 8], yb[N + 8];  // right-hand sides -> v = D^-1 L^-1 r -> solutions (reduced ordering)
    double sf[N + 8], sb[N + 8];  // backward-substitution accumulators
    double ref[N + 8];            // starting magnitude of every diagonal entry (pivot test)
    double rd[N + 8];             // 1 / d_k
    double zs[NV], lams[MI], nus[PE], dvec[MI];
    double rowq[NV], rowg[MI];    // dQ z, dG z
    double gcol[NWARP][NV];       // per-warp partial column sums of dG .* lam
    double acol[NV];              // column sums of dA .* nu
    double arow[NWARP][PE];       // per-warp partial row sums of dA z
    int apos[MI];                 // inequality -> slot among the active ones, or -1
    int ma, nt, fail;
};

__device__ __forceinline__ int tix(int I, int J) { return ((I * (I + 1)) >> 1) + J; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// 1/x to ~1 ulp: hardware estimate + two Newton steps (x is a checked pivot, never 0/inf/nan)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__device__ __forceinline__ double sum_over_g(double v) {  // lanes sharing t = lane & 3
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 16);
    return v;
}
__device__ __forceinline__ double sum_over_t(double v) {  // lanes sharing g = lane >> 2
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    return v;
}

// LDL' of the 8 x 8 diagonal block by warp 0: lane 0 eliminates in registers, then lanes 0..7 each build one
// column of inv(L11) (unit lower).  On exit the tile holds inv(L11) (zeros above the diagonal), rd[] = 1/d.
__device__ __forceinline__ void diag_block(double* D, const double* ref, double* rd, const bool positive, int* fail,
                                           const int lane) {
    if (lane == 0) {
        double a[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                if (c <= i) {
                    const double2 v = *reinterpret_cast<const double2*>(&D[i * 8 + c]);
                    a[i][c] = v.x;
                    a[i][c + 1] = v.y;
                }
            }
        }
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double d = a[k][k];
            const double mag = positive ? d : -d;
            if (!(mag > PIV_RTOL * ref[k])) {
                bad = true;
                d = positive ? 1.0 : -1.0;
            }
            const double r = fast_rcp(d);
            rd[k] = r;
            double lk[8];
#pragma unroll
            for (int i = k + 1; i < 8; ++i) lk[i] = a[i][k] * r;
#pragma unroll
            for (int i = k + 1; i < 8; ++i) {
#pragma unroll
                for (int c = k + 1; c <= i; ++c) a[i][c] = fma(-lk[i], a[c][k], a[i][c]);
            }
#pragma unroll
            for (int i = k + 1; i < 8; ++i) a[i][k] = lk[i];
        }
#pragma unroll
        for (int i = 1; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < i; ++c) D[i * 8 + c] = a[i][c];
        if (bad) *fail = 1;
    }
    __syncwarp();
    double x[8];
    if (lane < 8) {
        double l[8][8];
#pragma unroll
        for (int i = 1; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < i; ++c) l[i][c] = D[i * 8 + c];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) s = fma(-l[i][k], x[k], s);
            x[i] = (i >= lane) ? s : 0.0;
        }
    }
    __syncwarp();
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) D[i * 8 + lane] = x[i];
    }
}

__global__ void __launch_bounds__(THREADS, 4) qp_kkt_sqd_kernel(QpSolveArgs a, int* fb_list, int* fb_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Hdr& S = *reinterpret_cast<Hdr*>(smem_raw);
    double* const T = reinterpret_cast<double*>(smem_raw + sizeof(Hdr));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // g, t: DMMA fragment coordinates; the same pair addresses the tile-shaped global loads (column g, rows 2t, 2t+1)
    const int g = lane >> 2, t = lane & 3;
    const bool do_fwd = a.fwd != nullptr, do_rev = a.rev != nullptr;
#ifdef QP_PROFILE
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#define PROF(i)                  \
    do {                         \
        long long _n = clock64(); \
        pc[i] += _n - tprev;     \
        tprev = _n;              \
    } while (0)
#else
#define PROF(i)
#endif

    for (int64_t inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
        const double* Q = a.Q + (size_t)inst * NV * NV;
        const double* G = a.G + (size_t)inst * MI * NV;
        const double* A = a.A + (size_t)inst * PE * NV;
        // ---- vectors, active set
        if (tid < NV) {
            S.zs[tid] = a.z[(size_t)inst * NV + tid];
            S.yb[tid] = do_rev ? a.seed[(size_t)inst * NV + tid] : 0.0;
        } else if (tid < NV + PE) {
            S.nus[tid - NV] = a.nu[(size_t)inst * PE + tid - NV];
        } else if (warp == 3) {
            const double l0 = a.lam[(size_t)inst * MI + lane], l1 = a.lam[(size_t)inst * MI + 32 + lane];
            S.lams[lane] = l0;
            S.lams[32 + lane] = l1;
            const unsigned m0 = __ballot_sync(FULL, l0 != 0.0), m1 = __ballot_sync(FULL, l1 != 0.0);
            const unsigned lt = (1u << lane) - 1u;
            S.apos[lane] = (l0 != 0.0) ? __popc(m0 & lt) : -1;
            S.apos[32 + lane] = (l1 != 0.0) ? __popc(m0) + __popc(m1 & lt) : -1;
            if (lane == 0) {
                const int ma = __popc(m0) + __popc(m1);
                S.ma = ma;
                S.nt = (NV + ma + PE + 7) >> 3;
                S.fail = 0;
            }
        }
        // Q: lower-triangle tiles only; warp w owns tile rows w and 7-w (9 tiles each)
        double2 qv[16];
#pragma unroll
        for (int J = 0; J < 8; ++J) {
            if (J <= warp) qv[J] = ldg2(Q + (8 * J + g) * NV + 8 * warp + 2 * t);
            if (J <= 7 - warp) qv[8 + J] = ldg2(Q + (8 * J + g) * NV + 8 * (7 - warp) + 2 * t);
        }
        {   // L2 prefetch of the next instance of this CTA (inputs are streamed once from HBM)
            const int64_t nxt = inst + gridDim.x;
            if (nxt < a.B) {
                const char* bases[6] = {(const char*)(a.Q + (size_t)nxt * NV * NV), (const char*)(a.G + (size_t)nxt * MI * NV),
                                        do_fwd && a.dQ ? (const char*)(a.dQ + (size_t)nxt * NV * NV) : nullptr,
                                        do_fwd && a.dG ? (const char*)(a.dG + (size_t)nxt * MI * NV) : nullptr,
                                        (const char*)(a.A + (size_t)nxt * PE * NV),
                                        do_fwd && a.dA ? (const char*)(a.dA + (size_t)nxt * PE * NV) : nullptr};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    if (!bases[q]) continue;
                    const int lines = q < 4 ? 256 : 64;  // 128-byte lines
                    for (int l = tid; l < lines; l += THREADS)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(bases[q] + (size_t)l * 128));
                }
            }
        }
        __syncthreads();
        const int ma = S.ma, nt = S.nt;
        const int nred = NV + ma + PE, np = nt << 3;
        // ---- clear the tile rows below the z block, then fill
        {
            double2* zp = reinterpret_cast<double2*>(T + tix(NTZ, 0) * 64);
            const int cnt = (tix(nt, 0) - tix(NTZ, 0)) * 32;
            for (int i = tid; i < cnt; i += THREADS) zp[i] = make_double2(0.0, 0.0);
        }
        if (tid < np - NV) {
            S.yb[NV + tid] = 0.0;
            S.yf[NV + tid] = 0.0;
        }
        for (int i = tid; i < np; i += THREADS) {
            S.sf[i] = 0.0;
            S.sb[i] = 0.0;
        }
#pragma unroll
        for (int J = 0; J < 8; ++J) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int I = hh ? 7 - warp : warp;
                if (J <= I) {
                    const double2 v = qv[8 * hh + J];
                    double* tp = T + tix(I, J) * 64 + (2 * t) * 8 + g;
                    tp[0] = v.x;
                    tp[8] = v.y;
                    if (J == I) {
                        if (g == 2 * t) S.ref[8 * I + g] = v.x;
                        if (g == 2 * t + 1) S.ref[8 * I + g] = v.y;
                    }
                }
            }
        }
        __syncthreads();
        if (tid >= nred && tid < np) {  // identity padding (negative block: -1)
            T[tix(nt - 1, nt - 1) * 64 + (tid & 7) * 9] = -1.0;
        }
        // ---- G: D = G z - h for every row; active rows go to the matrix; warp w owns tile rows w, w+4
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int I = warp + 4 * hh, r0 = 8 * I + 2 * t;
            double2 gv[8];
#pragma unroll
            for (int J = 0; J < 8; ++J) gv[J] = ldg2(G + (8 * J + g) * MI + r0);
            const double h0 = (g == 0) ? a.h[(size_t)inst * MI + r0] : 0.0, h1 = (g == 0) ? a.h[(size_t)inst * MI + r0 + 1] : 0.0;
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int J = 0; J < 8; ++J) {
                const double zc = S.zs[8 * J + g];
                d0 = fma(gv[J].x, zc, d0);
                d1 = fma(gv[J].y, zc, d1);
            }
            d0 = sum_over_g(d0);
            d1 = sum_over_g(d1);
            const int a0 = S.apos[r0], a1 = S.apos[r0 + 1];
            if (a0 >= 0) {
                double* tp = T + tix(NTZ + (a0 >> 3), 0) * 64 + (a0 & 7) * 8 + g;
#pragma unroll
                for (int J = 0; J < 8; ++J) tp[J * 64] = gv[J].x;
            }
            if (a1 >= 0) {
                double* tp = T + tix(NTZ + (a1 >> 3), 0) * 64 + (a1 & 7) * 8 + g;
#pragma unroll
                for (int J = 0; J < 8; ++J) tp[J * 64] = gv[J].y;
            }
            if (g == 0) {
                d0 -= h0;
                d1 -= h1;
                S.dvec[r0] = d0;
                S.dvec[r0 + 1] = d1;
                if (a0 >= 0) T[tix(NTZ + (a0 >> 3), NTZ + (a0 >> 3)) * 64 + (a0 & 7) * 9] = d0 / S.lams[r0];
                else if (d0 == 0.0) S.fail = 1;  // lam_i = D_i = 0: singular column, the LU path reports it
                if (a1 >= 0) T[tix(NTZ + (a1 >> 3), NTZ + (a1 >> 3)) * 64 + (a1 & 7) * 9] = d1 / S.lams[r0 + 1];
                else if (d1 == 0.0) S.fail = 1;
            }
        }
        // ---- A (16 x 64): warp w owns tile columns 2w, 2w+1
        {
            double2 av[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int J = 2 * warp + (q >> 1), Ia = q & 1;
                av[q] = ldg2(A + (8 * J + g) * PE + 8 * Ia + 2 * t);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int J = 2 * warp + (q >> 1), Ia = q & 1;
                const int k0 = NV + ma + 8 * Ia + 2 * t, k1 = k0 + 1;
                T[tix(k0 >> 3, J) * 64 + (k0 & 7) * 8 + g] = av[q].x;
                T[tix(k1 >> 3, J) * 64 + (k1 & 7) * 8 + g] = av[q].y;
            }
        }
        // ---- forward right-hand side (QuadraticProgram.jl:429-433), symmetric-form scaling
        if (do_fwd) {
            const size_t b = (size_t)inst;
            if (a.dQ) {
                const double* X = a.dQ + b * NV * NV;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int r0 = 8 * (warp + 4 * hh) + 2 * t;
                    double2 v[8];
#pragma unroll
                    for (int J = 0; J < 8; ++J) v[J] = ldg2(X + (8 * J + g) * NV + r0);
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int J = 0; J < 8; ++J) {
                        const double zc = S.zs[8 * J + g];
                        s0 = fma(v[J].x, zc, s0);
                        s1 = fma(v[J].y, zc, s1);
                    }
                    s0 = sum_over_g(s0);
                    s1 = sum_over_g(s1);
                    if (g == 0) {
                        S.rowq[r0] = s0;
                        S.rowq[r0 + 1] = s1;
                    }
                }
            }
            if (a.dG) {
                const double* X = a.dG + b * MI * NV;
                double cs[8];
#pragma unroll
                for (int J = 0; J < 8; ++J) cs[J] = 0.0;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int r0 = 8 * (warp + 4 * hh) + 2 * t;
                    double2 v[8];
#pragma unroll
                    for (int J = 0; J < 8; ++J) v[J] = ldg2(X + (8 * J + g) * MI + r0);
                    const double l0 = S.lams[r0], l1 = S.lams[r0 + 1];
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int J = 0; J < 8; ++J) {
                        const double zc = S.zs[8 * J + g];
                        s0 = fma(v[J].x, zc, s0);
                        s1 = fma(v[J].y, zc, s1);
                        cs[J] = fma(v[J].x, l0, cs[J]);
                        cs[J] = fma(v[J].y, l1, cs[J]);
                    }
                    s0 = sum_over_g(s0);
                    s1 = sum_over_g(s1);
                    if (g == 0) {
                        S.rowg[r0] = s0;
                        S.rowg[r0 + 1] = s1;
                    }
                }
#pragma unroll
                for (int J = 0; J < 8; ++J) {
                    const double c = sum_over_t(cs[J]);
                    if (t == 0) S.gcol[warp][8 * J + g] = c;
                }
            }
            if (a.dA) {
                const double* X = a.dA + b * PE * NV;
                double2 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int J = 2 * warp + (q >> 1), Ia = q & 1;
                    v[q] = ldg2(X + (8 * J + g) * PE + 8 * Ia + 2 * t);
                }
#pragma unroll
                for (int Ia = 0; Ia < 2; ++Ia) {  // partial row sums over this warp's 16 columns
                    const double z0 = S.zs[16 * warp + g], z1 = S.zs[16 * warp + 8 + g];
                    double s0 = fma(v[Ia].x, z0, v[2 + Ia].x * z1), s1 = fma(v[Ia].y, z0, v[2 + Ia].y * z1);
                    s0 = sum_over_g(s0);
                    s1 = sum_over_g(s1);
                    if (g == 0) {
                        S.arow[warp][8 * Ia + 2 * t] = s0;
                        S.arow[warp][8 * Ia + 2 * t + 1] = s1;
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {  // full column sums of dA .* nu
                    double c = v[2 * jj].x * S.nus[2 * t] + v[2 * jj].y * S.nus[2 * t + 1] + v[2 * jj + 1].x * S.nus[8 + 2 * t] +
                               v[2 * jj + 1].y * S.nus[8 + 2 * t + 1];
                    c = sum_over_t(c);
                    if (t == 0) S.acol[16 * warp + 8 * jj + g] = c;
                }
            }
        }
        __syncthreads();
        if (tid < NV) {
            double v = 0.0;
            if (do_fwd) {
                const size_t b = (size_t)inst;
                if (a.dQ) v += S.rowq[tid];
                if (a.dq) v += a.dq[b * NV + tid];
                if (a.dG) v += (S.gcol[0][tid] + S.gcol[1][tid]) + (S.gcol[2][tid] + S.gcol[3][tid]);
                if (a.dA) v += S.acol[tid];
            }
            S.yf[tid] = v;
        } else if (do_fwd) {
            const size_t b = (size_t)inst;
            const int i = tid - NV, ar = S.apos[i];
            if (ar >= 0) {
                double v = a.dG ? S.rowg[i] : 0.0;
                if (a.dh) v -= a.dh[b * MI + i];
                S.yf[NV + ar] = v;
            }
            if (i < PE) {
                double v = a.dA ? (S.arow[0][i] + S.arow[1][i]) + (S.arow[2][i] + S.arow[3][i]) : 0.0;
                if (a.db) v -= a.db[b * PE + i];
                S.yf[NV + ma + i] = v;
            }
        }
        __syncthreads();
        PROF(0);

        // ---- blocked LDL' (block 8), both right-hand sides riding along
        for (int j = 0; j < nt; ++j) {
            const int c0 = j << 3;
            if (j == NTZ) {  // the Schur complement of the z block is complete: record its diagonal
                if (tid < np - NV) S.ref[NV + tid] = fabs(T[tix(NTZ + (tid >> 3), NTZ + (tid >> 3)) * 64 + (tid & 7) * 9]);
                __syncthreads();
            }
            double* Dt = T + tix(j, j) * 64;
            if (warp == 0) diag_block(Dt, &S.ref[c0], &S.rd[c0], j < NTZ, &S.fail, lane);
            __syncthreads();
            PROF(1);
            // panel: W(I,j) = A(I,j) inv(L11)'  (DMMA);  v_j = D^-1 inv(L11) r_j
            {
                const double2 bf = *reinterpret_cast<const double2*>(&Dt[g * 8 + 2 * t]);
                for (int I = j + 1 + warp; I < nt; I += NWARP) {
                    double2* tp = reinterpret_cast<double2*>(T + tix(I, j) * 64 + g * 8 + 2 * t);
                    const double2 af = *tp;
                    double2 c = make_double2(0.0, 0.0);
                    dmma(c.x, c.y, af.x, bf.x);
                    dmma(c.x, c.y, af.y, bf.y);
                    *tp = c;
                }
                if (warp == NWARP - 1) {
                    double* y = (lane & 8) ? S.yb : S.yf;
                    const int i = lane & 7;
                    double u = 0.0;
                    if (lane < 16) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) u = fma(Dt[i * 8 + k], y[c0 + k], u);
                        u *= S.rd[c0 + i];
                    }
                    __syncwarp();
                    if (lane < 16) y[c0 + i] = u;
                }
            }
            __syncthreads();
            PROF(2);
            // trailing update C(I,K) -= W(I,j) D^-1 W(K,j)' on the DMMA pipe; right-hand sides r_I -= W(I,j) v_j
            if (j < nt - 1) {
                for (int rr = c0 + 8 + tid; rr < np; rr += THREADS) {
                    const double2* wrow = reinterpret_cast<const double2*>(T + tix(rr >> 3, j) * 64 + (rr & 7) * 8);
                    double uf = S.yf[rr], ub = S.yb[rr];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 w2 = wrow[q];
                        uf = fma(-w2.x, S.yf[c0 + 2 * q], uf);
                        uf = fma(-w2.y, S.yf[c0 + 2 * q + 1], uf);
                        ub = fma(-w2.x, S.yb[c0 + 2 * q], ub);
                        ub = fma(-w2.y, S.yb[c0 + 2 * q + 1], ub);
                    }
                    S.yf[rr] = uf;
                    S.yb[rr] = ub;
                }
                const double nr0 = -S.rd[c0 + 2 * t], nr1 = -S.rd[c0 + 2 * t + 1];
                const int Tn = nt - 1 - j, npairs = (Tn * (Tn + 1)) >> 1;
                int ia = 0, ib = warp;  // pair index -> (ia >= ib) within the trailing triangle
                while (ib > ia) {
                    ib -= ia + 1;
                    ++ia;
                }
                for (int q = warp; q < npairs; q += NWARP) {
                    const int I = j + 1 + ia, K = j + 1 + ib;
                    double2 af = *reinterpret_cast<const double2*>(T + tix(I, j) * 64 + g * 8 + 2 * t);
                    const double2 bf = *reinterpret_cast<const double2*>(T + tix(K, j) * 64 + g * 8 + 2 * t);
                    double2* cp = reinterpret_cast<double2*>(T + tix(I, K) * 64 + g * 8 + 2 * t);
                    double2 c = *cp;
                    dmma(c.x, c.y, af.x * nr0, bf.x);
                    dmma(c.x, c.y, af.y * nr1, bf.y);
                    *cp = c;
                    ib += NWARP;
                    while (ib > ia) {
                        ib -= ia + 1;
                        ++ia;
                    }
                }
            }
            __syncthreads();
            PROF(3);
        }

        // ---- backward substitution L' x = v: warp 0 the forward-mode system, warp 1 the reverse-mode system
        if (warp < 2) {
            double* y = warp ? S.yb : S.yf;
            double* s = warp ? S.sb : S.sf;
            for (int j = nt - 1; j >= 0; --j) {
                const int c0 = j << 3;
                const double* Dt = T + tix(j, j) * 64;
                double xk = 0.0;
                if (lane < 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) xk = fma(Dt[k * 8 + lane], fma(-S.rd[c0 + k], s[c0 + k], y[c0 + k]), xk);
                }
                __syncwarp();
                if (lane < 8) y[c0 + lane] = xk;
                double x8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x8[k] = __shfl_sync(FULL, xk, k);
                for (int c = lane; c < c0; c += 32) {
                    const double* wt = T + tix(j, c >> 3) * 64 + (c & 7);
                    double v = s[c];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v = fma(wt[k * 8], x8[k], v);
                    s[c] = v;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        PROF(4);
        // ---- outputs (dz, dlam, dnu) = -x; inactive inequalities recovered from their singleton columns
        const bool failed = S.fail != 0;
        if (!failed) {
            double* rev = do_rev ? a.rev + (size_t)inst * N : nullptr;
            double* fwd = do_fwd ? a.fwd + (size_t)inst * N : nullptr;
            if (do_rev) {
                // inactive rows: out_lam_i = (G_i . x_z) / D_i   (G re-read: L2 resident)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int I = warp + 4 * hh, r0 = 8 * I + 2 * t;
                    double2 gv[8];
#pragma unroll
                    for (int J = 0; J < 8; ++J) gv[J] = ldg2(G + (8 * J + g) * MI + r0);
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int J = 0; J < 8; ++J) {
                        const double xc = S.yb[8 * J + g];
                        d0 = fma(gv[J].x, xc, d0);
                        d1 = fma(gv[J].y, xc, d1);
                    }
                    d0 = sum_over_g(d0);
                    d1 = sum_over_g(d1);
                    if (g == 0) {
                        const int a0 = S.apos[r0], a1 = S.apos[r0 + 1];
                        rev[NV + r0] = a0 >= 0 ? -S.yb[NV + a0] / S.lams[r0] : d0 / S.dvec[r0];
                        rev[NV + r0 + 1] = a1 >= 0 ? -S.yb[NV + a1] / S.lams[r0 + 1] : d1 / S.dvec[r0 + 1];
                    }
                }
                if (tid < NV) rev[tid] = -S.yb[tid];
                else if (tid < NV + PE) rev[NV + MI + tid - NV] = -S.yb[NV + ma + tid - NV];
            }
            if (do_fwd) {
                if (tid < NV) {
                    fwd[tid] = -S.yf[tid];
                    const int ar = S.apos[tid];
                    fwd[NV + tid] = ar >= 0 ? -S.yf[NV + ar] : 0.0;
                } else if (tid < NV + PE) {
                    fwd[NV + MI + tid - NV] = -S.yf[NV + ma + tid - NV];
                }
            }
            if (a.info && tid == 0) a.info[inst] = 0;
        } else if (tid == 0) {
            fb_list[atomicAdd(fb_count, 1)] = (int)inst;
        }
        __syncthreads();
        PROF(5);
    }
#ifdef QP_PROFILE
    if (a.prof && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) a.prof[i] = pc[i];
#endif
}

}  // namespace

// Launches the LDL' fast path followed by the pivoted-LU kernel over the instances it rejected (device-side
// list, no host round trip).  nt_cap = tile order of the largest reduced system of the batch.
int32_t qp_sqd_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, bool* handled) {
    *handled = false;
    const size_t smem = sizeof(Hdr) + (size_t)(nt_cap * (nt_cap + 1) / 2) * 64 * sizeof(double);
    if (smem > ctx->smem_optin) return 0;
    DO_CUDA(ctx, ctx->qp_fb.reserve(sizeof(int) * ((size_t)a.B + 1)));
    int* fb_count = ctx->qp_fb.as<int>();
    int* fb_list = fb_count + 1;
    DO_CUDA(ctx, cudaMemsetAsync(fb_count, 0, sizeof(int), ctx->stream));
    DO_CUDA(ctx, cudaFuncSetAttribute(qp_kkt_sqd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DO_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_kkt_sqd_kernel, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    QpSolveArgs aa = a;
    const bool profile = getenv("DIFFOPT_B200_PROFILE") != nullptr;
    long long* dprof = nullptr;
    if (profile) {
        DO_CUDA(ctx, cudaMalloc(&dprof, 8 * sizeof(long long)));
        DO_CUDA(ctx, cudaMemsetAsync(dprof, 0, 8 * sizeof(long long), ctx->stream));
        aa.prof = dprof;
    }
    qp_kkt_sqd_kernel<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(aa, fb_list, fb_count);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    if (profile) {
        long long h[8];
        int nfb = 0;
        DO_CUDA(ctx, cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaMemcpyAsync(&nfb, fb_count, sizeof nfb, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dprof);
        long long ninst = (a.B + grid - 1) / grid;
        fprintf(stderr,
                "[qp_sqd profile, CTA 0, %lld instances, %d CTA/SM, nt_cap %d, smem %zu, %d to LU] clocks/instance: "
                "assemble %lld diag %lld panel %lld update %lld backward %lld output %lld\n",
                ninst, per_sm, nt_cap, smem, nfb, h[0] / ninst, h[1] / ninst, h[2] / ninst, h[3] / ninst, h[4] / ninst,
                h[5] / ninst);
    }
    *handled = true;
    return qp_lu_launch_list(ctx, a, nt_cap, fb_list, fb_count);
}

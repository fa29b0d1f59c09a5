// Symmetric eigendecomposition of ONE mid-size PSD block (shared memory of one CTA holds its packed triangle: side
// 112 .. ~220) by the classical direct route instead of Jacobi sweeps -- what LAPACK's dsyevx does, laid out for one GPU:
//
//   1. psd_tridiag_kernel   Householder tridiagonalisation X = Q T Q' (dsytd2), one CTA, the packed lower triangle of X in
//                           shared memory (the cone's own vectorisation IS that layout: no unpacking), warp per row.
//   2. psd_bisect_kernel    all eigenvalues of T by multisection (4 Sturm counts per eigenvalue and step, division-free
//                           three-term recurrence with rescaling), then clusters of numerically coincident eigenvalues.
//   3. psd_invit_kernel     eigenvectors of T by inverse iteration (dstein: LU of T - lam I with partial pivoting, three
//                           solves from a random start), one warp per cluster; only members of one cluster
//                           (gap <= 1e-8 |T|) are orthogonalised against each other, all clusters run side by side.
//   4. psd_backtransform_kernel   U0 = Q Z: one warp per eigenvector, the vector in registers, reflectors streamed from L2.
//   5. psd_polish_*         one Newton-Schulz step U = U0 (3 I - U0'U0) / 2: removes the residual non-orthogonality
//                           eps |T| / gap <= 2e-8 that inverse iteration leaves between close (not clustered) eigenvalues.
//
// The sequential depth is d Householder steps + 26 multisection steps + 3 tridiagonal solves (one-sided Jacobi: ~2000
// dependent tournament rounds).  Reference: `LinearAlgebra.eigen` inside MathOptSetDistances' PSD projection gradient,
// reached from src/diff_opt.jl:509-519.
#include <math.h>

#include "common.cuh"

namespace {

constexpr double EPS = 2.220446049250313e-16;
#ifndef PSD_TD_THREADS
#define PSD_TD_THREADS 512
#endif
constexpr int TD_THREADS = PSD_TD_THREADS, TD_WARPS = TD_THREADS / 32, NQ = 7;  // a lane owns columns j = first + lane + 32 q, q < NQ
constexpr int MAX_SIDE = 32 * NQ;
constexpr double CLUSTER_RTOL = 1e-8;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) x += __shfl_xor_sync(FULL, x, q);
    return x;
}

__device__ __forceinline__ double rcp_newton(double x) {  // 1 / x to an ulp or two, x normal
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ double rsqrt_newton(double x) {  // 1 / sqrt(x) to an ulp or two, x normal
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    y = fma(y, e * fma(0.375, e, 0.5), y);
    e = fma(-x * y, y, 1.0);
    return fma(y, 0.5 * e, y);
}

// ---- 1. tridiagonalisation ------------------------------------------------------------------------------------------
// Shared-memory layout of the lower triangle: row i (i + 1 entries) starts at the EVEN offset roff(i), so a lane reads
// and writes two neighbouring entries with one 16-byte access; the slot after an odd-length row is padding.
__device__ __forceinline__ int roff(int i) {
    const int h = i >> 1;
    return (i & 1) ? 2 * (h + 1) * (h + 1) : 2 * h * (h + 1);
}

constexpr int NQ2 = (MAX_SIDE + 63) / 64;  // a lane owns the column pairs j0 + 2 lane + 64 q, q < NQ2

// H (n x n, column major): column k holds the Householder vector v_k in rows k+1 .. n-1 (v_k[k+1] = 1); tau[k];
// da[n] / eb[n-1]: diagonal and off-diagonal of T.  One step k (dsytd2): v_k from column k, p = tau A22 v_k,
// w_k = p - (tau/2)(p'v_k) v_k, A22 -= v_k w_k' + w_k v_k'.  The rank-2 update of step k-1 is DEFERRED into step k: column k is
// brought up to date on its own (a thread per row), which gives v_k; then ONE pass over the trailing triangle applies the
// pending update and accumulates A22 v_k on the updated entries while they are in registers -- the matrix crosses shared
// memory once per step instead of twice.  A warp takes rows, four at a time, its lanes the column pairs of a row; the
// mirrored half of the symmetric product accumulates per lane (column sums).  The kernel is bound by the dependent chain
// of its 5 block barriers per step on ONE SM.
__global__ void __launch_bounds__(TD_THREADS) psd_tridiag_kernel(const int n, const double* __restrict__ xtri, double* __restrict__ H,
                                                                 double* __restrict__ tau, double* __restrict__ da,
                                                                 double* __restrict__ eb) {
    extern __shared__ __align__(16) double sm[];
    const int nst = roff(n);
    double* A = sm;
    const int ldc = (n + 3) & ~1;      // vector length: n + one zero on either side of the live range, kept even
    double* vbuf = A + nst;            // 2 x ldc: v of the reflector being built / of the pending one
    double* wbuf = vbuf + 2 * ldc;     // 2 x ldc: w likewise
    double* pr = wbuf + 2 * ldc;       // ldc: row sums
    double* red = pr + ldc;            // TD_WARPS
    double* red2 = red + TD_WARPS;     // TD_WARPS
    double* sc = red2 + TD_WARPS;      // 2: diagonal and first sub-diagonal entry of the current column
    double* colpart = sc + 2;          // TD_WARPS x ldc
    int* ro = reinterpret_cast<int*>(colpart + TD_WARPS * ldc);  // row offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += TD_THREADS) ro[i] = roff(i);
    for (int i = tid; i < 2 * ldc; i += TD_THREADS) vbuf[i] = wbuf[i] = 0.0;
    for (int i = warp; i < n; i += TD_WARPS) {
        const double* src = xtri + (size_t)i * (i + 1) / 2;
        double* dst = A + roff(i);
        for (int j = lane; j <= (i | 1); j += 32) dst[j] = j <= i ? src[j] : 0.0;
    }
    __syncthreads();
#ifdef PSD_PROFILE
    long long pc[6] = {0, 0, 0, 0, 0, 0}, tp = clock64();
#define TPROF(i)                  \
    do {                          \
        long long _n = clock64(); \
        pc[i] += _n - tp;         \
        tp = _n;                  \
    } while (0)
#else
#define TPROF(i)
#endif
    int cur = 0;           // vbuf / wbuf half of the reflector being built; the other half holds the pending one
    bool pending = false;  // reflector k-1 not yet applied to rows and columns >= k
    for (int k = 0; k + 2 < n; ++k) {
        const int r0 = k + 1, j0 = r0 & ~1;
        double* const vn = vbuf + cur * ldc;
        double* const wn = wbuf + cur * ldc;
        const double* const vp = vbuf + (cur ^ 1) * ldc;
        const double* const wp = wbuf + (cur ^ 1) * ldc;
        // ---- column k brought up to date (pending reflector: v_{k-1}[k] = 1), its norm below the first sub-diagonal
        const int ic = k + tid;
        double xcol = 0.0;
        {
            double sq = 0.0;
            if (ic < n) {
                double* pa = A + ro[ic] + k;
                xcol = *pa;
                if (pending) {
                    xcol -= fma(vp[ic], wp[k], wp[ic] * vp[k]);
                    *pa = xcol;
                }
                if (ic >= k + 2) sq = xcol * xcol;
                else sc[ic - k] = xcol;
            }
            sq = warp_sum(sq);
            if (lane == 0) red2[warp] = sq;
        }
        __syncthreads();
        double ss = 0.0;
        {
            double s4[4] = {0.0, 0.0, 0.0, 0.0};  // four short chains instead of one of TD_WARPS additions
#pragma unroll
            for (int w = 0; w < TD_WARPS; ++w) s4[w & 3] += red2[w];
            ss = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        const double alpha = sc[1];
        double beta = alpha, t = 0.0, scal = 0.0;
        if (ss != 0.0) {
            const double h = fma(alpha, alpha, ss);
            if (h > 1e-280 && h < 1e280) {  // MUFU seeds + Newton steps: the scalar chain is on every step's critical path
                beta = -copysign(h * rsqrt_newton(h), alpha);
                t = (beta - alpha) * rcp_newton(beta);
                scal = rcp_newton(alpha - beta);
            } else {
                beta = -copysign(sqrt(h), alpha);
                t = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
        }
        if (ic < n + (n & 1)) {  // v_k, zero outside rows r0 .. n-1 (the pair grid starts at j0 >= k and ends at an even index)
            const double v = (ic < r0 || ic >= n) ? 0.0 : (ic == r0 ? 1.0 : xcol * scal);
            vn[ic] = v;
            if (ic >= r0 && ic < n) H[(size_t)k * n + ic] = v;
        }
        if (tid == 0) {
            tau[k] = t;
            eb[k] = beta;
            da[k] = sc[0];
        }
        __syncthreads();
        TPROF(0);
        const int nq = (n - j0 + 63) >> 6;  // pair columns in use
        {   // ---- one pass over rows >= r0, columns r0 .. row: pending update applied, then row sums and (per lane) column sums
            //      of A22 v_k on the updated entries; the diagonal lands in both sums and is taken out again below.  Four rows
            //      at a time: their loads are issued together (a pair beyond a row's end is read and masked), their sums share
            //      one transposing shuffle reduction.
            double2 cacc[NQ2];
#pragma unroll
            for (int q = 0; q < NQ2; ++q) cacc[q] = make_double2(0.0, 0.0);
            for (int ib = r0 + warp; ib < n; ib += 4 * TD_WARPS) {
                int iu[4];
                double* rowp[4];
                double vin[4], vip[4], wip[4], racc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = ib + u * TD_WARPS;
                    const bool ok = i < n;
                    iu[u] = ok ? i : -1;
                    rowp[u] = A + (ok ? ro[i] : 0) + 2 * lane;
                    vin[u] = ok ? vn[i] : 0.0;
                    vip[u] = (ok && pending) ? vp[i] : 0.0;
                    wip[u] = (ok && pending) ? wp[i] : 0.0;
                    racc[u] = 0.0;
                }
                const int imax = min(n - 1, ib + 3 * TD_WARPS);
#pragma unroll
                for (int q = 0; q < NQ2; ++q) {
                    const int jb = j0 + 64 * q;
                    if (jb > imax) break;
                    const int jp = jb + 2 * lane;
                    const bool in_range = jp < n;
                    const double2 vjn = in_range ? *reinterpret_cast<const double2*>(vn + jp) : make_double2(0.0, 0.0);
                    double2 vjp = make_double2(0.0, 0.0), wjp = make_double2(0.0, 0.0);
                    if (pending && in_range) {
                        vjp = *reinterpret_cast<const double2*>(vp + jp);
                        wjp = *reinterpret_cast<const double2*>(wp + jp);
                        if (jp < r0) vjp.x = wjp.x = 0.0;  // column k is already up to date
                    }
                    double2 av[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) av[u] = *reinterpret_cast<const double2*>(rowp[u] + jb);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double nx = fma(-vip[u], wjp.x, fma(-wip[u], vjp.x, av[u].x));
                        const double ny = fma(-vip[u], wjp.y, fma(-wip[u], vjp.y, av[u].y));
                        if (pending) {
                            if (jp + 1 <= iu[u]) *reinterpret_cast<double2*>(rowp[u] + jb) = make_double2(nx, ny);
                            else if (jp == iu[u]) rowp[u][jb] = nx;
                        }
                        const double ax = jp <= iu[u] ? nx : 0.0, ay = jp + 1 <= iu[u] ? ny : 0.0;
                        racc[u] = fma(ax, vjn.x, fma(ay, vjn.y, racc[u]));
                        cacc[q].x = fma(ax, vin[u], cacc[q].x);
                        cacc[q].y = fma(ay, vin[u], cacc[q].y);
                    }
                }
                double r2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const bool hi = lane & 16;
                    r2[u] = (hi ? racc[u + 2] : racc[u]) + __shfl_xor_sync(FULL, hi ? racc[u] : racc[u + 2], 16);
                }
                double r1;
                {
                    const bool hi = lane & 8;
                    r1 = (hi ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, hi ? r2[0] : r2[1], 8);
                }
                r1 += __shfl_xor_sync(FULL, r1, 4);
                r1 += __shfl_xor_sync(FULL, r1, 2);
                r1 += __shfl_xor_sync(FULL, r1, 1);
                const int i = ib + (lane >> 3) * TD_WARPS;  // lane 8 r holds the sum of row r of the group
                if ((lane & 7) == 0 && i < n) pr[i] = r1;
            }
#pragma unroll
            for (int q = 0; q < NQ2; ++q) {
                const int jp = j0 + 2 * lane + 64 * q;
                if (q < nq && jp < n + (n & 1)) *reinterpret_cast<double2*>(colpart + warp * ldc + jp) = cacc[q];
            }
        }
        __syncthreads();
        TPROF(1);
        const int i = r0 + tid;
        double pi_ = 0.0, part = 0.0;
        if (i < n) {
            const double vi = vn[i];
            double s4[4] = {fma(-A[ro[i] + i], vi, pr[i]), 0.0, 0.0, 0.0};
#pragma unroll
            for (int w = 0; w < TD_WARPS; ++w) s4[w & 3] += colpart[w * ldc + i];
            pi_ = t * ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            part = pi_ * vi;
        }
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        TPROF(2);
        double dot = 0.0;
        {
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int w = 0; w < TD_WARPS; ++w) s4[w & 3] += red[w];
            dot = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        if (i < n) wn[i] = fma(-0.5 * t * dot, vn[i], pi_);
        if (tid < 2) {  // zero outside the live range
            if (tid) wn[n] = 0.0;
            else if (j0 < r0) wn[j0] = 0.0;
        }
        __syncthreads();
        TPROF(3);
        pending = true;  // (t == 0 leaves w_k = 0: a pending update that changes nothing)
        cur ^= 1;
    }
    if (tid < 3 && n >= 3) {  // the last reflector (k = n-3) still has to reach the trailing 2 x 2 block
        const double* const vp = vbuf + (cur ^ 1) * ldc;
        const double* const wp = wbuf + (cur ^ 1) * ldc;
        const int i = tid == 0 ? n - 2 : n - 1, j = tid == 2 ? n - 1 : n - 2;
        if (pending) A[ro[i] + j] -= fma(vp[i], wp[j], wp[i] * vp[j]);
    }
    __syncthreads();
#ifdef PSD_PROFILE
    if (tid == 0) printf("[psd_tridiag profile] clocks per step: column + vector %lld, fused pass %lld, p+dot %lld, w %lld\n", pc[0] / (n - 2), pc[1] / (n - 2), pc[2] / (n - 2), pc[3] / (n - 2));
#endif
    if (tid == 0) {
        if (n >= 2) {
            da[n - 2] = A[ro[n - 2] + n - 2];
            eb[n - 2] = A[ro[n - 1] + n - 2];
        }
        da[n - 1] = A[ro[n - 1] + n - 1];
    }
}

// ---- 2. eigenvalues of T ---------------------------------------------------------------------------------------------
// Number of eigenvalues of the scaled T (|a| + |b| row sums <= 1) below x: sign changes of the Sturm sequence
// p_k = (a_k - x) p_{k-1} - b_{k-1}^2 p_{k-2} (growth <= 3 per step).  b^2 is floored at 1e-60 by the caller (a relative
// perturbation of 1e-30 of T: the sequence then never stays at an exact zero, which it would where T decouples and x hits
// a diagonal entry), and the pair is rescaled by a power of two every FOURTH step when it has left [1e-60, 1e100]: four
// steps shrink it by at most 1e-240, so it never underflows, and the rescaling leaves the dependent chain of one FMA per
// step almost alone.  A zero p_k counts as positive; its successor -b^2 p_{k-1} has the opposite sign of p_{k-1}, which
// gives the right number of sign changes.  Signs are counted on the integer pipe.
__device__ __forceinline__ int sturm_count(const int n, const double* a, const double* b2, const double x) {
    double p2 = 1.0, p1 = a[0] - x;
    int cnt = (__double2hiint(p1) >> 31) & 1;
    const double up = 0x1p+332, down = 0x1p-332;
    int k = 1;
#pragma unroll 2
    for (; k + 3 < n; k += 4) {
        const double pa = fma(a[k] - x, p1, -(b2[k - 1] * p2));
        const double pb = fma(a[k + 1] - x, pa, -(b2[k] * p1));
        const double pc = fma(a[k + 2] - x, pb, -(b2[k + 1] * pa));
        const double pd = fma(a[k + 3] - x, pc, -(b2[k + 2] * pb));
        const int h1 = __double2hiint(p1), ha = __double2hiint(pa), hb = __double2hiint(pb), hc = __double2hiint(pc), hd = __double2hiint(pd);
        cnt += (((h1 ^ ha) >> 31) & 1) + (((ha ^ hb) >> 31) & 1) + (((hb ^ hc) >> 31) & 1) + (((hc ^ hd) >> 31) & 1);
        const double m = fmax(fabs(pc), fabs(pd));
        const double f = m < 1e-60 ? up : (m > 1e100 ? down : 1.0);
        p2 = pc * f;
        p1 = pd * f;
    }
    for (; k < n; ++k) {
        const double pa = fma(a[k] - x, p1, -(b2[k - 1] * p2));
        cnt += ((__double2hiint(p1) ^ __double2hiint(pa)) >> 31) & 1;
        p2 = p1;
        p1 = pa;
    }
    return cnt;
}

constexpr int BS_WARPS = 4;

// lam[n] ascending; scale_out[0] = largest absolute row sum of T (0: T == 0).  One warp per eigenvalue: every step cuts
// [lo, hi] into 33 parts with one Sturm count per lane (11 steps: 2 x 33^-11 < eps / 4); warps spread over ceil(n/4) CTAs.
__global__ void __launch_bounds__(BS_WARPS * 32) psd_bisect_kernel(const int n, const double* __restrict__ da, const double* __restrict__ eb,
                                                                   double* __restrict__ lam, double* __restrict__ scale_out) {
    extern __shared__ double sm[];
    double* a = sm;
    double* b2 = a + n;
    __shared__ double red[BS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double g = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double r = fabs(da[i]) + (i > 0 ? fabs(eb[i - 1]) : 0.0) + (i + 1 < n ? fabs(eb[i]) : 0.0);
        g = fmax(g, r);
    }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) g = fmax(g, __shfl_xor_sync(FULL, g, q));
    if (lane == 0) red[warp] = g;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < BS_WARPS; ++w) s = fmax(s, red[w]);
    if (blockIdx.x == 0 && tid == 0) scale_out[0] = s;
    const int i = blockIdx.x * BS_WARPS + warp;
    if (s == 0.0 || !(s == s)) {  // zero (or not finite) matrix: eigenvalues 0, U = I downstream
        if (i < n && lane == 0) lam[i] = 0.0;
        return;
    }
    const double rs = 1.0 / s;
    for (int q = tid; q < n; q += blockDim.x) {
        a[q] = da[q] * rs;
        const double e = q + 1 < n ? eb[q] * rs : 0.0;
        b2[q] = fmax(e * e, 1e-60);
    }
    __syncthreads();
    if (i >= n) return;
    double lo = -1.0 - 1e-3, hi = 1.0 + 1e-3;
    for (int it = 0; it < 11; ++it) {
        if (hi - lo <= 4.0 * EPS) break;  // (uniform across the warp) the interval is down to rounding level
        const double x = lo + (lane + 1) * ((hi - lo) * (1.0 / 33.0));
        const int c = sturm_count(n, a, b2, x);
        const unsigned above = __ballot_sync(FULL, c > i);  // lanes whose point has more than i eigenvalues below it
        const int f = above ? __ffs(above) - 1 : 32;
        const double xf = __shfl_sync(FULL, x, f & 31), xb = __shfl_sync(FULL, x, (f + 31) & 31);
        if (f < 32) hi = xf;
        if (f > 0) lo = xb;
    }
    if (lane == 0) lam[i] = 0.5 * (lo + hi) * s;
}

// ---- 3. eigenvectors of T ---------------------------------------------------------------------------------------------
constexpr int IV_WARPS = 8;

__device__ __forceinline__ double hash_uniform(unsigned a, unsigned b) {  // deterministic start vector entries in (-1, 1)
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return ((double)h + 0.5) * (2.0 / 4294967296.0) - 1.0;
}


// Z (n x n, column major): column j = unit eigenvector of T for lam[j].  One warp per cluster (warp index = first member;
// a cluster = maximal run of eigenvalues whose neighbours are closer than CLUSTER_RTOL |T|).
__global__ void __launch_bounds__(IV_WARPS * 32) psd_invit_kernel(const int n, const double* __restrict__ da, const double* __restrict__ eb,
                                                                  const double* __restrict__ lam, const double* __restrict__ scale,
                                                                  double* __restrict__ Z) {
    extern __shared__ double sm[];
    double* ta = sm;          // diagonal of T / |T|
    double* te = ta + n;      // off-diagonal
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double s = scale[0];
    const bool degenerate = s == 0.0 || !(s == s);
    const double rs = degenerate ? 0.0 : 1.0 / s;  // T is used scaled to unit row sums: shifts and pivots are then O(1)
    for (int i = tid; i < n; i += blockDim.x) {
        ta[i] = da[i] * rs;
        te[i] = i + 1 < n ? eb[i] * rs : 0.0;
    }
    __syncthreads();
    const int first = blockIdx.x * IV_WARPS + warp;
    if (first >= n) return;
    if (degenerate) {
        for (int i = lane; i < n; i += 32) Z[(size_t)first * n + i] = i == first ? 1.0 : 0.0;
        return;
    }
    if (first > 0 && !(lam[first] - lam[first - 1] > CLUSTER_RTOL * s)) return;  // not the first of its cluster
    double* fr = te + n + (size_t)warp * 5 * n;  // reciprocal pivots of U
    double* fb = fr + n;                          // U first superdiagonal
    double* fc = fb + n;                          // L multipliers
    double* fd = fc + n;                          // U second superdiagonal
    double* x = fd + n;
    unsigned char* in = reinterpret_cast<unsigned char*>(te + n + (size_t)IV_WARPS * 5 * n) + (size_t)warp * n;
    const double tol = EPS;
    double xprev = 0.0;
    for (int j = first; j < n && (j == first || !(lam[j] - lam[j - 1] > CLUSTER_RTOL * s)); ++j) {
        double xj = lam[j] * rs;
        if (j > first) {  // dstein: members of a cluster get distinct shifts
            const double pert = 10.0 * EPS * fabs(xj);
            if (xj - xprev < pert) xj = xprev + pert;
        }
        xprev = xj;
        for (int i = lane; i < n; i += 32) x[i] = hash_uniform((unsigned)j, (unsigned)i);
        if (lane == 0) {
            // P L U = T - xj I with partial pivoting (dlagtf); pivots below tol are pushed to +-tol (dlagts, job -1) and kept
            // as reciprocals.  The running row (ak, bk) stays in registers: one reciprocal per step on the chain.
            double ak = ta[0] - xj, bk = te[0];
#pragma unroll 2
            for (int k = 0; k + 1 < n; ++k) {
                const double ck = te[k], an = ta[k + 1] - xj, bn = te[k + 1];
                if (fabs(ck) <= fabs(ak) || fabs(ck) < 1e-200) {
                    const double piv = fabs(ak) < tol ? (ak < 0.0 ? -tol : tol) : ak;
                    const double r = rcp_newton(piv), m = ck * r;
                    fr[k] = r;
                    fb[k] = bk;
                    fc[k] = m;
                    fd[k] = 0.0;
                    in[k] = 0;
                    ak = fma(-m, bk, an);
                    bk = bn;
                } else {
                    const double r = rcp_newton(ck), m = ak * r;
                    fr[k] = r;
                    fb[k] = an;
                    fc[k] = m;
                    fd[k] = bn;
                    in[k] = 1;
                    ak = fma(-m, an, bk);
                    bk = -m * bn;
                }
            }
            const double piv = fabs(ak) < tol ? (ak < 0.0 ? -tol : tol) : ak;
            fr[n - 1] = rcp_newton(piv);
            fb[n - 1] = 0.0;
            fd[n - 1] = 0.0;
        }
        __syncwarp();
        for (int it = 0; it < 3; ++it) {
            double nn = 0.0;
            for (int i = lane; i < n; i += 32) nn = fma(x[i], x[i], nn);
            nn = warp_sum(nn);
            const double sc = nn > 0.0 ? rsqrt(nn) : 0.0;
            for (int i = lane; i < n; i += 32) x[i] = nn > 0.0 ? x[i] * sc : (i == j ? 1.0 : 0.0);
            __syncwarp();
            if (lane == 0) {
                double prev = x[0];  // the entry the next step still changes, carried in a register
#pragma unroll 4
                for (int k = 1; k < n; ++k) {
                    const double xk = x[k], m = fc[k - 1];
                    if (!in[k - 1]) {
                        x[k - 1] = prev;
                        prev = fma(-m, prev, xk);
                    } else {
                        x[k - 1] = xk;
                        prev = fma(-m, xk, prev);
                    }
                }
                x[n - 1] = prev;
                double y1 = 0.0, y2 = 0.0;  // x[k+1], x[k+2]
#pragma unroll 4
                for (int k = n - 1; k >= 0; --k) {
                    const double xk = fma(-fd[k], y2, fma(-fb[k], y1, x[k])) * fr[k];
                    x[k] = xk;
                    y2 = y1;
                    y1 = xk;
                }
            }
            __syncwarp();
            for (int pass = 0; pass < (it == 2 ? 2 : 1); ++pass)
                for (int jp = first; jp < j; ++jp) {  // modified Gram-Schmidt against the cluster's earlier members
                    const double* zp = Z + (size_t)jp * n;
                    double dot = 0.0;
                    for (int i = lane; i < n; i += 32) dot = fma(zp[i], x[i], dot);
                    dot = warp_sum(dot);
                    for (int i = lane; i < n; i += 32) x[i] = fma(-dot, zp[i], x[i]);
                    __syncwarp();
                }
        }
        double nn = 0.0;
        for (int i = lane; i < n; i += 32) nn = fma(x[i], x[i], nn);
        nn = warp_sum(nn);
        const double sc = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
        for (int i = lane; i < n; i += 32) Z[(size_t)j * n + i] = x[i] * sc;
        __threadfence_block();
        __syncwarp();
    }
}

// ---- 4. back-transformation U0 = Q Z, Q = H_0 H_1 ... H_{n-3} --------------------------------------------------------------
constexpr int BT_PANEL = 16;  // reflectors staged in shared memory at a time

// One warp per eigenvector (its entries in registers).  The reflectors come through shared memory in panels of BT_PANEL,
// the next panel's loads in flight (registers) while the current one is applied.
__global__ void __launch_bounds__(256) psd_backtransform_kernel(const int n, const double* __restrict__ H, const double* __restrict__ tau,
                                                                const double* __restrict__ Z, double* __restrict__ U0) {
    extern __shared__ double sm[];
    double* st = sm;                 // tau
    double* pan = st + n;            // BT_PANEL x n
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += blockDim.x) st[i] = i + 2 < n ? tau[i] : 0.0;
    const int j = blockIdx.x * (blockDim.x >> 5) + warp;
    double x[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = lane + 32 * q;
        x[q] = (j < n && i < n) ? Z[(size_t)j * n + i] : 0.0;
    }
    double stage[BT_PANEL];  // thread i < n carries entry i of every reflector of the panel (coalesced, no index arithmetic)
    // panel p holds reflectors k = khi - BT_PANEL + 1 .. khi (those < 0 do not exist), applied from khi downwards
    auto fetch = [&](int khi) {
#pragma unroll
        for (int r = 0; r < BT_PANEL; ++r) {
            const int k = khi - r;
            stage[r] = (tid < n && k >= 0 && tid > k) ? H[(size_t)k * n + tid] : 0.0;
        }
    };
    auto commit = [&]() {
        if (tid < n) {
#pragma unroll
            for (int r = 0; r < BT_PANEL; ++r) pan[r * n + tid] = stage[r];
        }
    };
    int khi = n - 3;
    if (khi >= 0) fetch(khi);
    for (; khi >= 0; khi -= BT_PANEL) {
        __syncthreads();  // the previous panel is no longer read
        commit();
        __syncthreads();
        if (khi - BT_PANEL >= 0) fetch(khi - BT_PANEL);
        for (int r = 0; r < BT_PANEL && khi - r >= 0; ++r) {
            const double t = st[khi - r];
            if (t == 0.0) continue;
            const double* v = pan + r * n;
            double vq[NQ], dot = 0.0, dot2 = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = lane + 32 * q;
                vq[q] = i < n ? v[i] : 0.0;
                if (q & 1) dot2 = fma(vq[q], x[q], dot2);
                else dot = fma(vq[q], x[q], dot);
            }
            dot = warp_sum(dot + dot2) * t;
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[q] = fma(-dot, vq[q], x[q]);
        }
    }
    if (j < n) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = lane + 32 * q;
            if (i < n) U0[(size_t)j * n + i] = x[q];
        }
    }
}

// ---- 5. Newton-Schulz polish, projection ---------------------------------------------------------------------------------
// C (n x n, column major) = alpha D + beta op(A) diag(w) op(B) on the FP64 tensor pipe: 32 x 32 tile per CTA, 8 warps x two
// 8 x 8 blocks (mma.sync m8n8k4), operands staged in shared memory 32 deep (row stride 40: fragment loads take the minimum
// of two wavefronts).  TA: op(A) = A'; TB: op(B) = B'; w (optional): max(w_k, 0) scales the inner index;
// TRI: only the tiles on and above the diagonal are computed and C is the packed triangle (i <= j at j (j + 1) / 2 + i).
template <bool TA, bool TB, bool TRI>
__global__ void __launch_bounds__(256) psd_gemm_kernel(const int n, const double* __restrict__ A, const double* __restrict__ B,
                                                       const double* __restrict__ w, const double* __restrict__ D, const double alpha,
                                                       const double beta, double* __restrict__ C) {
    __shared__ double As[32][40], Bs[32][40];  // As[k][i], Bs[k][j]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    if (TRI && i0 > j0) return;
    const int bi = warp >> 1, bj = (warp & 1) * 2;  // this warp: blocks (bi, bj) and (bi, bj + 1) of the 4 x 4 block grid
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
    for (int k0 = 0; k0 < n; k0 += 32) {
        double va[4], vb[4];  // all eight loads of this thread in flight before the first store
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u, r = e & 31, c = e >> 5;
            if (TA) {  // op(A)(i, k) = A[k + i n]: r runs over k (contiguous)
                const int kk = k0 + r, ii = i0 + c;
                va[u] = (kk < n && ii < n) ? A[(size_t)ii * n + kk] * (w ? fmax(w[kk], 0.0) : 1.0) : 0.0;
            } else {   // A(i, k) = A[i + k n]: r runs over i (contiguous)
                const int ii = i0 + r, kk = k0 + c;
                va[u] = (kk < n && ii < n) ? A[(size_t)kk * n + ii] * (w ? fmax(w[kk], 0.0) : 1.0) : 0.0;
            }
            if (TB) {  // op(B)(k, j) = B[j + k n]: r runs over j (contiguous)
                const int jj = j0 + r, kb = k0 + c;
                vb[u] = (kb < n && jj < n) ? B[(size_t)kb * n + jj] : 0.0;
            } else {   // B(k, j) = B[k + j n]: r runs over k
                const int kb = k0 + r, jj = j0 + c;
                vb[u] = (kb < n && jj < n) ? B[(size_t)jj * n + kb] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u, r = e & 31, c = e >> 5;
            if (TA) As[r][c] = va[u];
            else As[c][r] = va[u];
            if (TB) Bs[c][r] = vb[u];
            else Bs[r][c] = vb[u];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; kk += 4) {
            const double a = As[kk + t][8 * bi + g];       // A fragment: row g, k-slot t
            const double b0 = Bs[kk + t][8 * bj + g];      // B fragments: k-slot t, column g
            const double b1 = Bs[kk + t][8 * bj + 8 + g];
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b0));
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b1));
        }
        __syncthreads();
    }
    // accumulator fragment: row g, columns 2t and 2t + 1 of the 8 x 8 block
#pragma unroll
    for (int blk = 0; blk < 2; ++blk)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = i0 + 8 * bi + g, j = j0 + 8 * (bj + blk) + 2 * t + v;
            const double acc = blk ? c1[v] : c0[v];
            if (i < n && j < n) {
                if (TRI) {
                    if (i <= j) C[(size_t)j * (j + 1) / 2 + i] = beta * acc;
                } else {
                    const size_t o = (size_t)j * n + i;
                    C[o] = (D ? alpha * D[o] : 0.0) + beta * acc;
                }
            }
        }
}

}  // namespace

size_t psd_tridiag_smem_bytes(int d) {
    const size_t h = (size_t)d >> 1;
    const size_t nst = (d & 1) ? 2 * (h + 1) * (h + 1) : 2 * h * (h + 1);  // roff(d)
    const size_t ldc = ((size_t)d + 3) & ~(size_t)1;
    return sizeof(double) * (nst + 5 * ldc + 2 * TD_WARPS + 2 + (size_t)TD_WARPS * ldc) + sizeof(int) * ((size_t)d + 2);
}

// the direct route serves sides the Jacobi kernel of one CTA cannot hold and whose packed triangle fits shared memory
bool psd_tridiag_supported(diffopt_b200_ctx* ctx, int d) { return d >= 3 && d <= MAX_SIDE && psd_tridiag_smem_bytes(d) <= ctx->smem_optin; }

// xtri: the cone's slice of v (packed triangle, d(d+1)/2); w0, w1, w2: three d x d scratch matrices; small: 4 d + 8 doubles;
// U (d x d) and lam (d, ascending) are the outputs.
int32_t psd_tridiag_eig_launch(diffopt_b200_ctx* ctx, int d, const double* xtri, double* w0, double* w1, double* w2, double* small,
                               double* U, double* lam) {
    double* tau = small;
    double* da = tau + d;
    double* eb = da + d;
    double* scale = eb + d;
    double *H = w0, *Z = w1, *U0 = w2, *S = w0;
    const size_t smem1 = psd_tridiag_smem_bytes(d);
    DO_CUDA(ctx, cudaFuncSetAttribute(psd_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    psd_tridiag_kernel<<<1, TD_THREADS, smem1, ctx->stream>>>(d, xtri, H, tau, da, eb);
    psd_bisect_kernel<<<(d + BS_WARPS - 1) / BS_WARPS, BS_WARPS * 32, sizeof(double) * 2 * (size_t)d, ctx->stream>>>(d, da, eb, lam, scale);
    const size_t smem3 = sizeof(double) * (2 * (size_t)d + (size_t)IV_WARPS * 5 * d) + (size_t)IV_WARPS * d + 16;
    DO_CUDA(ctx, cudaFuncSetAttribute(psd_invit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    psd_invit_kernel<<<(d + IV_WARPS - 1) / IV_WARPS, IV_WARPS * 32, smem3, ctx->stream>>>(d, da, eb, lam, scale, Z);
    psd_backtransform_kernel<<<(d + 7) / 8, 256, sizeof(double) * (size_t)(BT_PANEL + 1) * d, ctx->stream>>>(d, H, tau, Z, U0);
    const dim3 gg((unsigned)((d + 31) / 32), (unsigned)((d + 31) / 32));
    psd_gemm_kernel<true, false, false><<<gg, 256, 0, ctx->stream>>>(d, U0, U0, nullptr, nullptr, 0.0, 1.0, S);   // S = U0' U0
    psd_gemm_kernel<false, false, false><<<gg, 256, 0, ctx->stream>>>(d, U0, S, nullptr, U0, 1.5, -0.5, U);      // U = 1.5 U0 - 0.5 U0 S
    ctx->launches += 6;
    DO_CUDA(ctx, cudaGetLastError());
    return 0;
}

// pi(v) of one PSD block from its eigenpairs: vp (packed triangle) = U max(L, 0) U'
int32_t psd_projection_launch(diffopt_b200_ctx* ctx, int d, const double* U, const double* lam, double* vp) {
    const dim3 gg((unsigned)((d + 31) / 32), (unsigned)((d + 31) / 32));
    psd_gemm_kernel<false, true, true><<<gg, 256, 0, ctx->stream>>>(d, U, U, lam, nullptr, 0.0, 1.0, vp);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    return 0;
}

// Sparse direct path for ONE large KKT system with many right-hand sides (BASELINE config 3: MPC-structured QP,
// N = 240 000, 256 forward directions against one factorisation) -- the `LHS \ RHS` of QuadraticProgram.jl:486-492 for
// systems far beyond the dense kernels.  The reference refactorises for every direction (:438 runs once per
// forward_differentiate!); here the factorisation stays in the ctx and every block of right-hand sides reuses it.
//
// Method: reverse Cuthill-McKee ordering of the symmetrised pattern on the host (KKT matrices of stage-structured
// problems become narrow banded), LAPACK band storage built on the device, then
//   * dgbtf2-style LU with partial pivoting inside the band by one persistent CTA (read phase / barrier / write phase
//     per column, the active window of the band stays L2 resident), and
//   * dgbtrs-style sweeps with ONE WARP PER RIGHT-HAND SIDE: the warp keeps a sliding window of its column in shared
//     memory, prefetches the multipliers of the next column from L2 and never synchronises with other warps, so all
//     right-hand sides advance in parallel.
// Partial pivoting keeps this valid for the reference's nonsymmetric LHS = [Q G'L A'; G D 0; A 0 0] with its exact
// zeros on the diagonal.  A matrix whose RCM bandwidth exceeds BAND_MAX is rejected (-3): a fill-reducing supernodal
// factorisation for general patterns is the next step.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int BAND_MAX = 255;   // kl = ku <= BAND_MAX
constexpr int F_THREADS = 1024;  // factorisation CTA (> BAND_MAX: one thread per entry of the pivot column)
constexpr int S_WARPS = 4;      // right-hand sides per solve CTA
constexpr int WIN = 1024;       // sliding window (doubles) per warp: power of two >= kl + ku + 1 + 64 (refilled in 32-chunks)

// ---- host: reverse Cuthill-McKee on the symmetrised pattern ---------------------------------------------------------
void rcm_order(int64_t N, const std::vector<int64_t>& ptr, const std::vector<int32_t>& adj, std::vector<int32_t>& perm) {
    std::vector<int32_t> deg((size_t)N);
    for (int64_t i = 0; i < N; ++i) deg[(size_t)i] = (int32_t)(ptr[(size_t)i + 1] - ptr[(size_t)i]);
    std::vector<char> seen((size_t)N, 0), mark((size_t)N, 0);
    std::vector<int32_t> q, nbr;
    perm.clear();
    perm.reserve((size_t)N);
    auto bfs_last = [&](int32_t start) {  // a vertex of the last BFS level with small degree (pseudo-peripheral search)
        q.assign(1, start);
        mark[(size_t)start] = 1;
        for (size_t head = 0; head < q.size(); ++head) {
            const int32_t v = q[head];
            for (int64_t k = ptr[(size_t)v]; k < ptr[(size_t)v + 1]; ++k) {
                const int32_t w = adj[(size_t)k];
                if (!mark[(size_t)w] && !seen[(size_t)w]) {
                    mark[(size_t)w] = 1;
                    q.push_back(w);
                }
            }
        }
        for (int32_t v : q) mark[(size_t)v] = 0;
        int32_t best = q.back();
        for (size_t i = q.size() > 32 ? q.size() - 32 : 0; i < q.size(); ++i)
            if (deg[(size_t)q[i]] < deg[(size_t)best]) best = q[i];
        return best;
    };
    for (int64_t s0 = 0; s0 < N; ++s0) {
        if (seen[(size_t)s0]) continue;
        int32_t start = bfs_last((int32_t)s0);
        start = bfs_last(start);
        size_t head = perm.size();
        perm.push_back(start);
        seen[(size_t)start] = 1;
        while (head < perm.size()) {
            const int32_t v = perm[head++];
            nbr.clear();
            for (int64_t k = ptr[(size_t)v]; k < ptr[(size_t)v + 1]; ++k) {
                const int32_t w = adj[(size_t)k];
                if (!seen[(size_t)w]) {
                    seen[(size_t)w] = 1;
                    nbr.push_back(w);
                }
            }
            std::sort(nbr.begin(), nbr.end(), [&](int32_t a, int32_t b) { return deg[(size_t)a] < deg[(size_t)b]; });
            perm.insert(perm.end(), nbr.begin(), nbr.end());
        }
    }
    std::reverse(perm.begin(), perm.end());
}

// ---- device -----------------------------------------------------------------------------------------------------------
// band storage: AB is ldab x N column major, ldab = 2 kl + ku + 1, A(i, j) at AB[kv + i - j + j * ldab], kv = kl + ku
__global__ void band_scatter_kernel(int64_t N, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                    const double* __restrict__ nzval, const int32_t* __restrict__ inv, int trans, int kv, int64_t ldab,
                                    double* __restrict__ AB) {
    const int64_t col = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;  // 8 lanes per column of the CSC matrix
    const int l8 = threadIdx.x & 7;
    if (col >= N) return;
    for (int64_t k = colptr[col] - 1 + l8; k < colptr[col + 1] - 1; k += 8) {
        int64_t i = inv[rowval[k] - 1], j = inv[col];
        if (trans) {
            const int64_t t = i;
            i = j;
            j = t;
        }
        atomicAdd(&AB[j * ldab + kv + i - j], nzval[k]);  // duplicates are summed like SparseArrays does
    }
}

// LU with partial pivoting in band storage (LAPACK dgbtf2).  Per column: (1) pivot search, (2) everything a write of
// this step could clobber is read into registers / shared memory, barrier, (3) writes.
// SMEM_WINDOW: the kv + 1 columns the elimination can touch live in a circular shared-memory window of WC columns
// (column c at slot c & (WC - 1)); column j is written back when it retires and column j + kv + 1 is fetched one step
// ahead, so the per-column dependency chain runs at shared-memory instead of L2 latency.
template <bool SMEM_WINDOW>
__global__ void __launch_bounds__(F_THREADS, 1) band_lu_kernel(int N, int kl, int ku, double* __restrict__ AB, int* __restrict__ ipiv,
                                                               int* info, int WC) {
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ int s_idx[1 + F_THREADS / 32];
    __shared__ double s_piv, s_diag, s_rinv;
    __shared__ double s_l[BAND_MAX + 1];        // multipliers of the current column
    __shared__ double s_top[2 * BAND_MAX + 2];  // row j+jp of the window columns (the new pivot row)
    __shared__ double s_bot[2 * BAND_MAX + 2];  // row j of the window columns (moves to row j+jp)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarp = nthr >> 5;  // the launch picks the CTA size by band width (fewer warps: cheaper barriers)
    const int kv = kl + ku;
    const int ldab_i = 2 * kl + ku + 1;
    const size_t ldab = (size_t)ldab_i;
    const int wmask = WC - 1;
    auto col = [&](int c) -> double* { return SMEM_WINDOW ? s_dyn + (size_t)(c & wmask) * ldab : AB + (size_t)c * ldab; };
    constexpr int PD = 6;  // columns kept in flight by cp.async ahead of the window
    if (SMEM_WINDOW) {
        const int ncol = min(N, kv + 1 + PD);
        for (int e = tid; e < ncol * ldab_i; e += nthr) s_dyn[e] = AB[e];  // columns 0 .. kv+PD sit at their slots
        __syncthreads();
    }
    int ju = 0;
    for (int j = 0; j < N; ++j) {
        double* colj = col(j);
        const int km = min(kl, N - 1 - j);
        // column j + kv + 1 + PD starts its way into the window (its slot was freed PD + ... steps ago); the column needed
        // next step (j + kv + 1) was requested PD steps ago
        if (SMEM_WINDOW) {
            const int cin = j + kv + 1 + PD;
            if (cin < N && tid < ldab_i) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(col(cin) + tid);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(AB + (size_t)cin * ldab + tid) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // pivot search on a packed key: high word of |a| (sign 0, exponent, 12 mantissa bits) | (255 - row): one
        // redux.max picks the largest entry (to 2^-12 relative, i.e. threshold pivoting at 0.9998) and the first row on ties
        double mine = 0.0;
        unsigned key = 0u;
        if (tid <= km) {
            mine = colj[kv + tid];
            key = ((unsigned)__double2hiint(fabs(mine)) & 0xFFFFFF00u) | (unsigned)(255 - tid);
        }
        if (km < 32) {
            if (warp == 0) {
                key = __reduce_max_sync(0xffffffffu, key);
                if (lane == 0) s_idx[0] = (int)key;
            }
        } else {
            key = __reduce_max_sync(0xffffffffu, key);
            if (lane == 0) s_idx[1 + warp] = (int)key;
            __syncthreads();
            if (warp == 0) {
                key = lane < nwarp ? (unsigned)s_idx[1 + lane] : 0u;
                key = __reduce_max_sync(0xffffffffu, key);
                if (lane == 0) s_idx[0] = (int)key;
            }
        }
        if (tid == 0) s_diag = mine;
        __syncthreads();
        const unsigned kbest = (unsigned)s_idx[0];
        const int jp = 255 - (int)(kbest & 0xFFu);
        if (tid == jp) {
            s_piv = (kbest >> 8) ? mine : 0.0;  // zero or denormal column: singular
            s_rinv = (kbest >> 8) ? 1.0 / mine : 0.0;
            ipiv[j] = j + jp;
        }
        ju = max(ju, min(j + ku + jp, N - 1));
        const int nc = ju - j;  // window columns j+1 .. ju
        for (int c = 1 + tid; c <= nc; c += nthr) {
            const double* colc = col(j + c) + kv - c;
            s_top[c] = colc[jp];
            s_bot[c] = colc[0];
        }
        __syncthreads();
        const double piv = s_piv;
        if (piv == 0.0) {  // the whole column below the diagonal is exactly zero: singular
            if (tid == 0) *info = j + 1;
            return;
        }
        const double rinv = s_rinv;
        if (tid == 0) colj[kv] = piv;
        if (tid >= 1 && tid <= km) {
            const double v = (tid == jp ? s_diag : mine) * rinv;
            s_l[tid] = v;
            colj[kv + tid] = v;
        }
        __syncthreads();
        for (int c = 1 + warp; c <= nc; c += nwarp) {  // a warp per window column, lanes down the rows
            double* colc = col(j + c) + kv - c;
            const double top = s_top[c];
            for (int i = lane; i <= km; i += 32) {
                if (i == 0) {
                    colc[0] = top;
                } else {
                    const double old = (i == jp) ? s_bot[c] : colc[i];
                    colc[i] = fma(-s_l[i], top, old);
                }
            }
        }
        if (SMEM_WINDOW) {
            // column j is final: write it back (its slot is reused by column j + WC, requested no earlier than next step)
            if (tid < ldab_i) AB[(size_t)j * ldab + tid] = colj[tid];
            asm volatile("cp.async.wait_group %0;" ::"n"(PD - 1) : "memory");  // column j + kv + 1 has landed
        }
        __syncthreads();
    }
    if (tid == 0) *info = 0;
}

// Narrow bands (kl <= 31): the same elimination with ONE barrier per column.  Every warp loads the pivot column into
// registers and runs the pivot search, the Newton-refined reciprocal and the multipliers redundantly (packed-key redux
// + shuffles: no shared scalars, no extra barriers), then owns whole window columns: it reads a column's kl + 1 live
// rows into registers, takes the pivot-row and top-row entries by shuffle and writes the swapped + updated column back
// -- no cross-warp hazard inside a step.  Column j retires straight to global memory (U part from the window, pivot
// and multipliers from registers), so the window copy of the pivot column is never written while other warps read it.
// Variants measured on config 3 (N = 240 000, kl = ku = 19), factorisation time: three barriers per column 252 ms; this
// kernel 164 ms; the same with the retiring column leaving by a bulk (TMA) store one step later 172 ms; named-barrier
// look-ahead with an owner warp per column 159-176 ms; two warps with lanes = columns in the update 205-240 ms; pivot
// search by warp 0 only, published through shared memory behind a second barrier 164 ms.  The
// column step (~1300 clk) is bound by the dependent instruction chain of a warp (clock counters: pivot search 120,
// reciprocal + multipliers 150, one window column ~190 clk), not by arithmetic or memory.
__device__ __forceinline__ double band_fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

__global__ void __launch_bounds__(F_THREADS, 1) band_lu_warp_kernel(int N, int kl, int ku, double* __restrict__ AB, int* __restrict__ ipiv,
                                                                    int* info, int WC) {
    extern __shared__ __align__(16) double s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarp = nthr >> 5;
    const int kv = kl + ku;
    const int ldab_i = 2 * kl + ku + 1;
    const size_t ldab = (size_t)ldab_i;
    const int wmask = WC - 1;
    auto col = [&](int c) -> double* { return s_dyn + (size_t)(c & wmask) * ldab; };
    constexpr int PD = 6;  // columns kept in flight by cp.async ahead of the window
    {
        const int ncol = min(N, kv + 1 + PD);
        for (int e = tid; e < ncol * ldab_i; e += nthr) s_dyn[e] = AB[e];
        __syncthreads();
    }
    int ju = 0;
    for (int j = 0; j < N; ++j) {
        const double* colj = col(j);
        const int km = min(kl, N - 1 - j);
        {
            const int cin = j + kv + 1 + PD;
            if (cin < N && tid < ldab_i) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(col(cin) + tid);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(AB + (size_t)cin * ldab + tid) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // pivot search (same packed key as band_lu_kernel), redundantly in every warp
        const double mine = lane <= km ? colj[kv + lane] : 0.0;
        unsigned key = lane <= km ? (((unsigned)__double2hiint(fabs(mine)) & 0xFFFFFF00u) | (unsigned)(255 - lane)) : 0u;
        key = __reduce_max_sync(0xffffffffu, key);
        if (!(key >> 8)) {  // the whole column below the diagonal is zero (or denormal): singular
            if (tid == 0) *info = j + 1;
            return;
        }
        const int jp = 255 - (int)(key & 0xFFu);
        const double piv = __shfl_sync(0xffffffffu, mine, jp), diag = __shfl_sync(0xffffffffu, mine, 0);
        const double l = (lane == jp ? diag : mine) * band_fast_rcp(piv);  // multiplier of window row `lane` (lane >= 1)
        ju = max(ju, min(j + ku + jp, N - 1));
        const int nc = ju - j;
        if (warp == nwarp - 1) {  // column j retires (the last warp has the fewest window columns)
            double* out = AB + (size_t)j * ldab;
            for (int i = lane; i < kv; i += 32) out[i] = colj[i];
            if (lane == 0) {
                out[kv] = piv;
                ipiv[j] = j + jp;
            } else if (lane <= km) {
                out[kv + lane] = l;
            }
        }
        for (int c = 1 + warp; c <= nc; c += nwarp) {  // a warp per window column, lanes down the rows
            double* colc = col(j + c) + kv - c;
            const double old = lane <= km ? colc[lane] : 0.0;
            const double top = __shfl_sync(0xffffffffu, old, jp), bot = __shfl_sync(0xffffffffu, old, 0);
            if (lane <= km) colc[lane] = lane == 0 ? top : fma(-l, top, lane == jp ? bot : old);
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(PD - 1) : "memory");  // column j + kv + 1 has landed
        __syncthreads();
    }
    if (tid == 0) *info = 0;
}

// Y[i, r] = B[perm[i], r]: right-hand sides into the RCM ordering (so that the sweeps stream contiguously)
__global__ void band_gather_kernel(int64_t N, int64_t nrhs, const int32_t* __restrict__ perm, const double* __restrict__ B,
                                   double* __restrict__ Y) {
    const int64_t total = N * nrhs;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / N, i = e - r * N;
        Y[e] = B[r * N + perm[i]];
    }
}

// Solve with the factored band for nrhs right-hand sides, one warp per right-hand side, in place in Y (RCM order):
//   forward : y <- L^-1 P y   (interchanges ipiv, multipliers below the diagonal)
//   backward: x <- U^-1 y     (U has bandwidth kv = kl + ku), scattered to X in the ORIGINAL order through perm
// The warp keeps the active part of its column in a circular shared-memory window that is refilled 32 entries at a
// time one chunk ahead; multipliers are fetched one column ahead into registers behind an L1 prefetch PF columns
// ahead; pivots and permutation entries are fetched 32 at a time and broadcast by shuffles.
// QL = number of 32-row chunks of the L band (kl <= 32 QL); the U band (kv = 2 kl) takes 2 QL chunks.
template <int QL>
__global__ void __launch_bounds__(32 * S_WARPS) band_solve_kernel(int N, int kl, int ku, const double* __restrict__ AB,
                                                                  const int* __restrict__ ipiv, const int32_t* __restrict__ perm, int nrhs,
                                                                  double* __restrict__ Y, double* __restrict__ X) {
    __shared__ double s_win[S_WARPS][WIN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * S_WARPS + warp;
    if (r >= nrhs) return;
    double* win = s_win[warp];
    const int kv = kl + ku;
    const size_t ldab = (size_t)(2 * kl + ku + 1);
    double* y = Y + (size_t)r * N;
    double* x = X + (size_t)r * N;
    constexpr int MSK = WIN - 1;
    constexpr int PF = 12;  // L1 prefetch distance in columns
    const int lines_l = (kl * 8 + 8 + 127) / 128 + 1, lines_u = (kv * 8 + 8 + 127) / 128 + 1;
    // ---- forward sweep: window holds y[j .. j + kl], refilled from y one 32-chunk ahead
    int filled = min(N, kl + 1 + 32);  // window holds indices < filled
    for (int i = lane; i < filled; i += 32) win[i & MSK] = y[i];
    double chunk = (filled + lane < N) ? y[filled + lane] : 0.0;  // next 32 entries, in flight
    double lnext[QL];
#pragma unroll
    for (int q = 0; q < QL; ++q) {
        const int i = 1 + lane + 32 * q;
        lnext[q] = (i <= min(kl, N - 1)) ? AB[kv + i] : 0.0;
    }
    int pchunk = 0;  // ipiv[32 c + lane]
    __syncwarp();
    for (int j = 0; j < N; ++j) {
        if ((j & 31) == 0) {
            pchunk = (j + lane < N) ? ipiv[j + lane] : 0;
            // the chunk loaded 32 steps ago enters the window, the next one is requested
            if (j > 0) {
                if (filled + lane < N) win[(filled + lane) & MSK] = chunk;
                filled = min(N, filled + 32);
                chunk = (filled + lane < N) ? y[filled + lane] : 0.0;
            }
        }
        const int p = __shfl_sync(0xffffffffu, pchunk, j & 31);
        const int lm = min(kl, N - 1 - j);
        double lcur[QL];
#pragma unroll
        for (int q = 0; q < QL; ++q) lcur[q] = lnext[q];
        if (j + 1 < N) {
            const double* cn = AB + (size_t)(j + 1) * ldab + kv;
            const int lmn = min(kl, N - 2 - j);
#pragma unroll
            for (int q = 0; q < QL; ++q) {
                const int i = 1 + lane + 32 * q;
                lnext[q] = (i <= lmn) ? cn[i] : 0.0;
            }
            if (j + PF < N && lane < lines_l) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)(AB + (size_t)(j + PF) * ldab + kv) + lane * 128));
        }
        if (lane == 0 && p != j) {
            const double t0 = win[j & MSK], t1 = win[p & MSK];
            win[j & MSK] = t1;
            win[p & MSK] = t0;
        }
        __syncwarp();
        const double bj = win[j & MSK];
#pragma unroll
        for (int q = 0; q < QL; ++q) {
            const int i = 1 + lane + 32 * q;
            if (i <= lm) win[(j + i) & MSK] = fma(-lcur[q], bj, win[(j + i) & MSK]);
        }
        if (lane == 0) y[j] = bj;
        __syncwarp();
    }
    // ---- backward sweep: window holds y[j - kv .. j], refilled downwards one 32-chunk ahead
    int low = max(0, N - (kv + 1 + 32));  // window holds indices >= low
    for (int i = low + lane; i < N; i += 32) win[i & MSK] = y[i];
    chunk = (low - 32 + lane >= 0) ? y[low - 32 + lane] : 0.0;
    constexpr int UQ = 2 * QL;
    double unext[UQ];
    double dnext;
    {
        const double* cn = AB + (size_t)(N - 1) * ldab;
        dnext = cn[kv];
#pragma unroll
        for (int q = 0; q < UQ; ++q) {
            const int i = 1 + lane + 32 * q;
            unext[q] = (i <= min(kv, N - 1)) ? cn[kv - i] : 0.0;
        }
    }
    int permchunk = 0;
    __syncwarp();
    for (int j = N - 1, step = 0; j >= 0; --j, ++step) {
        if ((step & 31) == 0) {
            permchunk = (j - lane >= 0) ? perm[j - lane] : 0;
            if (step > 0) {
                if (low - 32 + lane >= 0) win[(low - 32 + lane) & MSK] = chunk;
                low = max(0, low - 32);
                chunk = (low - 32 + lane >= 0) ? y[low - 32 + lane] : 0.0;
            }
        }
        const int pj = __shfl_sync(0xffffffffu, permchunk, step & 31);
        const int um = min(kv, j);
        double ucur[UQ];
#pragma unroll
        for (int q = 0; q < UQ; ++q) ucur[q] = unext[q];
        const double dcur = dnext;
        if (j > 0) {
            const double* cn = AB + (size_t)(j - 1) * ldab;
            dnext = cn[kv];
            const int umn = min(kv, j - 1);
#pragma unroll
            for (int q = 0; q < UQ; ++q) {
                const int i = 1 + lane + 32 * q;
                unext[q] = (i <= umn) ? cn[kv - i] : 0.0;
            }
            if (j - PF >= 0 && lane < lines_u) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)(AB + (size_t)(j - PF) * ldab) + lane * 128));
        }
        __syncwarp();
        const double xj = win[j & MSK] / dcur;
#pragma unroll
        for (int q = 0; q < UQ; ++q) {
            const int i = 1 + lane + 32 * q;
            if (i <= um) win[(j - i) & MSK] = fma(-ucur[q], xj, win[(j - i) & MSK]);
        }
        if (lane == 0) x[pj] = xj;
        __syncwarp();
    }
}

// Register-window sweeps for narrow bands (kl <= 32, kv = kl + ku <= 64: the MPC case).  Same arithmetic as
// band_solve_kernel (dgbtrs), but the live part of the right-hand side sits in two (forward) / three (backward)
// registers per lane, pivot swaps and the broadcast of x_j are warp shuffles, and the multipliers of 32 columns at a
// time are staged for all warps of the CTA in a 3-slot shared-memory ring by cp.async two blocks ahead -- one
// __syncthreads per 32 columns, nothing but SHFL -> (select | DMUL) -> DFMA on the per-column dependency chain.
// U's diagonal is applied as a reciprocal computed off the chain (one rounding apart from dgbtrs's division).
#ifndef R_WARPS_DEF
#define R_WARPS_DEF 4
#endif
constexpr int R_WARPS = R_WARPS_DEF;
constexpr int R_LROW = 32, R_UROW = 64;
constexpr int R_SLOT = 32 * R_UROW + 32;  // doubles per ring slot (backward layout: 32 columns x 64 + 32 diagonals)

__global__ void __launch_bounds__(32 * R_WARPS) band_solve_reg_kernel(int N, int kl, int ku, const double* __restrict__ AB,
                                                                      const int* __restrict__ ipiv, const int32_t* __restrict__ perm,
                                                                      int nrhs, double* __restrict__ Y, double* __restrict__ X) {
    extern __shared__ __align__(16) double ring[];  // 3 * R_SLOT doubles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x * R_WARPS + warp;
    const bool active = r < nrhs;
    const int kv = kl + ku;
    const size_t ldab = (size_t)(2 * kl + ku + 1);
    double* y = Y + (size_t)(active ? r : 0) * N;
    double* x = X + (size_t)(active ? r : 0) * N;
    const int nblk = (N + 31) >> 5;
    auto cp8 = [](double* dst, const double* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    // ---------------- forward: y <- L^-1 P y ----------------
    auto load_L = [&](int b, int slot) {  // multipliers of columns 32 b .. 32 b + 31 : Ls[c][i - 1], zeros elsewhere
        double* Ls = ring + (size_t)slot * R_SLOT;
        if (b < nblk)
            for (int e = tid; e < 32 * R_LROW; e += 32 * R_WARPS) {
                const int c = e >> 5, i = e & 31, j = 32 * b + c;
                if (j < N && i < min(kl, N - 1 - j)) cp8(Ls + e, AB + (size_t)j * ldab + kv + 1 + i);
                else Ls[e] = 0.0;
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_L(0, 0);
    load_L(1, 1);
    double w0 = (active && lane < N) ? y[lane] : 0.0;
    double w1 = (active && 32 + lane < N) ? y[32 + lane] : 0.0;
    int pcur = lane < N ? ipiv[lane] : 0;
    for (int b = 0; b < nblk; ++b) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();  // block b has landed for every warp; everyone is done with block b - 1
        load_L(b + 2, (b + 2) % 3);
        const int j0 = 32 * b;
        const double nxt = (active && j0 + 64 + lane < N) ? y[j0 + 64 + lane] : 0.0;
        const int pnext = (j0 + 32 + lane < N) ? ipiv[j0 + 32 + lane] : 0;
        const double* Ls = ring + (size_t)(b % 3) * R_SLOT;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            if (j0 + jj < N) {
                const int p = __shfl_sync(0xffffffffu, pcur, jj) - j0;  // jj <= p <= jj + kl < 64
                const double xj = __shfl_sync(0xffffffffu, w0, jj);
                const double xp = __shfl_sync(0xffffffffu, p < 32 ? w0 : w1, p & 31);
                if (p != jj) {
                    if (lane == jj) w0 = xp;
                    if (p < 32) {
                        if (lane == p) w0 = xj;
                    } else if (lane == p - 32) {
                        w1 = xj;
                    }
                }
                const double l = Ls[jj * R_LROW + ((lane - jj - 1) & 31)];
                if (lane > jj) w0 = fma(-l, xp, w0);
                else w1 = fma(-l, xp, w1);
            }
        }
        if (active && j0 + lane < N) y[j0 + lane] = w0;
        w0 = w1;
        w1 = nxt;
        pcur = pnext;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---------------- backward: x <- U^-1 y, scattered through perm ----------------
    auto load_U = [&](int b, int slot) {  // Us[c][i - 1] = U(j - i, j), i = 1 .. kv; diagonals behind the 32 columns
        double* Us = ring + (size_t)slot * R_SLOT;
        if (b >= 0)
            for (int e = tid; e < 32 * R_UROW + 32; e += 32 * R_WARPS) {
                if (e < 32 * R_UROW) {
                    const int c = e >> 6, i = e & 63, j = 32 * b + c;
                    if (j < N && i < min(kv, j)) cp8(Us + e, AB + (size_t)j * ldab + kv - 1 - i);
                    else Us[e] = 0.0;
                } else {
                    const int c = e - 32 * R_UROW, j = 32 * b + c;
                    if (j < N) cp8(Us + e, AB + (size_t)j * ldab + kv);
                    else Us[e] = 1.0;
                }
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_U(nblk - 1, 0);
    load_U(nblk - 2, 1);
    auto ld = [&](int base) { return (active && base + lane >= 0 && base >= 0 && base + lane < N) ? y[base + lane] : 0.0; };
    double v2 = ld(32 * (nblk - 1)), v1 = ld(32 * (nblk - 2)), v0 = ld(32 * (nblk - 3));
    for (int b = nblk - 1, it = 0; b >= 0; --b, ++it) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        load_U(b - 2, (it + 2) % 3);
        const int j0 = 32 * b;
        const double nxt = ld(32 * (b - 3));
        const double* Us = ring + (size_t)(it % 3) * R_SLOT;
        const double rinv = 1.0 / Us[32 * R_UROW + lane];
#pragma unroll
        for (int jj = 31; jj >= 0; --jj) {
            if (j0 + jj < N) {
                const double xj = __shfl_sync(0xffffffffu, v2, jj) * __shfl_sync(0xffffffffu, rinv, jj);
                if (lane == jj) v2 = xj;
                const double ua = Us[jj * R_UROW + ((jj - lane - 1) & 63)];
                const double ub = Us[jj * R_UROW + jj + 31 - lane];
                if (lane < jj) v2 = fma(-ua, xj, v2);
                else v0 = fma(-ua, xj, v0);
                v1 = fma(-ub, xj, v1);
            }
        }
        if (active && j0 + lane < N) x[perm[j0 + lane]] = v2;
        v2 = v1;
        v1 = v0;
        v0 = nxt;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

int32_t sparse_band_setup(diffopt_b200_ctx* ctx, int64_t N, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int32_t trans, int64_t* bandwidth_out) {
    if (!ctx) return -1;
    SparseBandState& S = ctx->sparse;
    S.valid = false;
    if (N <= 0 || !colptr || !rowval || !nzval) BAD_ARG(ctx, "sparse_setup: bad argument");
    if (colptr[0] != 1) BAD_ARG(ctx, "sparse_setup: colptr must be 1-based (Julia SparseMatrixCSC)");
    if (N >= (int64_t)1 << 31) BAD_ARG(ctx, "sparse_setup: N too large");
    DeviceGuard guard_(ctx->device);
    const int64_t nnz = colptr[N] - 1;
    // symmetrised pattern (without the diagonal) as CSR
    std::vector<int64_t> ptr((size_t)N + 1, 0);
    for (int64_t c = 0; c < N; ++c)
        for (int64_t k = colptr[c] - 1; k < colptr[c + 1] - 1; ++k) {
            const int64_t r = rowval[k] - 1;
            if (r < 0 || r >= N) BAD_ARG(ctx, "sparse_setup: row index out of range");
            if (r != c) {
                ++ptr[(size_t)r + 1];
                ++ptr[(size_t)c + 1];
            }
        }
    for (int64_t i = 0; i < N; ++i) ptr[(size_t)i + 1] += ptr[(size_t)i];
    std::vector<int32_t> adj((size_t)ptr[(size_t)N]);
    {
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t c = 0; c < N; ++c)
            for (int64_t k = colptr[c] - 1; k < colptr[c + 1] - 1; ++k) {
                const int64_t r = rowval[k] - 1;
                if (r != c) {
                    adj[(size_t)fill[(size_t)r]++] = (int32_t)c;
                    adj[(size_t)fill[(size_t)c]++] = (int32_t)r;
                }
            }
    }
    std::vector<int32_t> perm;
    rcm_order(N, ptr, adj, perm);
    std::vector<int32_t> inv((size_t)N);
    for (int64_t i = 0; i < N; ++i) inv[(size_t)perm[(size_t)i]] = (int32_t)i;
    int64_t bw = 0;
    for (int64_t c = 0; c < N; ++c)
        for (int64_t k = colptr[c] - 1; k < colptr[c + 1] - 1; ++k) {
            const int64_t d = (int64_t)inv[(size_t)(rowval[k] - 1)] - inv[(size_t)c];
            bw = std::max(bw, d < 0 ? -d : d);
        }
    if (bandwidth_out) *bandwidth_out = bw;
    if (bw > BAND_MAX) {
        char buf[200];
        snprintf(buf, sizeof buf, "sparse_setup: bandwidth %lld after RCM exceeds %d (not a banded pattern)",
                 (long long)bw, BAND_MAX);
        ctx->err = buf;
        return -3;
    }
    const int kl = (int)bw, ku = (int)bw;
    const int64_t ldab = 2 * kl + ku + 1;
    const size_t abytes = sizeof(double) * (size_t)ldab * (size_t)N;
    DO_CUDA(ctx, S.AB.reserve(abytes));
    DO_CUDA(ctx, S.ipiv.reserve(sizeof(int) * (size_t)N));
    DO_CUDA(ctx, S.perm.reserve(sizeof(int32_t) * (size_t)N));
    DO_CUDA(ctx, ctx->in[0].reserve(sizeof(int64_t) * (size_t)(N + 1)));
    DO_CUDA(ctx, ctx->in[1].reserve(sizeof(int64_t) * (size_t)std::max<int64_t>(nnz, 1)));
    DO_CUDA(ctx, ctx->in[2].reserve(sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
    DO_CUDA(ctx, ctx->in[4].reserve(sizeof(int32_t) * (size_t)N));
    DO_CUDA(ctx, ctx->info.reserve(sizeof(int)));
    DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[0].ptr, colptr, sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[1].ptr, rowval, sizeof(int64_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[2].ptr, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(ctx->in[4].ptr, inv.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(S.perm.ptr, perm.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemsetAsync(S.AB.ptr, 0, abytes, ctx->stream));
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    {
        const int64_t blocks = (N * 8 + 255) / 256;
        band_scatter_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(N, ctx->in[0].as<int64_t>(), ctx->in[1].as<int64_t>(),
                                                                       ctx->in[2].as<double>(), ctx->in[4].as<int32_t>(), trans, kl + ku, ldab,
                                                                       S.AB.as<double>());
        ctx->launches++;
    }
    {
        // shared-memory window variant when kv + 2 columns of the band fit (narrow bands, the MPC case)
        int WC = 1;
        while (WC < kl + ku + 2 + 8 + 2) WC <<= 1;  // window + columns in flight (PD <= 8) + the retiring column
        const size_t wbytes = sizeof(double) * (size_t)WC * (size_t)ldab;
        if (wbytes + 16 * 1024 <= ctx->smem_optin && ldab <= F_THREADS) {
            DO_CUDA(ctx, cudaFuncSetAttribute(band_lu_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes));
            // one warp per window column in the update: as many warps as the window has columns (at least 8)
            int thr = 32 * std::min(16, std::max(8, kl + ku + 1));  // measured: 16 warps is the sweet spot (barrier cost vs update width)
            if (getenv("DIFFOPT_B200_BAND_THREADS")) thr = atoi(getenv("DIFFOPT_B200_BAND_THREADS"));
            if (kl <= 31 && !getenv("DIFFOPT_B200_BAND_LU_OLD")) {
                DO_CUDA(ctx, cudaFuncSetAttribute(band_lu_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes));
                band_lu_warp_kernel<<<1, thr, wbytes, ctx->stream>>>((int)N, kl, ku, S.AB.as<double>(), S.ipiv.as<int>(),
                                                                    ctx->info.as<int>(), WC);
            } else {
                band_lu_kernel<true><<<1, thr, wbytes, ctx->stream>>>((int)N, kl, ku, S.AB.as<double>(), S.ipiv.as<int>(),
                                                                      ctx->info.as<int>(), WC);
            }
        } else {
            band_lu_kernel<false><<<1, F_THREADS, 0, ctx->stream>>>((int)N, kl, ku, S.AB.as<double>(), S.ipiv.as<int>(),
                                                                    ctx->info.as<int>(), 1);
        }
        ctx->launches++;
    }
    DO_CUDA(ctx, cudaGetLastError());
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    int hinfo = 0;
    DO_CUDA(ctx, cudaMemcpyAsync(&hinfo, ctx->info.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    if (hinfo != 0) return hinfo;  // exactly singular: the reference throws SingularException
    S.N = N;
    S.kl = kl;
    S.ku = ku;
    S.valid = true;
    return 0;
}

int32_t sparse_band_solve(diffopt_b200_ctx* ctx, int64_t nrhs, const double* rhs, double* x_out, int32_t memspace) {
    if (!ctx) return -1;
    SparseBandState& S = ctx->sparse;
    if (!S.valid) BAD_ARG(ctx, "sparse_solve: no factorisation (call diffopt_b200_sparse_setup first)");
    if (nrhs <= 0 || !rhs || !x_out) BAD_ARG(ctx, "sparse_solve: bad argument");
    DeviceGuard guard_(ctx->device);
    const size_t bytes = sizeof(double) * (size_t)S.N * (size_t)nrhs;
    const void* dB = nullptr;
    void* dX = nullptr;
    DO_CUDA(ctx, stage_in(ctx, ctx->in[5], rhs, bytes, memspace, &dB));
    DO_CUDA(ctx, stage_out_prepare(ctx->out[2], x_out, bytes, memspace, &dX));
    DO_CUDA(ctx, S.work.reserve(bytes));
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    {
        int64_t gb = ((int64_t)S.N * nrhs + 255) / 256;
        if (gb > (int64_t)ctx->sm_count * 16) gb = (int64_t)ctx->sm_count * 16;
        band_gather_kernel<<<(unsigned)gb, 256, 0, ctx->stream>>>(S.N, nrhs, S.perm.as<int32_t>(), (const double*)dB, S.work.as<double>());
        ctx->launches++;
    }
    const int64_t blocks = (nrhs + S_WARPS - 1) / S_WARPS;
#define BAND_SOLVE(QL)                                                                                                             \
    band_solve_kernel<QL><<<(unsigned)blocks, 32 * S_WARPS, 0, ctx->stream>>>((int)S.N, S.kl, S.ku, S.AB.as<double>(), S.ipiv.as<int>(), \
                                                                              S.perm.as<int32_t>(), (int)nrhs, S.work.as<double>(),     \
                                                                              (double*)dX)
    if (S.kl <= 32 && S.kl + S.ku <= 64 && !getenv("DIFFOPT_B200_BAND_SOLVE_OLD")) {
        const size_t rbytes = sizeof(double) * 3 * R_SLOT;
        DO_CUDA(ctx, cudaFuncSetAttribute(band_solve_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rbytes));
        band_solve_reg_kernel<<<(unsigned)((nrhs + R_WARPS - 1) / R_WARPS), 32 * R_WARPS, rbytes, ctx->stream>>>(
            (int)S.N, S.kl, S.ku, S.AB.as<double>(), S.ipiv.as<int>(), S.perm.as<int32_t>(), (int)nrhs, S.work.as<double>(), (double*)dX);
    } else if (S.kl <= 32) BAND_SOLVE(1);
    else if (S.kl <= 64) BAND_SOLVE(2);
    else if (S.kl <= 128) BAND_SOLVE(4);
    else BAND_SOLVE(8);
#undef BAND_SOLVE
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    DO_CUDA(ctx, stage_out_finish(ctx, dX, x_out, bytes, memspace));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    return 0;
}

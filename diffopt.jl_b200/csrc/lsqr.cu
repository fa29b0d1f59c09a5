// Persistent cooperative LSQR kernel + the conic matrix-free operator.  See lsqr.cuh.
#include <cooperative_groups.h>

#include <algorithm>
#include <thread>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lsqr.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int LSQR_THREADS = 256;
constexpr int NSLOTS = 8;

#ifdef LSQR_PROF
__device__ long long g_prof[8192];
__device__ int g_prof_n;
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

struct Dev {
    cg::grid_group grid;
    double* partials;  // [NSLOTS][nblk]
    double* red;       // shared scratch (>= 2 * warps doubles)
    int tid, lane, warp, nwarp, nblk, gtid, gthreads;
    bool cluster = false;  // the whole grid is ONE thread-block cluster: hardware barrier instead of grid.sync()
    // batched launches (one problem per cluster of `nblk` CTAs, lsqr_batch_kernel): rank of this CTA inside its problem
    // (-1: the grid is one problem, rank = blockIdx.x) and whether the problem lives in a single CTA (barrier = bar.sync)
    int cta = -1;
    bool cta_only = false;
    __device__ __forceinline__ int rank() const { return cta >= 0 ? cta : (int)blockIdx.x; }
    // All-CTA barrier with release/acquire ordering of global memory.  The cluster barrier costs ~0.2 us (and
    // invalidates L1, so plain loads after it see the other CTAs' stores); cooperative grid.sync() costs 2-3 us.
    __device__ __forceinline__ void sync() {
#ifdef LSQR_PROF
        if (gtid == 0 && g_prof_n < 8190) g_prof[g_prof_n++] = gtimer();
#endif
        sync_();
#ifdef LSQR_PROF
        if (gtid == 0 && g_prof_n < 8190) g_prof[g_prof_n++] = gtimer();
#endif
    }
    __device__ __forceinline__ void sync_() {
        if (cta_only)
            __syncthreads();
        else if (cluster)
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        else
            grid.sync();
    }
};

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level sum of `v`, result stored as this CTA's partial in `slot`
__device__ void block_partial(Dev& d, int slot, double v) {
    v = warp_sum_all(v);
    if (d.lane == 0) d.red[d.warp] = v;
    __syncthreads();
    if (d.tid == 0) {
        double s = 0.0;
        for (int w = 0; w < d.nwarp; ++w) s += d.red[w];
        d.partials[slot * d.nblk + d.rank()] = s;
    }
    __syncthreads();
}

// after a grid.sync(): fixed-order sum of all CTA partials of `slot` (bitwise identical in every thread)
__device__ double total_of(Dev& d, int slot) {
    const double* p = d.partials + slot * d.nblk;
    double s = 0.0;
    for (int i = d.lane; i < d.nblk; i += 32) s += __ldcg(p + i);
    return warp_sum_all(s);
}

// dot product of sparse row `row` with x, computed by a group of T lanes (T = 1,2,4,...,32; groups are
// aligned inside the warp).  All 32 lanes must call this together.
template <bool kScaleNone = true>
__device__ __forceinline__ double row_dot(const CsrView& A, int row, bool valid, const double* __restrict__ x,
                                          int lane_in_group, int T) {
    double acc = 0.0;
    if (valid) {
        int s = A.rowptr[row], e = A.rowptr[row + 1];
        for (int k = s + lane_in_group; k < e; k += T) acc += A.val[k] * __ldcg(x + A.colind[k]);
    }
    for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

__device__ __forceinline__ int pick_T(const CsrView& A) {
    int nnz = A.rowptr[A.nrows];
    int avg = A.nrows > 0 ? nnz / A.nrows : 1;
    int T = 1;
    while (T < 32 && T * 2 <= avg) T <<= 1;
    return T;
}


// y_row = (sparse row) . x for every row of A, `epi(row, value)` called once per row (by the first lane of the row's
// lane group).  Each group of T lanes keeps RU rows in flight at once (RU independent rowptr -> (val, colind) -> x gather
// chains per lane) so that a persistent grid of a few hundred CTAs still covers the HBM latency on large matrices;
// consecutive groups take consecutive rows, so val / colind accesses of a warp stay contiguous.
template <int RU, class F>
__device__ __forceinline__ void spmv_rows(const Dev& d, const CsrView& A, const double* __restrict__ x, F&& epi) {
    const int T = pick_T(A);
    const int lig = d.lane & (T - 1), ngroups = d.gthreads / T, g = d.gtid / T;
    for (int base = 0; base < A.nrows; base += ngroups * RU) {
        int s[RU], len[RU];
        int maxlen = 0;
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int row = base + r * ngroups + g;
            s[r] = 0;
            len[r] = 0;
            if (row < A.nrows) {
                s[r] = A.rowptr[row];
                len[r] = A.rowptr[row + 1] - s[r];
            }
            maxlen = max(maxlen, len[r]);
        }
        maxlen = __reduce_max_sync(0xffffffffu, maxlen);  // keep the warp converged for the shuffles below
        double acc[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) acc[r] = 0.0;
        for (int k = lig; k < maxlen; k += T) {
            double av[RU];
            int ci[RU];
#pragma unroll
            for (int r = 0; r < RU; ++r) {
                const bool ok = k < len[r];
                av[r] = ok ? A.val[s[r] + k] : 0.0;
                ci[r] = ok ? A.colind[s[r] + k] : 0;
            }
#pragma unroll
            for (int r = 0; r < RU; ++r) acc[r] = fma(av[r], __ldcg(x + ci[r]), acc[r]);
        }
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            double t = acc[r];
            for (int o = T >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            const int row = base + r * ngroups + g;
            if (row < A.nrows && lig == 0) epi(row, t);
        }
    }
}

constexpr int SP_THREADS = 128;                   // threads of a streaming-SpMV CTA (8 CTAs per SM)
constexpr int ST_NPT = 8;                         // nonzeros per thread in flight
constexpr int ST_CHUNK = SP_THREADS * ST_NPT;     // nonzeros per row block
// cluster-synchronised kernel (small operators): one cluster of CL_THREADS-thread CTAs, same SpMV with bigger blocks
#ifndef CL_THREADS_DEF
#define CL_THREADS_DEF 512
#endif
constexpr int CL_THREADS = CL_THREADS_DEF;
constexpr int CL_NPT = 8;
constexpr int CL_CHUNK = CL_THREADS * CL_NPT;
constexpr int CL_MAX_CTAS = 16;
constexpr int ST_STRIDE = 2048;  // capacity of one partial slot = upper bound of every streaming grid

// Row-block ("CSR-stream") SpMV for matrices beyond L2.  The host cuts the rows into blocks of <= ST_CHUNK
// nonzeros (and <= SP_THREADS rows, or one longer row) and stores (first row, first nonzero) per block.  A CTA owns a
// CONTIGUOUS range of blocks, so the part of x it gathers from slides slowly and stays in L1 when the matrix has any
// locality (the kernels keep shared memory small to leave L1 its capacity).  Per block: NPT independent, fully
// coalesced (val, colind) loads per thread -- no dependence on row pointers, so the HBM pipe stays full regardless of
// the row lengths -- then the x gathers (read-only path) with the row's epilogue operands loaded underneath them
// (`pre(b, row)`), products parked in shared memory, one row per thread reduced in storage order (deterministic),
// `epi(b, row, value, operands)` once per row.
// sel(b, A, x, lb): matrix, gather vector and local block index of global block b.
// XCG: x was written earlier in the SAME launch (persistent kernels) -> gather with ld.global.cg, not the read-only path.
// XM: how x is gathered.  0: read-only path (x constant during the launch); 1: ld.global.cg (x was written earlier in the
// SAME launch by other CTAs: persistent kernels); 2: plain generic loads -- x is a copy staged in SHARED memory by the
// caller (one problem per CTA: the whole gather vector fits, and shared-memory gathers are not bound by the one-sector-
// per-request rate of L1TEX that caps the global-memory forms at ~50 % of HBM).  The matrix streams evict-first for 0, 2.
// cta / ncta: this CTA's rank among the CTAs that share the blocks.
template <int THREADS, int NPT, int XM, class Sel, class Pre, class Epi>
__device__ __forceinline__ void spmv_stream(int cta, int ncta, int nblk_total, double* prod, double* red, Sel&& sel, Pre&& pre,
                                            Epi&& epi) {
    constexpr int CHUNK = THREADS * NPT;
    constexpr bool XCG = XM == 1;
    const int tid = threadIdx.x;
    const int per = (nblk_total + ncta - 1) / ncta;
    const int b_end = min(nblk_total, (cta + 1) * per);
    for (int b = cta * per; b < b_end; ++b) {
        const CsrView* A;
        const double* x;
        int lb;
        sel(b, A, x, lb);
        const int2 d0 = __ldg(reinterpret_cast<const int2*>(A->blk) + lb);
        const int2 d1 = __ldg(reinterpret_cast<const int2*>(A->blk) + lb + 1);
        const int r0 = d0.x, s = d0.y, r1 = d1.x, cnt = d1.y - d0.y;
        if (cnt > CHUNK) {  // one long row: whole CTA, fixed-order reduction
            double acc = 0.0;
            for (int k = tid; k < cnt; k += THREADS)
                {
                const double* xp = x + (XCG ? __ldg(A->colind + s + k) : __ldcs(A->colind + s + k));
                acc = fma(XCG ? __ldg(A->val + s + k) : __ldcs(A->val + s + k), XM == 2 ? *xp : (XCG ? __ldcg(xp) : __ldg(xp)), acc);
            }
            acc = warp_sum_all(acc);
            if ((tid & 31) == 0) red[32 + (tid >> 5)] = acc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < THREADS / 32; ++w) t += red[32 + w];
                epi(b, r0, t, pre(b, r0));
            }
        } else {
            double av[NPT];
            int ci[NPT];
            int ps[NPT];  // slot of the product in `prod` (CSR position inside the block)
            const bool sorted = XM == 0 && A->spos != nullptr;  // column-ordered copy of the block (large operators)
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                const int k = tid + u * THREADS;
                av[u] = 0.0;
                ci[u] = 0;
                ps[u] = k;
                if (k < cnt) {  // streamed once per pass (evict-first) unless the operator lives in L2 (XCG kernels)
                    if (sorted) {
                        av[u] = __ldcs(A->sval + s + k);
                        ci[u] = __ldcs(A->scol + s + k);
                        ps[u] = __ldcs(A->spos + s + k);
                    } else {
                        av[u] = XCG ? __ldg(A->val + s + k) : __ldcs(A->val + s + k);
                        ci[u] = XCG ? __ldg(A->colind + s + k) : __ldcs(A->colind + s + k);
                    }
                }
            }
            // T threads per row (power of two, as many as the block's row count allows): rows of a few hundred nonzeros
            // (KKT matrices with dense blocks) are not summed by one thread; fixed order either way
            int T = 1;
            while (T < 32 && (r1 - r0) * T * 2 <= THREADS) T <<= 1;
            const int row = r0 + tid / T, sub = tid & (T - 1);
            int ra = 0, rb = 0;
            if (row < r1) {
                ra = __ldg(A->rowptr + row) - s;
                rb = __ldg(A->rowptr + row + 1) - s;
            }
            double xv[NPT];
#pragma unroll
            for (int u = 0; u < NPT; ++u)
                xv[u] = (tid + u * THREADS < cnt) ? (XM == 2 ? x[ci[u]] : (XCG ? __ldcg(x + ci[u]) : __ldg(x + ci[u]))) : 0.0;
            double4 opnd = make_double4(0.0, 0.0, 0.0, 0.0);
            if (row < r1 && sub == 0) opnd = pre(b, row);
#pragma unroll
            for (int u = 0; u < NPT; ++u)
                if (tid + u * THREADS < cnt) prod[ps[u]] = av[u] * xv[u];
            __syncthreads();
            double t = 0.0;
            for (int k = ra + sub; k < rb; k += T) t += prod[k];
            for (int o = T >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (row < r1 && sub == 0) epi(b, row, t, opnd);
        }
        __syncthreads();  // prod / red are free again
    }
}

// ---- SOC block of Dpi applied to y (symmetric block, so it serves Dpi and Dpi') ------------------
// group of `T` lanes handles cone `c`; writes out[off .. off+dim)
__device__ void soc_apply(const ConicOpView& op, int c, bool valid, const double* __restrict__ y, double* out,
                          int lig, int T) {
    int off = 0, dim = 0, cs = 0;
    double nx = 1.0, t = 0.0;
    if (valid) {
        off = op.soc_off[c];
        dim = op.soc_dim[c];
        cs = op.soc_case[c];
        nx = op.soc_nx[c];
        t = op.v[off];
    }
    double xy = 0.0;
    if (valid && cs == 2)
        for (int k = 1 + lig; k < dim; k += T) xy += op.v[off + k] * __ldcg(y + off + k);
    for (int o = T >> 1; o > 0; o >>= 1) xy += __shfl_xor_sync(0xffffffffu, xy, o);
    if (!valid) return;
    if (cs == 0) {
        for (int k = lig; k < dim; k += T) out[off + k] = __ldcg(y + off + k);
    } else if (cs == 1) {
        for (int k = lig; k < dim; k += T) out[off + k] = 0.0;
    } else {
        const double y0 = __ldcg(y + off);
        const double inv2 = 1.0 / (2.0 * nx);
        const double coef = t / (nx * nx) * xy;
        for (int k = lig; k < dim; k += T) {
            double r;
            if (k == 0)
                r = nx * y0 + xy;
            else {
                double xk = op.v[off + k];
                r = xk * y0 + (nx + t) * __ldcg(y + off + k) - coef * xk;
            }
            out[off + k] = r * inv2;
        }
    }
}

// ---- PSD block: out = Dpi*y (transpose=false) or Dpi'*y (true), whole grid cooperates ------------
// Dpi' y = vec(F(unvec(y)));  Dpi y = S' F T' y  (diag of unvec doubled on the way in, halved on the way out),
// F(X) = U (B o (U' X U)) U'.  Four GEMM phases on the FP64 tensor pipe (mma.sync m8n8k4), 3 barriers inside:
//   W1 = U' X        (X read straight from the triangle y)
//   W2 = (W1 U) o B
//   W1 = U W2
//   out = vec(W1 U') (upper-triangle tiles only, written straight into the triangle)
// Work unit = one 32 x 32 output tile of one cone (tiles of all cones in one list, so 512 small cones fill the grid as
// well as one 200 x 200 cone does); a 256-thread CTA stages 32-deep operand panels in shared memory ([k][i] layout,
// leading dimension 36: fragment reads are bank-conflict free), next panel prefetched into registers under the MMAs.
constexpr int PT = 32;        // tile side
constexpr int PLD = PT + 4;   // panel leading dimension (doubles)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// C(i0.., j0..) = sum_k la(i, k) * lb(k, j); st(i, j, value) for every in-range element of the tile.
// a_ic / b_kc: which index is contiguous in memory for the A / B operand (chooses the coalesced load mapping).
template <class LA, class LB, class ST>
__device__ __forceinline__ void psd_tile_gemm(int dd, int i0, int j0, bool a_ic, bool b_kc, double* As, double* Bs,
                                              LA&& la, LB&& lb, ST&& st) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int iw = (warp & 3) * 8, jw = (warp >> 2) * 16;
    const int f = tid & 31, s8 = tid >> 5;  // fast / slow index of the load mapping (slow: s8 + 8 r)
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    double ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int sl = s8 + 8 * r;
            const int ia = a_ic ? f : sl, ka = a_ic ? sl : f;
            const int kb = b_kc ? f : sl, jb = b_kc ? sl : f;
            ra[r] = (i0 + ia < dd && k0 + ka < dd) ? la(i0 + ia, k0 + ka) : 0.0;
            rb[r] = (k0 + kb < dd && j0 + jb < dd) ? lb(k0 + kb, j0 + jb) : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < dd; k0 += PT) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int sl = s8 + 8 * r;
            const int ia = a_ic ? f : sl, ka = a_ic ? sl : f;
            const int kb = b_kc ? f : sl, jb = b_kc ? sl : f;
            As[ka * PLD + ia] = ra[r];
            Bs[kb * PLD + jb] = rb[r];
        }
        __syncthreads();
        if (k0 + PT < dd) fetch(k0 + PT);
#pragma unroll
        for (int kk = 0; kk < PT; kk += 4) {
            const int kr = (kk + (lane & 3)) * PLD;
            const double av = As[kr + iw + (lane >> 2)];
            const double b0 = Bs[kr + jw + (lane >> 2)];
            const double b1 = Bs[kr + jw + 8 + (lane >> 2)];
            dmma884(c00, c01, av, b0);
            dmma884(c10, c11, av, b1);
        }
        __syncthreads();
    }
    const int i = i0 + iw + (lane >> 2), j = j0 + jw + 2 * (lane & 3);
    if (i < dd) {
        if (j < dd) st(i, j, c00);
        if (j + 1 < dd) st(i, j + 1, c01);
        if (j + 8 < dd) st(i, j + 8, c10);
        if (j + 9 < dd) st(i, j + 9, c11);
    }
}

__device__ void psd_apply_all(Dev& d, const ConicOpView& op, const double* __restrict__ y, double* out,
                              bool transpose) {
    if (op.npsd == 0) return;
    __shared__ double As[PT * PLD], Bs[PT * PLD];
    const double dscale = transpose ? 1.0 : 2.0, oscale = transpose ? 1.0 : 0.5;
    for (int phase = 0; phase < 4; ++phase) {
        for (int t = blockIdx.x; t < op.psd_ntiles; t += gridDim.x) {
            // cone of tile t: last c with toff[c] <= t
            int lo = 0, hi = op.npsd - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(op.psd_toff + mid) <= t) lo = mid; else hi = mid - 1;
            }
            const int c = lo, dd = op.psd_d[c], off = op.psd_off[c];
            const int nt = (dd + PT - 1) / PT, lt = t - __ldg(op.psd_toff + c);
            const int i0 = (lt % nt) * PT, j0 = (lt / nt) * PT;
            const long long uo = op.psd_uoff[c];
            const double* U = op.psd_U + uo;
            double* W1 = op.psd_w1 + uo;
            double* W2 = op.psd_w2 + uo;
            const double* yc = y + off;
            if (op.psd_ident[c]) {  // Dpi = I on this cone
                if (phase == 3)
                    for (int e = threadIdx.x; e < PT * PT; e += blockDim.x) {
                        const int i = i0 + e % PT, j = j0 + e / PT;
                        if (i <= j && j < dd) out[off + (long long)j * (j + 1) / 2 + i] = __ldcg(yc + (long long)j * (j + 1) / 2 + i);
                    }
                continue;
            }
            if (phase == 0) {  // W1 = U' X
                psd_tile_gemm(dd, i0, j0, false, true, As, Bs,
                    [&](int i, int k) { return U[k + (long long)i * dd]; },
                    [&](int k, int j) {
                        const int r = k < j ? k : j, cc = k < j ? j : k;
                        const double v = __ldcg(yc + (long long)cc * (cc + 1) / 2 + r);
                        return k == j ? v * dscale : v;
                    },
                    [&](int i, int j, double v) { W1[i + (long long)j * dd] = v; });
            } else if (phase == 1) {  // W2 = (W1 U) o B
                const double* Bm = op.psd_Bm + uo;
                psd_tile_gemm(dd, i0, j0, true, true, As, Bs,
                    [&](int i, int k) { return __ldcg(W1 + i + (long long)k * dd); },
                    [&](int k, int j) { return U[k + (long long)j * dd]; },
                    [&](int i, int j, double v) { W2[i + (long long)j * dd] = v * Bm[i + (long long)j * dd]; });
            } else if (phase == 2) {  // W1 = U W2
                psd_tile_gemm(dd, i0, j0, true, true, As, Bs,
                    [&](int i, int k) { return U[i + (long long)k * dd]; },
                    [&](int k, int j) { return __ldcg(W2 + k + (long long)j * dd); },
                    [&](int i, int j, double v) { W1[i + (long long)j * dd] = v; });
            } else if (i0 <= j0 + PT - 1) {  // out = vec(W1 U'), tiles touching the upper triangle
                psd_tile_gemm(dd, i0, j0, true, false, As, Bs,
                    [&](int i, int k) { return __ldcg(W1 + i + (long long)k * dd); },
                    [&](int k, int j) { return U[j + (long long)k * dd]; },
                    [&](int i, int j, double v) {
                        if (i <= j) out[off + (long long)j * (j + 1) / 2 + i] = i == j ? v * oscale : v;
                    });
            }
        }
        if (phase < 3) d.sync();
    }
    // caller syncs
}

// Dpi (or Dpi') applied to y -> out, for all cones; ends WITHOUT a grid sync when there is no PSD cone
// PSD = false compiles the PSD-cone phases out (the cluster / row-block variants and the streaming driver never run
// problems with PSD cones; their kernels stay small).
template <bool PSD>
__device__ void dpi_apply(Dev& d, const ConicOpView& op, const double* __restrict__ y, double* out, bool transpose) {
    for (int i = d.gtid; i < op.m; i += d.gthreads)
        if (op.kind[i] == 0) out[i] = op.diag[i] * __ldcg(y + i);
    {
        const int T = 8;
        const int lig = d.lane & (T - 1);
        const int ngroups = d.gthreads / T;
        const int g = d.gtid / T;
        for (int base = 0; base < op.nsoc; base += ngroups) {
            int c = base + g;
            soc_apply(op, c, c < op.nsoc, y, out, lig, T);
        }
    }
    if constexpr (PSD) psd_apply_all(d, op, y, out, transpose);
}

// ------------------------------------------------------------------------------------------------
// Operator application  dst = OP(src)*s_src + s_dst*dst  with ||dst||^2 reduced into `slot`.
// CSR:   OP = A (rows) or its transpose view (the caller hands the right CsrView).
// CONIC: OP = M or M'.
// All functions are called by the whole grid and contain grid syncs.

extern __shared__ double cl_prod[];  // CL_CHUNK products (dynamic shared memory of the cluster launch)

// CL = cluster-synchronised launch: the SpMVs run as row-block streams (every load of a thread's share is issued up
// front: two L2 latencies per product instead of two per lane-group step), see spmv_stream.
template <bool CL>
__device__ void csr_apply(Dev& d, const CsrView& A, const double* __restrict__ src, double s_src, double* dst,
                          double s_dst, int slot) {
    double acc = 0.0;
    if constexpr (CL) {
        spmv_stream<CL_THREADS, CL_NPT, 1>(
            d.rank(), d.nblk, A.nblk, cl_prod, d.red, [&](int b, const CsrView*& M, const double*& x, int& lb) { M = &A; x = src; lb = b; },
            [&](int, int row) { return make_double4(s_dst != 0.0 ? __ldcg(dst + row) : 0.0, 0.0, 0.0, 0.0); },
            [&](int, int row, double t, const double4& o) {
                t = t * s_src + s_dst * o.x;
                dst[row] = t;
                acc += t * t;
            });
    } else {
        spmv_rows<4>(d, A, src, [&](int row, double t) {
            t = t * s_src + s_dst * dst[row];
            dst[row] = t;
            acc += t * t;
        });
    }
    block_partial(d, slot, acc);
}

// dst = s_src * (M src) + s_dst * dst ; slots: slot (norm), slot+1 (last-row dot)
// `src_last` / `dst_last` are every thread's private copies of src[N-1] / dst[N-1] (all threads compute the last row
// identically): nobody reads those entries from memory, so thread 0 may store the new one right after the barrier
// that publishes the partial sums, and ||dst||^2 (returned in `norm2`) is complete without another barrier.
// SM (batched launches, one problem per CTA): the vectors the products gather from are first copied into shared memory
// (behind the CL_CHUNK products: n + m doubles), see spmv_stream XM = 2.
template <bool CL, bool PSD, bool SM = false>
__device__ void conic_apply(Dev& d, const ConicOpView& op, bool transpose, const double* __restrict__ src,
                            double s_src, double* dst, double s_dst, int slot, double src_last, double& dst_last,
                            double& norm2) {
    const int n = op.n, m = op.m, N = n + m + 1;
    double dotacc = 0.0;
    if constexpr (CL) {
        double* prod = cl_prod;
        double* xs_m = cl_prod + CL_CHUNK;  // staged gather vectors (SM): m entries, then n entries
        double* xs_n = xs_m + m;
        constexpr int XM = SM ? 2 : 1;
        // (one problem per CTA: every producer of these vectors ran on this SM, so the L1-allocating cp.async sees them;
        // the copies are in flight together, each thread waits for its own and the caller's barrier publishes them)
        auto stage = [&](double* to, const double* from, int len) {
            for (int i = d.tid; i < len; i += CL_THREADS)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(to + i)), "l"(from + i)
                             : "memory");
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
        };
        double acc = 0.0;
        if (!transpose) {
            dpi_apply<PSD>(d, op, src + n, op.wc, false);  // wc = Dpi t2
            d.sync();
            if constexpr (SM) {
                stage(xs_m, op.wc, m);
                stage(xs_n, src, n);
                __syncthreads();
            }
            const double t3 = src_last;
            const int nbt = op.At.nblk;
            spmv_stream<CL_THREADS, CL_NPT, XM>(
                d.rank(), d.nblk, nbt + op.A.nblk, prod, d.red,
                [&](int b, const CsrView*& A, const double*& x, int& lb) {
                    if (b < nbt) { A = &op.At; x = SM ? xs_m : op.wc; lb = b; }
                    else { A = &op.A; x = SM ? xs_n : src; lb = b - nbt; }
                },
                [&](int b, int row) {
                    return b < nbt ? make_double4(op.c[row], __ldcg(dst + row), __ldcg(src + row), 0.0)
                                   : make_double4(__ldcg(op.wc + row), __ldcg(src + n + row), op.b[row],
                                                  __ldcg(dst + n + row));
                },
                [&](int b, int row, double t, const double4& o) {
                    if (b < nbt) {  // rows 0..n-1:  (A' wc)_j + c_j t3
                        t = (t + o.x * t3) * s_src + s_dst * o.y;
                        dst[row] = t;
                        acc += t * t;
                        dotacc += o.x * o.z;
                    } else {        // rows n..n+m-1:  -(A t1)_i + t2_i - wc_i + b_i t3
                        t = (-t + o.y - o.x + o.z * t3) * s_src + s_dst * o.w;
                        dst[n + row] = t;
                        acc += t * t;
                        dotacc += o.z * o.x;
                    }
                });
            block_partial(d, slot, acc);
            block_partial(d, slot + 1, dotacc);
            d.sync();
            const double last = -total_of(d, slot + 1) * s_src + s_dst * dst_last;  // -(c't1 + b'wc)
            dst_last = last;
            if (d.gtid == 0) dst[N - 1] = last;
            norm2 = total_of(d, slot) + last * last;
        } else {
            const double u3 = src_last;
            if constexpr (SM) {
                stage(xs_n, src, n);
                __syncthreads();
            }
            spmv_stream<CL_THREADS, CL_NPT, XM>(  // wc = A u1 - u2 - b u3
                d.rank(), d.nblk, op.A.nblk, prod, d.red,
                [&](int b, const CsrView*& A, const double*& x, int& lb) { A = &op.A; x = SM ? xs_n : src; lb = b; },
                [&](int, int row) { return make_double4(__ldcg(src + n + row), op.b[row], 0.0, 0.0); },
                [&](int, int row, double t, const double4& o) {
                    op.wc[row] = t - o.x - o.y * u3;
                    dotacc += o.y * o.x;
                });
            d.sync();
            double* r2 = op.wc + m;
            dpi_apply<PSD>(d, op, op.wc, r2, true);
            if constexpr (SM) stage(xs_m, src + n, m);   // (src is not touched by dpi_apply: no barrier needed before)
            d.sync();
            spmv_stream<CL_THREADS, CL_NPT, XM>(  // rows 0..n-1: -(A' u2)_j - c_j u3
                d.rank(), d.nblk, op.At.nblk, prod, d.red,
                [&](int b, const CsrView*& A, const double*& x, int& lb) { A = &op.At; x = SM ? xs_m : src + n; lb = b; },
                [&](int, int row) { return make_double4(op.c[row], __ldcg(dst + row), __ldcg(src + row), 0.0); },
                [&](int, int row, double t, const double4& o) {
                    t = (-t - o.x * u3) * s_src + s_dst * o.y;
                    dst[row] = t;
                    acc += t * t;
                    dotacc += o.x * o.z;
                });
            for (int i = d.gtid; i < m; i += d.gthreads) {
                const double t = (__ldcg(r2 + i) + __ldcg(src + n + i)) * s_src + s_dst * __ldcg(dst + n + i);
                dst[n + i] = t;
                acc += t * t;
            }
            block_partial(d, slot, acc);
            block_partial(d, slot + 1, dotacc);
            d.sync();
            const double last = total_of(d, slot + 1) * s_src + s_dst * dst_last;
            dst_last = last;
            if (d.gtid == 0) dst[N - 1] = last;
            norm2 = total_of(d, slot) + last * last;
        }
        return;
    }
    if (!transpose) {
        // wc = Dpi t2
        dpi_apply<PSD>(d, op, src + n, op.wc, false);
        d.sync();
        const double t3 = src_last;
        double acc = 0.0;
        // rows 0..n-1:  (A' wc)_j + c_j t3
        spmv_rows<4>(d, op.At, op.wc, [&](int row, double t) {
            t = (t + op.c[row] * t3) * s_src + s_dst * dst[row];
            dst[row] = t;
            acc += t * t;
            dotacc += op.c[row] * __ldcg(src + row);
        });
        // rows n..n+m-1:  -(A t1)_i + t2_i - wc_i + b_i t3
        spmv_rows<4>(d, op.A, src, [&](int row, double t) {
            const double wci = __ldcg(op.wc + row);
            t = (-t + __ldcg(src + n + row) - wci + op.b[row] * t3) * s_src + s_dst * dst[n + row];
            dst[n + row] = t;
            acc += t * t;
            dotacc += op.b[row] * wci;
        });
        block_partial(d, slot, acc);
        block_partial(d, slot + 1, dotacc);
        d.sync();
        // last row: -(c't1 + b'wc)
        const double last = -total_of(d, slot + 1) * s_src + s_dst * dst_last;
        dst_last = last;
        if (d.gtid == 0) dst[N - 1] = last;
        norm2 = total_of(d, slot) + last * last;
    } else {
        // r = A u1 - u2 - b u3  -> wc
        const double u3 = src_last;
        spmv_rows<4>(d, op.A, src, [&](int row, double t) {
            const double u2 = __ldcg(src + n + row);
            op.wc[row] = t - u2 - op.b[row] * u3;
            dotacc += op.b[row] * u2;
        });
        d.sync();
        // out2 = Dpi' r + u2 : first Dpi' r into a second scratch = reuse psd-free path by writing to dst later.
        // We need dst's old value (s_dst * dst), so stage Dpi' r in op.wc's partner buffer: the x-part of dst is
        // disjoint, so compute into `tmp = psd_w-independent` region: use op.wc in place is unsafe (SOC reads all
        // entries of its cone) -> stage through dst? no.  Use the dedicated second scratch stored after wc.
        double* r2 = op.wc + m;  // second half of the 2m scratch
        dpi_apply<PSD>(d, op, op.wc, r2, true);
        d.sync();
        double acc = 0.0;
        // rows 0..n-1: -(A' u2)_j - c_j u3
        spmv_rows<4>(d, op.At, src + n, [&](int row, double t) {
            t = (-t - op.c[row] * u3) * s_src + s_dst * dst[row];
            dst[row] = t;
            acc += t * t;
            dotacc += op.c[row] * __ldcg(src + row);
        });
        for (int i = d.gtid; i < m; i += d.gthreads) {
            double t = (__ldcg(r2 + i) + __ldcg(src + n + i)) * s_src + s_dst * dst[n + i];
            dst[n + i] = t;
            acc += t * t;
        }
        block_partial(d, slot, acc);
        block_partial(d, slot + 1, dotacc);
        d.sync();
        const double last = total_of(d, slot + 1) * s_src + s_dst * dst_last;
        dst_last = last;
        if (d.gtid == 0) dst[N - 1] = last;
        norm2 = total_of(d, slot) + last * last;
    }
}

struct OpArgs {
    int kind;  // 0 csr, 1 conic
    CsrView fwd, adj;  // csr: matrix for A*v and for A'*u
    ConicOpView conic;
    int conic_trans;   // solve with M' instead of M
    int nrows, ncols;
};

// Returns ||dst||^2 (after the update).  The conic operator ends behind its own barrier; the CSR one needs a barrier
// here before its partial sums can be read.
template <bool CL, bool PSD, bool SM = false>
__device__ double op_apply(Dev& d, const OpArgs& o, bool adjoint, const double* src, double s_src, double* dst,
                           double s_dst, int slot, double src_last, double& dst_last) {
    if (o.kind == 0) {
        csr_apply<CL>(d, adjoint ? o.adj : o.fwd, src, s_src, dst, s_dst, slot);
        d.sync();
        return total_of(d, slot);
    }
    double norm2 = 0.0;
    conic_apply<CL, PSD, SM>(d, o.conic, adjoint != (o.conic_trans != 0), src, s_src, dst, s_dst, slot, src_last, dst_last, norm2);
    return norm2;
}

// The whole LSQR iteration (Paige-Saunders) for one operator; called by every thread of the CTAs that share the problem
// (the whole grid, one cluster, or -- batched launches -- one cluster / CTA per problem).  zero_below: right-hand sides of
// norm <= zero_below give x = 0 without iterating (ConicProgram.jl:369 tests 1e-4 on the reverse seed; -1 disables).
template <bool STREAM, bool PSD, bool SM>
__device__ void lsqr_body(Dev& d, const OpArgs& o, const double* __restrict__ rhs, const LsqrParams prm, const LsqrVectors vec,
                          const double zero_below) {
    const int nr = o.nrows, nc = o.ncols;
    double *u = vec.u, *v = vec.v, *w = vec.w, *x = vec.x;

    // private copies of the last entries of u and v (see conic_apply)
    double u_last = nr > 0 ? rhs[nr - 1] : 0.0, v_last = 0.0;
    // u = b, beta = ||b||
    {
        double acc = 0.0;
        for (int i = d.gtid; i < nr; i += d.gthreads) {
            double t = rhs[i];
            u[i] = t;
            acc += t * t;
        }
        for (int i = d.gtid; i < nc; i += d.gthreads) {
            x[i] = 0.0;
            v[i] = 0.0;
            w[i] = 0.0;
        }
        block_partial(d, 0, acc);
    }
    d.sync();
    double beta = sqrt(total_of(d, 0));
    if (beta <= zero_below) beta = 0.0;
    double alpha = 0.0;
    double su = 1.0, sv = 1.0;  // true u = su * u_mem, true v = sv * v_mem
    int istop = 0;
    long long itn = 0;
    double anorm = 0, acond = 0, ddnorm = 0, res2 = 0, xnorm = 0, xxnorm = 0, z = 0, sn2 = 0, cs2 = -1.0;
    double rnorm = beta, arnorm = 0.0;
    if (beta > 0) {
        su = 1.0 / beta;
        alpha = sqrt(op_apply<STREAM, PSD, SM>(d, o, true, u, su, v, 0.0, 2, u_last, v_last));   // v_mem = A' u_true
    }
    if (alpha > 0) sv = 1.0 / alpha;
    arnorm = alpha * beta;
    const double bnorm = beta;
    double rhobar = alpha, phibar = beta;
    const double ctol = prm.conlim > 0 ? 1.0 / prm.conlim : 0.0;
    double t1 = 0, t2 = 0, rho = 1.0, tau = 0;
    bool first = true;
    bool pending = false;  // an iteration whose x/w update and stop tests are still to be applied

    if (arnorm != 0.0) {
        while (true) {
            // ---- deferred vector update of the previous iteration (or w = v at start), fused with the next
            //      bidiagonalisation product
            double dd = 0.0;
            // (conic operator: the last entry of v is read from its register copy, see conic_apply)
            const int ilast = o.kind == 1 ? nc - 1 : -1;
            if (first) {
                for (int i = d.gtid; i < nc; i += d.gthreads) w[i] = sv * (i == ilast ? v_last : v[i]);
            } else {
                const double irho = 1.0 / rho;
                int i = d.gtid;
                for (; i + 3 * d.gthreads < nc; i += 4 * d.gthreads) {   // four independent elements in flight per thread
                    double wi[4], vi[4], xi[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = i + q * d.gthreads;
                        wi[q] = w[k];
                        vi[q] = k == ilast ? v_last : v[k];
                        xi[q] = x[k];
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = i + q * d.gthreads;
                        const double dk = wi[q] * irho;
                        dd += dk * dk;
                        x[k] = xi[q] + t1 * wi[q];
                        w[k] = sv * vi[q] + t2 * wi[q];
                    }
                }
                for (; i < nc; i += d.gthreads) {
                    double wi = w[i];
                    double dk = wi * irho;
                    dd += dk * dk;
                    x[i] += t1 * wi;
                    w[i] = sv * (i == ilast ? v_last : v[i]) + t2 * wi;
                }
            }
            block_partial(d, 4, dd);
            const bool more = itn < prm.maxiter;
            double unorm2 = 0.0;
            if (more) unorm2 = op_apply<STREAM, PSD, SM>(d, o, false, v, sv, u, -alpha * su, 0, v_last, u_last);  // u_mem = A v_true - alpha u_true
            else d.sync();  // the partial sums of the update above
            if (pending) {
                ddnorm += total_of(d, 4);
                acond = anorm * sqrt(ddnorm);
                const double test1 = rnorm / bnorm;
                const double test2 = arnorm / (anorm * rnorm);
                const double test3 = 1.0 / acond;
                const double t1_ = test1 / (1.0 + anorm * xnorm / bnorm);
                const double rtol = prm.btol + prm.atol * anorm * xnorm / bnorm;
                if (itn >= prm.maxiter) istop = 7;
                if (1.0 + test3 <= 1.0) istop = 6;
                if (1.0 + test2 <= 1.0) istop = 5;
                if (1.0 + t1_ <= 1.0) istop = 4;
                if (test3 <= ctol) istop = 3;
                if (test2 <= prm.atol) istop = 2;
                if (test1 <= rtol) istop = 1;
                pending = false;
            }
            first = false;
            if (istop > 0 || !more) break;
            itn += 1;
            beta = sqrt(unorm2);
            if (beta > 0) {
                su = 1.0 / beta;
                anorm = sqrt(anorm * anorm + alpha * alpha + beta * beta);
                alpha = sqrt(op_apply<STREAM, PSD, SM>(d, o, true, u, su, v, -beta * sv, 2, u_last, v_last));  // v_mem = A' u_true - beta v_true
                sv = alpha > 0 ? 1.0 / alpha : 1.0;
            } else {
                su = 1.0;  // u_mem is exactly zero
            }
            // plane rotations (damp = 0)
            const double rhobar1 = rhobar;
            rho = hypot(rhobar1, beta);
            const double cs = rhobar1 / rho, sn = beta / rho;
            const double theta = sn * alpha;
            rhobar = -cs * alpha;
            const double phi = cs * phibar;
            phibar = sn * phibar;
            tau = sn * phi;
            t1 = phi / rho;
            t2 = -theta / rho;
            const double delta = sn2 * rho, gambar = -cs2 * rho, rhs_ = phi - delta * z;
            const double zbar = rhs_ / gambar;
            xnorm = sqrt(xxnorm + zbar * zbar);
            const double gamma = hypot(gambar, theta);
            cs2 = gambar / gamma;
            sn2 = theta / gamma;
            z = rhs_ / gamma;
            xxnorm += z * z;
            rnorm = sqrt(phibar * phibar + res2);
            arnorm = alpha * fabs(tau);
            pending = true;
        }
    }
#ifdef LSQR_PROF
    if (d.gtid == 0) {
        const int n0 = g_prof_n > 400 ? 300 : 0;
        for (int k = n0; k + 1 < g_prof_n && k < n0 + 44; ++k)
            printf("%s %lld ns\n", (k & 1) ? "  compute" : "barrier", g_prof[k + 1] - g_prof[k]);
        g_prof_n = 0;
    }
#endif
    if (d.gtid == 0) {
        vec.stats[0] = (double)istop;
        vec.stats[1] = (double)itn;
        vec.stats[2] = rnorm;
        vec.stats[3] = arnorm;
        vec.stats[4] = anorm;
        vec.stats[5] = acond;
        vec.stats[6] = xnorm;
    }
}

template <int THREADS, int MINB, bool CLUSTER, bool STREAM, bool PSD>
__global__ void __launch_bounds__(THREADS, MINB) lsqr_kernel_t(OpArgs o, const double* __restrict__ rhs, LsqrParams prm,
                                                               LsqrVectors vec) {
    __shared__ double red[64];
    Dev d{cg::this_grid(), vec.partials, red, (int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5),
          (int)(blockDim.x >> 5), (int)gridDim.x, (int)(blockIdx.x * blockDim.x + threadIdx.x),
          (int)(gridDim.x * blockDim.x), CLUSTER};
    lsqr_body<STREAM, PSD, false>(d, o, rhs, prm, vec, -1.0);
}

// Lock-step batch: problem p = blockIdx.x / C is solved by the C CTAs of one cluster (C = 1: one CTA, the gather vectors
// staged in shared memory); every problem has its own operator, right-hand side and work vectors.  No barrier spans
// problems: a problem that converges early simply retires its CTAs.
template <bool SM>
__global__ void __launch_bounds__(CL_THREADS, 1) lsqr_batch_kernel(const OpArgs* __restrict__ ops, const double* const* __restrict__ rhs,
                                                                   LsqrParams prm, const LsqrVectors* __restrict__ vecs, const int C,
                                                                   const double zero_below) {
    __shared__ double red[64];
    __shared__ OpArgs so;
    __shared__ LsqrVectors sv;
    const int p = blockIdx.x / C, cta = blockIdx.x - p * C;
    for (int i = threadIdx.x; i < (int)(sizeof(OpArgs) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&so)[i] = reinterpret_cast<const int*>(ops + p)[i];
    for (int i = threadIdx.x; i < (int)(sizeof(LsqrVectors) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&sv)[i] = reinterpret_cast<const int*>(vecs + p)[i];
    __syncthreads();
    Dev d{cg::this_grid(), sv.partials, red, (int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5),
          (int)(blockDim.x >> 5), C, (int)(cta * blockDim.x + threadIdx.x), (int)(C * blockDim.x), true, cta, C == 1};
    lsqr_body<true, false, SM>(d, so, rhs[p], prm, sv, zero_below);
}

// stand-alone operator kernels (C-ABI conic_M_apply / conic_dpi_apply; also used by tests)
__global__ void __launch_bounds__(LSQR_THREADS) conic_M_kernel(ConicOpView op, int transpose, const double* t, double* out,
                                                              double* partials) {
    __shared__ double red[64];
    Dev d{cg::this_grid(), partials, red, (int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5),
          (int)(blockDim.x >> 5), (int)gridDim.x, (int)(blockIdx.x * blockDim.x + threadIdx.x),
          (int)(gridDim.x * blockDim.x)};
    double last = 0.0, norm2 = 0.0;
    conic_apply<false, true>(d, op, transpose != 0, t, 1.0, out, 0.0, 0, __ldcg(t + op.n + op.m), last, norm2);
}

__global__ void __launch_bounds__(LSQR_THREADS) conic_dpi_kernel(ConicOpView op, int transpose, const double* t, double* out,
                                                                double* partials) {
    __shared__ double red[64];
    Dev d{cg::this_grid(), partials, red, (int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5),
          (int)(blockDim.x >> 5), (int)gridDim.x, (int)(blockIdx.x * blockDim.x + threadIdx.x),
          (int)(gridDim.x * blockDim.x)};
    dpi_apply<true>(d, op, t, out, transpose != 0);
}


// ================================================================================================================
// Streaming LSQR for LARGE conic operators (working set beyond L2): the same iteration as lsqr_kernel, but every
// phase is its own kernel at full occupancy (the persistent kernel is capped at a few hundred CTAs by grid.sync and
// its register footprint, far too few loads in flight to cover HBM latency).  All Golub-Kahan scalars, the stop
// tests and the `done` flag live in a device struct: the host only enqueues iterations in batches and looks at the
// flag between batches; kernels of an iteration after convergence return at once.  Same arithmetic and the same
// deterministic reductions (per-CTA partials, fixed-order sums) as the persistent kernel.  No PSD cones here.
struct LsqrState {
    double alpha, beta, su, sv, rhobar, phibar, rho, t1, t2, tau;
    double anorm, acond, ddnorm, res2, xnorm, xxnorm, z, sn2, cs2, rnorm, arnorm, bnorm;
    double op_src, op_dst;  // scalars of the operator application in flight: dst = op_src * OP(src) + op_dst * dst
    long long itn;
    int istop, first, pending, done, more, do_op1;
};

constexpr int ST_THREADS = 256;
#define ST_DEV(partials)                                                                                          \
    __shared__ double red[64];                                                                                    \
    Dev d{cg::this_grid(), partials, red, (int)threadIdx.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5),     \
          (int)(blockDim.x >> 5), ST_STRIDE, (int)(blockIdx.x * blockDim.x + threadIdx.x),                        \
          (int)(gridDim.x * blockDim.x)}

__device__ double st_total(const double* partials, int slot, int stride, int count, double* red) {
    // fixed-order sum of `count` CTA partials by one 256-thread block
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partials[slot * stride + i];
    s = warp_sum_all(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    __syncthreads();
    return t;
}

// three totals in one pass (independent loads, one shared-memory exchange): the single-CTA scalar kernels are pure
// latency, three sequential st_total calls were most of it
__device__ void st_total3(const double* partials, int stride, int s0, int c0, int s1, int c1, int s2, int c2, double* red,
                          double& t0, double& t1, double& t2) {
    double a = 0.0, b = 0.0, c = 0.0;
    const int cmax = max(c0, max(c1, c2));
    for (int i = threadIdx.x; i < cmax; i += blockDim.x) {
        const double x0 = i < c0 ? partials[s0 * stride + i] : 0.0;
        const double x1 = i < c1 ? partials[s1 * stride + i] : 0.0;
        const double x2 = i < c2 ? partials[s2 * stride + i] : 0.0;
        a += x0;
        b += x1;
        c += x2;
    }
    a = warp_sum_all(a);
    b = warp_sum_all(b);
    c = warp_sum_all(c);
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[w] = a;
        red[8 + w] = b;
        red[16 + w] = c;
    }
    __syncthreads();
    t0 = t1 = t2 = 0.0;
    for (int k = 0; k < nw; ++k) {
        t0 += red[k];
        t1 += red[8 + k];
        t2 += red[16 + k];
    }
}

__global__ void __launch_bounds__(ST_THREADS) st_init_kernel(int nr, int nc, const double* __restrict__ rhs, LsqrVectors vec, LsqrState* S) {
    ST_DEV(vec.partials);
    double acc = 0.0;
    for (int i = d.gtid; i < nr; i += d.gthreads) {
        const double t = rhs[i];
        vec.u[i] = t;
        acc += t * t;
    }
    for (int i = d.gtid; i < nc; i += d.gthreads) {
        vec.x[i] = 0.0;
        vec.v[i] = 0.0;
        vec.w[i] = 0.0;
    }
    block_partial(d, 0, acc);
}

// after st_init_kernel: beta, su; sets up v = A' u (op1 with s_dst = 0)
__global__ void __launch_bounds__(ST_THREADS) st_init_fin_kernel(LsqrVectors vec, int stride, int cnt0, LsqrState* S) {
    __shared__ double red[64];
    const double t0 = st_total(vec.partials, 0, stride, cnt0, red);
    if (threadIdx.x == 0) {
        LsqrState s{};
        s.beta = sqrt(t0);
        s.su = 1.0;
        s.sv = 1.0;
        s.cs2 = -1.0;
        s.rho = 1.0;
        s.rnorm = s.beta;
        s.do_op1 = s.beta > 0 ? 1 : 0;
        if (s.beta > 0) s.su = 1.0 / s.beta;
        s.op_src = s.su;
        s.op_dst = 0.0;
        s.first = 1;
        *S = s;
    }
}

// x, w update of the previous iteration (LSQR's deferred vector update), ||w / rho||^2 into slot 4
__global__ void __launch_bounds__(ST_THREADS) st_update_kernel(int nc, LsqrVectors vec, LsqrState* S) {
    if (S->done) return;
    ST_DEV(vec.partials);
    const double sv = S->sv, t1 = S->t1, t2 = S->t2;
    double dd = 0.0;
    if (S->first) {
        for (int i = d.gtid; i < nc; i += d.gthreads) vec.w[i] = sv * vec.v[i];
    } else {
        const double irho = 1.0 / S->rho;
        for (int i = d.gtid; i < nc; i += d.gthreads) {
            const double wi = vec.w[i];
            const double dk = wi * irho;
            dd += dk * dk;
            vec.x[i] += t1 * wi;
            vec.w[i] = sv * vec.v[i] + t2 * wi;
        }
    }
    block_partial(d, 4, dd);
}

// ---- M src (non-transposed), phase 1: wc = Dpi src2
__global__ void __launch_bounds__(ST_THREADS) st_dpi_kernel(ConicOpView op, const double* __restrict__ y, double* out, int transpose,
                                                            const LsqrState* S, int which) {
    if (S->done || (which == 0 ? !S->more : !S->do_op1)) return;
    ST_DEV(nullptr);
    dpi_apply<false>(d, op, y, out, transpose != 0);
}

// ---- M src, phase 2: both sparse products, dst updated in place; slots 0 (norm) and 1 (last-row dot)
__global__ void __launch_bounds__(SP_THREADS, 8) st_M_rows_kernel(ConicOpView op, const double* __restrict__ src, double* dst, double* partials,
                                                               const LsqrState* S) {
    if (S->done || !S->more) return;
    ST_DEV(partials);
    const int n = op.n, m = op.m;
    const double s_src = S->op_src, s_dst = S->op_dst;
    const double t3 = __ldcg(src + n + m);
    double acc = 0.0, dotacc = 0.0;
    const int nbt = op.At.nblk;
    __shared__ double prod[ST_CHUNK];
    spmv_stream<SP_THREADS, ST_NPT, 0>(
        (int)blockIdx.x, (int)gridDim.x,
        nbt + op.A.nblk, prod, red,
        [&](int b, const CsrView*& A, const double*& x, int& lb) {
            if (b < nbt) { A = &op.At; x = op.wc; lb = b; }
            else { A = &op.A; x = src; lb = b - nbt; }
        },
        [&](int b, int row) {
            return b < nbt ? make_double4(op.c[row], dst[row], __ldcg(src + row), 0.0)
                           : make_double4(__ldcg(op.wc + row), __ldcg(src + n + row), op.b[row], dst[n + row]);
        },
        [&](int b, int row, double t, const double4& o) {
            if (b < nbt) {  // rows 0..n-1:  (A' wc)_j + c_j t3
                t = (t + o.x * t3) * s_src + s_dst * o.y;
                dst[row] = t;
                acc += t * t;
                dotacc += o.x * o.z;
            } else {        // rows n..n+m-1:  -(A t1)_i + t2_i - wc_i + b_i t3
                t = (-t + o.y - o.x + o.z * t3) * s_src + s_dst * o.w;
                dst[n + row] = t;
                acc += t * t;
                dotacc += o.z * o.x;
            }
        });
    block_partial(d, 0, acc);
    block_partial(d, 1, dotacc);
}

// ---- after M v: last row of u, pending stop tests, beta, scalars of the transposed application
__global__ void __launch_bounds__(ST_THREADS) st_mid_kernel(int N, LsqrVectors vec, int stride, int cnt_rows, int cnt_upd, LsqrParams prm,
                                                            LsqrState* S) {
    if (S->done) return;
    __shared__ double red[64];
    double tot0, tot1, tot4;
    st_total3(vec.partials, stride, 0, cnt_rows, 1, cnt_rows, 4, cnt_upd, red, tot0, tot1, tot4);
    if (threadIdx.x != 0) return;
    LsqrState s = *S;
    double unorm2 = tot0;
    if (s.more) {
        const double last = -tot1 * s.op_src + s.op_dst * vec.u[N - 1];
        vec.u[N - 1] = last;
        unorm2 += last * last;
    }
    if (s.pending) {
        s.ddnorm += tot4;
        s.acond = s.anorm * sqrt(s.ddnorm);
        const double test1 = s.rnorm / s.bnorm;
        const double test2 = s.arnorm / (s.anorm * s.rnorm);
        const double test3 = 1.0 / s.acond;
        const double t1_ = test1 / (1.0 + s.anorm * s.xnorm / s.bnorm);
        const double rtol = prm.btol + prm.atol * s.anorm * s.xnorm / s.bnorm;
        const double ctol = prm.conlim > 0 ? 1.0 / prm.conlim : 0.0;
        if (s.itn >= prm.maxiter) s.istop = 7;
        if (1.0 + test3 <= 1.0) s.istop = 6;
        if (1.0 + test2 <= 1.0) s.istop = 5;
        if (1.0 + t1_ <= 1.0) s.istop = 4;
        if (test3 <= ctol) s.istop = 3;
        if (test2 <= prm.atol) s.istop = 2;
        if (test1 <= rtol) s.istop = 1;
        s.pending = 0;
    }
    s.first = 0;
    if (s.istop > 0 || !s.more) {
        s.done = 1;
        *S = s;
        return;
    }
    s.itn += 1;
    s.beta = sqrt(unorm2);
    if (s.beta > 0) {
        s.su = 1.0 / s.beta;
        s.anorm = sqrt(s.anorm * s.anorm + s.alpha * s.alpha + s.beta * s.beta);
        s.op_src = s.su;
        s.op_dst = -s.beta * s.sv;
        s.do_op1 = 1;
    } else {
        s.su = 1.0;
        s.do_op1 = 0;
    }
    *S = s;
}

// ---- M' src, phase 1: wc = A src1 - src2 - b src3 ; slot 3 = b . src2
__global__ void __launch_bounds__(SP_THREADS, 8) st_Mt_A_kernel(ConicOpView op, const double* __restrict__ src, double* partials,
                                                             const LsqrState* S) {
    if (S->done || !S->do_op1) return;
    ST_DEV(partials);
    const int n = op.n, m = op.m;
    const double u3 = __ldcg(src + n + m);
    double dotacc = 0.0;
    __shared__ double prod[ST_CHUNK];
    spmv_stream<SP_THREADS, ST_NPT, 0>(
        (int)blockIdx.x, (int)gridDim.x,
        op.A.nblk, prod, red,
        [&](int b, const CsrView*& A, const double*& x, int& lb) { A = &op.A; x = src; lb = b; },
        [&](int, int row) { return make_double4(__ldcg(src + n + row), op.b[row], 0.0, 0.0); },
        [&](int, int row, double t, const double4& o) {
            op.wc[row] = t - o.x - o.y * u3;
            dotacc += o.y * o.x;
        });
    block_partial(d, 3, dotacc);
}

// ---- M' src, phase 3 (after Dpi' wc -> r2): rows of A' and the elementwise block; slots 2 (norm) and 5 (c . src1)
__global__ void __launch_bounds__(SP_THREADS, 8) st_Mt_rows_kernel(ConicOpView op, const double* __restrict__ src, double* dst, double* partials,
                                                                const LsqrState* S) {
    if (S->done || !S->do_op1) return;
    ST_DEV(partials);
    const int n = op.n, m = op.m;
    const double s_src = S->op_src, s_dst = S->op_dst;
    const double u3 = __ldcg(src + n + m);
    const double* r2 = op.wc + m;
    double acc = 0.0, dotacc = 0.0;
    __shared__ double prod[ST_CHUNK];
    spmv_stream<SP_THREADS, ST_NPT, 0>(
        (int)blockIdx.x, (int)gridDim.x,
        op.At.nblk, prod, red,
        [&](int b, const CsrView*& A, const double*& x, int& lb) { A = &op.At; x = src + n; lb = b; },
        [&](int, int row) { return make_double4(op.c[row], dst[row], __ldcg(src + row), 0.0); },
        [&](int, int row, double t, const double4& o) {
            t = (-t - o.x * u3) * s_src + s_dst * o.y;
            dst[row] = t;
            acc += t * t;
            dotacc += o.x * o.z;
        });
    for (int i = d.gtid; i < m; i += d.gthreads) {
        const double t = (__ldcg(r2 + i) + __ldcg(src + n + i)) * s_src + s_dst * dst[n + i];
        dst[n + i] = t;
        acc += t * t;
    }
    block_partial(d, 2, acc);
    block_partial(d, 5, dotacc);
}

// ---- after M' u: last row of v, alpha, plane rotations (or, at start-up, the initial scalars)
__global__ void __launch_bounds__(ST_THREADS) st_end_kernel(int N, LsqrVectors vec, int stride, int cnt_a, int cnt_rows, LsqrParams prm,
                                                            LsqrState* S, int startup) {
    if (S->done) return;
    __shared__ double red[64];
    double tot2, tot3, tot5;
    st_total3(vec.partials, stride, 2, cnt_rows, 3, cnt_a, 5, cnt_rows, red, tot2, tot3, tot5);
    if (threadIdx.x != 0) return;
    LsqrState s = *S;
    if (s.do_op1) {
        const double last = (tot3 + tot5) * s.op_src + s.op_dst * vec.v[N - 1];
        vec.v[N - 1] = last;
        s.alpha = sqrt(tot2 + last * last);
        if (startup) {
            if (s.alpha > 0) s.sv = 1.0 / s.alpha;
        } else {
            s.sv = s.alpha > 0 ? 1.0 / s.alpha : 1.0;
        }
    }
    if (startup) {
        s.arnorm = s.alpha * s.beta;
        s.bnorm = s.beta;
        s.rhobar = s.alpha;
        s.phibar = s.beta;
        s.more = s.itn < prm.maxiter ? 1 : 0;
        s.op_src = s.sv;
        s.op_dst = -s.alpha * s.su;
        if (s.arnorm == 0.0) s.done = 1;
        *S = s;
        return;
    }
    const double rhobar1 = s.rhobar;
    s.rho = hypot(rhobar1, s.beta);
    const double cs = rhobar1 / s.rho, sn = s.beta / s.rho;
    const double theta = sn * s.alpha;
    s.rhobar = -cs * s.alpha;
    const double phi = cs * s.phibar;
    s.phibar = sn * s.phibar;
    s.tau = sn * phi;
    s.t1 = phi / s.rho;
    s.t2 = -theta / s.rho;
    const double delta = s.sn2 * s.rho, gambar = -s.cs2 * s.rho, rhs_ = phi - delta * s.z;
    const double zbar = rhs_ / gambar;
    s.xnorm = sqrt(s.xxnorm + zbar * zbar);
    const double gamma = hypot(gambar, theta);
    s.cs2 = gambar / gamma;
    s.sn2 = theta / gamma;
    s.z = rhs_ / gamma;
    s.xxnorm += s.z * s.z;
    s.rnorm = sqrt(s.phibar * s.phibar + s.res2);
    s.arnorm = s.alpha * fabs(s.tau);
    s.pending = 1;
    s.more = s.itn < prm.maxiter ? 1 : 0;
    s.op_src = s.sv;
    s.op_dst = -s.alpha * s.su;
    *S = s;
}

__global__ void st_stats_kernel(const LsqrState* S, double* stats) {
    stats[0] = (double)S->istop;
    stats[1] = (double)S->itn;
    stats[2] = S->rnorm;
    stats[3] = S->arnorm;
    stats[4] = S->anorm;
    stats[5] = S->acond;
    stats[6] = S->xnorm;
}

CsrView view_of(const DevBuf& rp, const DevBuf& ci, const DevBuf& v, int64_t nrows, int64_t ncols,
                const DevBuf* blk = nullptr, int64_t nblk = 0, const DevBuf* sval = nullptr, const DevBuf* scol = nullptr,
                const DevBuf* spos = nullptr) {
    return CsrView{(int)nrows, (int)ncols, rp.as<int>(), ci.as<int>(), v.as<double>(),
                   blk ? blk->as<int>() : nullptr, (int)nblk,
                   sval ? sval->as<double>() : nullptr, scol ? scol->as<int>() : nullptr,
                   spos ? spos->as<unsigned short>() : nullptr};
}

}  // namespace

ConicOpView conic_view(diffopt_b200_ctx* ctx, bool stream_blocks = false) {
    ConicState& s = ctx->conic;
    ConicOpView o{};
    o.n = (int)s.n;
    o.m = (int)s.m;
    if (stream_blocks) {
        const bool so = s.A.sorted;
        o.A = view_of(s.A.rowptr, s.A.colind, s.A.val, s.m, s.n, &s.A.blk, s.A.nblk, so ? &s.A.sval : nullptr,
                      so ? &s.A.scol : nullptr, so ? &s.A.spos : nullptr);
        o.At = view_of(s.A.t_rowptr, s.A.t_colind, s.A.t_val, s.n, s.m, &s.A.t_blk, s.A.t_nblk, so ? &s.A.t_sval : nullptr,
                       so ? &s.A.t_scol : nullptr, so ? &s.A.t_spos : nullptr);
    } else {
        o.A = view_of(s.A.rowptr, s.A.colind, s.A.val, s.m, s.n, &s.A.cblk, s.A.ncblk);
        o.At = view_of(s.A.t_rowptr, s.A.t_colind, s.A.t_val, s.n, s.m, &s.A.t_cblk, s.A.t_ncblk);
    }
    o.b = s.b.as<double>();
    o.c = s.c.as<double>();
    o.diag = s.nn_scale.as<double>();
    o.kind = s.row_kind.as<signed char>();
    o.nsoc = (int)s.nsoc;
    o.soc_off = s.soc_off.as<int>();
    o.soc_dim = s.soc_dim.as<int>();
    o.soc_case = s.soc_case.as<int>();
    o.soc_nx = s.soc_nx.as<double>();
    o.v = s.v.as<double>();
    o.npsd = (int)s.npsd;
    o.psd_off = s.psd_off.as<int>();
    o.psd_d = s.psd_d.as<int>();
    o.psd_uoff = s.psd_uoff.as<long long>();
    o.psd_U = s.psd_U.as<double>();
    o.psd_Bm = s.psd_Bm.as<double>();
    o.psd_ident = s.psd_ident.as<int>();
    o.psd_toff = s.psd_toff.as<int>();
    o.psd_ntiles = (int)s.psd_ntiles;
    o.psd_w0 = s.psd_work.as<double>();
    o.psd_w1 = o.psd_w0 + s.psd_sumd2;
    o.psd_w2 = o.psd_w1 + s.psd_sumd2;
    o.wc = s.w1.as<double>();
    return o;
}

static int coop_grid(diffopt_b200_ctx* ctx, const void* kernel, int64_t work) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, LSQR_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    int64_t want = (work + LSQR_THREADS * 4 - 1) / (LSQR_THREADS * 4);
    // small problems: one CTA per SM keeps grid.sync cheap; large ones (beyond L2) need every resident CTA to cover HBM latency
    int64_t cap = (int64_t)ctx->sm_count * (work > (int64_t)4 << 20 ? per_sm : 1);
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

// Cluster-synchronised variant for operators that a single cluster can chew through (a few 10^5 nonzeros, everything
// L2 resident): ONE cluster of 16 (or 8) CTAs x 1024 threads, every phase boundary is a hardware cluster barrier.
constexpr int64_t CL_MAX_WORK = (int64_t)1 << 20;
constexpr int64_t CL_CLUSTER_WORK = (int64_t)512 << 10;

static int cluster_size_for(diffopt_b200_ctx* ctx, const void* kernel) {
    static int cached = -1;  // per process; every B200 in a box is the same part
    if (cached >= 0) return cached;
    cached = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
    }
    for (int cs : {16, 8}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)cs);
        cfg.blockDim = dim3(CL_THREADS);
        cfg.dynamicSmemBytes = CL_CHUNK * sizeof(double);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n >= 1) {
            cached = cs;
            break;
        }
        cudaGetLastError();
    }
    (void)ctx;
    return cached;
}

static int32_t lsqr_launch(diffopt_b200_ctx* ctx, OpArgs& o, int64_t work, const double* rhs_dev, LsqrParams prm,
                           double* x_dev, double* stats_host7) {
    LsqrWork& wk = ctx->lsqr;
    const size_t d = sizeof(double);
    DO_CUDA(ctx, wk.u.reserve(d * (size_t)o.nrows));
    DO_CUDA(ctx, wk.v.reserve(d * (size_t)o.ncols));
    DO_CUDA(ctx, wk.w.reserve(d * (size_t)o.ncols));
    const void* k_grid = (const void*)lsqr_kernel_t<LSQR_THREADS, 4, false, false, false>;
    const void* k_cluster = (const void*)lsqr_kernel_t<CL_THREADS, 1, true, true, false>;
    const void* k_gstream = (const void*)lsqr_kernel_t<CL_THREADS, 1, false, true, false>;
    // Three persistent variants, by size of the operator (`work` ~ nonzeros + vector lengths):
    //   cluster : <= CL_CLUSTER_WORK  one 16-CTA cluster, hardware barriers, row-block SpMV (16 SMs are enough)
    //   gstream : <= CL_MAX_WORK      one CTA per SM, grid.sync(), row-block SpMV (L1 gather rate of 16 SMs is not)
    //   grid    : larger / PSD cones  lane-group SpMV, up to 4 CTAs per SM
    const char* mode = getenv("DIFFOPT_B200_LSQR");
    const bool psd = o.kind == 1 && o.conic.npsd > 0;  // PSD applies are GEMM phases written for the lane-group grid
    int variant = 0;  // 0 grid, 1 cluster, 2 gstream
    if (!psd && work <= CL_MAX_WORK) variant = work <= CL_CLUSTER_WORK ? 1 : 2;
    if (mode && strcmp(mode, "grid") == 0) variant = 0;
    if (mode && strcmp(mode, "cluster") == 0 && !psd) variant = 1;
    if (mode && strcmp(mode, "gstream") == 0 && !psd) variant = 2;
    int cs = variant == 1 ? cluster_size_for(ctx, k_cluster) : 0;
    if (variant == 1 && cs == 0) variant = 2;
    const size_t dyn = variant ? CL_CHUNK * sizeof(double) : 0;
    // problems with PSD cones: same kernel compiled for 2 CTAs per SM (128 registers: the GEMM phases of the PSD apply
    // spill badly under the 64-register cap of the 4-CTA build)
    const void* k_psd = (const void*)lsqr_kernel_t<LSQR_THREADS, 2, false, false, true>;
    const void* k_lane = psd ? k_psd : k_grid;
    int grid = cs;
    if (variant == 0) grid = coop_grid(ctx, k_lane, work);
    if (variant == 2) {
        DO_CUDA(ctx, cudaFuncSetAttribute(k_gstream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        grid = ctx->sm_count;
    }
    {   // the row-block lists that match the launch
        auto pick = [&](CsrView& v, const CsrDev& M, bool t) {
            const bool g = variant == 2;
            v.blk = (t ? (g ? M.t_gblk : M.t_cblk) : (g ? M.gblk : M.cblk)).as<int>();
            v.nblk = (int)(t ? (g ? M.t_ngblk : M.t_ncblk) : (g ? M.ngblk : M.ncblk));
        };
        if (o.kind == 1) {
            pick(o.conic.A, ctx->conic.A, false);
            pick(o.conic.At, ctx->conic.A, true);
        } else {
            const bool fwd_is_t = o.fwd.rowptr == ctx->lsqr_mat.t_rowptr.as<int>();
            pick(o.fwd, ctx->lsqr_mat, fwd_is_t);
            pick(o.adj, ctx->lsqr_mat, !fwd_is_t);
        }
    }
    DO_CUDA(ctx, wk.tmp.reserve(d * (size_t)NSLOTS * (size_t)grid));
    DO_CUDA(ctx, wk.scal.reserve(d * 8));
    DO_CUDA(ctx, cudaMemsetAsync(wk.tmp.ptr, 0, d * (size_t)NSLOTS * (size_t)grid, ctx->stream));
    LsqrVectors vec{wk.u.as<double>(), wk.v.as<double>(), wk.w.as<double>(), x_dev, wk.tmp.as<double>(),
                    wk.scal.as<double>()};
    void* args[] = {&o, (void*)&rhs_dev, &prm, &vec};
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (variant == 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)cs);
        cfg.blockDim = dim3(CL_THREADS);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        DO_CUDA(ctx, cudaLaunchKernelExC(&cfg, k_cluster, args));
    } else if (variant == 2) {
        DO_CUDA(ctx, cudaLaunchCooperativeKernel(k_gstream, dim3(grid), dim3(CL_THREADS), args, dyn, ctx->stream));
    } else {
        DO_CUDA(ctx, cudaLaunchCooperativeKernel(k_lane, dim3(grid), dim3(LSQR_THREADS), args, 0, ctx->stream));
    }
    ctx->launches++;
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    double st[7];
    DO_CUDA(ctx, cudaMemcpyAsync(st, wk.scal.ptr, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    if (stats_host7) memcpy(stats_host7, st, sizeof st);
    return 0;
}

int32_t lsqr_run_csr(diffopt_b200_ctx* ctx, const CsrDev& M, bool trans, const double* rhs_dev, LsqrParams prm,
                     double* x_dev, double* stats_host7) {
    OpArgs o{};
    o.kind = 0;
    CsrView a = view_of(M.rowptr, M.colind, M.val, M.nrows, M.ncols, &M.cblk, M.ncblk);
    CsrView at = view_of(M.t_rowptr, M.t_colind, M.t_val, M.ncols, M.nrows, &M.t_cblk, M.t_ncblk);
    o.fwd = trans ? at : a;
    o.adj = trans ? a : at;
    o.nrows = (int)(trans ? M.ncols : M.nrows);
    o.ncols = (int)(trans ? M.nrows : M.ncols);
    return lsqr_launch(ctx, o, M.nnz + M.nrows + M.ncols, rhs_dev, prm, x_dev, stats_host7);
}

static int32_t lsqr_stream_conic(diffopt_b200_ctx* ctx, const double* rhs_dev, LsqrParams prm, double* x_dev, double* stats_host7) {
    ConicState& cs = ctx->conic;
    LsqrWork& wk = ctx->lsqr;
    ConicOpView op = conic_view(ctx, true);
    const int n = (int)cs.n, m = (int)cs.m, N = n + m + 1;
    const size_t dsz = sizeof(double);
    DO_CUDA(ctx, wk.u.reserve(dsz * (size_t)N));
    DO_CUDA(ctx, wk.v.reserve(dsz * (size_t)N));
    DO_CUDA(ctx, wk.w.reserve(dsz * (size_t)N));
    const int stride = ST_STRIDE;
    DO_CUDA(ctx, wk.tmp.reserve(dsz * (size_t)NSLOTS * (size_t)stride));
    DO_CUDA(ctx, wk.scal.reserve(dsz * 8 + sizeof(LsqrState) + 64));
    double* stats = wk.scal.as<double>();
    LsqrState* S = reinterpret_cast<LsqrState*>(stats + 8);
    LsqrVectors vec{wk.u.as<double>(), wk.v.as<double>(), wk.w.as<double>(), x_dev, wk.tmp.as<double>(), stats};
    auto blocks_for = [&](int64_t items) {
        int64_t b = (items + ST_THREADS - 1) / ST_THREADS;
        if (b > stride) b = stride;
        if (b < 1) b = 1;
        return (unsigned)b;
    };
    const unsigned g_vec = blocks_for(N), g_dpi = blocks_for(std::max<int64_t>(m, (int64_t)cs.nsoc * 8));
    // SpMV kernels: persistent over the row blocks, every resident CTA slot filled exactly once (8 CTAs per SM)
    auto blocks_spmv = [&](int64_t nblk) {
        int per_sm = 8;
        if (const char* e = getenv("DIFFOPT_B200_SPMV_CTAS")) per_sm = std::max(1, atoi(e));
        int64_t b = std::min<int64_t>(nblk, (int64_t)ctx->sm_count * per_sm);
        return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, stride));
    };
    const unsigned g_rows = blocks_spmv(cs.A.nblk + cs.A.t_nblk), g_a = blocks_spmv(cs.A.nblk);
    const unsigned g_trows = blocks_spmv(std::max<int64_t>(cs.A.t_nblk, ((int64_t)m + SP_THREADS * 8 - 1) / (SP_THREADS * 8)));
    cudaStream_t st = ctx->stream;
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    st_init_kernel<<<g_vec, ST_THREADS, 0, st>>>(N, N, rhs_dev, vec, S);
    st_init_fin_kernel<<<1, ST_THREADS, 0, st>>>(vec, stride, (int)g_vec, S);
    auto apply_Mt = [&](int startup) {  // v = su M' u + op_dst v
        st_Mt_A_kernel<<<g_a, SP_THREADS, 0, st>>>(op, vec.u, vec.partials, S);
        st_dpi_kernel<<<g_dpi, ST_THREADS, 0, st>>>(op, op.wc, op.wc + m, 1, S, 1);
        st_Mt_rows_kernel<<<g_trows, SP_THREADS, 0, st>>>(op, vec.u, vec.v, vec.partials, S);
        st_end_kernel<<<1, ST_THREADS, 0, st>>>(N, vec, stride, (int)g_a, (int)g_trows, prm, S, startup);
        ctx->launches += 4;
    };
    apply_Mt(1);
    ctx->launches += 2;
    LsqrState hs{};
    const long long batch = 16;
    for (long long it = 0; it <= prm.maxiter; it += batch) {
        for (long long k = 0; k < batch; ++k) {
            st_update_kernel<<<g_vec, ST_THREADS, 0, st>>>(N, vec, S);
            st_dpi_kernel<<<g_dpi, ST_THREADS, 0, st>>>(op, vec.v + n, op.wc, 0, S, 0);
            st_M_rows_kernel<<<g_rows, SP_THREADS, 0, st>>>(op, vec.v, vec.u, vec.partials, S);
            st_mid_kernel<<<1, ST_THREADS, 0, st>>>(N, vec, stride, (int)g_rows, (int)g_vec, prm, S);
            ctx->launches += 4;
            apply_Mt(0);
        }
        DO_CUDA(ctx, cudaMemcpyAsync(&hs, S, sizeof hs, cudaMemcpyDeviceToHost, st));
        DO_CUDA(ctx, cudaStreamSynchronize(st));
        if (hs.done) break;
    }
    st_stats_kernel<<<1, 1, 0, st>>>(S, stats);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    double stv[7];
    DO_CUDA(ctx, cudaMemcpyAsync(stv, stats, sizeof stv, cudaMemcpyDeviceToHost, st));
    DO_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
    if (stats_host7) memcpy(stats_host7, stv, sizeof stv);
    return 0;
}

int32_t lsqr_run_conic(diffopt_b200_ctx* ctx, const double* rhs_dev, LsqrParams prm, double* x_dev,
                       double* stats_host7) {
    ConicState& s = ctx->conic;
    OpArgs o{};
    o.kind = 1;
    o.conic = conic_view(ctx);
    o.conic_trans = 0;
    o.nrows = o.ncols = (int)(s.n + s.m + 1);
    int64_t work = s.A.nnz * 2 + s.n + 2 * s.m + s.psd_sumd2 * 8;
    // large operators without PSD cones: one kernel per phase at full occupancy (see lsqr_stream_conic)
    const char* mode = getenv("DIFFOPT_B200_LSQR");
    const bool want_stream = mode ? strcmp(mode, "stream") == 0 : work > ((int64_t)4 << 20);
    if (want_stream && s.npsd == 0 && !(mode && strcmp(mode, "persistent") == 0))
        return lsqr_stream_conic(ctx, rhs_dev, prm, x_dev, stats_host7);
    return lsqr_launch(ctx, o, work, rhs_dev, prm, x_dev, stats_host7);
}

static int32_t conic_small_launch(diffopt_b200_ctx* ctx, const void* kernel, const double* t_dev, bool transpose,
                                  double* out_dev) {
    ConicState& s = ctx->conic;
    ConicOpView op = conic_view(ctx);
    int64_t work = s.A.nnz * 2 + s.n + 2 * s.m + s.psd_sumd2 * 8;
    int grid = coop_grid(ctx, kernel, work);
    DO_CUDA(ctx, ctx->lsqr.tmp.reserve(sizeof(double) * (size_t)NSLOTS * (size_t)grid));
    double* partials = ctx->lsqr.tmp.as<double>();
    int tr = transpose ? 1 : 0;
    void* args[] = {&op, &tr, (void*)&t_dev, &out_dev, &partials};
    DO_CUDA(ctx, cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(LSQR_THREADS), args, 0, ctx->stream));
    ctx->launches++;
    return 0;
}

int32_t conic_apply_M(diffopt_b200_ctx* ctx, const double* t_dev, bool transpose, double* out_dev) {
    return conic_small_launch(ctx, (const void*)conic_M_kernel, t_dev, transpose, out_dev);
}
int32_t conic_apply_dpi(diffopt_b200_ctx* ctx, const double* t_dev, bool transpose, double* out_dev) {
    return conic_small_launch(ctx, (const void*)conic_dpi_kernel, t_dev, transpose, out_dev);
}

// Lock-step batch of conic problems (diffopt_b200_conic_batch_reverse): one persistent kernel, problem p on the C CTAs of
// cluster p.  `work` holds one region of `stride` doubles per problem: [rhs (N+1) | x (N) | u (N) | v (N) | w (N) |
// partials (NSLOTS * C) | stats (8)].  With C == 1 and `stage` the gather vectors are staged in shared memory.
int32_t lsqr_run_conic_batch(diffopt_b200_ctx* ctx, std::vector<ConicState>& states, int C, bool stage, double* work, size_t stride,
                             LsqrParams prm, double zero_below, DevBuf& ops_buf, DevBuf& vecs_buf, DevBuf& rhs_buf) {
    const int B = (int)states.size();
    if (B == 0) return 0;
    const int n = (int)states[0].n, m = (int)states[0].m, N = n + m + 1;
    std::vector<OpArgs> ops((size_t)B);
    std::vector<LsqrVectors> vecs((size_t)B);
    std::vector<const double*> rhs((size_t)B);
    for (int p = 0; p < B; ++p) {
        std::swap(ctx->conic, states[(size_t)p]);
        OpArgs o{};
        o.kind = 1;
        o.conic = conic_view(ctx);
        o.conic_trans = 0;
        o.nrows = o.ncols = N;
        ops[(size_t)p] = o;
        std::swap(ctx->conic, states[(size_t)p]);
        double* base = work + (size_t)p * stride;
        rhs[(size_t)p] = base;
        vecs[(size_t)p] = LsqrVectors{base + (N + 1) + N, base + (N + 1) + 2 * (size_t)N, base + (N + 1) + 3 * (size_t)N, base + (N + 1),
                                      base + (N + 1) + 4 * (size_t)N, base + (N + 1) + 4 * (size_t)N + (size_t)NSLOTS * C};
    }
    DO_CUDA(ctx, ops_buf.reserve(sizeof(OpArgs) * (size_t)B));
    DO_CUDA(ctx, vecs_buf.reserve(sizeof(LsqrVectors) * (size_t)B));
    DO_CUDA(ctx, rhs_buf.reserve(sizeof(double*) * (size_t)B));
    DO_CUDA(ctx, cudaMemcpyAsync(ops_buf.ptr, ops.data(), sizeof(OpArgs) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(vecs_buf.ptr, vecs.data(), sizeof(LsqrVectors) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaMemcpyAsync(rhs_buf.ptr, rhs.data(), sizeof(double*) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die at return
    stage = stage && C == 1;
    const size_t dyn = sizeof(double) * ((size_t)CL_CHUNK + (stage ? (size_t)(n + m) : 0));
    if (dyn > ctx->smem_optin) stage = false;
    const size_t dyn2 = sizeof(double) * ((size_t)CL_CHUNK + (stage ? (size_t)(n + m) : 0));
    const void* kern = stage ? (const void*)lsqr_batch_kernel<true> : (const void*)lsqr_batch_kernel<false>;
    DO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2));
    const OpArgs* dops = ops_buf.as<OpArgs>();
    const double* const* drhs = rhs_buf.as<const double*>();
    const LsqrVectors* dvecs = vecs_buf.as<LsqrVectors>();
    void* args[] = {(void*)&dops, (void*)&drhs, &prm, (void*)&dvecs, &C, &zero_below};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * C));
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = dyn2;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    DO_CUDA(ctx, cudaLaunchKernelExC(&cfg, kern, args));
    ctx->launches++;
    DO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return 0;
}

// Host: Julia CSC (1-based int64) -> device CSR of the matrix and of its transpose (0-based int32).
int32_t csr_from_csc_host(diffopt_b200_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* colptr,
                          const int64_t* rowval, const double* nzval, CsrDev& out) {
    if (nrows < 0 || ncols < 0 || !colptr) BAD_ARG(ctx, "csc: bad dimensions");
    const int64_t nnz = colptr[ncols] - colptr[0];
    if (colptr[0] != 1) BAD_ARG(ctx, "csc: colptr must be 1-based (Julia SparseMatrixCSC)");
    if (nnz < 0 || nnz > 2000000000LL || nrows > 2000000000LL || ncols > 2000000000LL)
        BAD_ARG(ctx, "csc: too large for int32 indices");
    // transpose view = CSC arrays reinterpreted: rows of A' are the columns of A
    std::vector<int> t_rowptr((size_t)ncols + 1), t_colind((size_t)nnz);
    for (int64_t j = 0; j <= ncols; ++j) t_rowptr[(size_t)j] = (int)(colptr[j] - 1);
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t r = rowval[k] - 1;
        if (r < 0 || r >= nrows) BAD_ARG(ctx, "csc: row index out of range");
        t_colind[(size_t)k] = (int)r;
    }
    // CSR of A by counting sort over rows
    std::vector<int> rowptr((size_t)nrows + 1, 0), colind((size_t)nnz);
    std::vector<double> val((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) rowptr[(size_t)t_colind[(size_t)k] + 1]++;
    for (int64_t i = 0; i < nrows; ++i) rowptr[(size_t)i + 1] += rowptr[(size_t)i];
    std::vector<int> next(rowptr.begin(), rowptr.end() - 1);
    for (int64_t j = 0; j < ncols; ++j)
        for (int k = t_rowptr[(size_t)j]; k < t_rowptr[(size_t)j + 1]; ++k) {
            int r = t_colind[(size_t)k];
            int dst = next[(size_t)r]++;
            colind[(size_t)dst] = (int)j;
            val[(size_t)dst] = nzval[k];
        }
    // row blocks of the streaming SpMV: <= ST_CHUNK nonzeros and <= SP_THREADS rows each, or one longer row
    auto row_blocks = [](const std::vector<int>& rp, int64_t nr, std::vector<int>& blk) {
        blk.clear();
        blk.push_back(0);
        int64_t r = 0;
        while (r < nr) {
            const int s0 = rp[(size_t)r];
            int64_t e = r;
            while (e < nr && rp[(size_t)e + 1] - s0 <= ST_CHUNK && e - r < SP_THREADS) ++e;
            if (e == r) e = r + 1;
            blk.push_back((int)e);
            r = e;
        }
        // interleave (first row, first nonzero) per block: one 8-byte load per block end for the issuing thread
        std::vector<int> d2(blk.size() * 2);
        for (size_t k = 0; k < blk.size(); ++k) {
            d2[2 * k] = blk[k];
            d2[2 * k + 1] = rp[(size_t)blk[k]];
        }
        blk.swap(d2);
    };
    // the same for the cluster-synchronised kernel: about one block per CTA and round, <= CL_CHUNK nonzeros each
    auto cluster_blocks = [](const std::vector<int>& rp, int64_t nr, int64_t CTAS, std::vector<int>& blk) {
        const int64_t nz = nr > 0 ? rp[(size_t)nr] : 0;
        const int64_t rounds = std::max<int64_t>(1, (nz + CTAS * CL_CHUNK - 1) / (CTAS * CL_CHUNK));
        const int64_t rows_round = std::max<int64_t>(1, (nr + CTAS * CL_THREADS - 1) / (CTAS * CL_THREADS));
        const int64_t want = CTAS * std::max(rounds, rows_round);
        double target = (double)std::max<int64_t>(nz / want + 1, 1);
        for (int attempt = 0; attempt < 24; ++attempt) {
            const int cap = (int)std::min<double>(target, (double)CL_CHUNK);
            blk.clear();
            blk.push_back(0);
            int64_t r = 0;
            while (r < nr) {
                const int s0 = rp[(size_t)r];
                int64_t e = r;
                while (e < nr && rp[(size_t)e + 1] - s0 <= cap && e - r < CL_THREADS) ++e;
                if (e == r) e = r + 1;
                blk.push_back((int)e);
                r = e;
            }
            if ((int64_t)blk.size() - 1 <= want || cap >= CL_CHUNK) break;
            target *= 1.03;
        }
        std::vector<int> d2(blk.size() * 2);
        for (size_t k = 0; k < blk.size(); ++k) {
            d2[2 * k] = blk[k];
            d2[2 * k + 1] = rp[(size_t)blk[k]];
        }
        blk.swap(d2);
    };
    std::vector<int> blk, t_blk, cblk, t_cblk, gblk, t_gblk;
    row_blocks(rowptr, nrows, blk);
    row_blocks(t_rowptr, ncols, t_blk);
    const int64_t cl_ctas = ctx->csr_cluster_ctas > 0 ? ctx->csr_cluster_ctas : CL_MAX_CTAS;
    cluster_blocks(rowptr, nrows, cl_ctas, cblk);
    cluster_blocks(t_rowptr, ncols, cl_ctas, t_cblk);
    cluster_blocks(rowptr, nrows, ctx->sm_count, gblk);
    cluster_blocks(t_rowptr, ncols, ctx->sm_count, t_gblk);
    out.ncblk = (int64_t)cblk.size() / 2 - 1;
    out.t_ncblk = (int64_t)t_cblk.size() / 2 - 1;
    out.ngblk = (int64_t)gblk.size() / 2 - 1;
    out.t_ngblk = (int64_t)t_gblk.size() / 2 - 1;
    out.nblk = (int64_t)blk.size() / 2 - 1;
    out.t_nblk = (int64_t)t_blk.size() / 2 - 1;
    // Large operators (the streaming driver): a second copy of (val, colind) with the nonzeros of every row block
    // ordered by column, plus each nonzero's position inside its block.  The gathers of a warp then fall into few
    // 128-byte lines when the operator has any locality (the L1TEX takes ~one line request per clock per SM, which is
    // what bounds the row-block SpMV), the products are scattered back to their CSR positions in shared memory, and the
    // row sums are unchanged bit for bit.
    std::vector<double> sval, t_sval;
    std::vector<int> scol, t_scol;
    std::vector<unsigned short> spos, t_spos;
    const char* sort_env = getenv("DIFFOPT_B200_SPMV_SORT");
    const bool want_sort = sort_env ? atoi(sort_env) != 0 : nnz >= ((int64_t)1 << 21);
    auto sort_blocks = [](const std::vector<int>& blk2, const std::vector<int>& ci, const double* va,
                          std::vector<double>& sv, std::vector<int>& sc, std::vector<unsigned short>& sp) {
        const size_t nb = blk2.size() / 2 - 1;
        const size_t nz = ci.size();
        sv.resize(nz); sc.resize(nz); sp.resize(nz);
        auto work = [&](size_t b0, size_t b1) {
            std::vector<int> idx;
            for (size_t b = b0; b < b1; ++b) {
                const int s0 = blk2[2 * b + 1], s1 = blk2[2 * b + 3], cnt = s1 - s0;
                if (cnt > ST_CHUNK) {  // one long row: streamed in place
                    for (int k = s0; k < s1; ++k) { sv[(size_t)k] = va[k]; sc[(size_t)k] = ci[(size_t)k]; sp[(size_t)k] = 0; }
                    continue;
                }
                idx.resize((size_t)cnt);
                for (int j = 0; j < cnt; ++j) idx[(size_t)j] = j;
                std::stable_sort(idx.begin(), idx.end(), [&](int a, int c) { return ci[(size_t)(s0 + a)] < ci[(size_t)(s0 + c)]; });
                for (int j = 0; j < cnt; ++j) {
                    const int k = s0 + idx[(size_t)j];
                    sv[(size_t)(s0 + j)] = va[k];
                    sc[(size_t)(s0 + j)] = ci[(size_t)k];
                    sp[(size_t)(s0 + j)] = (unsigned short)idx[(size_t)j];
                }
            }
        };
        const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; ++t) pool.emplace_back(work, nb * t / hw, nb * (t + 1) / hw);
        for (auto& th : pool) th.join();
    };
    if (want_sort && nnz > 0) {
        sort_blocks(blk, colind, val.data(), sval, scol, spos);
        sort_blocks(t_blk, t_colind, nzval, t_sval, t_scol, t_spos);
    }
    out.sorted = want_sort && nnz > 0;
    out.nrows = nrows;
    out.ncols = ncols;
    out.nnz = nnz;
    auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.reserve(bytes ? bytes : 8);
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpyAsync(b.ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
    };
    DO_CUDA(ctx, up(out.rowptr, rowptr.data(), sizeof(int) * rowptr.size()));
    DO_CUDA(ctx, up(out.colind, colind.data(), sizeof(int) * colind.size()));
    DO_CUDA(ctx, up(out.val, val.data(), sizeof(double) * val.size()));
    DO_CUDA(ctx, up(out.t_rowptr, t_rowptr.data(), sizeof(int) * t_rowptr.size()));
    DO_CUDA(ctx, up(out.t_colind, t_colind.data(), sizeof(int) * t_colind.size()));
    DO_CUDA(ctx, up(out.t_val, nzval, sizeof(double) * (size_t)nnz));
    if (out.sorted) {
        DO_CUDA(ctx, up(out.sval, sval.data(), sizeof(double) * sval.size()));
        DO_CUDA(ctx, up(out.scol, scol.data(), sizeof(int) * scol.size()));
        DO_CUDA(ctx, up(out.spos, spos.data(), sizeof(unsigned short) * spos.size()));
        DO_CUDA(ctx, up(out.t_sval, t_sval.data(), sizeof(double) * t_sval.size()));
        DO_CUDA(ctx, up(out.t_scol, t_scol.data(), sizeof(int) * t_scol.size()));
        DO_CUDA(ctx, up(out.t_spos, t_spos.data(), sizeof(unsigned short) * t_spos.size()));
    }
    DO_CUDA(ctx, up(out.blk, blk.data(), sizeof(int) * blk.size()));
    DO_CUDA(ctx, up(out.t_blk, t_blk.data(), sizeof(int) * t_blk.size()));
    DO_CUDA(ctx, up(out.cblk, cblk.data(), sizeof(int) * cblk.size()));
    DO_CUDA(ctx, up(out.t_cblk, t_cblk.data(), sizeof(int) * t_cblk.size()));
    DO_CUDA(ctx, up(out.gblk, gblk.data(), sizeof(int) * gblk.size()));
    DO_CUDA(ctx, up(out.t_gblk, t_gblk.data(), sizeof(int) * t_gblk.size()));
    DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die at return
    return 0;
}

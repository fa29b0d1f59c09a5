// Tuned batched KKT sensitivity kernel for the headline shape n=64, m=64, p=16 (KKT order N=144).
//
// Persistent CTAs (256 threads, two per SM when the reduced system is small enough) stream over the QP
// instances.  Per instance:
//   0. column-singleton elimination -- what the reference's sparse `\` (UMFPACK) does in its preprocessing:
//      an inequality with lam_i == 0 exactly makes column n+i of LHS = [Q G'diag(lam) A'; G diag(Gz-h) 0; A 0 0]
//      a singleton (only D_i = (Gz-h)_i is nonzero), so that unknown decouples exactly:
//        LHS  x = r_b :  x_lam_i = (r_b[n+i] - G_i x_z) / D_i   (computed after the reduced solve)
//        LHS' x = r_f :  x_lam_i = r_f[n+i] / D_i = 0            (r_f[n+i] = lam_i * (...) = 0)
//      The reduced system keeps z, the ACTIVE inequalities (lam_i != 0) and the equalities: order
//      n' = 64 + m_active + 16 (96 for the OptNet-style benchmark), padded to a multiple of 8 with identity.
//   1. assemble the reduced LHS (QuadraticProgram.jl:256-282) straight from HBM into shared memory as an
//      nt x nt grid of 8 x 8 tiles (row-major inside a tile = DMMA C/A fragment order); build the forward RHS
//      (:429-433) and the reverse RHS (:324-329);
//   2. right-looking blocked LU with partial pivoting, panel width 8:
//        - warp 0 holds the 8-column panel in REGISTERS and does the 8 pivot steps with warp collectives
//          (redux.max on the high words for the pivot search, pivot row broadcast through shared memory);
//          LAPACK-style row interchanges are tracked per row in registers and applied once per panel; the
//          forward RHS rides along as an extra, never-pivoted row (its multipliers are w = r_f U^-1);
//        - interchanges + U12 = L11^-1 A12 by one thread per column (the reverse RHS is column n');
//        - trailing update C -= L21 U12 on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA), 8 warps in a
//          4 x 2 cyclic tile decomposition, A/B fragments reused from registers;
//   3. blocked backward substitutions (explicit inverses of the 8 x 8 diagonal tiles) finish
//      LHS x_b = r_b and LHS' x_f = r_f from the one factorisation (:335 uses LHS, :438 uses LHS').
// The KKT matrix and its factors never touch HBM: algorithmic traffic is inputs + outputs only.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int NV = 64, MI = 64, PE = 16, N = 144, NTMAX = 18;
constexpr int THREADS = 256;
constexpr int NWARP = THREADS / 32;
constexpr unsigned FULL = 0xffffffffu;

// fixed-size part of the shared memory; the tile matrix K[nt*nt*64] follows it
struct __align__(16) SmemHdr {
    double Ub[NTMAX * 64];     // U row block of the current panel, tile J column-major: (k,n) at n*8+k
    double invU[NTMAX * 64];   // inverse of the upper-triangular diagonal tiles, row-major (also scratch)
    double invL[NTMAX * 64];   // inverse of the unit-lower diagonal tiles, row-major (also scratch)
    double y[N + 8];           // reverse RHS -> L^-1 P r_b -> x_b        (reduced ordering)
    double rf[N + 8];          // forward RHS -> w = U^-T r_f -> v        (reduced ordering)
    double rdiag[N + 8];       // 1 / U_kk
    double zs[NV], lams[MI], nus[PE];
    double dvec[MI];           // D = G z - h for every inequality
    double part[4 * 64];       // reduction scratch
    double prow[2][12];        // pivot-row broadcast staging (double buffered): 8 values, rinv, position
    double wblk[8];
    int perm[N + 8];           // position -> original (reduced) row
    int apos[MI];              // inequality -> slot among the active ones, or -1
    int alist[MI];             // active slot -> inequality
    int ipiv[8], ipiv_off[8];  // LAPACK-style interchanges of the current panel: row c0+k <-> ipiv[k]
    int info, ma, nt;
};

struct Ctx {
    SmemHdr& S;
    double* K;
    int nt;
    __device__ __forceinline__ int tile_off(int I, int J) const { return (J * nt + I) << 6; }
    __device__ __forceinline__ int row_off(int r) const { return ((r >> 3) << 6) + ((r & 7) << 3); }
    __device__ __forceinline__ int col_off(int c) const { return (((c >> 3) * nt) << 6) + (c & 7); }
    __device__ __forceinline__ int elem_off(int r, int c) const { return row_off(r) + col_off(c); }
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// explicit shared-window accesses for the panel's staging area (keeps generic->shared conversions out of the chain)
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_b32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_b32(unsigned addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// 1/x to ~1 ulp: hardware estimate + two Newton steps (no special-case slow path; x is a checked pivot)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// 8 per-lane values -> sums over the 32 lanes; on return lane L holds the total of value index (L >> 2) & 7
__device__ __forceinline__ double warp_reduce8(const double (&v)[8], int lane) {
    double a[4], b[2], c;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double keep = h16 ? v[i + 4] : v[i], send = h16 ? v[i] : v[i + 4];
        a[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double keep = h8 ? a[i + 2] : a[i], send = h8 ? a[i] : a[i + 2];
        b[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    {
        double keep = h4 ? b[1] : b[0], send = h4 ? b[0] : b[1];
        c = keep + __shfl_xor_sync(FULL, send, 4);
    }
    c += __shfl_xor_sync(FULL, c, 2);
    c += __shfl_xor_sync(FULL, c, 1);
    return c;
}

// ---- panel factorisation by one warp, panel held in registers -----------------------------------------
// NS = number of 32-row slots the panel still has.  np = padded order of the reduced system.
template <int NS>
__device__ __noinline__ void panel_factor(SmemHdr* Sp, double* Kp, const int nt, const int j, const int lane, const int np) {
    SmemHdr& S = *Sp;
    const Ctx X{S, Kp, nt};
    const int c0 = j << 3;
    const unsigned stage_base = (unsigned)__cvta_generic_to_shared(&S.prow[0][0]);
    double p[NS][8];
    double e[8];
    int dstpos[NS];
    unsigned live = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int r = c0 + lane + 32 * s;
        dstpos[s] = r;
        if (r < np) {
            live |= 1u << s;
            const double2* src = reinterpret_cast<const double2*>(&X.K[X.tile_off(r >> 3, j) + ((r & 7) << 3)]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double2 v = src[q];
                p[s][2 * q] = v.x;
                p[s][2 * q + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) p[s][c] = 0.0;
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) e[c] = S.rf[c0 + c];
    const unsigned valid = live;
    bool singular = false;
    double myrinv = 0.0;  // lane k keeps 1/pivot_k
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // pivot search: max |a| on the high 32 bits (relative precision 2^-17: a pivot within 1e-5 of the max)
        unsigned best = 0;
        double bv = 1.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            unsigned hi = (unsigned)__double2hiint(p[s][k]) & 0x7ffffff8u;
            unsigned key = ((live >> s) & 1u) ? (0x80000000u | hi | (unsigned)s) : 0u;
            if (key > best) {
                best = key;
                bv = p[s][k];
            }
        }
        const unsigned kmax = __reduce_max_sync(FULL, best);
        const double rloc = fast_rcp(bv);  // speculative: overlaps the reduction latency
        const int owner = __ffs(__ballot_sync(FULL, best == kmax)) - 1;
        const int sp = (int)(kmax & 7u);
        const bool zero_piv = (kmax & 0x7ffffff8u) == 0u;  // exactly zero (or denormal) pivot column
        singular |= zero_piv;
        const unsigned stage = stage_base + (k & 1) * 96;
        if (lane == owner) {
            // the owner publishes its pivot row, 1/pivot and the row's current position
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (s == sp) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c >= k) sts_f64(stage + 8 * c, p[s][c]);
                    sts_f64(stage + 64, zero_piv ? 0.0 : rloc);
                    sts_b32(stage + 72, dstpos[s]);
                    dstpos[s] = -1;  // marks "this lane's pivot row": fixed up below
                }
            }
            live &= ~(1u << sp);
        }
        __syncwarp();
        double pv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c >= k) pv[c] = lds_f64(stage + 8 * c);
        const double rinv = lds_f64(stage + 64);
        const int P = lds_b32(stage + 72);
        if (lane == k) {
            myrinv = rinv;
            S.ipiv[k] = P;
            S.ipiv_off[k] = X.row_off(P);
        }
        // LAPACK interchange k: the row sitting at position c0+k goes to the pivot row's position P
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int d = dstpos[s];
            d = (d == c0 + k) ? P : d;
            d = (d == -1) ? c0 + k : d;
            dstpos[s] = d;
        }
        // multipliers and rank-1 update of the live rows
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if ((live >> s) & 1u) {
                const double l = p[s][k] * rinv;
                p[s][k] = l;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c > k) p[s][c] = fma(-l, pv[c], p[s][c]);
            }
        }
        {
            const double l = e[k] * rinv;
            e[k] = l;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c > k) e[c] = fma(-l, pv[c], e[c]);
        }
    }
    // write the panel back in LAPACK layout (rows at their final positions; multipliers stored NEGATED)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const bool ok = (valid >> s) & 1u;
        const int pos = dstpos[s];
        if (ok) {
            const int kp = pos - c0;  // < 8: this is pivot row kp (entries c >= kp are U), else all multipliers
            double2* dst = reinterpret_cast<double2*>(&X.K[X.tile_off(pos >> 3, j) + ((pos & 7) << 3)]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double a = (2 * q < kp) ? -p[s][2 * q] : p[s][2 * q];
                double b = (2 * q + 1 < kp) ? -p[s][2 * q + 1] : p[s][2 * q + 1];
                dst[q] = make_double2(a, b);
            }
        }
    }
    if (lane < 8) {
        double w = e[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) w = (lane == c) ? e[c] : w;
        S.rf[c0 + lane] = w;
        S.wblk[lane] = w;
        S.rdiag[c0 + lane] = myrinv;
    }
    if (lane == 0 && singular && S.info == 0) S.info = c0 + 1;  // first panel with an exactly zero pivot
}

__device__ __forceinline__ void panel_dispatch(const Ctx& X, int j, int lane, int np) {
    const int M = np - (j << 3);
    if (M > 128) panel_factor<5>(&X.S, X.K, X.nt, j, lane, np);
    else if (M > 96) panel_factor<4>(&X.S, X.K, X.nt, j, lane, np);
    else if (M > 64) panel_factor<3>(&X.S, X.K, X.nt, j, lane, np);
    else if (M > 32) panel_factor<2>(&X.S, X.K, X.nt, j, lane, np);
    else panel_factor<1>(&X.S, X.K, X.nt, j, lane, np);
}

// list != nullptr: solve only the instances list[0 .. *count) (the ones the LDL' fast path rejected)
__global__ void __launch_bounds__(THREADS, 2) qp_kkt_n144_kernel(QpSolveArgs a, const int* __restrict__ list,
                                                                  const int* __restrict__ count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemHdr& S = *reinterpret_cast<SmemHdr*>(smem_raw);
    double* const K = reinterpret_cast<double*>(smem_raw + sizeof(SmemHdr));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool do_fwd = a.fwd != nullptr, do_rev = a.rev != nullptr;
#ifdef QP_PROFILE
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#define PROF(i)                                \
    do {                                       \
        long long _n = clock64();              \
        pc[i] += _n - tprev;                   \
        tprev = _n;                            \
    } while (0)
#else
#define PROF(i)
#endif

    const int64_t nwork = list ? (int64_t)*count : a.B;
    for (int64_t work = blockIdx.x; work < nwork; work += gridDim.x) {
        const int64_t inst = list ? (int64_t)list[work] : work;
        const size_t bm = (a.shared & 1) ? 0 : (size_t)inst, bd = (a.shared & 2) ? 0 : (size_t)inst;
        const double* Q = a.Q + bm * NV * NV;
        const double* G = a.G + bm * MI * NV;
        const double* A = a.A + bm * PE * NV;
        // ---- vectors; active set (column-singleton detection)
        if (tid < NV) S.zs[tid] = a.z[(size_t)inst * NV + tid];
        else if (tid < NV + MI) S.lams[tid - NV] = a.lam[(size_t)inst * MI + tid - NV];
        else if (tid < NV + MI + PE) S.nus[tid - NV - MI] = a.nu[(size_t)inst * PE + tid - NV - MI];
        if (warp == 7) {
            const double l0 = a.lam[(size_t)inst * MI + lane], l1 = a.lam[(size_t)inst * MI + 32 + lane];
            const unsigned m0 = __ballot_sync(FULL, l0 != 0.0), m1 = __ballot_sync(FULL, l1 != 0.0);
            const unsigned lt = (1u << lane) - 1u;
            const int p0 = __popc(m0 & lt), p1 = __popc(m0) + __popc(m1 & lt);
            S.apos[lane] = (l0 != 0.0) ? p0 : -1;
            S.apos[32 + lane] = (l1 != 0.0) ? p1 : -1;
            if (l0 != 0.0) S.alist[p0] = lane;
            if (l1 != 0.0) S.alist[p1] = 32 + lane;
            if (lane == 0) {
                const int ma = __popc(m0) + __popc(m1);
                S.ma = ma;
                S.nt = (NV + ma + PE + 7) >> 3;
                S.info = 0;
            }
        }
        {   // L2 prefetch of the next instance of this CTA (inputs are streamed once from HBM)
            const int64_t nxt = list ? a.B : inst + gridDim.x;
            if (nxt < a.B) {
                const size_t nm = (a.shared & 1) ? 0 : (size_t)nxt, nd = (a.shared & 2) ? 0 : (size_t)nxt;
                const char* bases[6] = {(const char*)(a.Q + nm * NV * NV), (const char*)(a.G + nm * MI * NV),
                                        (const char*)(a.A + nm * PE * NV),
                                        do_fwd && a.dQ ? (const char*)(a.dQ + nd * NV * NV) : nullptr,
                                        do_fwd && a.dG ? (const char*)(a.dG + nd * MI * NV) : nullptr,
                                        do_fwd && a.dA ? (const char*)(a.dA + nd * PE * NV) : nullptr};
                const int lines[6] = {256, 256, 64, 256, 256, 64};  // 128-byte lines
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (bases[q] && tid < lines[q])
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(bases[q] + (size_t)tid * 128));
            }
        }
        __syncthreads();
        const int ma = S.ma, nt = S.nt;
        const int nred = NV + ma + PE, np = nt << 3;
        const Ctx X{S, K, nt};
        // zero the whole reduced matrix, then fill
        for (int i = tid; i < nt * nt * 32; i += THREADS) reinterpret_cast<double2*>(K)[i] = make_double2(0.0, 0.0);
        if (tid < np) {
            S.y[tid] = (do_rev && tid < NV) ? a.seed[(size_t)inst * NV + tid] : 0.0;
            S.perm[tid] = tid;
            S.rf[tid] = 0.0;
        }
        // loads (coalesced column-major reads; thread keeps a fixed row r and 16 columns cg + 4 i)
        const int r = tid & 63, cg = tid >> 6;  // 4 column groups
        {
            double qv[16], gv[16], av[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = cg + 4 * i;
                qv[i] = __ldg(Q + c * NV + r);
                gv[i] = __ldg(G + c * MI + r);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = __ldg(A + tid + THREADS * i);
            __syncthreads();
            const double lam_r = S.lams[r];
            const int ar = S.apos[r];
            double dacc = 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = cg + 4 * i;
                K[X.elem_off(r, c)] = qv[i];
                if (ar >= 0) {
                    K[X.elem_off(NV + ar, c)] = gv[i];
                    K[X.elem_off(c, NV + ar)] = gv[i] * lam_r;
                }
                dacc = fma(gv[i], S.zs[c], dacc);
            }
            S.part[cg * 64 + r] = dacc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + THREADS * i, ii = idx & 15, c = idx >> 4;
                K[X.elem_off(NV + ma + ii, c)] = av[i];
                K[X.elem_off(c, NV + ma + ii)] = av[i];
            }
            if (tid >= nred && tid < np) K[X.elem_off(tid, tid)] = 1.0;  // identity padding
        }
        __syncthreads();
        if (tid < MI) {
            const double d = S.part[tid] + S.part[64 + tid] + S.part[128 + tid] + S.part[192 + tid] -
                             a.h[(size_t)inst * MI + tid];
            S.dvec[tid] = d;
            const int ar = S.apos[tid];
            if (ar >= 0) K[X.elem_off(NV + ar, NV + ar)] = d;
            else if (d == 0.0) atomicCAS(&S.info, 0, NV + tid + 1);  // lam_i == 0 and D_i == 0: singular column
        }
        __syncthreads();
        // ---- forward RHS (QuadraticProgram.jl:429-433): [dQ z + dq + dG'lam + dA'nu ; lam.(dG z - dh) ; dA z - db]
        if (do_fwd) {
            const size_t b = (size_t)inst;
            double rq = 0.0, rg = 0.0;
            double cv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) cv[i] = 0.0;
            if (a.dQ) {
                const double* Xp = a.dQ + bd * NV * NV;
#pragma unroll
                for (int i = 0; i < 16; ++i) rq = fma(__ldg(Xp + (cg + 4 * i) * NV + r), S.zs[cg + 4 * i], rq);
            }
            if (a.dG) {
                const double* Xp = a.dG + bd * MI * NV;
                const double lr = S.lams[r];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const double v = __ldg(Xp + (cg + 4 * i) * MI + r);
                    rg = fma(v, S.zs[cg + 4 * i], rg);
                    cv[i] = v * lr;
                }
            }
            S.part[cg * 64 + r] = rq;
            {   // column sums of dG .* lam over this warp's 32 rows: two 8-value transpose-reductions
                double c8[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) c8[i] = cv[8 * h + i];
                    const double csum = warp_reduce8(c8, lane);  // lane L: value (L>>2)&7 -> column cg+4*(8h+idx)
                    if ((lane & 3) == 0) S.invU[(warp & 1) * 64 + cg + 4 * (8 * h + ((lane >> 2) & 7))] = csum;
                }
            }
            __syncthreads();
            double r1 = 0.0;
            if (tid < NV)
                r1 = S.part[tid] + S.part[64 + tid] + S.part[128 + tid] + S.part[192 + tid] + S.invU[tid] + S.invU[64 + tid];
            __syncthreads();
            S.part[cg * 64 + r] = rg;
            // dA: 16 x 64, thread element idx = tid + 256 q: row ii = tid & 15, column (tid >> 4) + 16 q
            double ra = 0.0, ca[4] = {0.0, 0.0, 0.0, 0.0};
            if (a.dA) {
                const double* Xp = a.dA + bd * PE * NV;
                const int ii = tid & 15, c = tid >> 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double v = __ldg(Xp + tid + THREADS * q);
                    ra = fma(v, S.zs[c + 16 * q], ra);
                    ca[q] = v * S.nus[ii];
                }
            }
            ra += __shfl_xor_sync(FULL, ra, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) ca[q] += __shfl_xor_sync(FULL, ca[q], o);
            }
            if ((lane & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) S.invL[(tid >> 4) + 16 * q] = ca[q];  // column sums of dA .* nu
            }
            if (lane < 16) S.invU[128 + warp * 16 + lane] = ra;  // [warp][ii]
            __syncthreads();
            if (tid < NV) {
                double v = r1 + S.invL[tid];
                if (a.dq) v += a.dq[b * NV + tid];
                S.rf[tid] = v;
            } else if (tid < NV + MI) {
                const int i = tid - NV, ar = S.apos[i];
                if (ar >= 0) {
                    double v = S.part[i] + S.part[64 + i] + S.part[128 + i] + S.part[192 + i];
                    if (a.dh) v -= a.dh[b * MI + i];
                    S.rf[NV + ar] = S.lams[i] * v;
                }
            } else if (tid < N) {
                const int i = tid - NV - MI;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) v += S.invU[128 + w * 16 + i];
                if (a.db) v -= a.db[b * PE + i];
                S.rf[NV + ma + i] = v;
            }
        }
        __syncthreads();
        PROF(0);

        // ---- blocked LU
        for (int j = 0; j < nt; ++j) {
            const int c0 = j << 3;
            if (warp == 0) panel_dispatch(X, j, lane, np);
            __syncthreads();
            PROF(1);
            // (b) row interchanges (LAPACK dlaswp order) on every other column, then (c) U12 = L11^-1 A12;
            //     column np = the reverse RHS y
            {
                const double* Ld = &K[X.tile_off(j, j)];  // diagonal tile: negated multipliers below the diagonal
                const int c = tid;
                if (c <= np && (c < c0 || c >= c0 + 8)) {
                    double x[8];
                    if (c < np) {
                        double* colb = &K[X.col_off(c)];
                        const int r0 = X.row_off(c0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int po = S.ipiv_off[k], ko = r0 + 8 * k;  // rows c0..c0+7 share a tile: +8 per row
                            if (po != ko) {
                                const double t0 = colb[ko], t1 = colb[po];
                                colb[ko] = t1;
                                colb[po] = t0;
                            }
                        }
                        if (c >= c0 + 8) {
                            double* colp = &K[X.tile_off(j, c >> 3) + (c & 7)];
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i] = colp[i * 8];
#pragma unroll
                            for (int i = 1; i < 8; ++i) {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if (k < i) x[i] = fma(Ld[i * 8 + k], x[k], x[i]);
                                colp[i * 8] = x[i];
                            }
                            double2* ub = reinterpret_cast<double2*>(&S.Ub[((c >> 3) << 6) + ((c & 7) << 3)]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) ub[q] = make_double2(x[2 * q], x[2 * q + 1]);
                        }
                    } else if (do_rev) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int pr = S.ipiv[k];
                            if (pr != c0 + k) {
                                const double t0 = S.y[c0 + k], t1 = S.y[pr];
                                S.y[c0 + k] = t1;
                                S.y[pr] = t0;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = S.y[c0 + i];
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k < i) x[i] = fma(Ld[i * 8 + k], x[k], x[i]);
                            S.y[c0 + i] = x[i];
                        }
                    }
                } else if (tid == THREADS - 1) {  // permutation bookkeeping
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int pr = S.ipiv[k];
                        const int t0 = S.perm[c0 + k], t1 = S.perm[pr];
                        S.perm[c0 + k] = t1;
                        S.perm[pr] = t0;
                    }
                }
            }
            __syncthreads();
            PROF(2);
            // (d) trailing update on the DMMA pipe: C(I,J) += Lneg(I,j) * U(j,J)
            if (j < nt - 1) {
                const int wr = warp & 3, wc = warp >> 2;
                double2 af[5];
#pragma unroll
                for (int ii = 0; ii < 5; ++ii) {
                    const int I = j + 1 + wr + 4 * ii;
                    if (I < nt) af[ii] = *reinterpret_cast<const double2*>(&K[X.tile_off(I, j) + g * 8 + 2 * t]);
                }
                for (int J = j + 1 + wc; J < nt; J += 2) {
                    const double2 bf = *reinterpret_cast<const double2*>(&S.Ub[(J << 6) + g * 8 + 2 * t]);
                    double2 cc[5];
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < nt) cc[ii] = *reinterpret_cast<const double2*>(&K[X.tile_off(I, J) + g * 8 + 2 * t]);
                    }
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < nt) dmma(cc[ii].x, cc[ii].y, af[ii].x, bf.x);
                    }
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < nt) {
                            dmma(cc[ii].x, cc[ii].y, af[ii].y, bf.y);
                            *reinterpret_cast<double2*>(&K[X.tile_off(I, J) + g * 8 + 2 * t]) = cc[ii];
                        }
                    }
                }
                if (warp == 7 && do_rev) {  // y[r] += Lneg[r, panel] * y_blk
                    double yb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) yb[k] = S.y[c0 + k];
#pragma unroll
                    for (int s = 0; s < 5; ++s) {
                        const int rr = c0 + 8 + lane + 32 * s;
                        if (rr < np) {
                            const double2* lrow = reinterpret_cast<const double2*>(&K[X.tile_off(rr >> 3, j) + ((rr & 7) << 3)]);
                            double v = S.y[rr];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const double2 l2 = lrow[q];
                                v = fma(l2.x, yb[2 * q], v);
                                v = fma(l2.y, yb[2 * q + 1], v);
                            }
                            S.y[rr] = v;
                        }
                    }
                } else if (warp == 6 && do_fwd) {  // rf[c] -= w_blk * U[panel rows, c]
                    double wb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) wb[k] = S.wblk[k];
#pragma unroll
                    for (int s = 0; s < 5; ++s) {
                        const int c = c0 + 8 + lane + 32 * s;
                        if (c < np) {
                            const double2* ucol = reinterpret_cast<const double2*>(&S.Ub[((c >> 3) << 6) + ((c & 7) << 3)]);
                            double v = S.rf[c];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const double2 u2 = ucol[q];
                                v = fma(-wb[2 * q], u2.x, v);
                                v = fma(-wb[2 * q + 1], u2.y, v);
                            }
                            S.rf[c] = v;
                        }
                    }
                }
            }
            __syncthreads();
            PROF(3);
        }

        // ---- inverses of all diagonal tiles' triangles, in parallel (for the blocked backward substitutions)
        for (int item = tid; item < 2 * nt * 8; item += THREADS) {
            const int which = item / (nt * 8), rem = item % (nt * 8), jt = rem >> 3, c = rem & 7;
            const double* Ld = &K[X.tile_off(jt, jt)];
            double T[8][8];
            double x[8];
            if (which == 0) {  // column c of inv(U11)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k > i) T[i][k] = Ld[i * 8 + k];
#pragma unroll
                for (int i = 7; i >= 0; --i) {
                    double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k > i) v = fma(-T[i][k], x[k], v);
                    x[i] = v * S.rdiag[(jt << 3) + i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) S.invU[(jt << 6) + i * 8 + c] = x[i];
            } else {  // column c of inv(L11), unit lower, NEGATED strict part stored
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < i) T[i][k] = Ld[i * 8 + k];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < i) v = fma(T[i][k], x[k], v);
                    x[i] = v;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) S.invL[(jt << 6) + i * 8 + c] = x[i];
            }
        }
        __syncthreads();
        // ---- blocked backward substitutions: warp 0: U x = y ; warp 1: L' v = w
        if (warp == 0 && do_rev) {
            for (int j = nt - 1; j >= 0; --j) {
                const int c0 = j << 3;
                double xk = 0.0;
                if (lane < 8) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) xk = fma(S.invU[(j << 6) + lane * 8 + c], S.y[c0 + c], xk);
                }
                __syncwarp();
                if (lane < 8) S.y[c0 + lane] = xk;
                double x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = __shfl_sync(FULL, xk, k);
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const int rr = lane + 32 * s;
                    if (rr < c0) {
                        const double2* urow = reinterpret_cast<const double2*>(&K[X.tile_off(rr >> 3, j) + ((rr & 7) << 3)]);
                        double v = S.y[rr];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double2 u2 = urow[q];
                            v = fma(-u2.x, x[2 * q], v);
                            v = fma(-u2.y, x[2 * q + 1], v);
                        }
                        S.y[rr] = v;
                    }
                }
                __syncwarp();
            }
        } else if (warp == 1 && do_fwd) {
            for (int j = nt - 1; j >= 0; --j) {
                const int c0 = j << 3;
                double vk = 0.0;  // v_blk = inv(L11)' w_blk
                if (lane < 8) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) vk = fma(S.invL[(j << 6) + c * 8 + lane], S.rf[c0 + c], vk);
                }
                __syncwarp();
                if (lane < 8) S.rf[c0 + lane] = vk;
                double v8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v8[k] = __shfl_sync(FULL, vk, k);
#pragma unroll
                for (int s = 0; s < 5; ++s) {  // w[c] += sum_k Lneg[c0+k, c] v_k   for columns c < c0
                    const int c = lane + 32 * s;
                    if (c < c0) {
                        const double* lt = &K[X.tile_off(j, c >> 3) + (c & 7)];
                        double v = S.rf[c];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v = fma(lt[k * 8], v8[k], v);
                        S.rf[c] = v;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        PROF(4);
        // ---- outputs: (dz, dlam, dnu) = -x ;  x_f = P' v ; eliminated (inactive) inequalities recovered
        {
            double* rev = do_rev ? a.rev + (size_t)inst * N : nullptr;
            double* fwd = do_fwd ? a.fwd + (size_t)inst * N : nullptr;
            if (do_rev) {
                // inactive rows: out_lam_i = -(G_i . out_z) / D_i   (G re-read: L2 resident)
                double acc = 0.0;
                if (S.apos[r] < 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = cg + 4 * i;
                        acc = fma(__ldg(G + c * MI + r), -S.y[c], acc);
                    }
                }
                S.part[cg * 64 + r] = acc;
            }
            if (do_fwd && tid < MI && S.apos[tid] < 0) fwd[NV + tid] = 0.0;  // x_lam_i = r_f[n+i]/D_i = 0
            if (do_fwd && tid < np) {
                const int o = S.perm[tid];  // reduced index of the unknown stored at position tid
                const int full = o < NV ? o : (o < NV + ma ? NV + S.alist[o - NV] : (o < nred ? NV + MI + (o - NV - ma) : -1));
                if (full >= 0) fwd[full] = -S.rf[tid];
            }
            __syncthreads();
            if (do_rev) {
                if (tid < NV) rev[tid] = -S.y[tid];
                else if (tid < NV + MI) {
                    const int i = tid - NV, ar = S.apos[i];
                    if (ar >= 0) rev[tid] = -S.y[NV + ar];
                    else rev[tid] = -(S.part[i] + S.part[64 + i] + S.part[128 + i] + S.part[192 + i]) / S.dvec[i];
                } else if (tid < N) rev[tid] = -S.y[NV + ma + (tid - NV - MI)];
            }
            if (a.info && tid == 0) a.info[inst] = S.info;
            if (tid == 0 && S.info != 0) qp_report_sticky(a, inst);
        }
        __syncthreads();
        PROF(5);
    }
#ifdef QP_PROFILE
    if (a.prof && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) a.prof[i] = pc[i];
#endif
}

// max number of active inequalities (lam != 0) over the batch -> shared-memory size / CTAs per SM
__global__ void max_active_kernel(int64_t B, const double* __restrict__ lam, int* out) {
    int best = 0;
    for (int64_t b = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); b < B;
         b += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int lane = threadIdx.x & 31;
        int c = (lam[b * MI + lane] != 0.0) + (lam[b * MI + 32 + lane] != 0.0);
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        best = max(best, c);
    }
    if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(out, best);
}

}  // namespace

int32_t qp_sqd_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, bool* handled, int* max_active, int max_tag);
// shape-generic LDL' fast path (qp_batch_sqd_any.cu)
int32_t qp_sqd_any_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, bool* handled, int* max_active, int max_tag);
bool qp_sqd_any_supported(diffopt_b200_ctx* ctx, const QpSolveArgs& a);
int qp_sqd_any_nt(const QpSolveArgs& a, int active);
int qp_sqd_any_nt_limit(diffopt_b200_ctx* ctx, const QpSolveArgs& a);
int32_t qp_max_active_any_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int* dmax);

static int32_t lu_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, const int* list, const int* count,
                         bool* handled) {
    *handled = false;
    const size_t smem = sizeof(SmemHdr) + (size_t)nt_cap * nt_cap * 64 * sizeof(double);
    if (smem > ctx->smem_optin) return 0;  // falls back to the generic kernel
    *handled = true;
    int per_sm = 1;
    DO_CUDA(ctx, kernel_config((const void*)qp_kkt_n144_kernel, ctx->device, THREADS, smem, &per_sm));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    QpSolveArgs aa = a;
    const bool profile = !list && getenv("DIFFOPT_B200_PROFILE") != nullptr;
    long long* dprof = nullptr;
    if (profile) {
        DO_CUDA(ctx, cudaMalloc(&dprof, 8 * sizeof(long long)));
        DO_CUDA(ctx, cudaMemsetAsync(dprof, 0, 8 * sizeof(long long), ctx->stream));
        aa.prof = dprof;
    } else {
        aa.prof = nullptr;
    }
    qp_kkt_n144_kernel<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(aa, list, count);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    if (profile) {
        long long h[8];
        DO_CUDA(ctx, cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dprof);
        long long ninst = (a.B + grid - 1) / grid;
        fprintf(stderr,
                "[qp_n144 profile, CTA 0, %lld instances, %d CTA/SM, nt_cap %d, smem %zu] clocks/instance: assemble %lld "
                "panel %lld swap+trsm %lld update %lld backward %lld output %lld\n",
                ninst, per_sm, nt_cap, smem, h[0] / ninst, h[1] / ninst, h[2] / ninst, h[3] / ninst, h[4] / ninst,
                h[5] / ninst);
    }
    return 0;
}

// the pivoted LU kernel over a device-side list of instances (fallback of the LDL' fast path)
int32_t qp_lu_launch_list(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, const int* list, const int* count) {
    bool handled = false;
    (void)nt_cap;  // rejected instances may be larger than the fast path's configuration: always worst-case sized
    int32_t rc = lu_launch(ctx, a, NTMAX, list, count, &handled);
    if (rc == 0 && !handled) {
        ctx->err = "qp_batch: pivoted-LU fallback does not fit in shared memory";
        return -3;
    }
    return rc;
}

// DIFFOPT_B200_QP_KERNEL = generic | lu | (default) ldl : which kernel serves the batch.  The headline shape n=64, m=64,
// p=16 has its own LDL' and pivoted-LU kernels; every other shape whose worst case fits shared memory runs the
// shape-generic LDL' kernel with the generic pivoted-LU kernel behind it.
int32_t qp_batch_launch_tuned(diffopt_b200_ctx* ctx, const QpSolveArgs& a, bool* handled) {
    *handled = false;
    const char* force = getenv("DIFFOPT_B200_QP_KERNEL");
    // the tuned kernels stream dense directions; `ldl_any` runs the shape-generic pair on the headline shape too (measurement)
    const bool headline = a.n == NV && a.m == MI && a.p == PE && !a.rhs_pre && !(force && strcmp(force, "ldl_any") == 0);
    if (force && strcmp(force, "generic") == 0) return 0;
    if (!headline) {
        if (force && strcmp(force, "lu") == 0) return 0;
        if (!qp_sqd_any_supported(ctx, a)) return 0;
    }
    // the active-set hint below belongs to one shape: a new shape starts over (outstanding reports are drained first)
    const int64_t shape_key = ((int64_t)a.n << 40) | ((int64_t)a.m << 20) | (int64_t)a.p;
    if (shape_key != ctx->qp_hint_shape) {
        if (ctx->qp_hint_shape != -1) DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t& c : ctx->qp_hmax_call) c = -1;
        ctx->qp_hint = -1;
        ctx->qp_hint_shape = shape_key;
    }
    // The shared-memory configuration depends on the largest active set of the batch (reduced order 80 + active).
    // First call: measure it and wait for the answer.  Later calls: launch for the size seen by the PREVIOUS call
    // (no host round trip in the middle of the call); instances that do not fit that guess are handed to the
    // pivoted-LU kernel, and the size measured now configures the next call.
    DO_CUDA(ctx, ctx->qp_max.reserve(sizeof(int)));
    if (!ctx->qp_hmax_host) {
        DO_CUDA(ctx, cudaHostAlloc((void**)&ctx->qp_hmax_host, 4 * sizeof(int), cudaHostAllocDefault));
        for (cudaEvent_t& ev : ctx->qp_hmax_ev) DO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    int* dmax = ctx->qp_max.as<int>();
    const bool first = ctx->qp_hint < 0;
    // Hint for this launch: the newest report whose D2H copy has COMPLETED (its event says so).  Stream-ordered calls
    // (qp_batch_solve_async) return before their copies land, so a slot still in flight is left alone and the
    // previous hint stays in force -- the hint only sizes the launch, results never depend on it.
    if (!first) {
        int64_t newest = -1;
        for (int s = 0; s < 4; ++s) {
            if (ctx->qp_hmax_call[s] < 0 || cudaEventQuery(ctx->qp_hmax_ev[s]) != cudaSuccess) continue;
            if (ctx->qp_hmax_call[s] > newest) {
                newest = ctx->qp_hmax_call[s];
                ctx->qp_hint = ctx->qp_hmax_host[s] & 0xFF;  // low byte: active-set size, high bits: call tag
            }
            ctx->qp_hmax_call[s] = -1;
        }
    }
    const int64_t call = ctx->qp_calls++;
    // D2H copy of the word into a free ring slot + its event.  A slot stays taken until its copy has been seen to
    // complete, so with many calls in flight the ring holds the OLDEST outstanding reports (they complete first as
    // the stream drains) and newer calls skip the report instead of overwriting an unread one.
    int slot = -1;
    for (int s = 0; s < 4 && slot < 0; ++s)
        if (ctx->qp_hmax_call[s] < 0) slot = s;
    auto report_to_host = [&]() -> cudaError_t {
        if (slot < 0) return cudaSuccess;
        cudaError_t e = cudaMemcpyAsync(ctx->qp_hmax_host + slot, dmax, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) return e;
        ctx->qp_hmax_call[slot] = call;
        return cudaEventRecord(ctx->qp_hmax_ev[slot], ctx->stream);
    };
    // The largest active set of THIS batch configures a later call.  First call (and when the LDL' kernel does not
    // run): a small scan kernel; otherwise the LDL' kernel reports it itself while it assembles the instances.
    const bool want_ldl = !(force && strcmp(force, "lu") == 0);
    auto scan_active = [&]() -> cudaError_t {
        int64_t blocks = (a.B + 7) / 8;
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        if (!headline) return qp_max_active_any_launch(ctx, a, dmax) == 0 ? cudaSuccess : cudaErrorUnknown;
        max_active_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a.B, a.lam, dmax);
        ctx->launches++;
        return cudaGetLastError();
    };
    // the scan kernel needs a cleared word; the LDL' kernel's tagged values (call number << 8 | size) only grow
    ctx->qp_seq = (ctx->qp_seq + 1) & 0x3FFFFF;
    const bool clear = first || !want_ldl || ctx->qp_seq == 0;
    if (clear) DO_CUDA(ctx, cudaMemsetAsync(dmax, 0, sizeof(int), ctx->stream));
    if (first) {
        DO_CUDA(ctx, scan_active());
        DO_CUDA(ctx, report_to_host());
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->qp_hint = ctx->qp_hmax_host[slot] & 0xFF;
        ctx->qp_hmax_call[slot] = -1;
    }
    // (a shape whose worst case does not fit shared memory is configured for what fits: larger active sets go to the LU kernel)
    const int nt_cap = headline ? (NV + ctx->qp_hint + PE + 7) / 8 : std::min(qp_sqd_any_nt(a, ctx->qp_hint), qp_sqd_any_nt_limit(ctx, a));
    bool reported = first;
    if (want_ldl) {
        int32_t rc = headline ? qp_sqd_launch(ctx, a, nt_cap, handled, first ? nullptr : dmax, ctx->qp_seq << 8)
                              : qp_sqd_any_launch(ctx, a, nt_cap, handled, first ? nullptr : dmax, ctx->qp_seq << 8);
        if (rc != 0) return rc;
        if (*handled) {
            if (!first) DO_CUDA(ctx, report_to_host());
            ctx->qp_last_kernel = 2;
            ctx->qp_last_hint = ctx->qp_hint;
            return 0;
        }
    }
    if (!headline) return 0;  // (not reached: qp_sqd_any_supported covers the launch) the generic kernel takes the batch
    if (!reported) {
        if (!clear) DO_CUDA(ctx, cudaMemsetAsync(dmax, 0, sizeof(int), ctx->stream));
        DO_CUDA(ctx, scan_active());
        DO_CUDA(ctx, report_to_host());
    }
    ctx->qp_last_kernel = 1;
    ctx->qp_last_hint = ctx->qp_hint;
    if (!first) {  // the guess may be too small for the plain LU launch: size it for the worst case
        return lu_launch(ctx, a, NTMAX, nullptr, nullptr, handled);
    }
    return lu_launch(ctx, a, nt_cap, nullptr, nullptr, handled);
}

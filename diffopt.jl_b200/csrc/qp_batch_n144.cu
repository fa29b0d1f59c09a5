// Tuned batched KKT sensitivity kernel for the headline shape n=64, m=64, p=16 (KKT order N=144).
//
// One persistent CTA (512 threads) per SM streams over its QP instances.  Per instance:
//   1. assemble LHS = [Q G'diag(lam) A'; G diag(Gz-h) 0; A 0 0] (QuadraticProgram.jl:256-282) straight from HBM
//      into shared memory as an 18 x 18 grid of 8 x 8 tiles (row-major inside a tile = DMMA C/A fragment order);
//      build the forward RHS (:429-433) and the reverse RHS (:324-329) as two vectors that ride along;
//   2. right-looking blocked LU with partial pivoting, panel width 8:
//        - warp 0 holds the 8-column panel in REGISTERS and does the 8 pivot steps with warp collectives
//          (redux.max on the high words for the pivot search, shuffles for the pivot row); LAPACK-style
//          row interchanges are tracked per row in registers and applied once per panel;
//        - U12 = L11^-1 A12 (thread per column), written both in place and as column-major B tiles;
//        - trailing update C -= L21 U12 on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA), 16 warps in a
//          4 x 4 cyclic tile decomposition, A/B fragments reused from registers;
//        - the two RHS vectors get their L^-1 P / U^-T forward substitutions in the same sweep;
//   3. blocked backward substitutions (explicit inverses of the 8 x 8 diagonal tiles) finish
//      LHS x_b = r_b and LHS' x_f = r_f from the one factorisation (:335 uses LHS, :438 uses LHS').
// The KKT matrix and its factors never touch HBM: algorithmic traffic is inputs + outputs only.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NV = 64, MI = 64, PE = 16, N = 144, NT = 18;
constexpr int THREADS = 512;
constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) Smem {
    double K[NT * NT * 64];   // tile (I,J) at (J*NT + I)*64, element (r,c) at r*8+c
    double Ub[NT * 64];       // U row block of the current panel, tile J column-major: (k,n) at n*8+k
    double invU[NT * 64];     // inverse of the upper-triangular diagonal tiles, row-major
    double invL[NT * 64];     // inverse of the unit-lower diagonal tiles, row-major
    double y[N];              // reverse RHS -> L^-1 P r_b -> x_b
    double rf[N];             // forward RHS -> w = U^-T r_f -> v
    double rdiag[N];          // 1 / U_kk
    double zs[NV], lams[MI], nus[PE];
    double part[8 * 64];      // reduction scratch (row sums)
    double cpart[2 * 64];     // column-sum scratch [row half][col]
    double prow[2][12];       // pivot-row broadcast staging (double buffered): 8 values, rinv, position
    double wblk[8];
    int perm[N];              // position -> original row
    int mv_src[16], mv_dst[16];  // element offsets (row part) of the interchanges of the current panel
    int mv_srcrow[16], mv_dstrow[16];
    int nmoves, info;
};

__device__ __forceinline__ int tile_off(int I, int J) { return (J * NT + I) << 6; }
__device__ __forceinline__ int row_off(int r) { return ((r >> 3) << 6) + ((r & 7) << 3); }
__device__ __forceinline__ int col_off(int c) { return (((c >> 3) * NT) << 6) + (c & 7); }
__device__ __forceinline__ int elem_off(int r, int c) { return row_off(r) + col_off(c); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// 8 per-lane values -> sums over the 32 lanes; on return lane L holds the total of value index (L >> 2) & 7
// (transpose-reduce: 9 shuffles instead of 40)
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
    return c;  // value index = 4*h16 + 2*h8 + h4
}

// ---- panel factorisation by one warp, panel held in registers -----------------------------------------
// NS = number of 32-row slots the panel still has.  The forward-RHS row (rf) rides along as an extra row
// that is never chosen as pivot: its multipliers are exactly w = rf U^-1 for this block.
template <int NS>
__device__ __forceinline__ void panel_factor(Smem& S, const int j, const int lane) {
    const int c0 = j << 3;
    double p[NS][8];
    double e[8];
    int dstpos[NS];
    unsigned live = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int r = c0 + lane + 32 * s;
        dstpos[s] = r;
        if (r < N) {
            live |= 1u << s;
            const double2* src = reinterpret_cast<const double2*>(&S.K[tile_off(r >> 3, j) + ((r & 7) << 3)]);
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
        const double rloc = __drcp_rn(bv);  // speculative: overlaps the reduction latency
        const int owner = __ffs(__ballot_sync(FULL, best == kmax)) - 1;
        const int sp = (int)(kmax & 7u);
        double* stage = S.prow[k & 1];
        if (lane == owner) {
            // the owner publishes its pivot row, 1/pivot and the row's current position
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (s == sp) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c >= k) stage[c] = p[s][c];
                    stage[8] = (kmax & 0x7ffffff8u) == 0u ? 0.0 : rloc;
                    reinterpret_cast<int*>(stage + 9)[0] = dstpos[s];
                    dstpos[s] = -1;  // marks "this lane's pivot row": fixed up below
                }
            }
            live &= ~(1u << sp);
        }
        __syncwarp();
        double pv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c >= k) pv[c] = stage[c];
        const double rinv = stage[8];
        const int P = reinterpret_cast<const int*>(stage + 9)[0];
        if ((kmax & 0x7ffffff8u) == 0u && lane == 0 && S.info == 0) S.info = c0 + k + 1;  // exactly zero pivot
        // LAPACK interchange k: the row sitting at position c0+k goes to the pivot row's position P
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int d = dstpos[s];
            d = (d == c0 + k) ? P : d;
            d = (d == -1) ? c0 + k : d;
            dstpos[s] = d;
        }
        if (lane == 0) S.rdiag[c0 + k] = rinv;
        // multipliers and rank-1 update of the live rows (next column first: it is on the critical path)
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
    unsigned base = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int r = c0 + lane + 32 * s;
        const bool ok = (valid >> s) & 1u;
        const int pos = dstpos[s];
        if (ok) {
            const int kp = pos - c0;  // < 8: this is pivot row kp (entries c >= kp are U), else all multipliers
            double2* dst = reinterpret_cast<double2*>(&S.K[tile_off(pos >> 3, j) + ((pos & 7) << 3)]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double a = (2 * q < kp) ? -p[s][2 * q] : p[s][2 * q];
                double b = (2 * q + 1 < kp) ? -p[s][2 * q + 1] : p[s][2 * q + 1];
                dst[q] = make_double2(a, b);
            }
        }
        const bool moved = ok && pos != r;
        const unsigned mk = __ballot_sync(FULL, moved);
        if (moved) {
            const int idx = base + __popc(mk & ((1u << lane) - 1u));
            S.mv_src[idx] = row_off(r);
            S.mv_dst[idx] = row_off(pos);
            S.mv_srcrow[idx] = r;
            S.mv_dstrow[idx] = pos;
        }
        base += __popc(mk);
    }
    if (lane < 8) {
        double w = e[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) w = (lane == c) ? e[c] : w;
        S.rf[c0 + lane] = w;
        S.wblk[lane] = w;
    }
    if (lane == 0) S.nmoves = (int)base;
}

__device__ __forceinline__ void panel_dispatch(Smem& S, int j, int lane) {
    const int M = N - (j << 3);
    if (M > 128) panel_factor<5>(S, j, lane);
    else if (M > 96) panel_factor<4>(S, j, lane);
    else if (M > 64) panel_factor<3>(S, j, lane);
    else if (M > 32) panel_factor<2>(S, j, lane);
    else panel_factor<1>(S, j, lane);
}

__global__ void __launch_bounds__(THREADS, 1) qp_kkt_n144_kernel(QpSolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool do_fwd = a.fwd != nullptr, do_rev = a.rev != nullptr;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#define PROF(i)                                \
    do {                                       \
        long long _n = clock64();              \
        pc[i] += _n - tprev;                   \
        tprev = _n;                            \
    } while (0)

    for (int64_t inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
        const double* Q = a.Q + (size_t)inst * NV * NV;
        const double* G = a.G + (size_t)inst * MI * NV;
        const double* A = a.A + (size_t)inst * PE * NV;
        // ---- vectors, clears
        if (tid < NV) S.zs[tid] = a.z[(size_t)inst * NV + tid];
        else if (tid < NV + MI) S.lams[tid - NV] = a.lam[(size_t)inst * MI + tid - NV];
        else if (tid < NV + MI + PE) S.nus[tid - NV - MI] = a.nu[(size_t)inst * PE + tid - NV - MI];
        if (tid < N) {
            S.y[tid] = (do_rev && tid < NV) ? a.seed[(size_t)inst * NV + tid] : 0.0;
            S.perm[tid] = tid;
        }
        if (tid == 0) S.info = 0;
        // zero the structurally-zero blocks: tile rows 8..17 x tile cols 8..17
        for (int i = tid; i < 10 * 10 * 32; i += THREADS) {
            int J = 8 + i / 320, rem = i % 320;  // 10 tiles * 32 double2 per tile column
            reinterpret_cast<double2*>(&S.K[tile_off(8, J)])[rem] = make_double2(0.0, 0.0);
        }
        {   // L2 prefetch of the next instance of this CTA (inputs are streamed once from HBM)
            const int64_t nxt = inst + gridDim.x;
            if (nxt < a.B) {
                const char* bases[6] = {(const char*)(a.Q + (size_t)nxt * NV * NV), (const char*)(a.G + (size_t)nxt * MI * NV),
                                        (const char*)(a.A + (size_t)nxt * PE * NV),
                                        do_fwd && a.dQ ? (const char*)(a.dQ + (size_t)nxt * NV * NV) : nullptr,
                                        do_fwd && a.dG ? (const char*)(a.dG + (size_t)nxt * MI * NV) : nullptr,
                                        do_fwd && a.dA ? (const char*)(a.dA + (size_t)nxt * PE * NV) : nullptr};
                const int lines[6] = {256, 256, 64, 256, 256, 64};  // 128-byte lines
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (bases[q] && tid < lines[q])
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(bases[q] + (size_t)tid * 128));
            }
        }
        __syncthreads();
        // ---- assemble (coalesced column-major reads; thread keeps a fixed row r and 8 columns cg + 8 i)
        {
            const int r = tid & 63, cg = tid >> 6;
            const double lam_r = S.lams[r];
            double qv[8], gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cg + 8 * i;
                qv[i] = __ldg(Q + c * NV + r);
                gv[i] = __ldg(G + c * MI + r);
            }
            double av[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) av[i] = __ldg(A + tid + THREADS * i);
            double dacc = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cg + 8 * i;
                S.K[elem_off(r, c)] = qv[i];
                S.K[elem_off(NV + r, c)] = gv[i];
                S.K[elem_off(c, NV + r)] = gv[i] * lam_r;
                dacc = fma(gv[i], S.zs[c], dacc);
            }
            S.part[cg * 64 + r] = dacc;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + THREADS * i, ii = idx & 15, c = idx >> 4;
                S.K[elem_off(NV + MI + ii, c)] = av[i];
                S.K[elem_off(c, NV + MI + ii)] = av[i];
            }
        }
        __syncthreads();
        if (tid < MI) {
            double d = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) d += S.part[q * 64 + tid];
            S.K[elem_off(NV + tid, NV + tid)] = d - a.h[(size_t)inst * MI + tid];
        }
        __syncthreads();
        // ---- forward RHS (QuadraticProgram.jl:429-433): [dQ z + dq + dG'lam + dA'nu ; lam.(dG z - dh) ; dA z - db]
        if (do_fwd) {
            const size_t b = (size_t)inst;
            const int r = tid & 63, cg = tid >> 6;
            double rq = 0.0, rg = 0.0;
            double cv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cv[i] = 0.0;
            if (a.dQ) {
                const double* X = a.dQ + b * NV * NV;
#pragma unroll
                for (int i = 0; i < 8; ++i) rq = fma(__ldg(X + (cg + 8 * i) * NV + r), S.zs[cg + 8 * i], rq);
            }
            if (a.dG) {
                const double* X = a.dG + b * MI * NV;
                const double lr = S.lams[r];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double v = __ldg(X + (cg + 8 * i) * MI + r);
                    rg = fma(v, S.zs[cg + 8 * i], rg);
                    cv[i] = v * lr;
                }
            }
            S.part[cg * 64 + r] = rq;
            const double csum = warp_reduce8(cv, lane);   // lane L: column cg + 8*((L>>2)&7), rows of this warp
            if ((lane & 3) == 0) S.cpart[(warp & 1) * 64 + cg + 8 * ((lane >> 2) & 7)] = csum;
            __syncthreads();
            double r1 = 0.0;
            if (tid < NV) {
#pragma unroll
                for (int q = 0; q < 8; ++q) r1 += S.part[q * 64 + tid];
                r1 += S.cpart[tid] + S.cpart[64 + tid];
            }
            __syncthreads();
            S.part[cg * 64 + r] = rg;
            // dA: 16 x 64, thread (ii = tid & 15, c = tid >> 4 and +32)
            double ra = 0.0, ca0 = 0.0, ca1 = 0.0;
            if (a.dA) {
                const double* X = a.dA + b * PE * NV;
                const int ii = tid & 15, c = tid >> 4;
                const double v0 = __ldg(X + tid), v1 = __ldg(X + tid + THREADS);
                ra = v0 * S.zs[c] + v1 * S.zs[c + 32];
                ca0 = v0 * S.nus[ii];
                ca1 = v1 * S.nus[ii];
            }
            // reduce ra over the lanes with equal ii (lane bit 4) and ca over ii (lane bits 0..3)
            ra += __shfl_xor_sync(FULL, ra, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                ca0 += __shfl_xor_sync(FULL, ca0, o);
                ca1 += __shfl_xor_sync(FULL, ca1, o);
            }
            // scratch: invU / invL are free until the end of the factorisation
            if ((lane & 15) == 0) {
                S.invL[(tid >> 4)] = ca0;        // column c = tid >> 4
                S.invL[32 + (tid >> 4)] = ca1;   // column c + 32
            }
            if (lane < 16) S.invU[warp * 16 + lane] = ra;
            __syncthreads();
            if (tid < NV) {
                double v = r1 + S.invL[tid];
                if (a.dq) v += a.dq[b * NV + tid];
                S.rf[tid] = v;
            } else if (tid < NV + MI) {
                const int i = tid - NV;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) v += S.part[q * 64 + i];
                if (a.dh) v -= a.dh[b * MI + i];
                S.rf[tid] = S.lams[i] * v;
            } else if (tid < N) {
                const int i = tid - NV - MI;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 16; ++w) v += S.invU[w * 16 + i];
                if (a.db) v -= a.db[b * PE + i];
                S.rf[tid] = v;
            }
        } else if (tid < N) {
            S.rf[tid] = 0.0;
        }
        __syncthreads();
        PROF(0);

        // ---- blocked LU
        for (int j = 0; j < NT; ++j) {
            const int c0 = j << 3;
            if (warp == 0) panel_dispatch(S, j, lane);
            __syncthreads();
            PROF(1);
            // (b1) row interchanges on every other column: read phase (3 threads per column)
            const int nm = S.nmoves;
            const int ccol = tid % 160, part = tid / 160;  // column 144 = the reverse-RHS vector y
            const bool swapper = tid < 480 && ccol <= N && (ccol < c0 || ccol >= c0 + 8);
            double tmp[6];
            {
                if (swapper) {
                    if (ccol < N) {
                        const int cb = col_off(ccol);
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            const int i = part + 3 * q;
                            if (i < nm) tmp[q] = S.K[S.mv_src[i] + cb];
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            const int i = part + 3 * q;
                            if (i < nm) tmp[q] = S.y[S.mv_srcrow[i]];
                        }
                    }
                } else if (tid == 500) {  // permutation bookkeeping
                    int t2[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < nm) t2[i] = S.perm[S.mv_srcrow[i]];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < nm) S.perm[S.mv_dstrow[i]] = t2[i];
                }
            }
            __syncthreads();
            // (b2) write phase
            if (swapper) {
                if (ccol < N) {
                    const int cb = col_off(ccol);
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int i = part + 3 * q;
                        if (i < nm) S.K[S.mv_dst[i] + cb] = tmp[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int i = part + 3 * q;
                        if (i < nm) S.y[S.mv_dstrow[i]] = tmp[q];
                    }
                }
            }
            __syncthreads();
            // (c) U12 = L11^-1 A12, thread per trailing column (column 144 = y)
            if (tid <= N && tid >= c0 + 8) {
                const double* Ld = &S.K[tile_off(j, j)];  // diagonal tile: negated multipliers below the diagonal
                double x[8];
                if (tid < N) {
                    double* colp = &S.K[tile_off(j, tid >> 3) + (tid & 7)];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = colp[i * 8];
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < i) x[i] = fma(Ld[i * 8 + k], x[k], x[i]);
                        colp[i * 8] = x[i];
                    }
                    double2* ub = reinterpret_cast<double2*>(&S.Ub[((tid >> 3) << 6) + ((tid & 7) << 3)]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) ub[q] = make_double2(x[2 * q], x[2 * q + 1]);
                } else if (do_rev) {
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
            }
            __syncthreads();
            PROF(2);
            // (d) trailing update on the DMMA pipe: C(I,J) += Lneg(I,j) * U(j,J)
            if (j < NT - 1) {
                const int wr = warp & 3, wc = warp >> 2;
                double2 af[5];
#pragma unroll
                for (int ii = 0; ii < 5; ++ii) {
                    const int I = j + 1 + wr + 4 * ii;
                    if (I < NT) af[ii] = *reinterpret_cast<const double2*>(&S.K[tile_off(I, j) + g * 8 + 2 * t]);
                }
                for (int J = j + 1 + wc; J < NT; J += 4) {
                    const double2 bf = *reinterpret_cast<const double2*>(&S.Ub[(J << 6) + g * 8 + 2 * t]);
                    double2 cc[5];
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < NT) cc[ii] = *reinterpret_cast<const double2*>(&S.K[tile_off(I, J) + g * 8 + 2 * t]);
                    }
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < NT) dmma(cc[ii].x, cc[ii].y, af[ii].x, bf.x);
                    }
#pragma unroll
                    for (int ii = 0; ii < 5; ++ii) {
                        const int I = j + 1 + wr + 4 * ii;
                        if (I < NT) {
                            dmma(cc[ii].x, cc[ii].y, af[ii].y, bf.y);
                            *reinterpret_cast<double2*>(&S.K[tile_off(I, J) + g * 8 + 2 * t]) = cc[ii];
                        }
                    }
                }
                if (warp == 15 && do_rev) {  // y[r] += Lneg[r, panel] * y_blk
                    double yb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) yb[k] = S.y[c0 + k];
#pragma unroll
                    for (int s = 0; s < 5; ++s) {
                        const int r = c0 + 8 + lane + 32 * s;
                        if (r < N) {
                            const double2* lrow = reinterpret_cast<const double2*>(&S.K[tile_off(r >> 3, j) + ((r & 7) << 3)]);
                            double v = S.y[r];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const double2 l2 = lrow[q];
                                v = fma(l2.x, yb[2 * q], v);
                                v = fma(l2.y, yb[2 * q + 1], v);
                            }
                            S.y[r] = v;
                        }
                    }
                } else if (warp == 14 && do_fwd) {  // rf[c] -= w_blk * U[panel rows, c]
                    double wb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) wb[k] = S.wblk[k];
#pragma unroll
                    for (int s = 0; s < 5; ++s) {
                        const int c = c0 + 8 + lane + 32 * s;
                        if (c < N) {
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
        if (tid < 2 * NT * 8) {
            const int which = tid / (NT * 8), rem = tid % (NT * 8), jt = rem >> 3, c = rem & 7;
            const double* Ld = &S.K[tile_off(jt, jt)];
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
            for (int j = NT - 1; j >= 0; --j) {
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
                    const int r = lane + 32 * s;
                    if (r < c0) {
                        const double2* urow = reinterpret_cast<const double2*>(&S.K[tile_off(r >> 3, j) + ((r & 7) << 3)]);
                        double v = S.y[r];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double2 u2 = urow[q];
                            v = fma(-u2.x, x[2 * q], v);
                            v = fma(-u2.y, x[2 * q + 1], v);
                        }
                        S.y[r] = v;
                    }
                }
                __syncwarp();
            }
        } else if (warp == 1 && do_fwd) {
            for (int j = NT - 1; j >= 0; --j) {
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
                        const double* lt = &S.K[tile_off(j, c >> 3) + (c & 7)];
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
        // ---- outputs: (dz, dlam, dnu) = -x ;  x_f = P' v
        if (tid < N) {
            if (do_rev) a.rev[(size_t)inst * N + tid] = -S.y[tid];
        } else if (tid >= 256 && tid < 256 + N) {
            const int i = tid - 256;
            if (do_fwd) a.fwd[(size_t)inst * N + S.perm[i]] = -S.rf[i];
        }
        if (a.info && tid == 0) a.info[inst] = S.info;
        __syncthreads();
        PROF(5);
    }
    if (a.prof && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) a.prof[i] = pc[i];
}

}  // namespace

int32_t qp_batch_launch_tuned(diffopt_b200_ctx* ctx, const QpSolveArgs& a, bool* handled) {
    *handled = false;
    if (a.n != NV || a.m != MI || a.p != PE) return 0;
    const char* force = getenv("DIFFOPT_B200_QP_KERNEL");
    if (force && strcmp(force, "generic") == 0) return 0;
    if (sizeof(Smem) > ctx->smem_optin) return 0;
    *handled = true;
    DO_CUDA(ctx, cudaFuncSetAttribute(qp_kkt_n144_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(Smem)));
    int64_t grid = a.B < (int64_t)ctx->sm_count ? a.B : (int64_t)ctx->sm_count;
    QpSolveArgs aa = a;
    const bool profile = getenv("DIFFOPT_B200_PROFILE") != nullptr;
    long long* dprof = nullptr;
    if (profile) {
        DO_CUDA(ctx, cudaMalloc(&dprof, 8 * sizeof(long long)));
        DO_CUDA(ctx, cudaMemsetAsync(dprof, 0, 8 * sizeof(long long), ctx->stream));
        aa.prof = dprof;
    }
    qp_kkt_n144_kernel<<<(unsigned)grid, THREADS, sizeof(Smem), ctx->stream>>>(aa);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    if (profile) {
        long long h[8];
        DO_CUDA(ctx, cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dprof);
        long long ninst = (a.B + grid - 1) / grid;
        fprintf(stderr, "[qp_n144 profile, CTA 0, %lld instances] clocks/instance: assemble %lld panel %lld swap+trsm %lld "
                        "update %lld backward %lld output %lld\n",
                ninst, h[0] / ninst, h[1] / ninst, h[2] / ninst, h[3] / ninst, h[4] / ninst, h[5] / ninst);
    }
    return 0;
}

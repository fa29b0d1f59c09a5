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

#include "qp_sqd_dev.cuh"

int32_t qp_lu_launch_list(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, const int* list, const int* count);

namespace {

using namespace sqd;

constexpr int NV = 64, MI = 64, PE = 16, N = NV + MI + PE, NTZ = NV / 8;

struct __align__(16) Hdr {
    double yf[N + 8], yb[N + 8];  // right-hand sides -> v = D^-1 L^-1 r -> solutions (reduced ordering)
    double sf[N + 8], sb[N + 8];    // backward-substitution accumulators of the critical warps
    double sf2[N + 8], sb2[N + 8];  // ... and of their helper warps
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

// one 8-tile row of a column-major matrix (leading dimension ld): lane (g, t) gets rows r0, r0+1 of column 8J+g
__device__ __forceinline__ void load_row8(const double* M, const int ld, const int r0, const int g, double2 (&v)[8]) {
    if (M) {
#pragma unroll
        for (int J = 0; J < 8; ++J) v[J] = ldg2(M + (8 * J + g) * ld + r0);
    } else {
#pragma unroll
        for (int J = 0; J < 8; ++J) v[J] = make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(THREADS, 4) qp_kkt_sqd_kernel(QpSolveArgs a, int* fb_list, int* fb_count, const int nt_cap,
                                                                 int* max_active, const int max_tag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Hdr& S = *reinterpret_cast<Hdr*>(smem_raw);
    double* const T = reinterpret_cast<double*>(smem_raw + sizeof(Hdr));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // g, t: DMMA fragment coordinates; the same pair addresses the tile-shaped global loads (column g, rows 2t, 2t+1)
    const int g = lane >> 2, t = lane & 3;
    const int fo = el(g, 2 * t);  // fragment offset inside a tile
    const bool do_fwd = a.fwd != nullptr, do_rev = a.rev != nullptr;
#ifdef QP_PROFILE
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long sub[6] = {0, 0, 0, 0, 0, 0};
    long long asub[6] = {0, 0, 0, 0, 0, 0};
    long long tprev2 = 0;
    long long tsub = 0;
    long long tprev = clock64();
#define PROF(i)                  \
    do {                         \
        long long _n = clock64(); \
        pc[i] += _n - tprev;     \
        tprev = _n;              \
    } while (0)
#define ASUB(i)                  \
    do {                         \
        long long _n = clock64(); \
        asub[i] += _n - tprev2;  \
        tprev2 = _n;             \
    } while (0)
#else
#define PROF(i)
#define ASUB(i)
#endif
    const Vecs V = {S.yf, S.yb, S.sf, S.sb, S.sf2, S.sb2, S.ref, S.rd, &S.fail};

    for (int64_t inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
        const size_t b = (size_t)inst;
#ifdef QP_PROFILE
        tprev2 = clock64();
#endif
        const size_t bm = (a.shared & 1) ? 0 : b, bd = (a.shared & 2) ? 0 : b;  // shared matrices: one instance serves the batch
        const double* Q = a.Q + bm * NV * NV;
        const double* G = a.G + bm * MI * NV;
        const double* A = a.A + bm * PE * NV;
        // ---- vectors, active set
        if (tid < NV) {
            S.zs[tid] = a.z[b * NV + tid];
            S.yb[tid] = do_rev ? a.seed[b * NV + tid] : 0.0;
        } else if (tid < NV + PE) {
            S.nus[tid - NV] = a.nu[b * PE + tid - NV];
        } else if (warp == 3) {
            const double l0 = a.lam[b * MI + lane], l1 = a.lam[b * MI + 32 + lane];
            S.lams[lane] = l0;
            S.lams[32 + lane] = l1;
            const unsigned m0 = __ballot_sync(FULL, l0 != 0.0), m1 = __ballot_sync(FULL, l1 != 0.0);
            const unsigned lt = (1u << lane) - 1u;
            S.apos[lane] = (l0 != 0.0) ? __popc(m0 & lt) : -1;
            S.apos[32 + lane] = (l1 != 0.0) ? __popc(m0) + __popc(m1 & lt) : -1;
            if (lane == 0) {
                const int ma = __popc(m0) + __popc(m1);
                S.ma = ma;
                // largest active set of the batch (configures the next call); tagged with the call number in the high
                // bits so that the word never has to be cleared between calls
                if (max_active) atomicMax(max_active, max_tag | ma);
                S.nt = (NV + ma + PE + 7) >> 3;
                S.fail = 0;
            }
        }
        // Loads are software pipelined through two 8-tile register buffers: the next group is in flight while
        // the current one is consumed.  Q: lower-triangle tiles only; warp w owns tile rows w and 7-w (9 tiles).
        double2 qv[9], bufA[8], bufB[8], av[4];
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) {
            const bool first = s9 <= warp;
            const int I = first ? warp : 7 - warp, J = first ? s9 : s9 - warp - 1;
            qv[s9] = ldg2(Q + (8 * J + g) * NV + 8 * I + 2 * t);
        }
        load_row8(G, MI, 8 * warp + 2 * t, g, bufA);
        if (do_fwd) {  // this instance's direction data is consumed a few microseconds from now: pull it into L2
            const char* bases[3] = {a.dQ ? (const char*)(a.dQ + bd * NV * NV) : nullptr, a.dG ? (const char*)(a.dG + bd * MI * NV) : nullptr,
                                    a.dA ? (const char*)(a.dA + bd * PE * NV) : nullptr};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (!bases[q]) continue;
                for (int l = tid; l < (q < 2 ? 256 : 64); l += THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(bases[q] + (size_t)l * 128));
            }
        }
        __syncthreads();
        ASUB(0);
        const int ma = S.ma, nt = S.nt;
        const int nred = NV + ma + PE, np = nt << 3;
        if (nt > nt_cap) {  // larger than this launch was configured for: the pivoted-LU kernel takes it
            if (tid == 0) fb_list[atomicAdd(fb_count, 1)] = (int)inst;
            __syncthreads();
            continue;
        }
        // ---- clear what the fill below does not overwrite: the tiles of the (2,2) block and the padding rows
        {
            const int nneg = nt - NTZ;                      // tile rows below the z block
            const int cnt = ((nneg * (nneg + 1)) >> 1) * 32;  // double2 slots of the tiles (I >= NTZ, NTZ <= J <= I)
            for (int i = tid; i < cnt; i += THREADS) {
                const int tl = i >> 5;                        // tile number inside the (2,2) block, row major lower
                int I = 0;
                while (((I + 1) * (I + 2)) >> 1 <= tl) ++I;
                const int J = tl - ((I * (I + 1)) >> 1);
                reinterpret_cast<double2*>(T + tix(NTZ + I, NTZ + J) * 64)[i & 31] = make_double2(0.0, 0.0);
            }
            const int npad = np - nred;                       // rows nred .. np-1 of the last tile row, columns 0 .. NV-1
            for (int i = tid; i < npad * NV; i += THREADS) {
                const int r = nred + i / NV, c = i % NV;
                T[tix(nt - 1, c >> 3) * 64 + el(r & 7, c & 7)] = 0.0;
            }
        }
        if (tid < np - NV) {
            S.yb[NV + tid] = 0.0;
            S.yf[NV + tid] = 0.0;
        }
        for (int i = tid; i < np; i += THREADS) {
            S.sf[i] = 0.0;
            S.sb[i] = 0.0;
            S.sf2[i] = 0.0;
            S.sb2[i] = 0.0;
        }
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) {
            const bool first = s9 <= warp;
            const int I = first ? warp : 7 - warp, J = first ? s9 : s9 - warp - 1;
            double* tp = T + tix(I, J) * 64;
            tp[el(2 * t, g)] = qv[s9].x;
            tp[el(2 * t + 1, g)] = qv[s9].y;
            if (J == I) {
                if (g == 2 * t) S.ref[8 * I + g] = qv[s9].x;
                if (g == 2 * t + 1) S.ref[8 * I + g] = qv[s9].y;
            }
        }
        load_row8(G, MI, 8 * (warp + 4) + 2 * t, g, bufB);
        double hv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hv[q] = a.h[b * MI + 8 * (warp + 4 * (q >> 1)) + 2 * t + (q & 1)];
        __syncthreads();
        ASUB(1);
        if (tid < np - nred) {  // identity padding of rows nred .. np-1 (indexed from nred: nred may exceed the CTA size)
            const int r = (nred + tid) & 7;
            T[tix(nt - 1, nt - 1) * 64 + el(r, r)] = -1.0;
        }
        // ---- G: D = G z - h for every row; active rows go to the matrix; warp w owns tile rows w, w+4
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            double2(&gv)[8] = hh ? bufB : bufA;
            const int I = warp + 4 * hh, r0 = 8 * I + 2 * t;
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int J = 0; J < 8; ++J) {
                const double zc = S.zs[8 * J + g];
                d0 = fma(gv[J].x, zc, d0);
                d1 = fma(gv[J].y, zc, d1);
            }
            const int a0 = S.apos[r0], a1 = S.apos[r0 + 1];
            if (a0 >= 0) {
                double* tp = T + tix(NTZ + (a0 >> 3), 0) * 64 + el(a0 & 7, g);
#pragma unroll
                for (int J = 0; J < 8; ++J) tp[J * 64] = gv[J].x;
            }
            if (a1 >= 0) {
                double* tp = T + tix(NTZ + (a1 >> 3), 0) * 64 + el(a1 & 7, g);
#pragma unroll
                for (int J = 0; J < 8; ++J) tp[J * 64] = gv[J].y;
            }
            // next group in flight: dQ tile rows (forward mode)
            if (do_fwd) load_row8(a.dQ ? a.dQ + bd * NV * NV : nullptr, NV, r0, g, gv);
            d0 = sum_over_g(d0);
            d1 = sum_over_g(d1);
            if (g == 0) {
                d0 -= hv[2 * hh];
                d1 -= hv[2 * hh + 1];
                S.dvec[r0] = d0;
                S.dvec[r0 + 1] = d1;
                if (a0 >= 0) T[tix(NTZ + (a0 >> 3), NTZ + (a0 >> 3)) * 64 + el(a0 & 7, a0 & 7)] = d0 * fast_rcp(S.lams[r0]);
                else if (d0 == 0.0) S.fail = 1;  // lam_i = D_i = 0: singular column, the LU path reports it
                if (a1 >= 0) T[tix(NTZ + (a1 >> 3), NTZ + (a1 >> 3)) * 64 + el(a1 & 7, a1 & 7)] = d1 * fast_rcp(S.lams[r0 + 1]);
                else if (d1 == 0.0) S.fail = 1;
            }
        }
        ASUB(2);
        // ---- A (16 x 64): warp w owns tile columns 2w, 2w+1
#pragma unroll
        for (int q = 0; q < 4; ++q) av[q] = ldg2(A + (8 * (2 * warp + (q >> 1)) + g) * PE + 8 * (q & 1) + 2 * t);
        double vq = 0.0, vh = 0.0, vb = 0.0;  // dq / dh / db entries this thread will need
        if (do_fwd) {
            if (tid < NV) {
                if (a.dq) vq = a.dq[b * NV + tid];
            } else {
                if (a.dh) vh = a.dh[b * MI + tid - NV];
                if (a.db && tid - NV < PE) vb = a.db[b * PE + tid - NV];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int J = 2 * warp + (q >> 1), Ia = q & 1;
            const int k0 = NV + ma + 8 * Ia + 2 * t, k1 = k0 + 1;
            T[tix(k0 >> 3, J) * 64 + el(k0 & 7, g)] = av[q].x;
            T[tix(k1 >> 3, J) * 64 + el(k1 & 7, g)] = av[q].y;
        }
        ASUB(3);
        // ---- forward right-hand side (QuadraticProgram.jl:429-433), symmetric-form scaling
        if (do_fwd) {
            if (a.dA) {
#pragma unroll
                for (int q = 0; q < 4; ++q) av[q] = ldg2(a.dA + bd * PE * NV + (8 * (2 * warp + (q >> 1)) + g) * PE + 8 * (q & 1) + 2 * t);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) av[q] = make_double2(0.0, 0.0);
            }
            const double* dGp = a.dG ? a.dG + bd * MI * NV : nullptr;
            // dQ z
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                double2(&v)[8] = hh ? bufB : bufA;
                const int r0 = 8 * (warp + 4 * hh) + 2 * t;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int J = 0; J < 8; ++J) {
                    const double zc = S.zs[8 * J + g];
                    s0 = fma(v[J].x, zc, s0);
                    s1 = fma(v[J].y, zc, s1);
                }
                load_row8(dGp, MI, r0, g, v);
                s0 = sum_over_g(s0);
                s1 = sum_over_g(s1);
                if (g == 0) {
                    S.rowq[r0] = s0;
                    S.rowq[r0 + 1] = s1;
                }
            }
            // dG z (rows) and dG' lam (columns)
            double cs[8];
#pragma unroll
            for (int J = 0; J < 8; ++J) cs[J] = 0.0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                double2(&v)[8] = hh ? bufB : bufA;
                const int r0 = 8 * (warp + 4 * hh) + 2 * t;
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
            // dA z (partial rows over this warp's 16 columns) and dA' nu (full columns)
#pragma unroll
            for (int Ia = 0; Ia < 2; ++Ia) {
                const double z0 = S.zs[16 * warp + g], z1 = S.zs[16 * warp + 8 + g];
                double s0 = fma(av[Ia].x, z0, av[2 + Ia].x * z1), s1 = fma(av[Ia].y, z0, av[2 + Ia].y * z1);
                s0 = sum_over_g(s0);
                s1 = sum_over_g(s1);
                if (g == 0) {
                    S.arow[warp][8 * Ia + 2 * t] = s0;
                    S.arow[warp][8 * Ia + 2 * t + 1] = s1;
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                double c = av[2 * jj].x * S.nus[2 * t] + av[2 * jj].y * S.nus[2 * t + 1] + av[2 * jj + 1].x * S.nus[8 + 2 * t] +
                           av[2 * jj + 1].y * S.nus[8 + 2 * t + 1];
                c = sum_over_t(c);
                if (t == 0) S.acol[16 * warp + 8 * jj + g] = c;
            }
        }
        __syncthreads();
        ASUB(4);
        if (tid < NV) {
            double v = 0.0;
            if (do_fwd) v = (S.rowq[tid] + vq) + ((S.gcol[0][tid] + S.gcol[1][tid]) + (S.gcol[2][tid] + S.gcol[3][tid])) + S.acol[tid];
            S.yf[tid] = v;
        } else if (do_fwd) {
            const int i = tid - NV, ar = S.apos[i];
            if (ar >= 0) S.yf[NV + ar] = S.rowg[i] - vh;
            if (i < PE) S.yf[NV + ma + i] = ((S.arow[0][i] + S.arow[1][i]) + (S.arow[2][i] + S.arow[3][i])) - vb;
        }
        __syncthreads();
        PROF(0);

        // ---- blocked LDL' (block 8) with both right-hand sides riding along, then the backward substitution (qp_sqd_dev.cuh)
        factor<NTZ>(T, V, nt, np, NTZ, tid, lane, warp, g, t, fo SQD_SUB_ARG);
        __syncthreads();
        PROF(1);
        backward(T, V, nt, lane, warp, do_rev ? (const char*)G : nullptr, 256);
        __syncthreads();
        PROF(3);
        // ---- outputs (dz, dlam, dnu) = -x; inactive inequalities recovered from their singleton columns
        const bool failed = S.fail != 0;
        {   // pull the next instance's Q (lower-triangle lines) and G towards L2 while this one is written out
            const int64_t nxt = inst + gridDim.x;
            if (nxt < a.B) {
                const size_t nm = (a.shared & 1) ? 0 : (size_t)nxt;
                const char* qn = (const char*)(a.Q + nm * NV * NV);
                const char* gn = (const char*)(a.G + nm * MI * NV);
                for (int l = tid; l < 256; l += THREADS) {
                    if (16 * (l & 3) + 15 >= ((l >> 2) & ~7)) asm volatile("prefetch.global.L2 [%0];" ::"l"(qn + (size_t)l * 128));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)l * 128));
                }
                // ... and its vectors (z, lam, nu, h, seed, dq, dh, db): they gate the first barriers of the assembly
                if (tid < 32) {
                    const int v = tid >> 2, ln = tid & 3;  // 8 vectors x up to 4 lines of 128 B (plus one for misalignment)
                    const double* base = nullptr;
                    int len = 0;
                    switch (v) {
                        case 0: base = a.z + (size_t)nxt * NV; len = NV; break;
                        case 1: base = a.lam + (size_t)nxt * MI; len = MI; break;
                        case 2: base = a.h + (size_t)nxt * MI; len = MI; break;
                        case 3: base = a.nu + (size_t)nxt * PE; len = PE; break;
                        case 4: base = do_rev ? a.seed + (size_t)nxt * NV : nullptr; len = NV; break;
                        case 5: base = do_fwd && a.dq ? a.dq + (size_t)nxt * NV : nullptr; len = NV; break;
                        case 6: base = do_fwd && a.dh ? a.dh + (size_t)nxt * MI : nullptr; len = MI; break;
                        default: base = do_fwd && a.db ? a.db + (size_t)nxt * PE : nullptr; len = PE; break;
                    }
                    if (base && ln * 16 < len) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)base + (size_t)ln * 128));
                }
            }
        }
        if (!failed) {
            double* rev = do_rev ? a.rev + b * N : nullptr;
            double* fwd = do_fwd ? a.fwd + b * N : nullptr;
            if (do_rev) {
                // inactive rows: out_lam_i = (G_i . x_z) / D_i ; active rows: -w_i / lam_i.  Warp w: tile rows w, w+4
                load_row8(G, MI, 8 * warp + 2 * t, g, bufA);
                load_row8(G, MI, 8 * (warp + 4) + 2 * t, g, bufB);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    double2(&gv)[8] = hh ? bufB : bufA;
                    const int r0 = 8 * (warp + 4 * hh) + 2 * t;
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
                        const double o0 = a0 >= 0 ? -S.yb[NV + a0] * fast_rcp(S.lams[r0]) : d0 * fast_rcp(S.dvec[r0]);
                        const double o1 = a1 >= 0 ? -S.yb[NV + a1] * fast_rcp(S.lams[r0 + 1]) : d1 * fast_rcp(S.dvec[r0 + 1]);
                        *reinterpret_cast<double2*>(rev + NV + r0) = make_double2(o0, o1);
                    }
                }
            }
            if (warp >= 2) {
                const int u = tid - 64;
                if (do_rev) {
                    rev[u] = -S.yb[u];
                    if (u < PE) rev[NV + MI + u] = -S.yb[NV + ma + u];
                }
            } else {
                if (do_fwd) {
                    fwd[tid] = -S.yf[tid];
                    const int ar = S.apos[tid];
                    fwd[NV + tid] = ar >= 0 ? -S.yf[NV + ar] : 0.0;
                    if (tid < PE) fwd[NV + MI + tid] = -S.yf[NV + ma + tid];
                }
                if (a.info && tid == 0) a.info[inst] = 0;
            }
        } else if (tid == 0) {
            fb_list[atomicAdd(fb_count, 1)] = (int)inst;
        }
        __syncthreads();
        PROF(4);
    }
#ifdef QP_PROFILE
    if (a.prof && blockIdx.x == 0 && tid == 0) {
        for (int i = 0; i < 6; ++i) a.prof[i] = pc[i];
        for (int i = 0; i < 4; ++i) a.prof[8 + i] = sub[i];
        a.prof[17] = sub[4];
        a.prof[18] = sub[5];
        for (int i = 0; i < 5; ++i) a.prof[20 + i] = asub[i];
    }
    if (a.prof && blockIdx.x == 0 && tid == 32)
        for (int i = 0; i < 5; ++i) a.prof[12 + i] = sub[i];
#endif
}

}  // namespace

// Launches the LDL' fast path followed by the pivoted-LU kernel over the instances it rejected (device-side
// list, no host round trip).  nt_cap = tile order of the largest reduced system of the batch.
int32_t qp_sqd_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, int nt_cap, bool* handled, int* max_active, int max_tag) {
    *handled = false;
    const size_t smem = sizeof(Hdr) + (size_t)(nt_cap * (nt_cap + 1) / 2) * 64 * sizeof(double);
    if (smem > ctx->smem_optin) return 0;
    DO_CUDA(ctx, ctx->qp_fb.reserve(sizeof(int) * ((size_t)a.B + 1)));
    int* fb_count = ctx->qp_fb.as<int>();
    int* fb_list = fb_count + 1;
    DO_CUDA(ctx, cudaMemsetAsync(fb_count, 0, sizeof(int), ctx->stream));
    int per_sm = 1;
    DO_CUDA(ctx, kernel_config((const void*)qp_kkt_sqd_kernel, ctx->device, THREADS, smem, &per_sm));
    if (per_sm < 1) per_sm = 1;
    if (const char* cap = getenv("DIFFOPT_B200_SQD_PER_SM")) per_sm = atoi(cap) < per_sm && atoi(cap) > 0 ? atoi(cap) : per_sm;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    QpSolveArgs aa = a;
    const bool profile = getenv("DIFFOPT_B200_PROFILE") != nullptr;
    long long* dprof = nullptr;
    if (profile) {
        DO_CUDA(ctx, cudaMalloc(&dprof, 32 * sizeof(long long)));
        DO_CUDA(ctx, cudaMemsetAsync(dprof, 0, 32 * sizeof(long long), ctx->stream));
        aa.prof = dprof;
    }
    qp_kkt_sqd_kernel<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(aa, fb_list, fb_count, nt_cap, max_active, max_tag);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    if (profile) {
        long long h[32];
        int nfb = 0;
        DO_CUDA(ctx, cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaMemcpyAsync(&nfb, fb_count, sizeof nfb, cudaMemcpyDeviceToHost, ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dprof);
        long long ninst = (a.B + grid - 1) / grid;
        fprintf(stderr,
                "[qp_sqd profile, CTA 0, %lld instances, %d CTA/SM, nt_cap %d, smem %zu, %d to LU] clocks/instance: "
                "assemble %lld factor %lld backward %lld output %lld | warp0: wait F %lld, panel row %lld, fence %lld, arrive %lld, tile update %lld, elim %lld | "
                "warp1: wait E %lld, panel rows %lld, wait B %lld, rhs update %lld, pairs %lld | assemble: to sync1 %lld, to sync2 %lld, G %lld, A %lld, "
                "fwd rhs %lld\n",
                ninst, per_sm, nt_cap, smem, nfb, h[0] / ninst, h[1] / ninst, h[3] / ninst, h[4] / ninst, h[8] / ninst, h[17] / ninst, h[18] / ninst, h[9] / ninst,
                h[10] / ninst, h[11] / ninst, h[12] / ninst, h[13] / ninst, h[14] / ninst, h[15] / ninst, h[16] / ninst, h[20] / ninst, h[21] / ninst, h[22] / ninst,
                h[23] / ninst, h[24] / ninst);
    }
    *handled = true;
    return qp_lu_launch_list(ctx, a, nt_cap, fb_list, fb_count);
}

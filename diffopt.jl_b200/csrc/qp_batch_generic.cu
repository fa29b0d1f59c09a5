// Generic batched KKT sensitivity kernel: any (n, m, p).  One CTA per QP instance; the augmented KKT matrix lives in the
// CTA's shared memory when it fits (N = n+m+p <= 165) and in a per-CTA slice of a global scratch buffer otherwise (BIG:
// the same algorithm with strided loops -- correct for any N, slow, the safety net behind the LDL' fast path for larger
// problems):
//
//   assemble  LHS = [Q G'diag(lam) A'; G diag(Gz-h) 0; A 0 0]   (QuadraticProgram.jl:256-282)
//   augment   column N = reverse RHS [dl_dz;0;0]                 (:324-329)
//             row    N = forward RHS' (:429-433)
//   LU with partial (row) pivoting in shared memory; the extra column/row receive the
//   L- and U'-forward substitutions for free; one backward sweep finishes both
//   LHS x_b = r_b  and  LHS' x_f = r_f  from the SAME factorisation (:335, :438).
//
// This is the shape-generic correctness path (ragged shapes, no-G / no-A problems).  The
// n=64,m=64,p=16 headline shape is served by the tuned kernel in qp_batch_n144.cu.
#include "common.cuh"

namespace {

constexpr int GEN_THREADS = 512;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// accumulate, for one column-major R x n matrix X (global), the products
//   rowacc[r] += sum_j X[r,j] * zc[j]      (length R)
//   colacc[j] += sum_r X[r,j] * wr[r]      (length n)        (wr == nullptr -> skipped)
// rowscale (optional) multiplies the row sums' contribution: rowacc[r] += rowscale[r] * (...)
// Every output has ONE owner and a fixed summation order (results are reproducible bit for bit): rows are owned by the
// lanes of a warp (32 consecutive rows per warp, columns in order, four interleaved partial sums), columns by whole warps
// (lanes over the rows, fixed shuffle tree).  Called by all threads; the caller separates calls that share an
// accumulator with __syncthreads().
__device__ void accum_matvecs(const double* __restrict__ X, int R, int n, const double* zc, const double* wr,
                              double* rowacc, const double* rowscale, double* colacc) {
    if (X == nullptr || R == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = GEN_THREADS / 32;
    if (rowacc != nullptr) {
        for (int r0 = 32 * warp; r0 < R; r0 += 32 * NW) {
            const int r = r0 + lane;
            if (r >= R) continue;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int j = 0;
            for (; j + 3 < n; j += 4) {
                s0 = fma(X[(size_t)j * R + r], zc[j], s0);
                s1 = fma(X[(size_t)(j + 1) * R + r], zc[j + 1], s1);
                s2 = fma(X[(size_t)(j + 2) * R + r], zc[j + 2], s2);
                s3 = fma(X[(size_t)(j + 3) * R + r], zc[j + 3], s3);
            }
            for (; j < n; ++j) s0 = fma(X[(size_t)j * R + r], zc[j], s0);
            const double racc = (s0 + s1) + (s2 + s3);
            rowacc[r] += rowscale ? rowscale[r] * racc : racc;
        }
    }
    if (colacc != nullptr && wr != nullptr) {
        for (int j = warp; j < n; j += NW) {
            double c = 0.0;
            for (int r = lane; r < R; r += 32) c = fma(X[(size_t)j * R + r], wr[r], c);
            c = warp_sum(c);
            if (lane == 0) colacc[j] += c;
        }
    }
}

// list / count: optional device-side list of the instances to solve (the rejects of the LDL' fast path)
template <bool BIG>
__global__ void __launch_bounds__(GEN_THREADS, 1) qp_kkt_generic_kernel(QpSolveArgs a, const int* __restrict__ list, const int* __restrict__ count,
                                                                         double* __restrict__ gscratch) {
    extern __shared__ double smem[];
    const int n = a.n, m = a.m, p = a.p;
    const int N = n + m + p;
    const int ld = (N + 1) | 1;  // odd leading dimension: conflict-free row AND column access
    double* K = BIG ? gscratch + (size_t)blockIdx.x * ld * (N + 1) : smem;  // (N+1) x (N+1) augmented, column-major, ld
    double* zs = BIG ? smem : K + (size_t)ld * (N + 1);                       // n
    double* lams = zs + n;                  // m
    double* nus = lams + m;                 // p
    double* rf = nus + p;                   // N   forward RHS accumulator
    double* scal = rf + N;                  // [0]=rinv [1]=old_kk [2]=pivot value
    int* perm = reinterpret_cast<int*>(scal + 4);  // N
    int* ipiv = perm + N + 1;                      // [0] = pivot row of this step, [1] = info

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    const int64_t todo = list ? (int64_t)*count : a.B;
    for (int64_t it = blockIdx.x; it < todo; it += gridDim.x) {
        const int64_t inst = list ? (int64_t)list[it] : it;
        const size_t bm = (a.shared & 1) ? 0 : (size_t)inst;   // shared Q, G, A: one instance serves the batch
        const double* Q = a.Q + bm * n * n;
        const double* G = a.G ? a.G + bm * m * n : nullptr;
        const double* A = a.A ? a.A + bm * p * n : nullptr;
        const bool do_fwd = a.fwd != nullptr, do_rev = a.rev != nullptr;

        // ---- load vectors, clear K
        for (int i = tid; i < n; i += GEN_THREADS) zs[i] = a.z[(size_t)inst * n + i];
        for (int i = tid; i < m; i += GEN_THREADS) lams[i] = a.lam[(size_t)inst * m + i];
        for (int i = tid; i < p; i += GEN_THREADS) nus[i] = a.nu[(size_t)inst * p + i];
        for (int i = tid; i < N; i += GEN_THREADS) {
            rf[i] = 0.0;
            perm[i] = i;
        }
        for (int i = tid; i < ld * (N + 1); i += GEN_THREADS) K[i] = 0.0;  // (ld (N+1) < 2^31 for any N the scratch can hold)
        if (tid == 0) ipiv[1] = 0;
        __syncthreads();

        // ---- assemble
        for (int idx = tid; idx < n * n; idx += GEN_THREADS) {
            int i = idx % n, j = idx / n;
            K[i + j * ld] = Q[idx];
        }
        for (int idx = tid; idx < m * n; idx += GEN_THREADS) {
            int i = idx % m, j = idx / m;
            double v = G[idx];
            K[(n + i) + j * ld] = v;
            K[j + (n + i) * ld] = v * lams[i];
        }
        for (int idx = tid; idx < p * n; idx += GEN_THREADS) {
            int i = idx % p, j = idx / p;
            double v = A[idx];
            K[(n + m + i) + j * ld] = v;
            K[j + (n + m + i) * ld] = v;
        }
        // forward RHS (QuadraticProgram.jl:429-433) accumulated into rf via global reads
        if (do_fwd && a.rhs_pre) {  // assembled from sparse triplets (dq, dh, db already folded in)
            for (int i = tid; i < N; i += GEN_THREADS) {
                const double v = a.rhs_pre[(size_t)inst * N + i];
                rf[i] = (i >= n && i < n + m) ? lams[i - n] * v : v;
            }
        } else if (do_fwd) {
            const size_t b = (size_t)inst;
            const size_t bd = (a.shared & 2) ? 0 : b;
            accum_matvecs(a.dQ ? a.dQ + bd * n * n : nullptr, n, n, zs, nullptr, rf, nullptr, nullptr);
            __syncthreads();   // rf[0..n) changes owner between the calls
            accum_matvecs(a.dG ? a.dG + bd * m * n : nullptr, m, n, zs, lams, rf + n, lams, rf);
            __syncthreads();
            accum_matvecs(a.dA ? a.dA + bd * p * n : nullptr, p, n, zs, nus, rf + n + m, nullptr, rf);
        }
        __syncthreads();
        // diag(Gz - h); finish rf; place the two right-hand sides
        for (int i = tid; i < m; i += GEN_THREADS) {
            double d = 0.0;
            for (int j = 0; j < n; ++j) d += K[(n + i) + j * ld] * zs[j];
            K[(n + i) + (n + i) * ld] = d - a.h[(size_t)inst * m + i];
        }
        if (do_fwd) {
            for (int i = tid; i < N; i += GEN_THREADS) {
                double v = rf[i];
                if (a.rhs_pre) {
                } else if (i < n) {
                    if (a.dq) v += a.dq[(size_t)inst * n + i];
                } else if (i < n + m) {
                    if (a.dh) v -= lams[i - n] * a.dh[(size_t)inst * m + (i - n)];
                } else {
                    if (a.db) v -= a.db[(size_t)inst * p + (i - n - m)];
                }
                K[N + i * ld] = v;  // row N
            }
        }
        if (do_rev) {
            for (int i = tid; i < n; i += GEN_THREADS) K[i + N * ld] = a.seed[(size_t)inst * n + i];
        }
        __syncthreads();

        // ---- LU with partial pivoting (rows 0..N-1 eligible; row N and column N ride along)
        for (int k = 0; k < N; ++k) {
            if (warp == 0) {
                double best = -1.0;
                int bi = k;
                for (int i = k + lane; i < N; i += 32) {
                    double v = fabs(K[i + k * ld]);
                    if (v > best) {
                        best = v;
                        bi = i;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) {
                        best = ov;
                        bi = oi;
                    }
                }
                if (lane == 0) {
                    double pv = K[bi + k * ld];
                    ipiv[0] = bi;
                    scal[1] = K[k + k * ld];
                    scal[2] = pv;
                    if (best == 0.0 || !(best == best)) {
                        if (ipiv[1] == 0) ipiv[1] = k + 1;
                        scal[0] = 0.0;
                    } else {
                        scal[0] = 1.0 / pv;
                    }
                    int t = perm[k];
                    perm[k] = perm[bi];
                    perm[bi] = t;
                }
            }
            __syncthreads();
            const int pr = ipiv[0];
            const double rinv = scal[0];
            // swap rows k <-> pr in every column but k; scale column k (rows k+1..N)
            if (BIG) {
                if (pr != k)
                    for (int j = tid; j <= N; j += GEN_THREADS)
                        if (j != k) {
                            double t = K[k + (size_t)j * ld];
                            K[k + (size_t)j * ld] = K[pr + (size_t)j * ld];
                            K[pr + (size_t)j * ld] = t;
                        }
                for (int i = k + 1 + tid; i <= N; i += GEN_THREADS) {
                    double src = (i == pr) ? scal[1] : K[i + (size_t)k * ld];
                    K[i + (size_t)k * ld] = src * rinv;
                }
                if (tid == 0) K[k + (size_t)k * ld] = scal[2];
            } else if (tid <= N) {
                int j = tid;
                if (j != k && pr != k) {
                    double t = K[k + j * ld];
                    K[k + j * ld] = K[pr + j * ld];
                    K[pr + j * ld] = t;
                }
            } else if (tid >= 192 && tid < 192 + (N - k)) {
                int i = k + 1 + (tid - 192);
                double src = (i == pr) ? scal[1] : K[i + k * ld];
                K[i + k * ld] = src * rinv;
            } else if (tid == 191) {
                K[k + k * ld] = scal[2];
            }
            __syncthreads();
            // trailing update including the augmented row/column
            if (BIG) {
                for (int j = k + 1 + warp; j <= N; j += GEN_THREADS / 32) {
                    const double u = K[k + (size_t)j * ld];
                    for (int i = k + 1 + lane; i <= N; i += 32) K[i + (size_t)j * ld] -= K[i + (size_t)k * ld] * u;
                }
            } else {
                double l[6];
                const int i0 = k + 1 + lane;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    int i = i0 + 32 * q;
                    l[q] = (i <= N) ? K[i + k * ld] : 0.0;
                }
                for (int j = k + 1 + warp; j <= N; j += GEN_THREADS / 32) {
                    double u = K[k + j * ld];
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        int i = i0 + 32 * q;
                        if (i <= N) K[i + j * ld] -= l[q] * u;
                    }
                }
            }
            __syncthreads();
        }

        // ---- backward sweeps:  U x_b = y (column N)   and   L' v = w (row N)
        for (int k = N - 1; k >= 0; --k) {
            if (BIG) {
                if (do_rev) {
                    const double xb = K[k + (size_t)N * ld] / K[k + (size_t)k * ld];
                    for (int i = tid; i < k; i += GEN_THREADS) K[i + (size_t)N * ld] -= K[i + (size_t)k * ld] * xb;
                    if (tid == 0) rf[k] = xb;
                }
                if (do_fwd) {
                    const double vk = K[N + (size_t)k * ld];
                    for (int j = tid; j < k; j += GEN_THREADS) K[N + (size_t)j * ld] -= K[k + (size_t)j * ld] * vk;
                }
            } else if (tid < 256) {
                if (do_rev) {
                    double xb = K[k + N * ld] / K[k + k * ld];
                    if (tid < k) K[tid + N * ld] -= K[tid + k * ld] * xb;
                    if (tid == k) rf[k] = xb;  // rf reused as x_b store
                }
            } else {
                int j = tid - 256;
                if (do_fwd && j < k) K[N + j * ld] -= K[k + j * ld] * K[N + k * ld];
            }
            __syncthreads();
        }
        // ---- outputs: (dz, dlam, dnu) = -x ; x_f = P' v
        if (do_rev)
            for (int i = tid; i < N; i += GEN_THREADS) a.rev[(size_t)inst * N + i] = -rf[i];
        if (do_fwd)
            for (int i = tid; i < N; i += GEN_THREADS) a.fwd[(size_t)inst * N + perm[i]] = -K[N + i * ld];
        if (a.info && tid == 0) a.info[inst] = ipiv[1];
        if (tid == 0 && ipiv[1] != 0) qp_report_sticky(a, inst);
        __syncthreads();
    }
}

__global__ void qp_param_grads_kernel(int64_t B, int n, int m, int p, const double* __restrict__ z,
                                      const double* __restrict__ lam, const double* __restrict__ nu,
                                      const double* __restrict__ rev, int reduce, double* dQ, double* dq,
                                      double* dG, double* dh, double* dA, double* db) {
    // one thread per output element of one instance; loops over the batch when reducing
    const int N = n + m + p;
    const int64_t per = (int64_t)n * n + n + (int64_t)m * n + m + (int64_t)p * n + p;
    const int64_t total = reduce ? per : per * B;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t e = t % per;
        int64_t b0 = reduce ? 0 : t / per, b1 = reduce ? B : b0 + 1;
        double acc = 0.0;
        int kind;
        int64_t i = 0, j = 0;
        if (e < (int64_t)n * n) { kind = 0; i = e % n; j = e / n; }
        else if ((e -= (int64_t)n * n) < n) { kind = 1; i = e; }
        else if ((e -= n) < (int64_t)m * n) { kind = 2; i = e % m; j = e / m; }
        else if ((e -= (int64_t)m * n) < m) { kind = 3; i = e; }
        else if ((e -= m) < (int64_t)p * n) { kind = 4; i = e % p; j = e / p; }
        else { e -= (int64_t)p * n; kind = 5; i = e; }
        for (int64_t b = b0; b < b1; ++b) {
            const double* zb = z + b * n;
            const double* r = rev + b * N;
            switch (kind) {
                case 0: acc += 0.5 * (r[i] * zb[j] + zb[i] * r[j]); break;
                case 1: acc += r[i]; break;
                case 2: { double l = lam[b * m + i]; acc += l * r[n + i] * zb[j] + l * r[j]; } break;
                case 3: acc += -lam[b * m + i] * r[n + i]; break;
                case 4: acc += r[n + m + i] * zb[j] + nu[b * p + i] * r[j]; break;
                default: acc += -r[n + m + i]; break;
            }
        }
        double* dst = nullptr;
        int64_t bo = reduce ? 0 : b0;
        switch (kind) {
            case 0: if (dQ) dst = dQ + bo * n * n + i + j * n; break;
            case 1: if (dq) dst = dq + bo * n + i; break;
            case 2: if (dG) dst = dG + bo * m * n + i + j * m; break;
            case 3: if (dh) dst = dh + bo * m + i; break;
            case 4: if (dA) dst = dA + bo * p * n + i + j * p; break;
            default: if (db) dst = db + bo * p + i; break;
        }
        if (dst) *dst = acc;
    }
}

}  // namespace

size_t qp_generic_smem_bytes(int n, int m, int p) {
    int N = n + m + p;
    int ld = (N + 1) | 1;
    size_t d = (size_t)ld * (N + 1) + n + m + p + N + 4;
    return d * sizeof(double) + (size_t)(N + 1 + 4) * sizeof(int);
}

static int32_t generic_launch(diffopt_b200_ctx* ctx, const QpSolveArgs& a, const int* list, const int* count) {
    size_t smem = qp_generic_smem_bytes(a.n, a.m, a.p);
    const int N = a.n + a.m + a.p;
    const bool big = smem > ctx->smem_optin;
    int64_t grid = a.B < (int64_t)ctx->sm_count * 4 ? a.B : (int64_t)ctx->sm_count * 4;
    if (grid < 1) grid = 1;
    double* scratch = nullptr;
    if (big) {  // the matrix of every resident CTA in a global scratch buffer (<= 1 GiB), the vectors stay in shared memory
        const size_t ld = (size_t)((N + 1) | 1), per = ld * (size_t)(N + 1) * sizeof(double);
        smem = ((size_t)a.n + a.m + a.p + N + 4) * sizeof(double) + (size_t)(N + 1 + 4) * sizeof(int);
        if (smem > ctx->smem_optin || per > ((size_t)1 << 30) || ld * (size_t)(N + 1) >= ((size_t)1 << 31)) {
            char buf[160];
            snprintf(buf, sizeof buf, "qp_batch: N = n+m+p = %d is beyond the batched kernels (use diffopt_b200_sparse_setup / kkt_solve_csc)", N);
            ctx->err = buf;
            return -3;
        }
        const int64_t fit = (int64_t)(((size_t)1 << 30) / per);
        if (grid > fit) grid = fit < 1 ? 1 : fit;
        if (grid > (int64_t)ctx->sm_count * 2) grid = (int64_t)ctx->sm_count * 2;
        DO_CUDA(ctx, ctx->qp_scratch.reserve(per * (size_t)grid));
        scratch = ctx->qp_scratch.as<double>();
        DO_CUDA(ctx, cudaFuncSetAttribute(qp_kkt_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        qp_kkt_generic_kernel<true><<<(unsigned)grid, GEN_THREADS, smem, ctx->stream>>>(a, list, count, scratch);
    } else {
        DO_CUDA(ctx, cudaFuncSetAttribute(qp_kkt_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        qp_kkt_generic_kernel<false><<<(unsigned)grid, GEN_THREADS, smem, ctx->stream>>>(a, list, count, nullptr);
    }
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    return 0;
}

int32_t qp_batch_launch_generic(diffopt_b200_ctx* ctx, const QpSolveArgs& a) { return generic_launch(ctx, a, nullptr, nullptr); }

// the generic pivoted-LU kernel over a device-side list of instances (fallback of the shape-generic LDL' fast path)
int32_t qp_generic_launch_list(diffopt_b200_ctx* ctx, const QpSolveArgs& a, const int* list, const int* count) {
    return generic_launch(ctx, a, list, count);
}

int32_t qp_param_grads_launch(diffopt_b200_ctx* ctx, int64_t B, int n, int m, int p, const double* z,
                              const double* lam, const double* nu, const double* rev, int reduce,
                              double* dQ, double* dq, double* dG, double* dh, double* dA, double* db) {
    int64_t per = (int64_t)n * n + n + (int64_t)m * n + m + (int64_t)p * n + p;
    int64_t total = reduce ? per : per * B;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    qp_param_grads_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(B, n, m, p, z, lam, nu, rev, reduce, dQ, dq,
                                                                      dG, dh, dA, db);
    ctx->launches++;
    DO_CUDA(ctx, cudaGetLastError());
    return 0;
}

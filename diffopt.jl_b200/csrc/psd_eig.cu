// Symmetric eigendecomposition of the PSD-cone blocks of v = y - s (what `LinearAlgebra.eigen` does inside
// MathOptSetDistances' `projection_gradient_on_set(::PositiveSemidefiniteConeTriangle)`, reached from
// src/diff_opt.jl:509-519), followed by the quantities the operator-form Dpi needs: eigenvectors U, the
// matrix B of SURVEY.md C3 and pi(v) = vec(U max(L,0) U').
//
// Method: one-sided (Hestenes) Jacobi on G = (X + sigma I) V, sigma = |X|_F, V = I at the start.  The shift makes
// the matrix positive semidefinite, so right singular vectors are eigenvectors (without it a +l / -l eigenvalue
// pair shares one singular subspace).  Only columns are ever rotated, so every access is unit stride.
//   * d <= 111: the whole problem (G and V, 16 d^2 bytes) lives in the shared memory of one CTA; many cones per
//     launch, one warp per column pair, one __syncthreads per tournament round.
//   * larger d: block Jacobi.  Columns are cut into nblk blocks; in each outer round a CTA stages one block pair
//     from L2 into shared memory, orthogonalises it (all pairs in the first round of a sweep, cross pairs only
//     afterwards, so each column pair meets once per sweep) and writes it back; grid.sync() between outer rounds
//     (cooperative launch, nblk/2 CTAs).
// Eigenvalues are Rayleigh quotients v_i'g_i - sigma.  Eigenpair order is irrelevant to Dpi (B is built from the same
// order).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lsqr.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr double PSD_EIG_THRESHOLD = 1e-4;  // MathOptSetDistances' `l < 1e-4`
constexpr int PSD_MAX_SWEEPS = 30;

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) x += __shfl_xor_sync(0xffffffffu, x, q);
    return x;
}

__device__ __forceinline__ double fast_rcp(double x) {  // 1 / x, |relative error| ~ 1 ulp, x normal
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ double fast_rsqrt(double x) {  // 1 / sqrt(x), |relative error| ~ 1 ulp, x normal
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);                 // 1 - x y^2
    y = fma(y, e * fma(0.375, e, 0.5), y);          // third-order step
    e = fma(-x * y, y, 1.0);
    return fma(y, 0.5 * e, y);
}

// Orthogonalises columns (gp, gq) of G and applies the same rotation to (vp, vq) of V.  One warp.
__device__ __forceinline__ int rotate_pair(double* __restrict__ gp, double* __restrict__ gq, double* __restrict__ vp,
                                           double* __restrict__ vq, int d, int lane, double tol) {
    double a = 0.0, b = 0.0, g = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double x = gp[i], y = gq[i];
        a = fma(x, x, a);
        b = fma(y, y, b);
        g = fma(x, y, g);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    g = warp_sum(g);
    if (!(g * g > tol * tol * a * b)) return 0;
    // t = tan(theta) = 2 g sign(d) / (|d| + sqrt(d^2 + 4 g^2)), d = b - a; c = 1 / sqrt(1 + t^2), s = c t.
    // The scalar chain is the critical path of a tournament round (one warp retires a dependent instruction every ~10
    // clk): square root and reciprocals are MUFU seeds + Newton steps (~60 clk each) instead of the IEEE-rounded
    // library sequences (~250 clk each); c^2 + s^2 = 1 holds to an ulp or two, which is all Jacobi needs.
    const double dba = b - a, g2 = g + g;
    const double h2 = fma(dba, dba, g2 * g2);
    double t, c;
    if (h2 > 1e-280 && h2 < 1e280) {
        const double den = fabs(dba) + h2 * fast_rsqrt(h2);
        t = (dba >= 0.0 ? g2 : -g2) * fast_rcp(den);
        c = fast_rsqrt(fma(t, t, 1.0));
    } else {  // out of the range the flush-to-zero seeds cover
        const double zeta = dba / g2;
        t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
        c = rsqrt(fma(t, t, 1.0));
    }
    const double s = c * t;
    for (int i = lane; i < d; i += 32) {
        const double x = gp[i], y = gq[i];
        gp[i] = c * x - s * y;
        gq[i] = s * x + c * y;
        const double p = vp[i], q = vq[i];
        vp[i] = c * p - s * q;
        vq[i] = s * p + c * q;
    }
    return 1;
}

// Round-robin tournament ("circle method") over `np` players (np even): the k-th pair of round r.
__device__ __forceinline__ void circle_pair(int np, int r, int k, int& p, int& q) {
    const int m1 = np - 1;
    if (k == 0) {
        p = m1;
        q = r % m1;
    } else {
        p = (r + k) % m1;
        q = (r + m1 - k) % m1;
    }
}

// |X|_F of each cone from its triangle (off-diagonal entries count twice).
__global__ void psd_sigma_kernel(const int* __restrict__ poff, const int* __restrict__ pd, const double* __restrict__ v,
                                 double* __restrict__ sigma) {
    const int c = blockIdx.x, d = pd[c], off = poff[c];
    __shared__ double red[32];
    double acc = 0.0;
    for (int col = threadIdx.x >> 5; col < d; col += blockDim.x >> 5) {
        const long long base = off + (long long)col * (col + 1) / 2;
        for (int r = threadIdx.x & 31; r <= col; r += 32) {
            const double a = v[base + r];
            acc += (r == col ? 1.0 : 2.0) * a * a;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        sigma[c] = sqrt(s);
    }
}

// G = unvec(v) + sigma I, V = I   (grid.y = cone)
__global__ void psd_init_kernel(const int* __restrict__ poff, const int* __restrict__ pd,
                                const long long* __restrict__ uoff, const double* __restrict__ v,
                                const double* __restrict__ sigma, double* __restrict__ Gall, double* __restrict__ Vall) {
    const int c = blockIdx.y, d = pd[c], off = poff[c];
    const double sg = sigma[c];
    double* G = Gall + uoff[c];
    double* V = Vall + uoff[c];
    const long long dd = (long long)d * d;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < dd; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % d), j = (int)(e / d);
        const int r = i < j ? i : j, cc = i < j ? j : i;
        G[e] = v[off + (long long)cc * (cc + 1) / 2 + r] + (i == j ? sg : 0.0);
        V[e] = i == j ? 1.0 : 0.0;
    }
}

// Small cones: one CTA per cone, everything in shared memory.
__global__ void __launch_bounds__(1024) psd_jacobi_small_kernel(const int* __restrict__ list, const int* __restrict__ pd,
                                                                const long long* __restrict__ uoff, double* Gall,
                                                                double* Vall, int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int c = list[blockIdx.x], d = pd[c];
    const int np = d + (d & 1);  // players, padded to even
    double* Gs = sm;
    double* Vs = sm + (size_t)d * d;
    double* G = Gall + uoff[c];
    double* V = Vall + uoff[c];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int e = tid; e < d * d; e += nt) {
        Gs[e] = G[e];
        Vs[e] = V[e];
    }
    __shared__ int s_rot;
    if (tid == 0) s_rot = 0;
    __syncthreads();
    const double tol = 4.4e-16 * sqrt((double)d);
    int sweep = 0;
    for (; sweep < PSD_MAX_SWEEPS; ++sweep) {
        int rot = 0;
        for (int r = 0; r < np - 1; ++r) {
            for (int k = warp; k < np / 2; k += nw) {
                int p, q;
                circle_pair(np, r, k, p, q);
                if (p < d && q < d)
                    rot += rotate_pair(Gs + (size_t)p * d, Gs + (size_t)q * d, Vs + (size_t)p * d, Vs + (size_t)q * d, d,
                                       lane, tol);
            }
            __syncthreads();
        }
        if (lane == 0 && rot) atomicAdd(&s_rot, rot);
        __syncthreads();
        const int total = s_rot;
        __syncthreads();
        if (tid == 0) s_rot = 0;
        if (total == 0) break;
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += nt) {
        G[e] = Gs[e];
        V[e] = Vs[e];
    }
    if (tid == 0 && sweeps_out) sweeps_out[c] = sweep;
}

// Large cone: block Jacobi, cooperative launch with nblk/2 CTAs.  Block j holds columns [j*b, (j+1)*b).
__global__ void __launch_bounds__(1024) psd_jacobi_block_kernel(int d, int b, int nblk, double* G, double* V,
                                                                int* counters /* PSD_MAX_SWEEPS+1, zeroed */,
                                                                int* sweeps_out) {
    extern __shared__ double sm[];
    cg::grid_group grid = cg::this_grid();
    double* Gs = sm;                       // 2b columns of length d
    double* Vs = sm + (size_t)2 * b * d;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const double tol = 4.4e-16 * sqrt((double)d);
    int sweep = 0;
    for (; sweep < PSD_MAX_SWEEPS; ++sweep) {
        int rot = 0;
        for (int r = 0; r < nblk - 1; ++r) {
            int bi, bj;
            circle_pair(nblk, r, blockIdx.x, bi, bj);
            // stage the two blocks (columns beyond d do not exist)
            const int ci = bi * b, cj = bj * b;
            // a warp per column, eight independent loads per lane and matrix in flight (the staging was the longest
            // part of an outer round when every element waited for its own L2 round trip)
            for (int lc = warp; lc < 2 * b; lc += nw) {
                const int gc = lc < b ? ci + lc : cj + lc - b;
                if (gc >= d) continue;
                const double* gsrc = G + (size_t)gc * d;
                const double* vsrc = V + (size_t)gc * d;
                double* gdst = Gs + (size_t)lc * d;
                double* vdst = Vs + (size_t)lc * d;
                for (int i0 = lane; i0 < d; i0 += 256) {
                    double tg[8], tv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + 32 * u;
                        tg[u] = i < d ? gsrc[i] : 0.0;
                        tv[u] = i < d ? vsrc[i] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + 32 * u;
                        if (i < d) {
                            gdst[i] = tg[u];
                            vdst[i] = tv[u];
                        }
                    }
                }
            }
            __syncthreads();
            if (r == 0) {  // all pairs of the 2b local columns
                for (int ir = 0; ir < 2 * b - 1; ++ir) {
                    for (int k = warp; k < b; k += nw) {
                        int p, q;
                        circle_pair(2 * b, ir, k, p, q);
                        const int gp = p < b ? ci + p : cj + p - b, gq = q < b ? ci + q : cj + q - b;
                        if (gp < d && gq < d)
                            rot += rotate_pair(Gs + (size_t)p * d, Gs + (size_t)q * d, Vs + (size_t)p * d,
                                               Vs + (size_t)q * d, d, lane, tol);
                    }
                    __syncthreads();
                }
            } else {  // cross pairs only
                for (int ir = 0; ir < b; ++ir) {
                    for (int k = warp; k < b; k += nw) {
                        const int p = k, q = b + (k + ir) % b;
                        if (ci + p < d && cj + q - b < d)
                            rot += rotate_pair(Gs + (size_t)p * d, Gs + (size_t)q * d, Vs + (size_t)p * d,
                                               Vs + (size_t)q * d, d, lane, tol);
                    }
                    __syncthreads();
                }
            }
            for (int lc = warp; lc < 2 * b; lc += nw) {
                const int gc = lc < b ? ci + lc : cj + lc - b;
                if (gc >= d) continue;
                double* gdst = G + (size_t)gc * d;
                double* vdst = V + (size_t)gc * d;
                const double* gsrc = Gs + (size_t)lc * d;
                const double* vsrc = Vs + (size_t)lc * d;
                for (int i = lane; i < d; i += 32) {
                    gdst[i] = gsrc[i];
                    vdst[i] = vsrc[i];
                }
            }
            if (r == nblk - 2 && lane == 0 && rot) atomicAdd(&counters[sweep], rot);
            grid.sync();
        }
        if (*(volatile int*)&counters[sweep] == 0) break;
    }
    if (blockIdx.x == 0 && tid == 0 && sweeps_out) *sweeps_out = sweep;
}

// lam_i = v_i' g_i - sigma   (one warp per column; grid.y = cone)
__global__ void psd_lambda_kernel(const int* __restrict__ pd, const long long* __restrict__ uoff,
                                  const long long* __restrict__ loff, const double* __restrict__ Gall,
                                  const double* __restrict__ Vall, const double* __restrict__ sigma,
                                  double* __restrict__ lam) {
    const int c = blockIdx.y, d = pd[c];
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= d) return;
    const double* g = Gall + uoff[c] + (size_t)col * d;
    const double* w = Vall + uoff[c] + (size_t)col * d;
    double acc = 0.0;
    for (int i = lane; i < d; i += 32) acc = fma(w[i], g[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) lam[loff[c] + col] = acc - sigma[c];
}

// B, the identity flag and pi(v) = vec(U max(L,0) U')   (grid.y = cone)
__global__ void psd_finish_kernel(const int* __restrict__ poff, const int* __restrict__ pd,
                                  const long long* __restrict__ uoff, const long long* __restrict__ loff,
                                  const double* __restrict__ lam, const double* __restrict__ Vall,
                                  double* __restrict__ Bm, int* __restrict__ ident, double* __restrict__ vp, const int vp_max_d) {
    extern __shared__ double ls[];  // eigenvalues of this cone
    const int c = blockIdx.y, d = pd[c], off = poff[c];
    const double* V = Vall + uoff[c];
    double* Bc = Bm + uoff[c];
    for (int i = threadIdx.x; i < d; i += blockDim.x) ls[i] = lam[loff[c] + i];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int allpos = 1;
        for (int i = 0; i < d; ++i)
            if (!(ls[i] >= 0.0)) allpos = 0;
        ident[c] = allpos;
    }
    const long long dd = (long long)d * d;
    const long long stride = (long long)gridDim.x * blockDim.x, first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long e = first; e < dd; e += stride) {
        const int i = (int)(e % d), j = (int)(e / d);
        const double li = ls[i], lj = ls[j];
        const bool ni = li < PSD_EIG_THRESHOLD, nj = lj < PSD_EIG_THRESHOLD;
        double bv;
        if (!ni && !nj) bv = 1.0;
        else if (ni && nj) bv = 0.0;
        else {
            const double lp = ni ? fmax(lj, 0.0) : fmax(li, 0.0);    // positive side
            const double lm = ni ? -fmin(li, 0.0) : -fmin(lj, 0.0);  // negative side
            bv = lp / (lm + lp);
        }
        Bc[e] = bv;
    }
    if (d > vp_max_d) return;  // larger cones: pi(v) by the tiled product of psd_projection_launch
    const long long tri = (long long)d * (d + 1) / 2;
    for (long long e = first; e < tri; e += stride) {
        int cc = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
        while ((long long)(cc + 1) * (cc + 2) / 2 <= e) ++cc;
        while ((long long)cc * (cc + 1) / 2 > e) --cc;
        const int r = (int)(e - (long long)cc * (cc + 1) / 2);
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            const double l = ls[k];
            if (l > 0.0) acc = fma(V[r + (size_t)k * d] * l, V[cc + (size_t)k * d], acc);
        }
        vp[off + e] = acc;
    }
}

}  // namespace

// Host driver, called from diffopt_b200_conic_setup after v = y - s is on the device.
bool psd_tridiag_supported(diffopt_b200_ctx* ctx, int d);
int32_t psd_tridiag_eig_launch(diffopt_b200_ctx* ctx, int d, const double* xtri, double* w0, double* w1, double* w2, double* small,
                               double* U, double* lam);
int32_t psd_projection_launch(diffopt_b200_ctx* ctx, int d, const double* U, const double* lam, double* vp);

// Which eigensolver takes a cone of side d: 0 one-CTA Jacobi (d <= 111), 1 tridiagonalisation + bisection + inverse iteration
// (psd_tridiag.cu), 2 block Jacobi over a cooperative grid.  DIFFOPT_B200_PSD=jacobi keeps every larger cone on the block Jacobi.
static int psd_method(diffopt_b200_ctx* ctx, int d, size_t smem_cap) {
    if ((size_t)16 * d * d <= smem_cap) return 0;
    const char* force = getenv("DIFFOPT_B200_PSD");
    if (!(force && strcmp(force, "jacobi") == 0) && psd_tridiag_supported(ctx, d)) return 1;
    return 2;
}

int32_t psd_eig_launch(diffopt_b200_ctx* ctx, const std::vector<int>& h_d, const std::vector<long long>& h_uoff,
                       const std::vector<int>& h_off) {
    ConicState& S = ctx->conic;
    const int npsd = (int)h_d.size();
    if (npsd == 0) return 0;
    // per-cone eigenvalue offsets, list of small cones, scratch: sigma[npsd], counters
    std::vector<long long> loff((size_t)npsd);
    std::vector<int> small;
    long long sumd = 0;
    int maxd = 0, maxd_small = 0;
    const size_t smem_cap = ctx->smem_optin > 4096 ? ctx->smem_optin - 2048 : 0;
    for (int c = 0; c < npsd; ++c) {
        loff[(size_t)c] = sumd;
        sumd += h_d[(size_t)c];
        maxd = h_d[(size_t)c] > maxd ? h_d[(size_t)c] : maxd;
        if ((size_t)16 * h_d[(size_t)c] * h_d[(size_t)c] <= smem_cap) {
            small.push_back(c);
            maxd_small = h_d[(size_t)c] > maxd_small ? h_d[(size_t)c] : maxd_small;
        }
    }
    DO_CUDA(ctx, S.psd_lam.reserve(sizeof(double) * (size_t)(sumd + npsd + 1)));
    DO_CUDA(ctx, S.psd_loff.reserve(sizeof(long long) * (size_t)npsd + sizeof(int) * (small.size() + 1)));
    DO_CUDA(ctx, cudaMemcpyAsync(S.psd_loff.ptr, loff.data(), sizeof(long long) * (size_t)npsd, cudaMemcpyHostToDevice,
                                 ctx->stream));
    int* d_small = reinterpret_cast<int*>(S.psd_loff.as<long long>() + npsd);
    if (!small.empty())
        DO_CUDA(ctx, cudaMemcpyAsync(d_small, small.data(), sizeof(int) * small.size(), cudaMemcpyHostToDevice,
                                     ctx->stream));
    int nlarge = 0, ntri = 0, maxd_tri = 0;
    for (int c = 0; c < npsd; ++c) {
        const int meth = psd_method(ctx, h_d[(size_t)c], smem_cap);
        nlarge += meth == 2;
        ntri += meth == 1;
        if (meth == 1 && h_d[(size_t)c] > maxd_tri) maxd_tri = h_d[(size_t)c];
    }
    DO_CUDA(ctx, ctx->in[14].reserve(sizeof(int) * (size_t)(nlarge + 1) * (PSD_MAX_SWEEPS + 2)));
    DO_CUDA(ctx, cudaMemsetAsync(ctx->in[14].ptr, 0, sizeof(int) * (size_t)(nlarge + 1) * (PSD_MAX_SWEEPS + 2),
                                 ctx->stream));
    double* lam = S.psd_lam.as<double>();
    double* sigma = lam + sumd;
    double* G = S.psd_work.as<double>();  // first of the three d x d scratch matrices per cone
    double* V = S.psd_U.as<double>();
    const int* poff = S.psd_off.as<int>();
    const int* pd = S.psd_d.as<int>();
    const long long* uoff = S.psd_uoff.as<long long>();

    const bool any_jacobi = !small.empty() || nlarge > 0;  // the Jacobi kernels work on G = (X + sigma I) V, V = I
    // cones of the one-CTA Jacobi form pi(v) inside psd_finish_kernel, larger ones by the tiled product
    const int vp_max_d = (int)floor(sqrt((double)smem_cap / 16.0));
    if (any_jacobi) {
        psd_sigma_kernel<<<npsd, 256, 0, ctx->stream>>>(poff, pd, S.v.as<double>(), sigma);
        int chunks = (int)(((long long)maxd * maxd + 255) / 256);
        if (chunks > 64) chunks = 64;
        psd_init_kernel<<<dim3((unsigned)chunks, (unsigned)npsd), 256, 0, ctx->stream>>>(poff, pd, uoff, S.v.as<double>(),
                                                                                          sigma, G, V);
        ctx->launches += 2;
    }
    if (!small.empty()) {
        const size_t smem = (size_t)16 * maxd_small * maxd_small;
        int warps = (maxd_small + 1) / 2;
        warps = warps < 4 ? 4 : (warps > 32 ? 32 : warps);
        DO_CUDA(ctx, cudaFuncSetAttribute(psd_jacobi_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        psd_jacobi_small_kernel<<<(unsigned)small.size(), warps * 32, smem, ctx->stream>>>(d_small, pd, uoff, G, V,
                                                                                            nullptr);
        ctx->launches++;
    }
    int li = 0;
    for (int c = 0; c < npsd; ++c) {
        const int d = h_d[(size_t)c];
        if (psd_method(ctx, d, smem_cap) != 2) continue;
        int b = (int)(smem_cap / ((size_t)32 * d));  // 2 matrices x 2b columns x d doubles
        if (b > 32) b = 32;
        if (b > d / 20) b = d / 20 < 4 ? 4 : d / 20;  // measured on B200 (d = 200): 10 CTAs x 10-column blocks beat 4 x 25
        if (const char* e = getenv("DIFFOPT_B200_PSD_BLOCK")) {  // tuning knob: block width of the block Jacobi
            const int w = atoi(e);
            if (w >= 1 && (size_t)32 * w * d <= smem_cap && w <= 32) b = w;
        }
        if (b < 1) BAD_ARG(ctx, "conic_setup: PSD side too large for the block Jacobi eigensolver");
        int nblk = (d + b - 1) / b;
        nblk += nblk & 1;
        b = (d + nblk - 1) / nblk;  // even out the blocks
        const int ctas = nblk / 2;
        if (ctas > ctx->sm_count) BAD_ARG(ctx, "conic_setup: PSD side too large for the block Jacobi eigensolver");
        const size_t smem = (size_t)32 * b * d;
        int warps = b < 4 ? 4 : b;
        DO_CUDA(ctx, cudaFuncSetAttribute(psd_jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        int dd = d, bb = b, nb = nblk;
        double* Gc = G + h_uoff[(size_t)c];
        double* Vc = V + h_uoff[(size_t)c];
        int* counters = ctx->in[14].as<int>() + (size_t)li * (PSD_MAX_SWEEPS + 2);
        int* sweeps_out = counters + PSD_MAX_SWEEPS + 1;
        void* args[] = {&dd, &bb, &nb, &Gc, &Vc, &counters, &sweeps_out};
        DO_CUDA(ctx, cudaLaunchCooperativeKernel((void*)psd_jacobi_block_kernel, dim3((unsigned)ctas),
                                                 dim3((unsigned)warps * 32), args, smem, ctx->stream));
        ctx->launches++;
        ++li;
    }
    if (any_jacobi) {
        psd_lambda_kernel<<<dim3((unsigned)((maxd + 7) / 8), (unsigned)npsd), 256, 0, ctx->stream>>>(
            pd, uoff, S.psd_loff.as<long long>(), G, V, sigma, lam);
        ctx->launches++;
    }
    if (ntri > 0) {  // after the Rayleigh quotients above (which cover every cone): these cones get their own lam and U
        DO_CUDA(ctx, S.psd_tri.reserve(sizeof(double) * (size_t)(5 * maxd_tri + 8)));
        double* work = S.psd_work.as<double>();
        for (int c = 0; c < npsd; ++c) {
            const int d = h_d[(size_t)c];
            if (psd_method(ctx, d, smem_cap) != 1) continue;
            const long long uo = h_uoff[(size_t)c];
            if (int32_t rc = psd_tridiag_eig_launch(ctx, d, S.v.as<double>() + h_off[(size_t)c], work + uo, work + S.psd_sumd2 + uo,
                                                    work + 2 * S.psd_sumd2 + uo, S.psd_tri.as<double>(), V + uo, lam + loff[(size_t)c]))
                return rc;
        }
    }
    int fchunks = (int)(((long long)maxd * maxd + 255) / 256);
    if (fchunks > 2 * ctx->sm_count) fchunks = 2 * ctx->sm_count;
    psd_finish_kernel<<<dim3((unsigned)fchunks, (unsigned)npsd), 256, sizeof(double) * (size_t)maxd, ctx->stream>>>(
        poff, pd, uoff, S.psd_loff.as<long long>(), lam, V, S.psd_Bm.as<double>(), S.psd_ident.as<int>(),
        S.vp.as<double>(), vp_max_d);
    ctx->launches += 1;
    for (int c = 0; c < npsd; ++c)
        if (h_d[(size_t)c] > vp_max_d)
            if (int32_t rc = psd_projection_launch(ctx, h_d[(size_t)c], V + h_uoff[(size_t)c], lam + loff[(size_t)c],
                                                   S.vp.as<double>() + h_off[(size_t)c]))
                return rc;
    if (getenv("DIFFOPT_B200_PSD_DEBUG") && nlarge > 0) {  // sweeps and rotations per sweep of the block Jacobi
        std::vector<int> h((size_t)nlarge * (PSD_MAX_SWEEPS + 2));
        DO_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->in[14].ptr, sizeof(int) * h.size(), cudaMemcpyDeviceToHost,
                                     ctx->stream));
        DO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int l = 0; l < nlarge; ++l) {
            const int* r = h.data() + (size_t)l * (PSD_MAX_SWEEPS + 2);
            fprintf(stderr, "psd block jacobi #%d: sweeps=%d rotations:", l, r[PSD_MAX_SWEEPS + 1]);
            for (int k = 0; k <= r[PSD_MAX_SWEEPS + 1] && k < PSD_MAX_SWEEPS; ++k) fprintf(stderr, " %d", r[k]);
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

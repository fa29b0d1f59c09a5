// Microbenchmark of 8x8 LDL' diagonal-block variants (one warp, tile in shared memory).
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#define FULL 0xffffffffu
constexpr double PIV_RTOL = 1e-12;
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
// ---- V0: lane 0 eliminates, lanes 0..7 invert
__device__ __noinline__ void v0(double* D, double* ref, double* rd, const bool positive, int* fail, const int lane, long long* tmid) {
    if (lane == 0) {
        double a[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; c += 2)
                if (c <= i) { const double2 v = *reinterpret_cast<const double2*>(&D[i * 8 + c]); a[i][c] = v.x; a[i][c + 1] = v.y; }
        double thr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) thr[k] = PIV_RTOL * ref[k];
        bool bad = false;
        double d = a[0][0];
        if (!((positive ? d : -d) > thr[0])) { bad = true; d = positive ? 1.0 : -1.0; }
        double r = fast_rcp(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            rd[k] = r;
            double rn = 0.0;
            if (k < 7) {
                double dn = fma(-(a[k + 1][k] * a[k + 1][k]), r, a[k + 1][k + 1]);
                if (!((positive ? dn : -dn) > thr[k + 1])) { bad = true; dn = positive ? 1.0 : -1.0; }
                rn = fast_rcp(dn);
            }
            double lk[8];
#pragma unroll
            for (int i = k + 1; i < 8; ++i) lk[i] = a[i][k] * r;
#pragma unroll
            for (int i = k + 2; i < 8; ++i) a[i][k + 1] = fma(-lk[i], a[k + 1][k], a[i][k + 1]);
#pragma unroll
            for (int c = k + 2; c < 8; ++c)
#pragma unroll
                for (int i = c; i < 8; ++i) a[i][c] = fma(-lk[i], a[c][k], a[i][c]);
#pragma unroll
            for (int i = k + 1; i < 8; ++i) D[i * 8 + k] = lk[i];
            r = rn;
        }
        if (bad) *fail = 1;
    }
    __syncwarp();
    if (tmid) *tmid = clock64();
    double x[8];
    if (lane < 8) {
        double l[8][8];
#pragma unroll
        for (int i = 1; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < i; ++c) l[i][c] = D[i * 8 + c];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int i = k + 1; i < 8; ++i) x[i] = fma(-l[i][k], x[k], x[i]);
    }
    __syncwarp();
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) D[i * 8 + lane] = x[i];
    }
}
// ---- V1: lane i owns row i (lanes 0..7); shuffles broadcast the pivot and the pivot column.
// Also builds inv(L) rows on the fly: lane i keeps row i of inv(L) (forward substitution with the multipliers).
__device__ __noinline__ void v1(double* D, double* ref, double* rd, const bool positive, int* fail, const int lane, long long* tmid) {
    const int i = lane & 7;
    double a[8];
    {
        const double2* row = reinterpret_cast<const double2*>(&D[i * 8]);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const double2 v = row[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
    }
    const double thr = PIV_RTOL * ref[i];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // column k entries of the rows below (unnormalised), gathered before the reciprocal is known
        double tc[8];
#pragma unroll
        for (int c = k + 1; c < 8; ++c) tc[c] = __shfl_sync(FULL, a[k], c);
        double d = __shfl_sync(FULL, a[k], k);
        const double th = __shfl_sync(FULL, thr, k);
        if (!((positive ? d : -d) > th)) { bad = true; d = positive ? 1.0 : -1.0; }
        const double r = fast_rcp(d);
        if (lane == k) rd[k] = r;
        const double l = a[k] * r;
#pragma unroll
        for (int c = k + 1; c < 8; ++c) a[c] = fma(-l, tc[c], a[c]);
        a[k] = l;
    }
    if (bad && lane == 0) *fail = 1;
    if (tmid) *tmid = clock64();
    // inverse of unit lower L (rows in lanes): X = L^-1, row i: x_i = e_i - sum_{k<i} l_ik x_k  (row vectors)
    // lane i needs rows k < i of X: sequential over k with shuffles
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = (c == i) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        // row k of X is final once rows < k were folded in; broadcast it and fold into rows i > k
#pragma unroll
        for (int c = 0; c <= k; ++c) {
            const double xk = __shfl_sync(FULL, x[c], k);
            if (i > k) x[c] = fma(-a[k], xk, x[c]);
        }
    }
    if (lane < 8) {
        double2* row = reinterpret_cast<double2*>(&D[i * 8]);
#pragma unroll
        for (int q = 0; q < 4; ++q) row[q] = make_double2(x[2 * q], x[2 * q + 1]);
    }
    __syncwarp();
}
// ---- V2: lane 0 only; inverse of L built alongside the elimination; pivot checks after the chain
__device__ __noinline__ void v2(double* D, double* ref, double* rd, const bool positive, int* fail, const int lane, long long* tmid) {
    if (lane == 0) {
        double a[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; c += 2)
                if (c <= i) { const double2 v = *reinterpret_cast<const double2*>(&D[i * 8 + c]); a[i][c] = v.x; a[i][c + 1] = v.y; }
        double x[8][8];
        double r = fast_rcp(a[0][0]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            rd[k] = r;
            double rn = 0.0;
            if (k < 7) {
                const double dn = fma(-(a[k + 1][k] * a[k + 1][k]), r, a[k + 1][k + 1]);
                a[k + 1][k + 1] = dn;
                rn = fast_rcp(dn);
            }
            double lk[8];
#pragma unroll
            for (int i = k + 1; i < 8; ++i) lk[i] = a[i][k] * r;
#pragma unroll
            for (int i = k + 2; i < 8; ++i) a[i][k + 1] = fma(-lk[i], a[k + 1][k], a[i][k + 1]);
#pragma unroll
            for (int c = k + 2; c < 8; ++c)
#pragma unroll
                for (int i = (c == k + 1 ? c + 1 : c); i < 8; ++i)
                    if (!(i == c && c == k + 1)) a[i][c] = fma(-lk[i], a[c][k], a[i][c]);
#pragma unroll
            for (int i = k + 1; i < 8; ++i) {
#pragma unroll
                for (int c = 0; c < k; ++c) x[i][c] = fma(-lk[i], x[k][c], x[i][c]);
                x[i][k] = -lk[i];
            }
            r = rn;
        }
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) ok = ok && ((positive ? a[k][k] : -a[k][k]) > PIV_RTOL * ref[k]);
        if (!ok) *fail = 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double v0 = c < i ? x[i][c] : (c == i ? 1.0 : 0.0), v1 = c + 1 < i ? x[i][c + 1] : (c + 1 == i ? 1.0 : 0.0);
                *reinterpret_cast<double2*>(&D[i * 8 + c]) = make_double2(v0, v1);
            }
        }
    }
    __syncwarp();
    if (tmid) *tmid = clock64();
}
// independent-DFMA issue rate of one warp
__global__ void fma_rate(double* out, long long* t, double s) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = s + i + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 256; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], s, 1.0);
    }
    const long long t1 = clock64();
    double a = 0; for (int i = 0; i < 8; ++i) a += x[i];
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) t[0] = t1 - t0;
}
__device__ __forceinline__ void panel_row(double* p, const double* Ld, const double* scale) {
    double w[8];
    {
        const double2* pp = reinterpret_cast<const double2*>(p);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const double2 v = pp[q]; w[2 * q] = v.x; w[2 * q + 1] = v.y; }
    }
#pragma unroll
    for (int c = 1; c < 8; ++c) {
        double l[8];
#pragma unroll
        for (int k = 0; k < c; k += 2) { const double2 v = *reinterpret_cast<const double2*>(&Ld[c * 8 + k]); l[k] = v.x; l[k + 1] = v.y; }
        double acc = w[c], acc2 = 0.0;
#pragma unroll
        for (int k = 0; k < c; ++k) { if (k & 1) acc2 = fma(-w[k], l[k], acc2); else acc = fma(-w[k], l[k], acc); }
        w[c] = acc + acc2;
    }
    if (scale) {
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] *= scale[c];
    }
    double2* pp = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int q = 0; q < 4; ++q) pp[q] = make_double2(w[2 * q], w[2 * q + 1]);
}
__global__ void panel_bench(const double* M, double* out, long long* t, int reps, int nl) {
    __shared__ __align__(16) double D[64], W[32 * 8];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64; i += 32) D[i] = M[i] * 0.1;
    for (int i = lane; i < 256; i += 32) W[i] = M[i & 63];
    __syncwarp();
    long long tot = 0;
    for (int rep = 0; rep < reps; ++rep) {
        const long long t0 = clock64();
        if (lane < nl) panel_row(W + lane * 8, D, nullptr);
        __syncwarp();
        tot += clock64() - t0;
    }
    if (lane == 0) t[0] = tot / reps;
    for (int i = lane; i < 256; i += 32) out[i & 63] = W[i];
}
template <int V>
__global__ void bench(const double* M, double* out, long long* t, int reps) {
    __shared__ __align__(16) double D[64];
    __shared__ double ref[8], rd[8];
    __shared__ int fail;
    const int lane = threadIdx.x;
    long long tot = 0, tot1 = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = lane; i < 64; i += 32) D[i] = M[i] * (1.0 + 1e-9 * rep);
        if (lane < 8) ref[lane] = fabs(M[lane * 9]);
        if (lane == 0) fail = 0;
        __syncwarp();
        long long tm = 0;
        const long long t0 = clock64();
        if (V == 0) v0(D, ref, rd, true, &fail, lane, &tm); else if (V == 1) v1(D, ref, rd, true, &fail, lane, &tm); else v2(D, ref, rd, true, &fail, lane, &tm);
        const long long t1 = clock64();
        tot += t1 - t0; tot1 += tm - t0;
        __syncwarp();
    }
    if (lane == 0) { t[0] = tot / reps; t[1] = tot1 / reps; t[2] = fail; }
    for (int i = lane; i < 64; i += 32) out[i] = D[i];
    if (lane < 8) out[64 + lane] = rd[lane];
}
int main() {
    double h[64], L[64] = {0};
    // SPD test matrix
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 8; ++k) s += sin(1.0 + i * 8 + k) * sin(1.0 + j * 8 + k); h[i * 8 + j] = s / 8 + (i == j ? 0.5 : 0.0); }
    double *dM, *dout; long long* dt;
    cudaMalloc(&dM, sizeof h); cudaMalloc(&dout, 72 * 8); cudaMallocManaged(&dt, 64);
    cudaMemcpy(dM, h, sizeof h, cudaMemcpyHostToDevice);
    // reference: inverse of unit-lower L of LDL' on the host
    double a[64]; for (int i = 0; i < 64; ++i) a[i] = h[i] * (1.0 + 1e-9 * 999);
    double d[8];
    for (int k = 0; k < 8; ++k) { d[k] = a[k * 9]; for (int i = k + 1; i < 8; ++i) { double l = a[i * 8 + k] / d[k]; for (int c = k + 1; c <= i; ++c) a[i * 8 + c] -= l * a[c * 8 + k]; L[i * 8 + k] = l; } }
    double X[64] = {0};
    for (int c = 0; c < 8; ++c) { double x[8]; for (int i = 0; i < 8; ++i) { double s = (i == c); for (int k = 0; k < i; ++k) s -= L[i * 8 + k] * x[k]; x[i] = (i >= c) ? s : 0; X[i * 8 + c] = x[i]; } }
    for (int v = 0; v < 3; ++v) {
        if (v == 0) bench<0><<<1, 32>>>(dM, dout, dt, 1000); else if (v == 1) bench<1><<<1, 32>>>(dM, dout, dt, 1000); else bench<2><<<1, 32>>>(dM, dout, dt, 1000);
        cudaDeviceSynchronize();
        double o[72]; cudaMemcpy(o, dout, sizeof o, cudaMemcpyDeviceToHost);
        double err = 0, errd = 0;
        for (int i = 0; i < 64; ++i) err = fmax(err, fabs(o[i] - X[i]));
        for (int k = 0; k < 8; ++k) errd = fmax(errd, fabs(o[64 + k] * d[k] - 1.0));
        printf("variant %d: %lld clk total, %lld clk elimination part, fail %lld, max err invL %.2e, rd %.2e (%s)\n", v, dt[0], dt[1], dt[2], err, errd, cudaGetErrorString(cudaGetLastError()));
    }
    for (int nl = 8; nl <= 32; nl *= 4) { panel_bench<<<1, 32>>>(dM, dout, dt, 1000, nl); cudaDeviceSynchronize(); printf("panel_row, %d lanes: %lld clk\n", nl, dt[0]); }
    for (int nl = 1; nl <= 32; nl *= 32) { fma_rate<<<1, nl>>>(dout, dt, 1.0000001); cudaDeviceSynchronize(); printf("independent DFMA, %d active lane(s): %.2f clk per warp instruction\n", nl, (double)dt[0] / (256 * 8)); }
    return 0;
}

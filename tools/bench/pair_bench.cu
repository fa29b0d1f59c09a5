// Microbenchmark of trailing-update code shapes on shared-memory tiles (one or three warps).
#include <cuda_runtime.h>
#include <stdio.h>
#define FULL 0xffffffffu
__host__ __device__ __forceinline__ int tix(int I, int J) { return ((I * (I + 1)) >> 1) + J; }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16(double2& c01, double2& c23, const double2 a_top, const double2 a_bot, const double2 b) {
    // m16n8k8: A regs a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); B b0 (t, g) b1 (t+4, g); C c0,c1 (g; 2t,2t+1) c2,c3 (g+8; ..)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c01.x), "+d"(c01.y), "+d"(c23.x), "+d"(c23.y)
                 : "d"(a_top.x), "d"(a_bot.x), "d"(a_top.y), "d"(a_bot.y), "d"(b.x), "d"(b.y));
}
constexpr int NT = 12;
template <int V>
__global__ void bench(double* out, long long* t, int reps, int j) {
    extern __shared__ __align__(16) double T[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, tt = lane & 3, fo = g * 8 + 2 * tt;
    for (int i = tid; i < tix(NT, 0) * 64; i += blockDim.x) T[i] = 1e-3 * ((i * 37) % 101 - 50);
    __syncthreads();
    const double nr0 = -0.5, nr1 = -0.25;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const int nrows = NT - 1 - j;  // rows I = j+1 .. NT-1 (all pairs K = j+1..I)
        if (V == 0) {
            for (int idx = warp; idx < nrows; idx += nw) {
                const int I = NT - 1 - idx;
                double2 af = *reinterpret_cast<const double2*>(T + tix(I, j) * 64 + fo);
                af.x *= nr0; af.y *= nr1;
                const double* bp = T + tix(j + 1, j) * 64 + fo;
                double* cp = T + tix(I, j + 1) * 64 + fo;
                int K = j + 1;
                for (; K < I; K += 2) {
                    const double2 b0 = *reinterpret_cast<const double2*>(bp);
                    const double2 b1 = *reinterpret_cast<const double2*>(bp + (K + 1) * 64);
                    bp += (2 * K + 3) * 64;
                    double2 ca = *reinterpret_cast<double2*>(cp), cb = *reinterpret_cast<double2*>(cp + 64);
                    dmma(ca.x, ca.y, af.x, b0.x); dmma(cb.x, cb.y, af.x, b1.x);
                    dmma(ca.x, ca.y, af.y, b0.y); dmma(cb.x, cb.y, af.y, b1.y);
                    *reinterpret_cast<double2*>(cp) = ca; *reinterpret_cast<double2*>(cp + 64) = cb;
                    cp += 128;
                }
                if (K == I) {
                    const double2 b0 = *reinterpret_cast<const double2*>(bp);
                    double2 ca = *reinterpret_cast<double2*>(cp);
                    dmma(ca.x, ca.y, af.x, b0.x); dmma(ca.x, ca.y, af.y, b0.y);
                    *reinterpret_cast<double2*>(cp) = ca;
                }
            }
        } else if (V == 1) {
            // column-wise with row pairs on m16n8k8: for column K, rows I >= K in pairs; B fragment reused down the column
            // work item = (K, row pair); dealt to warps round robin over columns
            for (int K = j + 1 + warp; K < NT; K += nw) {
                const double2 bf = *reinterpret_cast<const double2*>(T + tix(K, j) * 64 + fo);
                int I = K;
                if ((NT - K) & 1) {  // odd count: the diagonal tile alone (m8n8k4 x 2)
                    double2 af = bf; af.x *= nr0; af.y *= nr1;
                    double2* cp = reinterpret_cast<double2*>(T + tix(K, K) * 64 + fo);
                    double2 c = *cp;
                    dmma(c.x, c.y, af.x, bf.x); dmma(c.x, c.y, af.y, bf.y);
                    *cp = c;
                    I = K + 1;
                }
                for (; I + 1 < NT; I += 2) {
                    double2 a0 = *reinterpret_cast<const double2*>(T + tix(I, j) * 64 + fo);
                    double2 a1 = *reinterpret_cast<const double2*>(T + tix(I + 1, j) * 64 + fo);
                    a0.x *= nr0; a0.y *= nr1; a1.x *= nr0; a1.y *= nr1;
                    double2* c0p = reinterpret_cast<double2*>(T + tix(I, K) * 64 + fo);
                    double2* c1p = reinterpret_cast<double2*>(T + tix(I + 1, K) * 64 + fo);
                    double2 c0 = *c0p, c1 = *c1p;
                    dmma16(c0, c1, a0, a1, bf);
                    *c0p = c0; *c1p = c1;
                }
            }
        } else {
            // (V == 2 or 3) row-pair oriented m16n8k8: rows (I, I+1) share A fragments (scaled once), loop over K <= I
            for (int idx = warp; 2 * idx < nrows; idx += nw) {
                const int I1 = NT - 1 - 2 * idx, I0 = I1 - 1;  // I0 may be j (invalid) when nrows is odd
                double2 a1 = *reinterpret_cast<const double2*>(T + tix(I1, j) * 64 + fo);
                a1.x *= nr0; a1.y *= nr1;
                if (I0 > j) {
                    double2 a0 = *reinterpret_cast<const double2*>(T + tix(I0, j) * 64 + fo);
                    a0.x *= nr0; a0.y *= nr1;
                    const double* bp = T + tix(j + 1, j) * 64 + fo;
                    double* c0p = T + tix(I0, j + 1) * 64 + fo;
                    double* c1p = T + tix(I1, j + 1) * 64 + fo;
                    int K = j + 1;
                    if (V == 3) {
                        for (; K + 3 <= I0; K += 4) {  // four columns per iteration
                            const double2 b0 = *reinterpret_cast<const double2*>(bp);
                            const double2 b1 = *reinterpret_cast<const double2*>(bp + (K + 1) * 64);
                            const double2 b2 = *reinterpret_cast<const double2*>(bp + (2 * K + 3) * 64);
                            const double2 b3 = *reinterpret_cast<const double2*>(bp + (3 * K + 6) * 64);
                            bp += (4 * K + 10) * 64;
                            double2 c00 = *reinterpret_cast<double2*>(c0p), c10 = *reinterpret_cast<double2*>(c1p);
                            double2 c01 = *reinterpret_cast<double2*>(c0p + 64), c11 = *reinterpret_cast<double2*>(c1p + 64);
                            double2 c02 = *reinterpret_cast<double2*>(c0p + 128), c12 = *reinterpret_cast<double2*>(c1p + 128);
                            double2 c03 = *reinterpret_cast<double2*>(c0p + 192), c13 = *reinterpret_cast<double2*>(c1p + 192);
                            dmma16(c00, c10, a0, a1, b0);
                            dmma16(c01, c11, a0, a1, b1);
                            dmma16(c02, c12, a0, a1, b2);
                            dmma16(c03, c13, a0, a1, b3);
                            *reinterpret_cast<double2*>(c0p) = c00; *reinterpret_cast<double2*>(c1p) = c10;
                            *reinterpret_cast<double2*>(c0p + 64) = c01; *reinterpret_cast<double2*>(c1p + 64) = c11;
                            *reinterpret_cast<double2*>(c0p + 128) = c02; *reinterpret_cast<double2*>(c1p + 128) = c12;
                            *reinterpret_cast<double2*>(c0p + 192) = c03; *reinterpret_cast<double2*>(c1p + 192) = c13;
                            c0p += 256; c1p += 256;
                        }
                    }
                    for (; K + 1 <= I0; K += 2) {  // two columns per iteration
                        const double2 b0 = *reinterpret_cast<const double2*>(bp);
                        const double2 b1 = *reinterpret_cast<const double2*>(bp + (K + 1) * 64);
                        bp += (2 * K + 3) * 64;
                        double2 c00 = *reinterpret_cast<double2*>(c0p), c10 = *reinterpret_cast<double2*>(c1p);
                        double2 c01 = *reinterpret_cast<double2*>(c0p + 64), c11 = *reinterpret_cast<double2*>(c1p + 64);
                        dmma16(c00, c10, a0, a1, b0);
                        dmma16(c01, c11, a0, a1, b1);
                        *reinterpret_cast<double2*>(c0p) = c00; *reinterpret_cast<double2*>(c1p) = c10;
                        *reinterpret_cast<double2*>(c0p + 64) = c01; *reinterpret_cast<double2*>(c1p + 64) = c11;
                        c0p += 128; c1p += 128;
                    }
                    if (K == I0) {
                        const double2 b0 = *reinterpret_cast<const double2*>(bp);
                        bp += (K + 1) * 64;
                        double2 c00 = *reinterpret_cast<double2*>(c0p), c10 = *reinterpret_cast<double2*>(c1p);
                        dmma16(c00, c10, a0, a1, b0);
                        *reinterpret_cast<double2*>(c0p) = c00; *reinterpret_cast<double2*>(c1p) = c10;
                        c1p += 64;
                    }
                    // tile (I1, I1): b = W(I1, j) itself
                    double2 bf = *reinterpret_cast<const double2*>(T + tix(I1, j) * 64 + fo);
                    double2 c = *reinterpret_cast<double2*>(c1p);
                    dmma(c.x, c.y, a1.x, bf.x); dmma(c.x, c.y, a1.y, bf.y);
                    *reinterpret_cast<double2*>(c1p) = c;
                } else {
                    double2 bf = *reinterpret_cast<const double2*>(T + tix(I1, j) * 64 + fo);
                    double2* cp = reinterpret_cast<double2*>(T + tix(I1, I1) * 64 + fo);
                    double2 c = *cp;
                    dmma(c.x, c.y, a1.x, bf.x); dmma(c.x, c.y, a1.y, bf.y);
                    *cp = c;
                }
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) t[0] = (t1 - t0) / reps;
    double s = 0;
    for (int i = tid; i < tix(NT, 0) * 64; i += blockDim.x) s += T[i];
    out[tid] = s;
}
int main() {
    double* out; long long* t;
    cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&t, 64);
    const int smem = tix(NT, 0) * 64 * 8;
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int j = 0; j <= 6; j += 3)
        for (int nw = 1; nw <= 4; nw += (nw == 1 ? 2 : 1)) {
            const int npairs = (NT - 1 - j) * (NT - j) / 2;
            long long r[4]; double h[4][1];
            for (int v = 0; v < 4; ++v) {
                if (v == 0) bench<0><<<1, 32 * nw, smem>>>(out, t, 200, j);
                if (v == 1) bench<1><<<1, 32 * nw, smem>>>(out, t, 200, j);
                if (v == 2) bench<2><<<1, 32 * nw, smem>>>(out, t, 200, j);
                if (v == 3) bench<3><<<1, 32 * nw, smem>>>(out, t, 200, j);
                cudaDeviceSynchronize();
                r[v] = t[0];
                cudaMemcpy(h[v], out, 8, cudaMemcpyDeviceToHost);
            }
            printf("j=%d warps=%d pairs=%d : m8n8k4 rows %lld clk (%.0f/pair)  m16n8k8 cols %lld clk (%.0f/pair)  m16n8k8 row pairs %lld (%.0f/pair)  x4 cols %lld (%.0f/pair)  chk %.6e %.6e %.6e %.6e %s\n", j, nw,
                   npairs, r[0], (double)r[0] / npairs * nw, r[1], (double)r[1] / npairs * nw, r[2], (double)r[2] / npairs * nw, r[3], (double)r[3] / npairs * nw, h[0][0], h[1][0], h[2][0], h[3][0], cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

// Measures the FP64 denominators the roofline needs on this GPU (MEASURED_PEAKS.json has none):
//   (a) cuBLAS DGEMM 8192^3 (best of 10 and 4 s sustained)  -- "FP64 tensor roofline" per SURVEY 8d
//   (b) raw mma.sync m8n8k4 / m16n8k16 .f64 issue throughput (DMMA pipe)
//   (c) raw DFMA throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu -lcublas
// Prints one JSON line.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dmma884_kernel(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                  "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_best(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    // raw pipes: 8 CTAs x 256 threads per SM? use 4 CTAs of 256 threads per SM (32 warps/SM)
    int iters = 4096;
    dim3 grid(sms * 4), block(256);
    float ms884 = time_best([&] { dmma884_kernel<<<grid, block>>>(out, iters); }, 5);
    double fl884 = (double)grid.x * (block.x / 32) * iters * 8.0 * (2.0 * 8 * 8 * 4);
    float ms16816 = time_best([&] { dmma16816_kernel<<<grid, block>>>(out, iters); }, 5);
    double fl16816 = (double)grid.x * (block.x / 32) * iters * 4.0 * (2.0 * 16 * 8 * 16);
    float msfma = time_best([&] { dfma_kernel<<<grid, block>>>(out, iters); }, 5);
    double flfma = (double)grid.x * block.x * iters * 16.0 * 2.0;
    CK(cudaGetLastError());

    // cuBLAS DGEMM
    const int n = 8192;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n));
    CK(cudaMalloc(&B, sizeof(double) * n * n));
    CK(cudaMalloc(&C, sizeof(double) * n * n));
    CK(cudaMemset(A, 0, sizeof(double) * n * n));
    CK(cudaMemset(B, 0, sizeof(double) * n * n));
    cublasHandle_t h;
    cublasCreate(&h);
    double one = 1.0, zero = 0.0;
    auto gemm = [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); };
    float msg = time_best(gemm, 10);
    double flg = 2.0 * n * (double)n * n;
    // sustained: back to back for ~4 s
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int reps = (int)(4000.0f / msg) + 1;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) gemm();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float mss;
    CK(cudaEventElapsedTime(&mss, e0, e1));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f, "
           "\"dfma_tflops\": %.2f, \"dgemm8192_tflops_burst\": %.2f, \"dgemm8192_tflops_sustained\": %.2f}\n",
           prop.name, sms, fl884 / ms884 * 1e-9, fl16816 / ms16816 * 1e-9, flfma / msfma * 1e-9, flg / msg * 1e-9,
           flg * reps / mss * 1e-9);
    return 0;
}

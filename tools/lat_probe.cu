// Latency probe (single warp): dependent-chain latencies of the instructions the panel factorisation leans on.
#include <cuda_runtime.h>
#include <stdio.h>
#define REP 256
__global__ void probe(double* out, long long* t, double seed) {
    __shared__ double sm[64];
    int lane = threadIdx.x;
    sm[lane] = seed + lane; sm[lane + 32] = seed;
    __syncwarp();
    double x = seed, y = seed * 0.5;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) x = fma(x, y, 1.0);
    t1 = clock64(); t[0] = (t1 - t0);
    // DMMA chain (accumulator dependent)
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    t1 = clock64(); t[1] = (t1 - t0);
    // redux chain
    unsigned k = (unsigned)lane + (unsigned)c0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) k = __reduce_max_sync(0xffffffffu, k + lane) ;
    t1 = clock64(); t[2] = (t1 - t0);
    // shfl chain (32-bit)
    unsigned s = k;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) s = __shfl_sync(0xffffffffu, s + 1, (lane + 1) & 31);
    t1 = clock64(); t[3] = (t1 - t0);
    // ballot+ffs chain
    unsigned b = s;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) b = __ffs(__ballot_sync(0xffffffffu, (b + lane) & 1)) + b;
    t1 = clock64(); t[4] = (t1 - t0);
    // LDS chain (pointer chase)
    int idx = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) idx = ((int)sm[idx & 63]) & 31;
    t1 = clock64(); t[5] = (t1 - t0);
    // drcp chain
    double r = x + 3.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) r = __drcp_rn(r) + 1.5;
    t1 = clock64(); t[6] = (t1 - t0) * 4;
    // STS -> syncwarp -> LDS round trip
    double q = r;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) { if (lane == (i & 31)) sm[32] = q; __syncwarp(); q = sm[32] + 1.0; __syncwarp(); }
    t1 = clock64(); t[7] = (t1 - t0);
    // DMUL chain
    double m = q;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) m = m * 1.0000001;
    t1 = clock64(); t[8] = (t1 - t0);
    // rcp.approx.ftz.f64 + 2 Newton steps chain
    double fr = m + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double rr;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(fr));
        double e = fma(-fr, rr, 1.0);
        rr = fma(rr, e, rr);
        e = fma(-fr, rr, 1.0);
        rr = fma(rr, e, rr);
        fr = rr + 1.5;
    }
    t1 = clock64(); t[10] = (t1 - t0) * 4;
    // bare MUFU.RCP64H chain
    double fq = fr;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double rr;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(fq));
        fq = rr;
    }
    t1 = clock64(); t[11] = (t1 - t0) * 4;
    m += fr + fq;
    // __syncthreads cost with 512 threads is measured separately
    out[lane] = x + c0 + c1 + k + s + b + idx + r + q + m;
}
__global__ void barprobe(long long* t) {
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) t[9] = t1 - t0;
}
int main() {
    double* out; long long* t;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&t, 16 * 8);
    probe<<<1, 32>>>(out, t, 1.000001);
    barprobe<<<1, 512>>>(t);
    cudaDeviceSynchronize();
    const char* names[] = {"DFMA", "DMMA.8x8x4", "REDUX.max", "SHFL", "ballot+ffs(+add)", "LDS chase(+cvt)", "DRCP(+add)", "STS/syncwarp/LDS", "DMUL", "syncthreads(512)", "fast_rcp(+add)", "MUFU.RCP64H"};
    for (int i = 0; i < 12; ++i) printf("%-20s %.1f clk\n", names[i], (double)t[i] / REP);
    return 0;
}

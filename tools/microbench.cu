// Latency micro-benchmarks that drive the BA kernel design (FP64 dependent-op latency, barriers, L2 gathers, grid.sync).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench tools/microbench.cu ; run on the B200.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_dfma(double* out, long long* cyc, double a, double b, int n) {
    double x = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc, float a, float b, int n) {
    float x = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_drcp(double* out, long long* cyc, int n) {
    double x = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = 1.0 / x + 0.5;
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_tput(double* out, long long* cyc, double a, double b, int n) {   // 8 independent chains per thread
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = out[j];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = next[p];
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ldg(const int* next, double* out, long long* cyc, int n) {
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = next[p];
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int n) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 << (i % 5));
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_gridsync(long long* cyc, int n) {
    cg::grid_group g = cg::this_grid();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) g.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sincos(double* out, long long* cyc, int n) {
    double x = out[0] + 0.3;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { double s, c; sincos(x, &s, &c); x = s + c; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* d; long long* c; float* f; int* nxt;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 64); cudaMalloc(&f, 1 << 20); cudaMalloc(&nxt, 4 << 20);
    cudaMemset(d, 0, 1 << 20); cudaMemset(f, 0, 1 << 20);
    int* h = new int[1 << 20];
    for (int i = 0; i < (1 << 20); ++i) h[i] = (int)(((long long)i * 7919 + 104729) & ((1 << 20) - 1));
    cudaMemcpy(nxt, h, 4 << 20, cudaMemcpyHostToDevice);
    long long cy;
    const int N = 4096;
    auto rep = [&](const char* name, int n, double per = 1.0) { cudaDeviceSynchronize(); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("%-34s %8.1f cycles/op\n", name, (double)cy / n / per); };
    for (int threads : {32, 256, 1024}) {
        printf("--- %d threads, 1 CTA\n", threads);
        k_dfma<<<1, threads>>>(d, c, 1.0000001, 1e-9, N); rep("dependent DFMA", N);
        k_ffma<<<1, threads>>>(f, c, 1.0000001f, 1e-9f, N); rep("dependent FFMA", N);
        k_dfma_tput<<<1, threads>>>(d, c, 1.0000001, 1e-9, N); rep("DFMA, 8 indep chains (per op)", N, 8);
        k_drcp<<<1, threads>>>(d, c, N); rep("dependent 1.0/x + add", N);
        k_sincos<<<1, threads>>>(d, c, 512); rep("dependent sincos(double)+add", 512);
        k_lds<<<1, threads>>>(d, c, N); rep("dependent LDS", N);
        k_ldg<<<1, threads>>>(nxt, d, c, N); rep("dependent LDG (4 MB chase, L2)", N);
        k_bar<<<1, threads>>>(d, c, N); rep("__syncthreads", N);
        k_shfl<<<1, threads>>>(d, c, N); rep("shfl_xor + DADD dependent", N);
    }
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int n = 1000;
    void* args[] = {(void*)&c, (void*)&n};
    cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(nsm), dim3(256), args, 0, 0); rep("grid.sync (148 CTAs x 256 thr)", n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(nsm), dim3(256), args, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("grid.sync wall: %.3f us each (%d SMs)\n", ms * 1e3 / n, nsm);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock rate attr %d kHz\n", clk);
    return 0;
}

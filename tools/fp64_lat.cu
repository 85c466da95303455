// Latency microbenchmarks behind the pivot-chain model of the reduced solve (ba.cu): dependent DFMA / DMUL / DADD, LDS.64,
// MUFU.RCP64H-based 1/x, st->ld shared round trip, __syncwarp, bar.sync at several CTA sizes.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_lat tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double seed, int nthreads_active) {
    __shared__ double sh[1024];
    const int tid = threadIdx.x;
    sh[tid] = seed + tid;
    __syncthreads();
    long long t0, t1;
    double x = seed, y = seed * 0.5;
    // dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fma(x, y, y);
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0);
    // dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = x * y;
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0);
    // dependent DADD
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = x + y;
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0);
    // dependent 1/x
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = 1.0 / (x + 3.0);
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0);
    // pointer-chasing LDS.64
    int idx = tid & 31;
    unsigned base; asm("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(base) : "l"(sh));
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + idx * 8)); idx = ((int)v) & 31; }
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0);
    // st -> syncwarp -> ld round trip in shared memory
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + (tid & 31) * 8), "d"(x) : "memory");
        __syncwarp();
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(base + ((tid + 1) & 31) * 8));
        x += 1.0;
    }
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0);
    // CTA barrier
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0);
    // divergent 9-lane update like the pivot warp's (two paths)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        if (tid < 9) {
            double val = sh[tid + 40], a0 = sh[tid], a1 = sh[tid + 1], a2 = sh[tid + 2];
            if (tid < 6) { double b0 = sh[tid + 9], b1 = sh[tid + 10], b2 = sh[tid + 11]; val -= a0 * (x * b0 + y * b1 + b2) + a1 * (y * b0 + b1) + a2 * b0; }
            else val -= a0 * x + a1 * y + a2;
            sh[tid + 40] = val;
        }
        __syncwarp();
        x += sh[40 + (i & 7)];
    }
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0);
    // empty clock pair
    t0 = clock64(); t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0);
    if (x == 12345.678) out[9] = (long long)x;
}
int main() {
    long long* d; cudaMalloc(&d, 128);
    for (int threads : {32, 128, 512}) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, threads>>>(d, 1.0000001, threads);
        long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        printf("threads %3d: DFMA %.1f | DMUL %.1f | DADD %.1f | 1/x(+add) %.1f | LDS chase %.1f | st-syncwarp-ld-add %.1f | bar.sync %.1f | 9-lane divergent update+syncwarp+ld %.1f | clock pair %lld (cycles each)\n",
               threads, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 16.0, h[4] / 32.0, h[5] / 16.0, h[6] / 16.0, h[7] / 16.0, h[8]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

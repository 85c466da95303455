// Standalone timing of the reduced-system solve (ba_chol_solve_smem) on a synthetic banded SPD system of C4's shape.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSE2_SOLVE_STAMPS -I se2lam_b200/csrc -o tools/solve_bench tools/solve_bench.cu
#ifndef SE2_SOLVE_STAMPS
#define SE2_SOLVE_STAMPS 1
#endif
#include "../se2lam_b200/csrc/common.cu"
#include "../se2lam_b200/csrc/ba.cu"
#include "../se2lam_b200/csrc/ba_band.cu"
#include <cstdio>
#include <vector>
__global__ void __launch_bounds__(512) tw_kernel(Dev d) {
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    ldlt_twisted_solve(d, d.S, d.bs, &bar, 0, 1);
}
int main() {
    const int nf = 49, n = 3 * nf, band = 5;
    std::vector<double> S((size_t)n * n, 0.0), bs(n);
    std::vector<int> colmax(n);
    for (int a = 0; a < nf; ++a) for (int r = 0; r < 3; ++r) colmax[3 * a + r] = 3 * std::min(a + band, nf - 1) + 2;
    for (int i = 0; i < n; ++i) { bs[i] = 0.01 * (i % 7) - 0.02; for (int j = 0; j <= i; ++j) if (i / 3 - j / 3 <= band) S[(size_t)i * n + j] = (i == j) ? 50.0 + (i % 5) : 0.3 / (1 + i - j); }
    Dev d{};
    d.n = n; d.nf = nf;
    double *dS, *dbs, *ddx; int* dcm; LMState* st;
    cudaMalloc(&dS, sizeof(double) * ((size_t)n * n + n + 16)); cudaMalloc(&ddx, sizeof(double) * n); cudaMalloc(&dcm, sizeof(int) * n); cudaMalloc(&st, sizeof(LMState));
    dbs = dS + (size_t)n * n;
    cudaMemcpy(dS, S.data(), sizeof(double) * S.size(), cudaMemcpyHostToDevice); cudaMemcpy(dbs, bs.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(dcm, colmax.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
    d.S = dS; d.bs = dbs; d.dxp = ddx; d.colmax = dcm; d.st = st;
    const size_t smem = ldlt_smem_bytes(n);
    cudaFuncSetAttribute(ba_chol_solve_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {512, 256, 128}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            ba_chol_solve_smem<<<1, threads, smem>>>(d);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long st_[128]; cudaMemcpyFromSymbol(st_, g_stamps, sizeof st_);
            if (rep == 2) {
                printf("threads %d: kernel %.1f us (%s) | init %lld | first invert %lld | factor %lld | backsub %lld | copy-out %lld cycles\n", threads, ms * 1e3,
                       cudaGetErrorString(cudaGetLastError()), st_[1] - st_[0], st_[2] - st_[1], st_[3] - st_[2], st_[4] - st_[3], st_[5] - st_[4]);
#if SE2_SOLVE_STAMPS >= 2
                long long ws[256]; cudaMemcpyFromSymbol(ws, g_wstamps, sizeof ws);
                {
                    const long long* q = ws;   // warp 0
                    printf("   step 10 pivot warp: ok-flag %lld | W loads + update %lld | syncwarp %lld | re-read + minors %lld | 1/det + W %lld | publish %lld | (end %lld) | barrier %lld | total %lld\n",
                           q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6], q[8] - q[7], q[8] - q[0]);
                    for (int wv : {1, 2, 5, 15}) if (wv * 32 < threads) { q = ws + 16 * wv; printf("   step 10 warp %2d: start skew %lld | ok-flag %lld | row pass %lld | barrier wait %lld\n", wv, q[0] - ws[0], q[1] - q[0], q[7] - q[1], q[8] - q[7]); }
                }
#endif
            }
        }
    }
    {
        // two-sided variant on CTAs 0 / 1
        const int m0 = 22, w = band, m1 = nf - m0 - w, nb1 = nf - m0;
        std::vector<int> cm1(3 * nb1);
        for (int b = 0; b < nb1; ++b) for (int r = 0; r < 3; ++r) cm1[3 * b + r] = 3 * std::min(b + band, nb1 - 1) + 2;
        int* dcm1; double* tb; unsigned* fl;
        cudaMalloc(&dcm1, sizeof(int) * cm1.size()); cudaMalloc(&tb, sizeof(double) * TW_BUF_DOUBLES); cudaMalloc(&fl, 16);
        cudaMemcpy(dcm1, cm1.data(), sizeof(int) * cm1.size(), cudaMemcpyHostToDevice);
        d.tw_m0 = m0; d.tw_w = w; d.tw_cmax1 = dcm1; d.tw_buf = tb; d.tw_flag = fl;
        cudaFuncSetAttribute(tw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(fl, 0, 16); cudaMemset(ddx, 0, sizeof(double) * n);
            cudaEventRecord(e0);
            tw_kernel<<<2, 512, smem>>>(d);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long st_[128]; cudaMemcpyFromSymbol(st_, g_stamps, sizeof st_);
            if (rep == 2) {
                printf("twisted (m0 %d, w %d, m1 %d): kernel %.1f us (%s)\n", m0, w, m1, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
                for (int c = 0; c < 2; ++c) {
                    const long long* q = st_ + 64 * c;
                    printf("   CTA %d: stage %lld | factor own %lld | %s %lld | wait/merge %lld | %s %lld | total %lld cycles\n", c, q[51] - q[50], q[52] - q[51],
                           c ? "publish" : "wait peer", q[53] - q[52], q[54] - q[53], c ? "backsolve" : "sep factor + backsolve", q[55] - q[54], q[55] - q[50]);
                }
                printf("   CTA0 start->CTA1 start skew %lld\n", st_[64 + 50] - st_[50]);
            }
        }
    }
    std::vector<double> x(n); cudaMemcpy(x.data(), ddx, sizeof(double) * n, cudaMemcpyDeviceToHost);
    // residual check
    double worst = 0;
    for (int i = 0; i < n; ++i) { double r = -bs[i]; for (int j = 0; j < n; ++j) r += (j <= i ? S[(size_t)i * n + j] : S[(size_t)j * n + i]) * x[j]; worst = std::max(worst, std::fabs(r)); }
    printf("max residual %.3e\n", worst);
    return 0;
}

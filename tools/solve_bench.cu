// Standalone timing of the reduced-system solve (ba_chol_solve_smem) on a synthetic banded SPD system of C4's shape.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSE2_SOLVE_STAMPS -I se2lam_b200/csrc -o tools/solve_bench tools/solve_bench.cu
#define SE2_SOLVE_STAMPS 1
#include "../se2lam_b200/csrc/common.cu"
#include "../se2lam_b200/csrc/ba.cu"
#include <cstdio>
#include <vector>
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
            long long st_[64]; cudaMemcpyFromSymbol(st_, g_stamps, sizeof st_);
            if (rep == 2) {
                printf("threads %d: kernel %.1f us (%s) | init %lld | first invert %lld | factor %lld | backsub %lld | copy-out %lld cycles\n", threads, ms * 1e3,
                       cudaGetErrorString(cudaGetLastError()), st_[1] - st_[0], st_[2] - st_[1], st_[3] - st_[2], st_[4] - st_[3], st_[5] - st_[4]);
                for (int kb = 0; kb < 8; ++kb)
                    printf("   step %d: read W %lld | barrier X %lld... pivot work %lld | end barrier %lld\n", kb, 0LL, st_[11 + 4 * kb] - st_[10 + 4 * kb],
                           st_[12 + 4 * kb] - st_[11 + 4 * kb], st_[13 + 4 * kb] - st_[12 + 4 * kb]);
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

// Partitioned block-band LDL^T of the reduced pose system (large windows, e.g. BASELINE config 5: 2000 keyframes).
// See ba_band.cu for the algorithm. Internal to libse2gpu.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace se2band {

struct Part {          // one partition: interior block columns [a0, a1), separators on both sides
    int a0, a1;        // interior pose blocks
    int l3, r3;        // scalar width of the left / right separator (0 at the ends)
    long long ws;      // offset of this partition's workspace (doubles)
    long long ws_elems;
};

struct Plan {
    bool active = false;
    int n = 0, nf = 0, w = 0, bw = 0, p = 0, nT = 0, bwT = 0, max_m3 = 0;
    size_t band_elems = 0;               // n * (bw + 1) doubles of band storage for S
    size_t smem_part = 0, smem_sep = 0;
    Part* d_parts = nullptr;
    double* d_work = nullptr;
    int* d_ok = nullptr;                 // [p + 1]: per partition, then the final verdict
    std::vector<Part> parts;
};

// bmax[a] = last block row coupled to block column a (monotone non-decreasing, >= a). Returns true and fills `pl` when the
// envelope is narrow enough for the partitioned solver; false (pl.active = false) otherwise.
bool plan(Plan& pl, int nf, const std::vector<int>& bmax, int smem_optin);
void release(Plan& pl);
// S in band storage (row r holds columns r-bw..r at S[r*(bw+1) + c - r + bw]), bs [n]; dxp [n] receives the solution (zeros
// when the system is not positive definite), *solve_ok 1/0. Three launches on `s`.
int solve(const Plan& pl, const double* Sband, const double* bs, double* dxp, int* solve_ok, cudaStream_t s);

}  // namespace se2band

// ORB matchers on sm_100a — kernels + C ABI (include/se2gpu.h: se2gpu_hamming_distance, se2gpu_match_by_window,
// se2gpu_match_by_projection, se2gpu_search_by_bow).
//
// Replaces se2lam::ORBmatcher (reference src/ORBmatcher.cpp) and the Frame keypoint grid it queries
// (src/Frame.cpp:64-77, 209-286). Split of work:
//   * data-parallel part (one warp per query): which database keypoints fall in the query's grid window
//     (GetFeaturesInArea semantics incl. its cell-range rounding and level filter) and their 256-bit
//     Hamming distances (__popc over 8 x 32 bit == DescriptorDistance :110-126), emitted IN THE
//     REFERENCE'S CANDIDATE ORDER (grid column, grid row, insertion index);
//   * order-dependent part (one warp, queries in sequence): the greedy best / second-best resolution with
//     the "already matched better" skip and steal-back (:308-346, :415-449, :187-246) and the rotation
//     histogram (:350-372). It consumes the precomputed candidate lists, so the sequential loop is short.
#include <climits>
#include <vector>

#include "common.h"

namespace {

using se2gpu::fail;

constexpr int TH_HIGH = 100, TH_LOW = 75, HISTO_LENGTH = 30;   // ORBmatcher.cpp:45-47
constexpr int GRID_ROWS = 48, GRID_COLS = 64;                  // Frame.h:26-27

__device__ __forceinline__ int hamming256(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
    int dsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) dsum += __popc(a[i] ^ b[i]);
    return dsum;
}

__global__ void k_hamming_pairs(const uint32_t* a, const uint32_t* b, int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hamming256(a + 8 * (size_t)i, b + 8 * (size_t)i);
}

// Frame::PosInGrid (Frame.cpp:209-220): cell of every database keypoint, -1 if outside the grid
__global__ void k_grid_cell(const se2gpu_keypoint* kp, int n, se2gpu_grid_params g, int* cell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kp[i].x, g.min_x), g.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kp[i].y, g.min_y), g.inv_h));
    cell[i] = (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
}

// grid order = (cell column, cell row, index): rank by counting (n is ~1e3)
__global__ void k_grid_order(const int* cell, int n, int* order, int* n_valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ci = cell[i];
    if (ci < 0) return;
    int r = 0;
    for (int j = 0; j < n; ++j) { const int cj = cell[j]; r += (cj >= 0) && (cj < ci || (cj == ci && j < i)); }
    order[r] = i;
    atomicAdd(n_valid, 1);
}

// Frame::GetFeaturesInArea (Frame.cpp:222-286) for one query per warp + DescriptorDistance of every hit.
// cand[q*cap + k] = (i2, dist) in the reference's iteration order; ncand[q] = hits.
__global__ void __launch_bounds__(256) k_candidates(const float* __restrict__ qxy, const int* __restrict__ qmin_level,
                                                    const int* __restrict__ qmax_level, const float* __restrict__ qr,
                                                    const uint8_t* __restrict__ qvalid, const uint32_t* __restrict__ qdesc, int nq,
                                                    const se2gpu_keypoint* __restrict__ kp, const uint32_t* __restrict__ desc,
                                                    const int* __restrict__ cell, const int* __restrict__ order,
                                                    const int* __restrict__ n_valid, const uint8_t* __restrict__ db_skip,
                                                    se2gpu_grid_params g, int cap, int2* __restrict__ cand, int* __restrict__ ncand) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    int count = 0;
    if (!qvalid || qvalid[q]) {
        const float x = qxy[2 * q], y = qxy[2 * q + 1], r = qr[q];
        const int minLevel = qmin_level[q], maxLevel = qmax_level[q];
        int x0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.min_x), r), g.inv_w));
        int x1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.min_x), r), g.inv_w));
        int y0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.min_y), r), g.inv_h));
        int y1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.min_y), r), g.inv_h));
        x0 = max(0, x0); x1 = min(GRID_COLS - 1, x1); y0 = max(0, y0); y1 = min(GRID_ROWS - 1, y1);
        const bool empty = x0 >= GRID_COLS || x1 < 0 || y0 >= GRID_ROWS || y1 < 0;
        const bool check = !(minLevel == -1 && maxLevel == -1);
        const int nv = *n_valid;
        if (!empty)
            for (int k0 = 0; k0 < nv; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                int i2 = -1;
                if (k < nv) {
                    i2 = order[k];
                    const int c = cell[i2], cx = c / GRID_ROWS, cy = c - cx * GRID_ROWS;
                    const se2gpu_keypoint p = kp[i2];
                    hit = cx >= x0 && cx <= x1 && cy >= y0 && cy <= y1;
                    if (check) hit = hit && p.octave >= minLevel && p.octave <= maxLevel;
                    hit = hit && !(fabsf(__fsub_rn(p.x, x)) > r || fabsf(__fsub_rn(p.y, y)) > r);
                    if (db_skip) hit = hit && !db_skip[i2];
                }
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const int pos = count + __popc(bal & ((1u << lane) - 1));
                    if (pos < cap) cand[(size_t)q * cap + pos] = make_int2(i2, hamming256(qdesc + 8 * (size_t)q, desc + 8 * (size_t)i2));
                }
                count += __popc(bal);
            }
    }
    if (lane == 0) ncand[q] = min(count, cap);
}

struct Top2 { int d1, p1, d2, p2; };  // two smallest (dist, position) pairs, lexicographic
__device__ __forceinline__ void top2_push(Top2& t, int dist, int pos) {
    if (dist < t.d1 || (dist == t.d1 && pos < t.p1)) { t.d2 = t.d1; t.p2 = t.p1; t.d1 = dist; t.p1 = pos; }
    else if (dist < t.d2 || (dist == t.d2 && pos < t.p2)) { t.d2 = dist; t.p2 = pos; }
}
__device__ __forceinline__ Top2 top2_warp(Top2 t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 u;
        u.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o); u.p1 = __shfl_xor_sync(0xffffffffu, t.p1, o);
        u.d2 = __shfl_xor_sync(0xffffffffu, t.d2, o); u.p2 = __shfl_xor_sync(0xffffffffu, t.p2, o);
        top2_push(t, u.d1, u.p1);
        top2_push(t, u.d2, u.p2);
    }
    return t;
}

__device__ void three_maxima(const int* hist, int& ind1, int& ind2, int& ind3) {  // ORBmatcher.cpp:64-105
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// MatchByWindow's sequential part, one warp. work: [n2] vMatchesDistance, [n2] vnMatches21, [n1] bin_of
__global__ void __launch_bounds__(32) k_resolve_window(const se2gpu_keypoint* __restrict__ kp1, const se2gpu_keypoint* __restrict__ kp2,
                                                       int n1, int n2, int min_level, int max_level, float nnratio, int cap,
                                                       const int2* __restrict__ cand, const int* __restrict__ ncand,
                                                       int* __restrict__ work, int* __restrict__ matches12, float* __restrict__ prev,
                                                       int* __restrict__ nmatches_out) {
    __shared__ int hist[HISTO_LENGTH];
    const int lane = threadIdx.x;
    int* vdist = work; int* m21 = work + n2; int* bin_of = work + 2 * n2;
    for (int i = lane; i < n2; i += 32) { vdist[i] = INT_MAX; m21[i] = -1; }
    for (int i = lane; i < n1; i += 32) { matches12[i] = -1; bin_of[i] = -1; }
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    __syncwarp();
    int nmatches = 0;
    const float factor = (float)HISTO_LENGTH / 360.0f;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = kp1[i1].octave;
        if (level1 > max_level || level1 < min_level) continue;
        const int nc = ncand[i1];
        if (nc == 0) continue;
        Top2 t{INT_MAX, INT_MAX, INT_MAX, INT_MAX};
        for (int k = lane; k < nc; k += 32) {
            const int2 c = cand[(size_t)i1 * cap + k];
            if (vdist[c.x] <= c.y) continue;
            top2_push(t, c.y, k);
        }
        t = top2_warp(t);
        const int bestDist = t.d1, bestDist2 = t.d2;
        if (bestDist <= TH_LOW && (float)bestDist < (float)bestDist2 * nnratio) {   // :329-330 (INT_MAX*ratio stays huge)
            const int bestIdx2 = cand[(size_t)i1 * cap + t.p1].x;
            if (lane == 0) {
                if (m21[bestIdx2] >= 0) { matches12[m21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2; m21[bestIdx2] = i1; vdist[bestIdx2] = bestDist; nmatches++;
                float rot = __fsub_rn(kp1[i1].angle, kp2[bestIdx2].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.f);
                int bin = (int)roundf(__fmul_rn(rot, factor));
                if (bin == HISTO_LENGTH) bin = 0;
                bin_of[i1] = bin; hist[bin]++;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        for (int k = 0; k < n1; ++k) {
            const int b = bin_of[k];
            if (b < 0 || b == i1 || b == i2 || b == i3) continue;
            if (matches12[k] >= 0) { matches12[k] = -1; nmatches--; }
        }
        *nmatches_out = nmatches;
    }
    __syncwarp();
    for (int k = lane; k < n1; k += 32)
        if (matches12[k] >= 0) { prev[2 * k] = kp2[matches12[k]].x; prev[2 * k + 1] = kp2[matches12[k]].y; }
}

// MatchByProjection's sequential part (:415-449), one warp. work: [n] vMatchesDistance
__global__ void __launch_bounds__(32) k_resolve_projection(const se2gpu_keypoint* __restrict__ kp, int n, int nmp, float nnratio, int cap,
                                                           const int2* __restrict__ cand, const int* __restrict__ ncand,
                                                           int* __restrict__ work, int* __restrict__ matches, int* __restrict__ nmatches_out) {
    const int lane = threadIdx.x;
    for (int i = lane; i < n; i += 32) { work[i] = INT_MAX; matches[i] = -1; }
    __syncwarp();
    int nmatches = 0;
    for (int i = 0; i < nmp; ++i) {
        const int nc = ncand[i];
        if (nc == 0) continue;
        Top2 t{INT_MAX, INT_MAX, INT_MAX, INT_MAX};
        for (int k = lane; k < nc; k += 32) {
            const int2 c = cand[(size_t)i * cap + k];
            if (work[c.x] <= c.y) continue;
            top2_push(t, c.y, k);
        }
        t = top2_warp(t);
        if (t.d1 <= TH_HIGH) {
            const int bestIdx = cand[(size_t)i * cap + t.p1].x;
            const int bestLevel = kp[bestIdx].octave;
            const int bestLevel2 = (t.d2 == INT_MAX) ? -1 : kp[cand[(size_t)i * cap + t.p2].x].octave;
            if (bestLevel == bestLevel2 && (float)t.d1 > nnratio * (float)t.d2) continue;
            if (lane == 0) {
                if (matches[bestIdx] >= 0) { matches[bestIdx] = -1; nmatches--; }
                matches[bestIdx] = i; work[bestIdx] = t.d1; nmatches++;
            }
            __syncwarp();
        }
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// SearchByBoW (:128-276), one warp walks the two ascending node lists; lanes share the inner candidate loop
__global__ void __launch_bounds__(32) k_search_by_bow(const float* __restrict__ angle1, const uint32_t* __restrict__ d1, const uint8_t* __restrict__ mp1, int n1,
                                                      const int* __restrict__ node1, int nnode1, const int* __restrict__ ptr1, const int* __restrict__ feat1,
                                                      const float* __restrict__ angle2, const uint32_t* __restrict__ d2, const uint8_t* __restrict__ mp2, int n2,
                                                      const int* __restrict__ node2, int nnode2, const int* __restrict__ ptr2, const int* __restrict__ feat2,
                                                      int mp_only, float nnratio, int check_ori, uint8_t* __restrict__ matched2, int* __restrict__ bin_of,
                                                      int* __restrict__ matches12, int* __restrict__ nmatches_out) {
    __shared__ int hist[HISTO_LENGTH];
    const int lane = threadIdx.x;
    for (int i = lane; i < n1; i += 32) { matches12[i] = -1; bin_of[i] = -1; }
    for (int i = lane; i < n2; i += 32) matched2[i] = 0;
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    __syncwarp();
    const float factor = (float)HISTO_LENGTH / 360.0f;
    int nmatches = 0, a = 0, b = 0;
    while (a < nnode1 && b < nnode2) {
        const int na = node1[a], nb = node2[b];
        if (na == nb) {
            for (int i1 = ptr1[a]; i1 < ptr1[a + 1]; ++i1) {
                const int idx1 = feat1[i1];
                if (mp_only && !mp1[idx1]) continue;
                Top2 t{INT_MAX, INT_MAX, INT_MAX, INT_MAX};
                for (int i2 = ptr2[b] + lane; i2 < ptr2[b + 1]; i2 += 32) {
                    const int idx2 = feat2[i2];
                    if (mp_only && !mp2[idx2]) continue;
                    if (matched2[idx2]) continue;
                    top2_push(t, hamming256(d1 + 8 * (size_t)idx1, d2 + 8 * (size_t)idx2), i2);
                }
                t = top2_warp(t);
                if (t.d1 < TH_LOW && (float)t.d1 < nnratio * (float)t.d2) {
                    const int bestIdx2 = feat2[t.p1];
                    if (lane == 0) {
                        matches12[idx1] = bestIdx2; matched2[bestIdx2] = 1;
                        if (check_ori) {
                            float rot = __fsub_rn(angle1[idx1], angle2[bestIdx2]);
                            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                            int bin = (int)roundf(__fmul_rn(rot, factor));
                            if (bin == HISTO_LENGTH) bin = 0;
                            bin_of[idx1] = bin; hist[bin]++;
                        }
                        nmatches++;
                    }
                    __syncwarp();
                }
            }
            ++a; ++b;
        } else if (na < nb) {
            while (a < nnode1 && node1[a] < nb) ++a;
        } else {
            while (b < nnode2 && node2[b] < na) ++b;
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (check_ori) {
            int i1, i2, i3;
            three_maxima(hist, i1, i2, i3);
            // mapMatches12.erase(idx) + nmatches-- for every entry of the losing bins (:262-272)
            for (int k = 0; k < n1; ++k) {
                const int bb = bin_of[k];
                if (bb < 0 || bb == i1 || bb == i2 || bb == i3) continue;
                matches12[k] = -1; nmatches--;
            }
        }
        *nmatches_out = nmatches;
    }
}

struct Scratch {   // RAII device scratch for the synchronous host-buffer entry points
    std::vector<void*> p;
    ~Scratch() { for (void* q : p) cudaFree(q); }
    template <class T> T* get(size_t n) { T* r = nullptr; if (cudaMalloc((void**)&r, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr; p.push_back(r); return r; }
    template <class T> T* up(const T* h, size_t n) { T* r = get<T>(n); if (r && n) cudaMemcpy(r, h, n * sizeof(T), cudaMemcpyHostToDevice); return r; }
};

}  // namespace

extern "C" {

int se2gpu_hamming_distance(const uint8_t* a, const uint8_t* b, int n, int* out, int device) {
    if (n < 0 || (n && (!a || !b || !out))) return fail(SE2GPU_ERR_INVALID, "bad arguments");
    int rc = se2gpu::select_device(device);
    if (rc != SE2GPU_OK) return rc;
    if (n == 0) return SE2GPU_OK;
    Scratch s;
    uint32_t* da = (uint32_t*)s.up(a, (size_t)n * 32); uint32_t* db = (uint32_t*)s.up(b, (size_t)n * 32); int* dout = s.get<int>(n);
    if (!da || !db || !dout) return fail(SE2GPU_ERR_CUDA, "alloc failed");
    SE2_LAUNCH(k_hamming_pairs, (n + 255) / 256, 256, 0, 0, da, db, n, dout);
    SE2_CUDA(cudaMemcpy(out, dout, sizeof(int) * n, cudaMemcpyDeviceToHost));
    return SE2GPU_OK;
}

int se2gpu_match_by_window(const se2gpu_keypoint* kp1, const uint8_t* desc1, int n1, const se2gpu_keypoint* kp2,
                           const uint8_t* desc2, int n2, float* prev, se2gpu_grid_params grid, int win_size,
                           int level_offset, int min_level, int max_level, float nnratio, int* matches12, int device) {
    if (n1 < 0 || n2 < 0) return fail(SE2GPU_ERR_INVALID, "negative sizes");
    int rc = se2gpu::select_device(device);
    if (rc != SE2GPU_OK) return rc;
    if (n1 == 0) return 0;
    if (n2 == 0) { for (int i = 0; i < n1; ++i) matches12[i] = -1; return 0; }
    Scratch s;
    auto* dk1 = s.up(kp1, n1); auto* dk2 = s.up(kp2, n2);
    uint32_t* dd1 = (uint32_t*)s.up(desc1, (size_t)n1 * 32); uint32_t* dd2 = (uint32_t*)s.up(desc2, (size_t)n2 * 32);
    float* dprev = s.up(prev, (size_t)2 * n1);
    // per-query window parameters (ORBmatcher.cpp:299-305)
    std::vector<int> qmin(n1), qmax(n1);
    std::vector<float> qr(n1, (float)win_size);
    for (int i = 0; i < n1; ++i) { const int l = kp1[i].octave; qmin[i] = l - level_offset > 0 ? l - level_offset : 0; qmax[i] = l + level_offset; }
    int* dqmin = s.up(qmin.data(), n1); int* dqmax = s.up(qmax.data(), n1); float* dqr = s.up(qr.data(), n1);
    int* cell = s.get<int>(n2); int* order = s.get<int>(n2); int* nvalid = s.get<int>(1);
    const int cap = n2;
    int2* cand = s.get<int2>((size_t)n1 * cap); int* ncand = s.get<int>(n1);
    int* work = s.get<int>((size_t)2 * n2 + n1); int* dm = s.get<int>(n1); int* dn = s.get<int>(1);
    if (!dk1 || !dk2 || !dd1 || !dd2 || !dprev || !dqmin || !dqmax || !dqr || !cell || !order || !nvalid || !cand || !ncand || !work || !dm || !dn)
        return fail(SE2GPU_ERR_CUDA, "alloc failed");
    SE2_CUDA(cudaMemset(nvalid, 0, sizeof(int)));
    SE2_LAUNCH(k_grid_cell, (n2 + 127) / 128, 128, 0, 0, dk2, n2, grid, cell);
    SE2_LAUNCH(k_grid_order, (n2 + 127) / 128, 128, 0, 0, cell, n2, order, nvalid);
    SE2_LAUNCH(k_candidates, (n1 * 32 + 255) / 256, 256, 0, 0, dprev, dqmin, dqmax, dqr, (const uint8_t*)nullptr, dd1, n1, dk2, dd2, cell, order,
               nvalid, (const uint8_t*)nullptr, grid, cap, cand, ncand);
    SE2_LAUNCH(k_resolve_window, 1, 32, 0, 0, dk1, dk2, n1, n2, min_level, max_level, nnratio, cap, cand, ncand, work, dm, dprev, dn);
    int nm = 0;
    SE2_CUDA(cudaMemcpy(matches12, dm, sizeof(int) * n1, cudaMemcpyDeviceToHost));
    SE2_CUDA(cudaMemcpy(prev, dprev, sizeof(float) * 2 * n1, cudaMemcpyDeviceToHost));
    SE2_CUDA(cudaMemcpy(&nm, dn, sizeof(int), cudaMemcpyDeviceToHost));
    return nm;
}

int se2gpu_match_by_projection(const se2gpu_keypoint* kf_kp, const uint8_t* kf_desc, int n_kf, const uint8_t* kf_observed,
                               const uint8_t* mp_valid, const float* mp_uv, int n_mp, const int* mp_octave,
                               const uint8_t* mp_desc, se2gpu_grid_params grid, int win_size, int level_offset,
                               float nnratio, int* matches_idx_mp, int device) {
    if (n_kf < 0 || n_mp < 0) return fail(SE2GPU_ERR_INVALID, "negative sizes");
    int rc = se2gpu::select_device(device);
    if (rc != SE2GPU_OK) return rc;
    for (int i = 0; i < n_kf; ++i) matches_idx_mp[i] = -1;
    if (n_kf == 0 || n_mp == 0) return 0;
    Scratch s;
    auto* dk = s.up(kf_kp, n_kf);
    uint32_t* dd = (uint32_t*)s.up(kf_desc, (size_t)n_kf * 32); uint32_t* dmpd = (uint32_t*)s.up(mp_desc, (size_t)n_mp * 32);
    uint8_t* dobs = s.up(kf_observed, n_kf); uint8_t* dvalid = s.up(mp_valid, n_mp); float* duv = s.up(mp_uv, (size_t)2 * n_mp);
    std::vector<int> qmin(n_mp), qmax(n_mp);
    std::vector<float> qr(n_mp);
    for (int i = 0; i < n_mp; ++i) {   // ORBmatcher.cpp:400-404
        const int pl = mp_octave[i];
        qr[i] = (float)(pl * win_size); qmin[i] = pl > level_offset ? pl - level_offset : 0; qmax[i] = pl + level_offset;
    }
    int* dqmin = s.up(qmin.data(), n_mp); int* dqmax = s.up(qmax.data(), n_mp); float* dqr = s.up(qr.data(), n_mp);
    int* cell = s.get<int>(n_kf); int* order = s.get<int>(n_kf); int* nvalid = s.get<int>(1);
    const int cap = n_kf;
    int2* cand = s.get<int2>((size_t)n_mp * cap); int* ncand = s.get<int>(n_mp);
    int* work = s.get<int>(n_kf); int* dm = s.get<int>(n_kf); int* dn = s.get<int>(1);
    if (!dk || !dd || !dmpd || !dobs || !dvalid || !duv || !dqmin || !dqmax || !dqr || !cell || !order || !nvalid || !cand || !ncand || !work || !dm || !dn)
        return fail(SE2GPU_ERR_CUDA, "alloc failed");
    SE2_CUDA(cudaMemset(nvalid, 0, sizeof(int)));
    SE2_LAUNCH(k_grid_cell, (n_kf + 127) / 128, 128, 0, 0, dk, n_kf, grid, cell);
    SE2_LAUNCH(k_grid_order, (n_kf + 127) / 128, 128, 0, 0, cell, n_kf, order, nvalid);
    SE2_LAUNCH(k_candidates, (n_mp * 32 + 255) / 256, 256, 0, 0, duv, dqmin, dqmax, dqr, dvalid, dmpd, n_mp, dk, dd, cell, order, nvalid, dobs, grid,
               cap, cand, ncand);
    SE2_LAUNCH(k_resolve_projection, 1, 32, 0, 0, dk, n_kf, n_mp, nnratio, cap, cand, ncand, work, dm, dn);
    int nm = 0;
    SE2_CUDA(cudaMemcpy(matches_idx_mp, dm, sizeof(int) * n_kf, cudaMemcpyDeviceToHost));
    SE2_CUDA(cudaMemcpy(&nm, dn, sizeof(int), cudaMemcpyDeviceToHost));
    return nm;
}

int se2gpu_search_by_bow(const se2gpu_bow_kf* k1, const se2gpu_bow_kf* k2, int mp_only, float nnratio, int check_orientation,
                         int* matches12, int device) {
    if (!k1 || !k2 || k1->n < 0 || k2->n < 0) return fail(SE2GPU_ERR_INVALID, "bad arguments");
    int rc = se2gpu::select_device(device);
    if (rc != SE2GPU_OK) return rc;
    for (int i = 0; i < k1->n; ++i) matches12[i] = -1;
    if (k1->n == 0 || k2->n == 0 || k1->n_node == 0 || k2->n_node == 0) return 0;
    Scratch s;
    auto up_kf = [&](const se2gpu_bow_kf* k, const float*& ang, const uint32_t*& d, const uint8_t*& mp, const int*& node, const int*& ptr, const int*& feat) {
        ang = s.up(k->angle, k->n); d = (const uint32_t*)s.up(k->desc, (size_t)k->n * 32); mp = s.up(k->has_mp, k->n);
        node = s.up(k->node, k->n_node); ptr = s.up(k->ptr, k->n_node + 1); feat = s.up(k->feat, k->ptr[k->n_node]);
        return ang && d && mp && node && ptr && feat;
    };
    const float *a1, *a2; const uint32_t *d1, *d2; const uint8_t *m1, *m2; const int *nd1, *nd2, *p1, *p2, *f1, *f2;
    if (!up_kf(k1, a1, d1, m1, nd1, p1, f1) || !up_kf(k2, a2, d2, m2, nd2, p2, f2)) return fail(SE2GPU_ERR_CUDA, "alloc failed");
    uint8_t* matched2 = s.get<uint8_t>(k2->n); int* bin_of = s.get<int>(k1->n); int* dm = s.get<int>(k1->n); int* dn = s.get<int>(1);
    if (!matched2 || !bin_of || !dm || !dn) return fail(SE2GPU_ERR_CUDA, "alloc failed");
    SE2_LAUNCH(k_search_by_bow, 1, 32, 0, 0, a1, d1, m1, k1->n, nd1, k1->n_node, p1, f1, a2, d2, m2, k2->n, nd2, k2->n_node, p2, f2, mp_only, nnratio,
               check_orientation, matched2, bin_of, dm, dn);
    int nm = 0;
    SE2_CUDA(cudaMemcpy(matches12, dm, sizeof(int) * k1->n, cudaMemcpyDeviceToHost));
    SE2_CUDA(cudaMemcpy(&nm, dn, sizeof(int), cudaMemcpyDeviceToHost));
    return nm;
}

}  // extern "C"

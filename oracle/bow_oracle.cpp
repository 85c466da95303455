// Bag-of-words ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatements of
//   * DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(feature, word_id, weight, nid, levelsup)
//     (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1220-1262) with FORB::distance = Hamming distance
//     (Thirdparty/DBoW2/DBoW2/FORB.cpp) on a flattened tree, and of the BowVector / FeatureVector assembly of
//     transform(features, v, fv, levelsup) (:1150-1216, TF_IDF weighting + L1 norm: the ORBvoc settings), and
//   * MapPoint::updateMainKFandDescriptor's choice of the main descriptor (reference src/MapPoint.cpp:228-272).
// Pinned by independent brute-force numpy restatements in tests/test_bow_oracle.py (the real ORBvoc.bin is not available
// offline; synthetic k-ary trees stand in for it). Only tests/ may load this.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
inline int hamming(const uint8_t* a, const uint8_t* b) {   // FORB::distance / ORBmatcher::DescriptorDistance
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}
}  // namespace

extern "C" {

// TemplatedVocabulary.h:1220-1262
void voc_oracle_transform(int n_nodes, const uint8_t* node_desc, const int* child_ptr, const int* children, const int* word_of,
                          const double* weight_of, int levels, const uint8_t* feat, int n, int levelsup, int* word_id, double* weight,
                          int* node_id) {
    (void)n_nodes;
    for (int f = 0; f < n; ++f) {
        const uint8_t* feature = feat + 32 * (size_t)f;
        const int nid_level = levels - levelsup;
        int nid = -1;
        if (nid_level <= 0) nid = 0;  // root
        int final_id = 0;             // root
        int current_level = 0;
        do {
            ++current_level;
            const int c0 = child_ptr[final_id], c1 = child_ptr[final_id + 1];
            final_id = children[c0];
            double best_d = hamming(feature, node_desc + 32 * (size_t)final_id);
            for (int k = c0 + 1; k < c1; ++k) {
                const int id = children[k];
                const double d = hamming(feature, node_desc + 32 * (size_t)id);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_ptr[final_id + 1] > child_ptr[final_id]);   // !isLeaf()
        word_id[f] = word_of[final_id];
        weight[f] = weight_of[final_id];
        if (node_id) node_id[f] = nid;
    }
}

// MapPoint.cpp:245-267
void median_descriptor_oracle(const uint8_t* desc, const int* ptr, int M, int* best_idx, int* best_median) {
    for (int m = 0; m < M; ++m) {
        const int N = ptr[m + 1] - ptr[m];
        if (N <= 0) { best_idx[m] = -1; if (best_median) best_median[m] = INT_MAX; continue; }
        const uint8_t* d0 = desc + 32 * (size_t)ptr[m];
        std::vector<float> Distances((size_t)N * N);
        for (int i = 0; i < N; i++) {
            Distances[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                int distij = hamming(d0 + 32 * (size_t)i, d0 + 32 * (size_t)j);
                Distances[(size_t)i * N + j] = distij;
                Distances[(size_t)j * N + i] = distij;
            }
        }
        int bestMedian = INT_MAX, bestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + (size_t)i * N, Distances.begin() + (size_t)(i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            int median = vDists[0.5 * (N - 1)];
            if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        }
        best_idx[m] = bestIdx;
        if (best_median) best_median[m] = bestMedian;
    }
}

}  // extern "C"

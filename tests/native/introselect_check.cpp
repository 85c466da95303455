// Pins se2lam_b200/csrc/introselect.h against this toolchain's real std::nth_element
// (the call cv::KeyPointsFilter::retainBest makes). Exit code 0 = identical permutations everywhere.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../se2lam_b200/csrc/introselect.h"

struct KP { float response; int id; };
struct Greater { bool operator()(const KP& a, const KP& b) const { return a.response > b.response; } };

static long g_heap_cases = 0;

static bool check(const std::vector<int>& scores, int nth) {
    const int n = (int)scores.size();
    std::vector<KP> ref(n);
    std::vector<uint32_t> mine(n);
    for (int i = 0; i < n; ++i) { ref[i] = KP{(float)scores[i], i}; mine[i] = ((uint32_t)scores[i] << 24) | (uint32_t)i; }
    std::nth_element(ref.begin(), ref.begin() + nth, ref.end(), Greater());
    se2gpu::kp_nth_element(mine.data(), n, nth);
    for (int i = 0; i < n; ++i)
        if ((int)(mine[i] & 0xFFFFFF) != ref[i].id) {
            fprintf(stderr, "mismatch n=%d nth=%d at %d: mine id %u ref id %d\n", n, nth, i, mine[i] & 0xFFFFFF, ref[i].id);
            return false;
        }
    return true;
}

// median-of-3 killer sequence (Musser) drives introselect into its heap-select fallback
static std::vector<int> killer(int n) {
    std::vector<int> v(n);
    int k = n / 2;
    for (int i = 0; i < k; ++i) {
        if (i % 2 == 0) v[i] = i + 1;
        else v[i] = k + i + (k % 2 ? 0 : 1);
        v[k + i] = 2 * (i + 1);
    }
    for (auto& x : v) x = 250 - (x % 250);
    return v;
}

int main() {
    std::mt19937 rng(12345);
    long cases = 0;
    for (int rep = 0; rep < 20000; ++rep) {
        int n = 1 + rng() % 700;
        int range = 1 + rng() % (rep % 3 == 0 ? 4 : (rep % 3 == 1 ? 40 : 230));
        std::vector<int> s(n);
        for (auto& x : s) x = 7 + rng() % range;
        if (rep % 7 == 0) std::sort(s.begin(), s.end());
        if (rep % 11 == 0) std::sort(s.rbegin(), s.rend());
        int nth = rng() % n;
        if (!check(s, nth)) return 1;
        ++cases;
    }
    for (int n = 4; n < 3000; n += 37) {
        auto s = killer(n);
        for (int nth : {0, n / 3, n / 2, n - 1}) { if (!check(s, nth)) return 1; ++cases; }
        // organ pipe
        std::vector<int> o(n);
        for (int i = 0; i < n; ++i) o[i] = 10 + std::min(i, n - 1 - i) % 240;
        if (!check(o, n / 2)) return 1;
        ++cases;
    }
    // monotone sequences with many ties at all sizes (forces deep recursion / degenerate pivots)
    for (int n = 1; n < 400; ++n) {
        std::vector<int> s(n);
        for (int i = 0; i < n; ++i) s[i] = 20 + (i * 7) % 5;
        for (int nth = 0; nth < n; nth += 1 + n / 13) { if (!check(s, nth)) return 1; ++cases; }
    }
    printf("introselect: %ld cases identical to std::nth_element\n", cases);
    return 0;
}

// Host check of se2lam_b200/csrc/fast_screen.h — the source pass A of orb_fast_cells_tma8 is compiled from — with the
// packed-SIMD instructions emulated:
//   1. screen4 == the scalar definition of the FAST-9-16 quick reject, per pixel, for random and extreme patches
//   2. the screen never rejects a true FAST corner (scalar 9-contiguous-arc definition, cv::FAST [upstream OpenCV fast.cpp])
//   3. a CTA's pass A simulated thread by thread (ItemWalk, inside_mask8, seg_offset, the kernel's word addressing): every
//      interior pixel of a cell is screened exactly once, the warps' list segments neither overlap nor overflow
// Prints "OK <checks>" and exits 0, or a diagnostic and exits 1.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../se2lam_b200/csrc/fast_screen.h"

static const int RING[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static bool scalar_screen(const uint8_t* p, int pw, int t) {   // p = centre pixel
    const int v = p[0];
    bool dk = true, br = true;
    for (int k = 0; k < 8; k += 2) {   // opposite pairs (0,8) (2,10) (4,12) (6,14)
        const int a = p[RING[k][1] * pw + RING[k][0]], b = p[RING[k + 8][1] * pw + RING[k + 8][0]];
        dk = dk && (v - a > t || v - b > t);
        br = br && (a - v > t || b - v > t);
    }
    return dk || br;
}

static bool scalar_corner(const uint8_t* p, int pw, int t) {
    const int v = p[0];
    int d[32];
    for (int k = 0; k < 16; ++k) d[k] = d[k + 16] = p[RING[k][1] * pw + RING[k][0]] - v;
    for (int s = 0; s < 16; ++s) {
        bool allb = true, alld = true;
        for (int k = 0; k < 9; ++k) { allb = allb && d[s + k] > t; alld = alld && d[s + k] < -t; }
        if (allb || alld) return true;
    }
    return false;
}

static uint32_t ldw(const uint8_t* base, long word) { uint32_t w; memcpy(&w, base + 4 * word, 4); return w; }

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

int main() {
    std::mt19937 rng(12345);
    long checks = 0;
    // ---- 1 + 2: screen4 on a 7-row x 3-word neighbourhood
    for (int iter = 0; iter < 200000; ++iter) {
        const int pw = 12;
        uint8_t buf[7 * 12];
        const int mode = iter % 5;
        for (auto& b : buf) {
            switch (mode) {
                case 0: b = (uint8_t)(rng() & 255); break;
                case 1: b = (uint8_t)(100 + (rng() % 50)); break;                       // differences near the thresholds
                case 2: b = (rng() & 1) ? 255 : 0; break;                               // extremes: |diff| = 255
                case 3: b = (uint8_t)((rng() % 3) * 21 + 90); break;                    // ties exactly at t = 20 +- 1
                default: b = (uint8_t)(((rng() >> 8) & 1) ? (rng() & 255) : 128); break;
            }
        }
        const int ts[4] = {20, 7, 1, 254};
        const int t = ts[iter % 4];
        const uint8_t* row = buf + 3 * pw;   // centre row, words 0..2; the screened word is word 1 (bytes 4..7)
        const unsigned m = fastpx::screen4(ldw(buf + 0 * pw, 1), ldw(buf + 6 * pw, 1), ldw(buf + 1 * pw, 0), ldw(buf + 1 * pw, 1), ldw(buf + 1 * pw, 2),
                                           ldw(buf + 5 * pw, 0), ldw(buf + 5 * pw, 1), ldw(buf + 5 * pw, 2), ldw(row, 0), ldw(row, 1), ldw(row, 2),
                                           fastpx::screen_T1(t), fastpx::screen_U1(t));
        for (int j = 0; j < 4; ++j) {
            const bool want = scalar_screen(row + 4 + j, pw, t);
            if ((((m >> j) & 1u) != 0) != want) FAIL("screen4 mismatch: iter %d pixel %d t %d got %u want %d", iter, j, t, (m >> j) & 1u, (int)want);
            if (scalar_corner(row + 4 + j, pw, t) && !want) FAIL("screen rejects a true corner: iter %d pixel %d t %d", iter, j, t);
            ++checks;
        }
    }
    // ---- inside_mask8, exhaustively over the argument range the kernels produce (every alignment shift of the TMA box)
    for (int shift = 0; shift < 16; ++shift)
        for (int cw = 1; cw <= 300; ++cw) {
            std::vector<int> seen(cw, 0);
            for (int h = 0; h < fastpx::pairs_per_row(cw, shift); ++h) {
                const int x0 = fastpx::pair_x0(h, shift);
                if (x0 > cw - 1 || x0 < -7) FAIL("pair %d of a %d px row (shift %d) starts at %d", h, cw, shift, x0);
                unsigned want = 0;
                for (int j = 0; j < 8; ++j) if (x0 + j >= 0 && x0 + j < cw) { want |= 1u << j; seen[x0 + j]++; }
                if (fastpx::inside_mask8(x0, cw) != want) FAIL("inside_mask8(%d, %d) = %x, want %x", x0, cw, fastpx::inside_mask8(x0, cw), want);
                ++checks;
            }
            for (int x = 0; x < cw; ++x) if (seen[x] != 1) FAIL("pairs cover pixel %d of %d (shift %d) %d times", x, cw, shift, seen[x]);
            std::fill(seen.begin(), seen.end(), 0);
            for (int g = 0; g < fastpx::groups_per_row(cw, shift); ++g) {
                const int x0 = fastpx::group_x0(g, shift);
                if (x0 > cw - 1 || x0 < -3) FAIL("group %d of a %d px row (shift %d) starts at %d", g, cw, shift, x0);
                const unsigned m = fastpx::inside_mask8(x0, cw) & 0xFu;
                for (int j = 0; j < 4; ++j) { const bool in = x0 + j >= 0 && x0 + j < cw; if ((((m >> j) & 1u) != 0) != in) FAIL("group mask"); if (in) seen[x0 + j]++; }
            }
            for (int x = 0; x < cw; ++x) if (seen[x] != 1) FAIL("groups cover pixel %d of %d (shift %d) %d times", x, cw, shift, seen[x]);
        }
    // ---- 3: pass A of one CTA (256 threads = 8 warps), simulated thread by thread exactly as the kernel addresses the patch
    const int NT = 256, NW = 8;
    const int sizes[][2] = {{122, 75}, {101, 62}, {103, 61}, {85, 50}, {93, 50}, {75, 41}, {61, 33}, {49, 26}, {1, 1}, {5, 3}, {6, 9}, {230, 225}, {250, 249}, {7, 200}, {13, 1}, {229, 17}};
    int cell_no = 0;
    for (const auto& sz : sizes) {
      for (int rep = 0; rep < 4; ++rep) {
        const int cw = sz[0], ch = sz[1];
        const int shift = (cell_no++ * 7 + rep * 5) % 16;                       // alignment shift of the TMA box start
        const int pw = (shift + cw + 6 + 15) & ~15, pww = pw / 4, bh = ch + 6 + (int)(rng() % 3);
        if ((size_t)pw * bh > 65535 || pw > 256 || bh > 256) continue;   // the library takes such cells to orb_fast_cells / _big (16-bit list entries, box <= 256 x 256)
        std::vector<uint8_t> smem((size_t)pw * bh + 4096, 0xAB);   // bytes behind the patch = the kernel's score plane (garbage to pass A)
        for (int y = 0; y < bh; ++y) for (int x = 0; x < pw; ++x) smem[(size_t)y * pw + x] = (uint8_t)((((x / 5) ^ (y / 4)) & 1) * 60 + 80 + (int)(rng() % 25));
        const uint8_t* patch = smem.data();
        const uint8_t* p0 = patch + 3 * pw + 3 + shift;
        const int t = 20;
        const int h0 = fastpx::first_pair(shift), G2 = fastpx::pairs_per_row(cw, shift), nitems = ch * G2;
        std::vector<int> visited((size_t)cw * ch, 0), cand((size_t)cw * ch, 0);
        std::vector<int> owner((size_t)8 * nitems, -1);
        std::vector<int> nseg(NW, 0);
        for (int tid = 0; tid < NT; ++tid) {
            const int wid = tid / 32, lane = tid % 32;
            fastpx::ItemWalk it;
            it.init(tid, NT, G2);
            const int seg = fastpx::seg_offset(wid, nitems, NW);
            for (int it0 = wid * 32; it0 < nitems; it0 += NW * 32) {
                const bool active = it.y < ch;
                if (active != (it0 + lane < nitems)) FAIL("activity test differs from the item bound (cw %d ch %d tid %d)", cw, ch, tid);
                if (active) {
                    if (it.y * G2 + it.h != it0 + lane) FAIL("ItemWalk left its item sequence (cw %d ch %d tid %d)", cw, ch, tid);
                    const int x0 = fastpx::pair_x0(it.h, shift);
                    const long cp = (long)(it.y + 3) * pww + 2 * (it.h + h0);
                    if (cp & 1) FAIL("odd word index for a 64-bit load");
                    const uint32_t n3x = ldw(patch, cp - 3 * pww), n3y = ldw(patch, cp - 3 * pww + 1), s3x = ldw(patch, cp + 3 * pww), s3y = ldw(patch, cp + 3 * pww + 1);
                    const uint32_t n2x = ldw(patch, cp - 2 * pww), n2y = ldw(patch, cp - 2 * pww + 1), s2x = ldw(patch, cp + 2 * pww), s2y = ldw(patch, cp + 2 * pww + 1);
                    const uint32_t zx = ldw(patch, cp), zy = ldw(patch, cp + 1);
                    const uint32_t n2l = ldw(patch, cp - 2 * pww - 1), n2r = ldw(patch, cp - 2 * pww + 2), s2l = ldw(patch, cp + 2 * pww - 1), s2r = ldw(patch, cp + 2 * pww + 2);
                    const uint32_t zl = ldw(patch, cp - 1), zr = ldw(patch, cp + 2);
                    if (4 * (cp + 3 * pww + 1) + 4 > (long)smem.size() || cp - 3 * pww < 0) FAIL("pass A reads outside shared memory");
                    unsigned m = fastpx::screen4(n3x, s3x, n2l, n2x, n2y, s2l, s2x, s2y, zl, zx, zy, fastpx::screen_T1(t), fastpx::screen_U1(t)) |
                                 fastpx::screen4(n3y, s3y, n2x, n2y, n2r, s2x, s2y, s2r, zx, zy, zr, fastpx::screen_T1(t), fastpx::screen_U1(t)) << 4;
                    const unsigned in = fastpx::inside_mask8(x0, cw);
                    for (int j = 0; j < 8; ++j) if ((in >> j) & 1u) visited[(size_t)it.y * cw + x0 + j]++;
                    m &= in;
                    const int e0 = it.y * pw + x0;
                    while (m) {
                        const int j = __builtin_ffs((int)m) - 1;
                        m &= m - 1;
                        const int off = e0 + j;
                        if (off < 0 || off > 65535 || off % pw != x0 + j || off / pw != it.y) FAIL("bad list entry");
                        cand[(size_t)it.y * cw + x0 + j] = 1;
                        const int slot = seg + nseg[wid]++;
                        if (slot >= 8 * nitems) FAIL("list overflow");
                        if (owner[slot] != -1) FAIL("segments of warps %d and %d overlap", owner[slot], wid);
                        owner[slot] = wid;
                    }
                }
                it.next();
            }
        }
        for (int w = 0; w + 1 < NW; ++w)
            if (fastpx::seg_offset(w, nitems, NW) + nseg[w] > fastpx::seg_offset(w + 1, nitems, NW)) FAIL("segment of warp %d overflows into the next one", w);
        for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x) {
                if (visited[(size_t)y * cw + x] != 1) FAIL("pixel (%d,%d) of a %dx%d cell screened %d times", x, y, cw, ch, visited[(size_t)y * cw + x]);
                const bool want = scalar_screen(p0 + y * pw + x, pw, t);
                if ((cand[(size_t)y * cw + x] != 0) != want) FAIL("candidate set differs at (%d,%d) of a %dx%d cell", x, y, cw, ch);
                ++checks;
            }
      }
    }
    printf("OK %ld\n", checks);
    return 0;
}

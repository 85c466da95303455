// FAST-9-16 necessary condition on packed ring differences, 4 pixels (one 32-bit patch word) at a time.
//
// Shared by the CUDA kernels (orb.cu: pass A of orb_fast_cells_tma8) and by a host unit test (tests/native/fast_screen_host.cpp),
// which compiles THIS source with the packed-SIMD intrinsics emulated and checks it against the scalar definition of cv::FAST's
// quick reject (reference src/ORBextractor.cpp:616-623 -> cv::FAST(cell, th, true) [upstream OpenCV fast.cpp]).
//
// A 9-pixel arc of the 16-pixel Bresenham ring always contains one pixel of each opposite pair (k, k+8). For a pixel to be a
// corner at threshold t with a DARKER arc (all 9 ring pixels < v - t) every one of the 4 tested pairs (0,8) (4,12) (2,10) (6,14)
// therefore needs a member with v - p > t; for a BRIGHTER arc a member with p - v > t. With e_k = 256 + v - p_k (in [1, 511]):
//   darker possible   <=>  min over pairs of max(e_k, e_k+8) > 256 + t
//   brighter possible <=>  max over pairs of min(e_k, e_k+8) < 256 - t
// Two pixels ride in the two s16 halves of a word, so one VIMNMX(3).S16x2 evaluates two pixels at once.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SE2_HD __host__ __device__ __forceinline__
#else
#define SE2_HD inline
#endif

namespace fastpx {
// packed-SIMD primitives: the PTX instruction in device code, an emulation of its semantics in host code
// (prmt.b32 default mode, per-halfword signed min/max, wrapping per-halfword add)
SE2_HD int16_t lo16(unsigned a) { return (int16_t)(a & 0xFFFF); }
SE2_HD int16_t hi16(unsigned a) { return (int16_t)(a >> 16); }
SE2_HD unsigned pack16(int lo, int hi) { return ((unsigned)lo & 0xFFFFu) | ((unsigned)hi << 16); }
SE2_HD unsigned perm(unsigned a, unsigned b, unsigned s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, s);
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
SE2_HD unsigned maxs2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    return pack16(lo16(a) > lo16(b) ? lo16(a) : lo16(b), hi16(a) > hi16(b) ? hi16(a) : hi16(b));
#endif
}
SE2_HD unsigned mins2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vmins2(a, b);
#else
    return pack16(lo16(a) < lo16(b) ? lo16(a) : lo16(b), hi16(a) < hi16(b) ? hi16(a) : hi16(b));
#endif
}
SE2_HD unsigned min3s2(unsigned a, unsigned b, unsigned c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_s16x2(a, b, c);
#else
    return mins2(mins2(a, b), c);
#endif
}
SE2_HD unsigned max3s2(unsigned a, unsigned b, unsigned c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    return maxs2(maxs2(a, b), c);
#endif
}
SE2_HD unsigned add2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return pack16((lo16(a) + lo16(b)) & 0xFFFF, (hi16(a) + hi16(b)) & 0xFFFF);
#endif
}
}  // namespace fastpx

namespace fastpx {

// threshold constants of screen4 (per halfword): D + T1 >= 0 (s16) <=> D > 256 + t ;  U1 + ~B >= 0 <=> B < 256 - t
SE2_HD unsigned screen_T1(int t) { return (unsigned)(0x10000 - (257 + t)) * 0x10001u; }
SE2_HD unsigned screen_U1(int t) { return (unsigned)(256 - t) * 0x10001u; }

// The 4 pixels of patch word `zc` (row y, bytes x..x+3). Neighbouring words: zl / zr = the words left / right of zc on row y,
// n3 / s3 = the word above / below at rows y-3 / y+3, n2l n2c n2r / s2l s2c s2r = the three words at rows y-2 / y+2.
// Returns bit j set <=> pixel j may be a FAST corner at the threshold encoded in T1/U1 (necessary condition).
SE2_HD unsigned screen4(unsigned n3, unsigned s3, unsigned n2l, unsigned n2c, unsigned n2r, unsigned s2l, unsigned s2c, unsigned s2r,
                        unsigned zl, unsigned zc, unsigned zr, unsigned T1, unsigned U1) {
    // 4-byte spans starting at dx = -3, +3 (row y) and dx = -2, +2 (rows y-2, y+2) of the word's first pixel
    const unsigned w_m3 = perm(zl, zc, 0x4321), w_p3 = perm(zc, zr, 0x6543);
    const unsigned nw_ = perm(n2l, n2c, 0x5432), ne_ = perm(n2c, n2r, 0x5432);
    const unsigned sw_ = perm(s2l, s2c, 0x5432), se_ = perm(s2c, s2r, 0x5432);
    unsigned m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int hlf = 0; hlf < 2; ++hlf) {
        const unsigned sel = hlf ? 0x4342u : 0x4140u;      // pixels (0,1) or (2,3) -> the two s16 halves
        const unsigned cb = perm(zc, 0, sel) + 0x01000100u;                                  // 256 + v
        // e = 256 + v - ring (both halves stay in [1,511], so the plain subtraction never borrows)
        const unsigned e0 = cb - perm(s3, 0, sel), e8 = cb - perm(n3, 0, sel);               // (0,+3) (0,-3)
        const unsigned e4 = cb - perm(w_p3, 0, sel), e12 = cb - perm(w_m3, 0, sel);          // (+3,0) (-3,0)
        const unsigned e2 = cb - perm(se_, 0, sel), e10 = cb - perm(nw_, 0, sel);            // (+2,+2) (-2,-2)
        const unsigned e6 = cb - perm(ne_, 0, sel), e14 = cb - perm(sw_, 0, sel);            // (+2,-2) (-2,+2)
        const unsigned D = mins2(min3s2(maxs2(e0, e8), maxs2(e4, e12), maxs2(e2, e10)), maxs2(e6, e14));
        const unsigned B = maxs2(max3s2(mins2(e0, e8), mins2(e4, e12), mins2(e2, e10)), mins2(e6, e14));
        const unsigned r = maxs2(add2(D, T1), add2(~B, U1));                                 // >= 0 per half <=> may be a corner
        const unsigned ok = ~r & 0x80008000u;
        m |= ((ok >> 15) & 1u) << (2 * hlf) | ((ok >> 31) & 1u) << (2 * hlf + 1);
    }
    return m;
}

// bits j of an 8-pixel item (interior x = x0 + j) that lie inside the cell interior [0, cw); needs x0 <= cw - 1 and x0 >= -7
SE2_HD unsigned inside_mask8(int x0, int cw) {
    const int lo = x0 < 0 ? -x0 : 0;
    const int hi = cw - x0 < 8 ? cw - x0 : 8;        // 1 <= hi <= 8
    return (0xFFu << lo) & (0xFFu >> (8 - hi)) & 0xFFu;
}

// ---- pass A work distribution of orb_fast_cells_tma8 (shared with the host test, which simulates the CTA's threads)
// The patch row starts `shift` bytes (0..15) left of the cell's first apron pixel (the TMA box start is 16 B aligned), so interior
// pixel x sits at patch byte x + 3 + shift. An item = row y of the cell x one 8-byte pair of patch words (2p, 2p+1), i.e. interior
// x = 8p - 3 - shift .. +7; the pairs that touch the interior are p = first_pair .. first_pair + pairs_per_row - 1.
SE2_HD int first_pair(int shift) { return (shift + 3) >> 3; }
SE2_HD int pairs_per_row(int cw, int shift) { return ((shift + 2 + cw) >> 3) - first_pair(shift) + 1; }
SE2_HD int pair_x0(int h, int shift) { return 8 * (h + first_pair(shift)) - 3 - shift; }   // interior x of bit 0 of item (., h): -7 .. cw-1
// the same for the 4-pixel groups of orb_fast_cells_tma (one patch word per item)
SE2_HD int first_group(int shift) { return (shift + 3) >> 2; }
SE2_HD int groups_per_row(int cw, int shift) { return ((shift + 2 + cw) >> 2) - first_group(shift) + 1; }
SE2_HD int group_x0(int g, int shift) { return 4 * (g + first_group(shift)) - 3 - shift; }   // -3 .. cw-1
// thread tid visits items tid, tid + nthreads, tid + 2 nthreads, ... without a division per item
struct ItemWalk {
    int y, h, dY, dH, G2;
    SE2_HD void init(int tid, int nthreads, int g2) { G2 = g2; y = tid / g2; h = tid - y * g2; dY = nthreads / g2; dH = nthreads - dY * g2; }
    SE2_HD void next() { h += dH; y += dY; if (h >= G2) { h -= G2; ++y; } }
};
// candidate-list segment of warp w: the 32-item chunks are dealt round-robin over the CTA's warps, a warp's segment holds
// 8 entries per item it owns; segments are packed back to back (total 8 * nitems entries)
SE2_HD int seg_offset(int w, int nitems, int nwarps) {
    const int T = 32 * nwarps, full = nitems / T, rem = nitems - full * T;
    return 8 * (w * full * 32 + (rem < w * 32 ? rem : w * 32));
}

}  // namespace fastpx

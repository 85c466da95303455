// ORB front-end on sm_100a — kernels + C ABI (include/se2gpu.h, se2gpu_orb_*).
//
// Replaces se2lam::ORBextractor::operator() (reference src/ORBextractor.cpp:727-788) and the OpenCV
// primitives underneath it, for batches of frames, bit-exactly (keypoint order included):
//   orb_pyr0 / orb_resize   ComputePyramid :790-831 (copyMakeBorder REFLECT_101, resize INTER_LINEAR: 11-bit
//                           fixed point, each level from the previous one)
//   orb_pyr0_undistort      optional: cv::undistort of Frame.cpp:22 folded into level 0 (se2gpu_orb_set_undistort)
//   orb_fast_cells          per-cell cv::FAST(…,fastTh)/FAST(…,7) :608-623 — one CTA per grid cell: the cell and
//                           its 3 px apron are staged in shared memory once; a branch-free necessary condition on
//                           every pixel, survivors compacted, then the threshold-free arc score
//                           M = max over the 16 nine-pixel arcs of min(+-diff) (corner at t <=> M > t, score = M-1)
//                           at full lane occupancy, cell-local 3x3 strict NMS into a bitmap, raster-ordered emission
//   orb_fast_cells_big      the same result for cells too large for the shared-memory candidate list
//   orb_select              quota redistribution :631-679 + KeyPointsFilter::retainBest twice :687-710 — the
//                           libstdc++ introselect permutation is reproduced exactly, warp-cooperatively (introselect.h)
//   orb_blur                GaussianBlur 7x7 sigma 2 :769 (float32 separable, fused multiply-add, RNE to u8)
//   orb_orient_describe     IC_Angle :130-157 + computeOrbDescriptor :160-200, one warp per keypoint; keypoint
//                           records are written as 28-byte cv::KeyPoint and 32-byte descriptors
// Data layout in HBM (per frame): every pyramid level is a bordered u8 plane (w+32) x (h+32) with row pitch
// rounded up to 32 B; two copies (plain, blurred); candidates are packed 32-bit records x:12|y:12|score:8.
#include <cuda.h>   // CUtensorMap + the cuTensorMapEncodeTiled prototype only: the entry point is resolved at run time (no libcuda link)

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"
#include "fast_screen.h"
#include "introselect.h"

namespace {

using se2gpu::fail;

constexpr int EDGE = 16;           // EDGE_THRESHOLD, ORBextractor.cpp:83
constexpr int HALF_PATCH = 15;     // HALF_PATCH_SIZE :82
constexpr int PATCH = 31;          // PATCH_SIZE :81
constexpr int MAX_LEVELS = 16;
constexpr int FAST_THREADS = 256;
constexpr int ORB_LANES = 4;
constexpr int BLUR_TW = 128, BLUR_TH = 36;   // one warp per tile; BLUR_TH + 6 warm-up rows = 6 turns of the 7-row ring

__device__ const signed char d_pattern[1024] = {
#include "orb_pattern_31.inc"
};
__constant__ int c_umax[16];
__constant__ int c_vmax[16];   // c_vmax[|u|] = largest |v| of the patch disc in column u (the disc is stored row-wise as umax[|v|])
__constant__ float c_gauss[7];

struct LevelGeo {
    int w, h, pitch;
    size_t plane_off;   // byte offset of the bordered plane inside a frame's plane block
    int nDesired, cols, rows, nCells, nfeaturesCell;
    int cell_base;      // first cell of this level in the cell table
    int kp_off, kp_cap; // slot range in the per-frame level keypoint buffer
    float scale, kp_size;
    int tab_off;        // offset of this level's resize tables (xofs | yofs) in the int table
    int tile_base, tiles_x, tiles_y;
    int fbw, fbh;       // orb_fast_cells_tma: TMA box of this level's FAST cells (bytes per row: multiple of 16; rows)
};

struct CellGeo {
    int level, x0, y0, x1, y1;  // interior [x0,x1) x [y0,y1) in level ROI coordinates
    int skipped;                // the reference's `continue` cells (:579-580, :603-604)
    int cand_off, cand_cap;     // slot range in the per-frame candidate buffer
};

struct TileGeo { int level, x0, y0; };  // blur tile origin in bordered-plane coordinates

struct CellHdr { int n_base, n_a, n_b, pad; };

struct OrbDev {  // passed by value to kernels
    int nlevels, nfeatures, fast_th, t_lo;
    int n_cells, n_tiles;
    const LevelGeo* levels;
    const CellGeo* cells;
    const TileGeo* tiles;
    const int* itab;            // xofs/yofs tables
    const short* stab;          // ialpha/ibeta tables (2 per entry)
    uint8_t* plain;             // [B][frame_plane_bytes]
    uint8_t* blurred;
    size_t frame_plane_bytes;
    uint32_t* cand;             // [B][cand_total]
    size_t cand_total;
    CellHdr* hdr;               // [B][n_cells]
    uint32_t* lkp;              // [B][lkp_total]
    int lkp_total;
    int* lcount;                // [B][nlevels]
    int* err;                   // device error flag
    int frame0;                 // first frame of this launch group (pipelined host path processes the batch in chunks)
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// level 0: copyMakeBorder(image, temp, 16,16,16,16, BORDER_REFLECT_101). The border is 16 px wide, so when the caller's
// rows are 16 B aligned the interior is a stream of aligned 16-byte copies: blockIdx.x < gridDim.x-1 does those, one
// thread = one vector x PYR0_ROWS rows. The last blockIdx.x column fills what is left of each row (the two 16 px
// borders, a ragged interior tail, the pitch padding — everything when the input is unaligned) one word per thread.
// The level geometry comes in as a kernel parameter (constant bank), not through a dependent global load.
constexpr int PYR0_ROWS = 4;
__global__ void __launch_bounds__(256) orb_pyr0(OrbDev d, LevelGeo L, const uint8_t* __restrict__ imgs, int stride, size_t frame_stride, int nvec) {
    const int tid = threadIdx.y * 64 + threadIdx.x;
    const int f = blockIdx.z + d.frame0;
    const uint8_t* img = imgs + f * frame_stride;
    uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
    const int rows = L.h + 2 * EDGE;
    if (blockIdx.x + 1 < gridDim.x) {
        const int vx = blockIdx.x * 64 + threadIdx.x;
        const int y0 = (blockIdx.y * 4 + threadIdx.y) * PYR0_ROWS;
        if (vx >= nvec) return;
        uint4 v[PYR0_ROWS];
#pragma unroll
        for (int k = 0; k < PYR0_ROWS; ++k)
            if (y0 + k < rows) v[k] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)reflect101(y0 + k - EDGE, L.h) * stride) + vx);
#pragma unroll
        for (int k = 0; k < PYR0_ROWS; ++k)
            if (y0 + k < rows) *reinterpret_cast<uint4*>(plane + (size_t)(y0 + k) * L.pitch + EDGE + 16 * vx) = v[k];
        return;
    }
    // remainder words of the rows [16*blockIdx.y, +16): columns [0,16) and [16 + 16*nvec, pitch)
    const int tail0 = EDGE + 16 * nvec, nw = 4 + (L.pitch - tail0) / 4;
    for (int item = tid; item < 4 * PYR0_ROWS * nw; item += 256) {
        const int r = item / nw, wi = item - r * nw;
        const int y = blockIdx.y * 4 * PYR0_ROWS + r;
        if (y >= rows) break;
        const int x0 = wi < 4 ? 4 * wi : tail0 + 4 * (wi - 4);
        const uint8_t* row = img + (size_t)reflect101(y - EDGE, L.h) * stride;
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int x = x0 + q;
            const uint32_t px = (x < L.w + 2 * EDGE) ? row[reflect101(x - EDGE, L.w)] : 0u;
            word |= px << (8 * q);
        }
        *reinterpret_cast<uint32_t*>(plane + (size_t)y * L.pitch + x0) = word;
    }
}

// level 0 with the lens undistortion of Frame::Frame folded in (reference src/Frame.cpp:22: cv::undistort(im, img, Kcam, Dcam)
// right before the extractor): every pixel of the bordered level-0 plane is
// remap(INTER_LINEAR, BORDER_CONSTANT 0) of the raw frame through the fixed-point map of the (reflected) pixel - integer
// source coordinates in m1, 5+5 fraction bits in m2, weights (32-a)(32-b)*32 ... (sum 2^15), (sum + 2^14) >> 15. The map
// depends on the camera only and is built once per frame size on the host with OpenCV's double arithmetic
// (build_undistort_map). One thread = 4 output pixels, one word store; the undistorted image is never materialised.
__global__ void __launch_bounds__(128) orb_pyr0_undistort(OrbDev d, LevelGeo L, const uint8_t* __restrict__ imgs, int stride, size_t frame_stride,
                                                          const short2* __restrict__ m1, const uint16_t* __restrict__ m2) {
    const int x4 = (blockIdx.x * 128 + threadIdx.x) * 4, y = blockIdx.y;
    const int f = blockIdx.z + d.frame0;
    if (x4 >= L.pitch) return;
    const uint8_t* img = imgs + f * frame_stride;
    const int dy = reflect101(y - EDGE, L.h);
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int x = x4 + q;
        uint32_t o = 0;
        if (x < L.w + 2 * EDGE) {
            const int idx = dy * L.w + reflect101(x - EDGE, L.w);
            const short2 s = __ldg(m1 + idx);
            const int fr = __ldg(m2 + idx), a = fr & 31, b = fr >> 5;
            const int sx = s.x, sy = s.y;
            const bool x0 = (unsigned)sx < (unsigned)L.w, x1 = (unsigned)(sx + 1) < (unsigned)L.w;
            const bool y0 = (unsigned)sy < (unsigned)L.h, y1 = (unsigned)(sy + 1) < (unsigned)L.h;
            const uint8_t* p = img + (ptrdiff_t)sy * stride + sx;
            const int t00 = (x0 && y0) ? p[0] : 0, t01 = (x1 && y0) ? p[1] : 0, t10 = (x0 && y1) ? p[stride] : 0, t11 = (x1 && y1) ? p[stride + 1] : 0;
            const int val = t00 * ((32 - a) * (32 - b) * 32) + t01 * (a * (32 - b) * 32) + t10 * ((32 - a) * b * 32) + t11 * (a * b * 32);
            o = (uint32_t)min(max((val + (1 << 14)) >> 15, 0), 255);
        }
        word |= o << (8 * q);
    }
    uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
    *reinterpret_cast<uint32_t*>(plane + (size_t)y * L.pitch + x4) = word;
}

// level l>0: resize(level l-1 -> level l, INTER_LINEAR) + copyMakeBorder(REFLECT_101) in one pass, cv::resize's
// fixed-point arithmetic (11-bit coefficients, horizontal pass kept at full precision, vertical pass >>4, >>16, +2 >>2).
// One CTA = a 128 x RESIZE_TR tile of the bordered output plane, three stages in shared memory:
//   0  the source rows/columns the tile touches (a contiguous box, also across the reflected border) as 16 B vectors
//   1  horizontal pass h[s][x] = src[s][sx]*a0 + src[s][sx+1]*a1 for every staged source row s, ONCE per (row, column)
//      (consecutive output rows share source rows; cv::resize does the same with its row buffers)
//   2  vertical pass from two h rows per output row, 4 pixels per thread, one 32-bit store
// Column terms (xofs, ialpha) and row terms are computed once per tile (one thread per column / row) and shared through small tables.
constexpr int RESIZE_TR = 32;
struct ResizeRow { int s0, s1, b0, b1; };
struct ResizeCol { int sx0, sx1; short a0, a1; int valid; };   // 16 B
__global__ void __launch_bounds__(256) orb_resize(OrbDev d, LevelGeo L, LevelGeo S, int max_rows, int raw_pitch) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    __shared__ ResizeRow rowinfo[RESIZE_TR];
    __shared__ __align__(16) ResizeCol colinfo[128];
    __shared__ int s_lo[2], s_hi[2];     // [1] source rows
    __shared__ int s_cmin[4], s_cmax[4]; // source column range per column warp
    int* hbuf = reinterpret_cast<int*>(rs_smem);                       // [max_rows][128]
    uint8_t* raw = rs_smem + (size_t)max_rows * 128 * sizeof(int);     // [max_rows][raw_pitch]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x4 = blockIdx.x * 128 + tx * 4, y0 = blockIdx.y * RESIZE_TR;
    const int f = blockIdx.z + d.frame0;
    const int W = L.w + 2 * EDGE, H = L.h + 2 * EDGE;
    const int* xofs = d.itab + L.tab_off;
    const int* yofs = xofs + L.w;
    const short2* ialpha = reinterpret_cast<const short2*>(d.stab + 2 * (size_t)L.tab_off);
    const short2* ibeta = ialpha + L.w;
    // column terms of the tile's 128 output columns: warps 0..3, one column per thread (the 8 thread rows share them through
    // shared memory instead of each recomputing its 4 columns); row terms: warp 4
    if (tid < 128) {
        const int x = blockIdx.x * 128 + tid;
        ResizeCol c{0, 0, 0, 0, 0};
        int cmin = 0x7fffffff, cmax = -1;
        if (x < W) {
            const int dx = reflect101(x - EDGE, L.w);
            c.sx0 = __ldg(xofs + dx);
            c.sx1 = min(c.sx0 + 1, S.w - 1);
            const short2 aa = __ldg(ialpha + dx);
            c.a0 = aa.x; c.a1 = aa.y; c.valid = 1;
            cmin = c.sx0; cmax = c.sx1;
        }
        colinfo[tid] = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o)); cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); }
        if (tx == 0) { s_cmin[ty] = cmin; s_cmax[ty] = cmax; }
    } else if (ty == 4) {   // row terms of the tile's RESIZE_TR output rows and their source row range
        const int y = y0 + tx;
        int rmin = 0x7fffffff, rmax = -1;
        if (y < H) {
            const int dy = reflect101(y - EDGE, L.h);
            const int sy = __ldg(yofs + dy);
            const short2 bb = __ldg(ibeta + dy);
            ResizeRow r{min(max(sy, 0), S.h - 1), min(max(sy + 1, 0), S.h - 1), bb.x, bb.y};
            rowinfo[tx] = r;
            rmin = r.s0; rmax = r.s1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o)); rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o)); }
        if (tx == 0) { s_lo[1] = rmin; s_hi[1] = rmax; }
    }
    __syncthreads();
    const int c_min = min(min(s_cmin[0], s_cmin[1]), min(s_cmin[2], s_cmin[3])), c_max = max(max(s_cmax[0], s_cmax[1]), max(s_cmax[2], s_cmax[3]));
    const int c_lo = c_min & ~15, nvec = c_max < 0 ? 0 : (c_max - c_lo) / 16 + 1;
    const int r_lo = s_lo[1], nsr = s_hi[1] - r_lo + 1;
    if (nsr > max_rows || nvec * 16 > raw_pitch) { if (tid == 0) *d.err = 3; return; }   // sized on the host from the scale factor
    const uint8_t* src = d.plain + f * d.frame_plane_bytes + S.plane_off + (size_t)EDGE * S.pitch + EDGE;
    // stage 0: 16 vectors per source row and pass (a 128-column tile at scale <= 1.9 spans <= 16 vectors), no division
    if (nvec <= 16) {
        const int v = tid & 15;
        if (v < nvec)
            for (int r = tid >> 4; r < nsr; r += 16)
                *reinterpret_cast<uint4*>(raw + r * raw_pitch + 16 * v) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(r_lo + r) * S.pitch + c_lo) + v);
    } else {
        for (int i = tid; i < nsr * nvec; i += 256) {
            const int r = i / nvec, v = i - r * nvec;
            *reinterpret_cast<uint4*>(raw + r * raw_pitch + 16 * v) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(r_lo + r) * S.pitch + c_lo) + v);
        }
    }
    // this thread's 4 columns
    int sx0[4], sx1[4], a0[4], a1[4];
    unsigned vmask = 0;                       // byte mask of the valid columns
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const ResizeCol c = colinfo[4 * tx + q];
        sx0[q] = c.valid ? c.sx0 - c_lo : 0; sx1[q] = c.valid ? c.sx1 - c_lo : 0;
        a0[q] = c.a0; a1[q] = c.a1;
        vmask |= c.valid ? (0xFFu << (8 * q)) : 0u;
    }
    __syncthreads();
    // stage 1: horizontal pass, stored pre-shifted (the vertical pass uses S >> 4 only)
    if (x4 < W) {
        for (int r = ty; r < nsr; r += 8) {
            const uint8_t* rp = raw + r * raw_pitch;
            int4 hv;
            hv.x = (rp[sx0[0]] * a0[0] + rp[sx1[0]] * a1[0]) >> 4;
            hv.y = (rp[sx0[1]] * a0[1] + rp[sx1[1]] * a1[1]) >> 4;
            hv.z = (rp[sx0[2]] * a0[2] + rp[sx1[2]] * a1[2]) >> 4;
            hv.w = (rp[sx0[3]] * a0[3] + rp[sx1[3]] * a1[3]) >> 4;
            *reinterpret_cast<int4*>(hbuf + r * 128 + 4 * tx) = hv;
        }
    }
    __syncthreads();
    // stage 2: vertical pass. With coefficients in [0, 2048] (pairs summing to 2048 +- 1) and 8-bit pixels the result is in
    // [0, 255] by construction ((2049 * (255 * 2049 >> 4) >> 16) + 2 >> 2 = 255): cv's saturate_cast never fires, no clamp here.
    if (x4 >= L.pitch) return;
    uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
#pragma unroll
    for (int k = 0; k < RESIZE_TR / 8; ++k) {
        const int r = ty + 8 * k, y = y0 + r;
        if (y >= H) break;
        uint32_t word = 0;
        if (x4 < W) {
            const ResizeRow ri = rowinfo[r];
            const int4 A = *reinterpret_cast<const int4*>(hbuf + (ri.s0 - r_lo) * 128 + 4 * tx);
            const int4 B = *reinterpret_cast<const int4*>(hbuf + (ri.s1 - r_lo) * 128 + 4 * tx);
            const unsigned v0 = (unsigned)((((ri.b0 * A.x) >> 16) + ((ri.b1 * B.x) >> 16) + 2) >> 2);
            const unsigned v1 = (unsigned)((((ri.b0 * A.y) >> 16) + ((ri.b1 * B.y) >> 16) + 2) >> 2);
            const unsigned v2 = (unsigned)((((ri.b0 * A.z) >> 16) + ((ri.b1 * B.z) >> 16) + 2) >> 2);
            const unsigned v3 = (unsigned)((((ri.b0 * A.w) >> 16) + ((ri.b1 * B.w) >> 16) + 2) >> 2);
            word = (v0 | (v1 << 8) | (v2 << 16) | (v3 << 24)) & vmask;
        }
        *reinterpret_cast<uint32_t*>(plane + (size_t)y * L.pitch + x4) = word;
    }
}

// orb_resize with a shared-memory diet (the resize is bound by shared-memory wavefronts: 0.62 per cycle and SM at level 1, 38 % of
// them bank-conflict replays of the byte gathers, profiles/r02c_ncu_orb.md): the horizontal pass reads a 12-byte window per row and
// thread (3 word loads instead of 8 byte loads; PRMT picks the byte pair, IDP.2A does src[sx]*a0 + src[sx+1]*a1) and its results are
// kept as 16-bit values (<= 32 640), which halves the wavefronts of the store and of the two loads per output row of the vertical pass.
// The arithmetic is that of orb_resize term by term; the planes are bit-identical. Values that do not fit the window form (scale
// factors above ~2.6, negative coefficients) take the byte path inside the same kernel.
// PDL: launched with programmatic stream serialisation (cudaLaunchAttributeProgrammaticStreamSerialization). Every CTA releases its
// dependents at once (griddepcontrol.launch_dependents), so the CTAs of the NEXT level's launch are scheduled into the SM slots
// the last wave of this launch frees and run their table set-up there; they read this level's plane only behind
// griddepcontrol.wait, which returns when this whole grid has completed and its writes are visible. The pyramid is a chain of seven
// dependent, short launches: this overlaps each launch's ramp-up with its predecessor's tail.
template <bool PDL>
__global__ void __launch_bounds__(256) orb_resize_w(OrbDev d, LevelGeo L, LevelGeo S, int max_rows, int raw_pitch) {
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) uint8_t rs_smem[];
    __shared__ ResizeRow rowinfo[RESIZE_TR];
    __shared__ __align__(16) ResizeCol colinfo[128];
    __shared__ int s_lo[2], s_hi[2];     // [1] source rows
    __shared__ int s_cmin[4], s_cmax[4]; // source column range per column warp
    uint16_t* hbuf = reinterpret_cast<uint16_t*>(rs_smem);                  // [max_rows][128], 16 bit: (255 * 2048) >> 4 = 32 640
    uint8_t* raw = rs_smem + (size_t)max_rows * 128 * sizeof(uint16_t);     // [max_rows][raw_pitch] (+16 B of slack behind the last row)
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x4 = blockIdx.x * 128 + tx * 4, y0 = blockIdx.y * RESIZE_TR;
    const int f = blockIdx.z + d.frame0;
    const int W = L.w + 2 * EDGE, H = L.h + 2 * EDGE;
    const int* xofs = d.itab + L.tab_off;
    const int* yofs = xofs + L.w;
    const short2* ialpha = reinterpret_cast<const short2*>(d.stab + 2 * (size_t)L.tab_off);
    const short2* ibeta = ialpha + L.w;
    // column terms of the tile's 128 output columns: warps 0..3, one column per thread (the 8 thread rows share them through
    // shared memory instead of each recomputing its 4 columns); row terms: warp 4
    if (tid < 128) {
        const int x = blockIdx.x * 128 + tid;
        ResizeCol c{0, 0, 0, 0, 0};
        int cmin = 0x7fffffff, cmax = -1;
        if (x < W) {
            const int dx = reflect101(x - EDGE, L.w);
            c.sx0 = __ldg(xofs + dx);
            c.sx1 = min(c.sx0 + 1, S.w - 1);
            const short2 aa = __ldg(ialpha + dx);
            c.a0 = aa.x; c.a1 = aa.y; c.valid = 1;
            cmin = c.sx0; cmax = c.sx1;
        }
        colinfo[tid] = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o)); cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); }
        if (tx == 0) { s_cmin[ty] = cmin; s_cmax[ty] = cmax; }
    } else if (ty == 4) {   // row terms of the tile's RESIZE_TR output rows and their source row range
        const int y = y0 + tx;
        int rmin = 0x7fffffff, rmax = -1;
        if (y < H) {
            const int dy = reflect101(y - EDGE, L.h);
            const int sy = __ldg(yofs + dy);
            const short2 bb = __ldg(ibeta + dy);
            ResizeRow r{min(max(sy, 0), S.h - 1), min(max(sy + 1, 0), S.h - 1), bb.x, bb.y};
            rowinfo[tx] = r;
            rmin = r.s0; rmax = r.s1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o)); rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o)); }
        if (tx == 0) { s_lo[1] = rmin; s_hi[1] = rmax; }
    }
    __syncthreads();
    const int c_min = min(min(s_cmin[0], s_cmin[1]), min(s_cmin[2], s_cmin[3])), c_max = max(max(s_cmax[0], s_cmax[1]), max(s_cmax[2], s_cmax[3]));
    const int c_lo = c_min & ~15, nvec = c_max < 0 ? 0 : (c_max - c_lo) / 16 + 1;
    const int r_lo = s_lo[1], nsr = s_hi[1] - r_lo + 1;
    if (nsr > max_rows || nvec * 16 > raw_pitch) { if (tid == 0) *d.err = 3; return; }   // sized on the host from the scale factor
    const uint8_t* src = d.plain + f * d.frame_plane_bytes + S.plane_off + (size_t)EDGE * S.pitch + EDGE;
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");     // the source level is complete and visible from here on
    // stage 0: 16 vectors per source row and pass (a 128-column tile at scale <= 1.9 spans <= 16 vectors), no division
    if (nvec <= 16) {
        const int v = tid & 15;
        if (v < nvec)
            for (int r = tid >> 4; r < nsr; r += 16)
                *reinterpret_cast<uint4*>(raw + r * raw_pitch + 16 * v) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(r_lo + r) * S.pitch + c_lo) + v);
    } else {
        for (int i = tid; i < nsr * nvec; i += 256) {
            const int r = i / nvec, v = i - r * nvec;
            *reinterpret_cast<uint4*>(raw + r * raw_pitch + 16 * v) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(r_lo + r) * S.pitch + c_lo) + v);
        }
    }
    // this thread's 4 columns
    int sx0[4], sx1[4], a0[4], a1[4];
    unsigned vmask = 0;                       // byte mask of the valid columns
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const ResizeCol c = colinfo[4 * tx + q];
        sx0[q] = c.valid ? c.sx0 - c_lo : 0; sx1[q] = c.valid ? c.sx1 - c_lo : 0;
        a0[q] = c.a0; a1[q] = c.a1;
        vmask |= c.valid ? (0xFFu << (8 * q)) : 0u;
    }
    __syncthreads();
    // the 4 columns' source bytes sx0, sx0+1 lie in one 12-byte window of the staged row (3 aligned words from word w0 on) as long
    // as the scale factor is below ~2.6; byte pairs come out of the window with one PRMT (selector fixed per column), and
    // src[sx0]*a0 + src[sx0+1]*a1 is one IDP.2A. Where cv::resize clamps sx1 to sx0 (right edge) a1 is 0, so the byte behind sx0
    // may be anything. Three 32-bit loads per row instead of eight byte loads: the kernel is bound by shared-memory wavefronts.
    unsigned wq[4], selq[4];
    bool pairq[4];
    int w0 = 0;
    bool window_ok = true;
    {
        int mn = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < 4; ++q) if ((vmask >> (8 * q)) & 1u) mn = min(mn, sx0[q]);
        if (mn == 0x7fffffff) mn = 0;
        w0 = mn >> 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool valid = (vmask >> (8 * q)) & 1u;
            const int o = valid ? sx0[q] - 4 * w0 : 0;
            if (valid && (o + 1 > 11 || (sx1[q] != sx0[q] + 1 && a1[q] != 0))) window_ok = false;
            pairq[q] = o >= 7;
            const int o2 = o - (pairq[q] ? 4 : 0);
            selq[q] = (unsigned)o2 | ((unsigned)(o2 + 1) << 4);
            wq[q] = ((unsigned)a0[q] & 0xFFFFu) | ((unsigned)a1[q] << 16);
            if (a0[q] < 0 || a1[q] < 0) window_ok = false;     // cv's coefficients are in [0, 2048]; anything else takes the byte path
        }
    }
    // stage 1: horizontal pass, stored pre-shifted (the vertical pass uses S >> 4 only)
    if (x4 < W) {
        if (window_ok) {
            for (int r = ty; r < nsr; r += 8) {
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw + r * raw_pitch) + w0;
                const uint32_t W0 = rw[0], W1 = rw[1], W2 = rw[2];
                unsigned hq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned bytes = __byte_perm(pairq[q] ? W1 : W0, pairq[q] ? W2 : W1, selq[q]);
                    hq[q] = __dp2a_lo(wq[q], bytes, 0u) >> 4;
                }
                *reinterpret_cast<uint2*>(hbuf + r * 128 + 4 * tx) = make_uint2(hq[0] | (hq[1] << 16), hq[2] | (hq[3] << 16));
            }
        } else {
            for (int r = ty; r < nsr; r += 8) {
                const uint8_t* rp = raw + r * raw_pitch;
                const unsigned h0 = (unsigned)((rp[sx0[0]] * a0[0] + rp[sx1[0]] * a1[0]) >> 4), h1 = (unsigned)((rp[sx0[1]] * a0[1] + rp[sx1[1]] * a1[1]) >> 4);
                const unsigned h2 = (unsigned)((rp[sx0[2]] * a0[2] + rp[sx1[2]] * a1[2]) >> 4), h3 = (unsigned)((rp[sx0[3]] * a0[3] + rp[sx1[3]] * a1[3]) >> 4);
                *reinterpret_cast<uint2*>(hbuf + r * 128 + 4 * tx) = make_uint2((h0 & 0xFFFFu) | (h1 << 16), (h2 & 0xFFFFu) | (h3 << 16));
            }
        }
    }
    __syncthreads();
    // stage 2: vertical pass. With coefficients in [0, 2048] (pairs summing to 2048 +- 1) and 8-bit pixels the result is in
    // [0, 255] by construction ((2049 * (255 * 2049 >> 4) >> 16) + 2 >> 2 = 255): cv's saturate_cast never fires, no clamp here.
    if (x4 >= L.pitch) return;
    uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
#pragma unroll
    for (int k = 0; k < RESIZE_TR / 8; ++k) {
        const int r = ty + 8 * k, y = y0 + r;
        if (y >= H) break;
        uint32_t word = 0;
        if (x4 < W) {
            const ResizeRow ri = rowinfo[r];
            const uint2 A = *reinterpret_cast<const uint2*>(hbuf + (ri.s0 - r_lo) * 128 + 4 * tx);
            const uint2 B = *reinterpret_cast<const uint2*>(hbuf + (ri.s1 - r_lo) * 128 + 4 * tx);
            const unsigned v0 = (unsigned)((((ri.b0 * (int)(A.x & 0xFFFFu)) >> 16) + ((ri.b1 * (int)(B.x & 0xFFFFu)) >> 16) + 2) >> 2);
            const unsigned v1 = (unsigned)((((ri.b0 * (int)(A.x >> 16)) >> 16) + ((ri.b1 * (int)(B.x >> 16)) >> 16) + 2) >> 2);
            const unsigned v2 = (unsigned)((((ri.b0 * (int)(A.y & 0xFFFFu)) >> 16) + ((ri.b1 * (int)(B.y & 0xFFFFu)) >> 16) + 2) >> 2);
            const unsigned v3 = (unsigned)((((ri.b0 * (int)(A.y >> 16)) >> 16) + ((ri.b1 * (int)(B.y >> 16)) >> 16) + 2) >> 2);
            word = (v0 | (v1 << 8) | (v2 << 16) | (v3 << 24)) & vmask;
        }
        *reinterpret_cast<uint32_t*>(plane + (size_t)y * L.pitch + x4) = word;
    }
}

// FAST-9-16 on packed ring differences. For centre v and ring pixel p the s16x2 word
//   q = (256 + v - p) | (256 + p - v) << 16  =  p * 0xFFFF + ((256 + v) | (256 - v) << 16)      (one IMAD; both halves in [1,511])
// carries the "darker" and the "brighter" test side by side, so Blackwell's packed 3-input min/max (VIMNMX3.S16x2)
// evaluates both polarities at once. Arc score M = max over the 16 nine-pixel arcs of min(+-(v - ring)); corner at
// t <=> M > t, and cv::FAST's stored score == M-1 whatever t was.
__device__ __forceinline__ unsigned fast_cv(int v) { return (unsigned)(256 + v) | ((unsigned)(256 - v) << 16); }
__device__ __forceinline__ unsigned fast_q(unsigned p, unsigned cv) { return p * 0xFFFFu + cv; }

__device__ __forceinline__ int fast_arc_score(const uint8_t* __restrict__ p, int pw) {
    const unsigned cv = fast_cv(p[0]);
    unsigned q[16];
    q[0] = fast_q(p[3 * pw], cv);       q[1] = fast_q(p[3 * pw + 1], cv);   q[2] = fast_q(p[2 * pw + 2], cv);    q[3] = fast_q(p[pw + 3], cv);
    q[4] = fast_q(p[3], cv);            q[5] = fast_q(p[-pw + 3], cv);      q[6] = fast_q(p[-2 * pw + 2], cv);   q[7] = fast_q(p[-3 * pw + 1], cv);
    q[8] = fast_q(p[-3 * pw], cv);      q[9] = fast_q(p[-3 * pw - 1], cv);  q[10] = fast_q(p[-2 * pw - 2], cv);  q[11] = fast_q(p[-pw - 3], cv);
    q[12] = fast_q(p[-3], cv);          q[13] = fast_q(p[pw - 3], cv);      q[14] = fast_q(p[2 * pw - 2], cv);   q[15] = fast_q(p[3 * pw - 1], cv);
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(q[k], q[(k + 1) & 15], q[(k + 2) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a9 = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned b9 = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a9, b9);
    }
    return (int)max(best & 0xFFFFu, best >> 16) - 256;
}

// one CTA per (cell, frame): cv::FAST(cell, fastTh, NMS) and, if that yields <= 3 keypoints, cv::FAST(cell, 7, NMS)
// (ORBextractor.cpp:616-623). The cell and its 3 px apron are staged once in shared memory with aligned 32-bit loads.
//   A  every pixel: necessary condition (a 9-arc contains one pixel of each opposite pair (k, k+8), so for one polarity all 4
//      tested pairs need a member beyond t), branch-free on two pixels per s16x2 word; survivors are compacted into a
//      shared list (warp scan + one shared atomic per 128 pixels)
//   B  list entries, full warps: arc score -> score plane (same pitch as the patch, 1 px apron)
//   C  list entries: strict 3x3 maximum -> one bit per pixel in a raster-order bitmap
//   D  one warp: exclusive scan of the bitmap words' popcounts = raster-order output slots
//   E  one thread per bitmap word: emit (score | y | x)
__global__ void __launch_bounds__(FAST_THREADS) orb_fast_cells(OrbDev d) {
    extern __shared__ uint8_t smem[];
    __shared__ int s_total, s_ncand;
    const CellGeo c = d.cells[blockIdx.x];
    const int f = blockIdx.y + d.frame0;
    CellHdr* hdr = d.hdr + (size_t)f * d.n_cells + blockIdx.x;
    const int cw = c.x1 - c.x0, ch = c.y1 - c.y0;
    if (c.skipped || cw <= 0 || ch <= 0) {
        if (threadIdx.x == 0) { hdr->n_base = 0; hdr->n_a = 0; hdr->n_b = 0; }
        return;
    }
    const LevelGeo& L = d.levels[c.level];
    const uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
    // patch columns start at the 4-aligned bordered-plane column ax0 <= x0-3 (plane base and pitch are 32 B aligned)
    const int bx0 = c.x0 - 3 + EDGE, by0 = c.y0 - 3 + EDGE;
    const int ax0 = bx0 & ~3, shift = bx0 - ax0;
    const int ph = ch + 6;
    const int pww = (shift + cw + 6 + 3) >> 2, pw = pww * 4;
    const int nbits = pw * ch, nwords = (nbits + 31) >> 5;
    const int gx0 = (3 + shift) >> 2, G = ((3 + shift + cw - 1) >> 2) - gx0 + 1, nitems = ch * G;   // pass A work items
    uint8_t* patch = smem;
    uint8_t* score = smem + ((pw * ph + 15) & ~15);                                       // [(ch+2) x pw], pixel (x,y) at (y+1)*pw + x+1
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(score + ((pw * (ch + 2) + 15) & ~15)); // [nwords], bit y*pw + x
    uint32_t* woff = bitmap + nwords;                                                     // [nwords] exclusive popcount scan
    uint16_t* list = reinterpret_cast<uint16_t*>(woff + nwords);                          // [<= cw*ch] entries y*pw + x
    for (int i = threadIdx.x; i < pww * ph; i += FAST_THREADS) {
        const int py = i / pww, pxw = i - py * pww;
        reinterpret_cast<uint32_t*>(patch)[i] = *reinterpret_cast<const uint32_t*>(plane + (size_t)(by0 + py) * L.pitch + ax0 + 4 * pxw);
    }
    constexpr int NW = FAST_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t* p0 = patch + 3 * pw + 3 + shift;
    int thr = d.fast_th;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < (pw * (ch + 2) + 3) / 4; i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0u;
        for (int i = threadIdx.x; i < nwords; i += FAST_THREADS) bitmap[i] = 0u;
        if (threadIdx.x == 0) s_ncand = 0;
        __syncthreads();
        // A: one thread = the 4 pixels of one patch word (group), items = (row, group) pairs flattened over the warp
        {
            const uint32_t* pwords = reinterpret_cast<const uint32_t*>(patch);
            const unsigned T1 = (unsigned)(0x10000 - (257 + thr)) * 0x10001u;   // D + T1 >= 0 (s16)  <=>  D > 256 + thr
            const unsigned U1 = (unsigned)(256 - thr) * 0x10001u;               // U1 + ~B >= 0       <=>  B < 256 - thr
            for (int it0 = wid * 32; it0 < nitems; it0 += NW * 32) {
                const int item = it0 + lane;
                unsigned m = 0;
                int y = 0, col0 = 0;
                if (item < nitems) {
                    y = item / G;
                    const int g = item - y * G + gx0;
                    col0 = 4 * g - (3 + shift);                    // interior x of the group's first pixel
                    const uint32_t* c = pwords + (y + 3) * pww + g;
                    const uint32_t n3 = c[-3 * pww], s3 = c[3 * pww];
                    const uint32_t n2a = c[-2 * pww - 1], n2b = c[-2 * pww], n2c = c[-2 * pww + 1];
                    const uint32_t s2a = c[2 * pww - 1], s2b = c[2 * pww], s2c = c[2 * pww + 1];
                    const uint32_t za = c[-1], zb = c[0], zc = c[1];
                    // 4-byte spans starting at dx = -3, -2, +2, +3 of the group's first pixel
                    const uint32_t w_m3 = __byte_perm(za, zb, 0x4321), w_p3 = __byte_perm(zb, zc, 0x6543);
                    const uint32_t nw_ = __byte_perm(n2a, n2b, 0x5432), ne_ = __byte_perm(n2b, n2c, 0x5432);
                    const uint32_t sw_ = __byte_perm(s2a, s2b, 0x5432), se_ = __byte_perm(s2b, s2c, 0x5432);
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const unsigned sel = hlf ? 0x4342u : 0x4140u;      // pixels (0,1) or (2,3) -> the two s16 halves
                        const unsigned cb = __byte_perm(zb, 0, sel) + 0x01000100u;          // 256 + v
                        // e = 256 + v - ring (both halves stay in [1,511], so the plain subtraction never borrows)
                        const unsigned e0 = cb - __byte_perm(s3, 0, sel), e8 = cb - __byte_perm(n3, 0, sel);      // (0,+3) (0,-3)
                        const unsigned e4 = cb - __byte_perm(w_p3, 0, sel), e12 = cb - __byte_perm(w_m3, 0, sel); // (+3,0) (-3,0)
                        const unsigned e2 = cb - __byte_perm(se_, 0, sel), e10 = cb - __byte_perm(nw_, 0, sel);   // (+2,+2) (-2,-2)
                        const unsigned e6 = cb - __byte_perm(ne_, 0, sel), e14 = cb - __byte_perm(sw_, 0, sel);   // (+2,-2) (-2,+2)
                        const unsigned D = __vmins2(__vimin3_s16x2(__vmaxs2(e0, e8), __vmaxs2(e4, e12), __vmaxs2(e2, e10)), __vmaxs2(e6, e14));
                        const unsigned B = __vmaxs2(__vimax3_s16x2(__vmins2(e0, e8), __vmins2(e4, e12), __vmins2(e2, e10)), __vmins2(e6, e14));
                        const unsigned r = __vmaxs2(__vadd2(D, T1), __vadd2(~B, U1));       // >= 0 per half <=> may be a corner
                        const unsigned ok = ~r & 0x80008000u;
                        m |= ((ok >> 15) & 1u) << (2 * hlf) | ((ok >> 31) & 1u) << (2 * hlf + 1);
                    }
                    // pixels of the group outside the cell interior
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col0 + j < 0 || col0 + j >= cw) m &= ~(1u << j);
                }
                const int cnt = __popc(m);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int tot = __shfl_sync(0xffffffffu, inc, 31);
                if (tot) {
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&s_ncand, tot);
                    base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        list[base++] = (uint16_t)(y * pw + col0 + j);
                    }
                }
            }
        }
        __syncthreads();
        const int ncand = s_ncand;
        // B
        for (int i = threadIdx.x; i < ncand; i += FAST_THREADS) {
            const int off = list[i];
            const int m = fast_arc_score(p0 + off, pw);
            if (m > thr) score[off + pw + 1] = (uint8_t)(m - 1);
        }
        __syncthreads();
        // C
        for (int i = threadIdx.x; i < ncand; i += FAST_THREADS) {
            const int off = list[i];
            const uint8_t* q = score + off + pw + 1;
            const int sc = q[0];
            if (sc && sc > q[-pw - 1] && sc > q[-pw] && sc > q[-pw + 1] && sc > q[-1] && sc > q[1] && sc > q[pw - 1] && sc > q[pw] && sc > q[pw + 1])
                atomicOr(&bitmap[off >> 5], 1u << (off & 31));
        }
        __syncthreads();
        // D
        if (wid == 0) {
            int run = 0;
            for (int b0 = 0; b0 < nwords; b0 += 32) {
                const int v = (b0 + lane < nwords) ? __popc(bitmap[b0 + lane]) : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                if (b0 + lane < nwords) woff[b0 + lane] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_total = run;
        }
        __syncthreads();
        if (s_total > 3 || thr == 7) break;
        thr = 7;                 // cellKeyPoints.size() <= 3: clear and retry with the fixed fallback threshold
        __syncthreads();
    }
    // E
    uint32_t* out = d.cand + (size_t)f * d.cand_total + c.cand_off;
    for (int j = threadIdx.x; j < nwords; j += FAST_THREADS) {
        uint32_t w = bitmap[j];
        int pos = woff[j];
        while (w) {
            const int off = j * 32 + __ffs(w) - 1;
            w &= w - 1;
            const int y = off / pw, x = off - y * pw;
            if (pos < c.cand_cap) out[pos] = ((uint32_t)score[off + pw + 1] << 24) | ((uint32_t)(c.y0 + y) << 12) | (uint32_t)(c.x0 + x);
            else *d.err = 1;
            ++pos;
        }
    }
    if (threadIdx.x == 0) { hdr->n_base = s_total; hdr->n_a = s_total; hdr->n_b = s_total; }
}

// ---- TMA (cp.async.bulk.tensor, SASS UTMALDG) + mbarrier helpers
struct FastMaps { CUtensorMap m[MAX_LEVELS]; };   // one 3-D map (x bytes, rows, frames) per pyramid level, box = that level's cell patch
__device__ __forceinline__ unsigned orb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void orb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(orb_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// waits for the phase with the given parity; gives up after ~0.2 s of SM clocks (a copy that never lands - a broken tensor map -
// must end in an error code, not in a hung GPU) and returns false
__device__ __forceinline__ bool orb_mbar_wait(unsigned long long* bar, unsigned parity) {
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(orb_smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > 400000000LL) return false;
    }
}
// one box of the 3-D tensor (x, y, z) -> shared memory; completion is signalled on `bar` with the box's byte count
__device__ __forceinline__ void orb_tma_box3(void* dst_smem, const CUtensorMap* map, int x, int y, int z, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(orb_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(orb_smem_u32(dst_smem)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(orb_smem_u32(bar)) : "memory");
}

// orb_fast_cells with the cell staged by the TMA unit: ONE cp.async.bulk.tensor.3d per CTA pulls the cell and its 3 px apron
// (box fbw x fbh of the level's tensor map, origin at the 16 B aligned column left of the cell's first apron pixel; the patch
// pitch is the level's constant fbw) while the CTA clears its score plane and bitmap;
// the 256 threads no longer issue the ~10 address divisions + LDG + STS each of the staging loop. Pass A walks its
// (row, 4-pixel group) items with an incremental (y, group) pair instead of a division per item, and masks the group's
// pixels outside the cell with two shifts. Passes B-E are those of orb_fast_cells; results are bit-identical.
__global__ void __launch_bounds__(FAST_THREADS) orb_fast_cells_tma(OrbDev d, const __grid_constant__ FastMaps maps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int s_total, s_ncand;
    __shared__ __align__(8) unsigned long long s_bar;
    const CellGeo c = d.cells[blockIdx.x];
    const int f = blockIdx.y + d.frame0;
    CellHdr* hdr = d.hdr + (size_t)f * d.n_cells + blockIdx.x;
    const int cw = c.x1 - c.x0, ch = c.y1 - c.y0;
    if (c.skipped || cw <= 0 || ch <= 0) {
        if (threadIdx.x == 0) { hdr->n_base = 0; hdr->n_a = 0; hdr->n_b = 0; }
        return;
    }
    const LevelGeo& L = d.levels[c.level];
    const int pw = L.fbw, pww = pw >> 2, bh = L.fbh;          // patch pitch = box width
    uint8_t* smem = smem_raw + ((128u - (orb_smem_u32(smem_raw) & 127u)) & 127u);   // TMA destination: 128 B aligned
    const int nbits = pw * ch, nwords = (nbits + 31) >> 5;
    // the box starts at the 16 B aligned bordered-plane column ax0 <= x0-3 (TMA tile coordinates of 1-byte elements must be 16 B
    // aligned in the innermost dimension: an unaligned start raises an illegal-instruction fault, tools/tma_probe.cu)
    const int bx0 = c.x0 - 3 + EDGE, ax0 = bx0 & ~15, shift = bx0 - ax0;
    const int g0 = fastpx::first_group(shift), G = fastpx::groups_per_row(cw, shift), nitems = ch * G;   // pass A work items
    uint8_t* patch = smem;                                                                // [bh x pw], pixel (x,y) of the cell at (y+3)*pw + x+3
    uint8_t* score = smem + ((pw * bh + 15) & ~15);                                       // [(ch+2) x pw], pixel (x,y) at (y+1)*pw + x+1
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(score + ((pw * (ch + 2) + 15) & ~15)); // [nwords], bit y*pw + x
    uint32_t* woff = bitmap + nwords;                                                     // [nwords] exclusive popcount scan
    uint16_t* list = reinterpret_cast<uint16_t*>(woff + nwords);                          // [<= cw*ch] entries y*pw + x
    if (threadIdx.x == 0) orb_mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) orb_tma_box3(patch, &maps.m[c.level], ax0, c.y0 - 3 + EDGE, f, (unsigned)(pw * bh), &s_bar);
    constexpr int NW = FAST_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t* p0 = patch + 3 * pw + 3 + shift;
    // pass A item walk: item = y*G + g; this thread's items are threadIdx.x, +256, +512, ...
    const int y_first = (int)threadIdx.x / G, g_first = (int)threadIdx.x - y_first * G;
    const int dY = FAST_THREADS / G, dG = FAST_THREADS - dY * G;
    int thr = d.fast_th;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < (pw * (ch + 2) + 3) / 4; i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0u;
        for (int i = threadIdx.x; i < nwords; i += FAST_THREADS) bitmap[i] = 0u;
        if (threadIdx.x == 0) s_ncand = 0;
        if (pass == 0 && !orb_mbar_wait(&s_bar, 0)) { *d.err = 4; return; }   // the patch has landed (async-proxy writes are visible after the wait)
        __syncthreads();
        // A: one thread = the 4 pixels of one patch word (group)
        {
            const uint32_t* pwords = reinterpret_cast<const uint32_t*>(patch);
            const unsigned T1 = fastpx::screen_T1(thr), U1 = fastpx::screen_U1(thr);
            int y = y_first, g = g_first;
            for (int it0 = wid * 32; it0 < nitems; it0 += NW * 32) {
                unsigned m = 0;
                const int col0 = fastpx::group_x0(g, shift);       // interior x of the group's first pixel
                if (y < ch) {                                      // <=> it0 + lane < nitems
                    const uint32_t* c = pwords + (y + 3) * pww + g + g0;
                    const uint32_t n3 = c[-3 * pww], s3 = c[3 * pww];
                    const uint32_t n2a = c[-2 * pww - 1], n2b = c[-2 * pww], n2c = c[-2 * pww + 1];
                    const uint32_t s2a = c[2 * pww - 1], s2b = c[2 * pww], s2c = c[2 * pww + 1];
                    const uint32_t za = c[-1], zb = c[0], zc = c[1];
                    // necessary condition (fast_screen.h), then drop the group's pixels outside the cell interior
                    m = fastpx::screen4(n3, s3, n2a, n2b, n2c, s2a, s2b, s2c, za, zb, zc, T1, U1) & fastpx::inside_mask8(col0, cw) & 0xFu;
                }
                const int cnt = __popc(m);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int tot = __shfl_sync(0xffffffffu, inc, 31);
                if (tot) {
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&s_ncand, tot);
                    base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
                    const int e0 = y * pw + col0;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        list[base++] = (uint16_t)(e0 + j);
                    }
                }
                g += dG; y += dY;
                if (g >= G) { g -= G; ++y; }
            }
        }
        __syncthreads();
        const int ncand = s_ncand;
        // B
        for (int i = threadIdx.x; i < ncand; i += FAST_THREADS) {
            const int off = list[i];
            const int m = fast_arc_score(p0 + off, pw);
            if (m > thr) score[off + pw + 1] = (uint8_t)(m - 1);
        }
        __syncthreads();
        // C
        for (int i = threadIdx.x; i < ncand; i += FAST_THREADS) {
            const int off = list[i];
            const uint8_t* q = score + off + pw + 1;
            const int sc = q[0];
            if (sc && sc > q[-pw - 1] && sc > q[-pw] && sc > q[-pw + 1] && sc > q[-1] && sc > q[1] && sc > q[pw - 1] && sc > q[pw] && sc > q[pw + 1])
                atomicOr(&bitmap[off >> 5], 1u << (off & 31));
        }
        __syncthreads();
        // D
        if (wid == 0) {
            int run = 0;
            for (int b0 = 0; b0 < nwords; b0 += 32) {
                const int v = (b0 + lane < nwords) ? __popc(bitmap[b0 + lane]) : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                if (b0 + lane < nwords) woff[b0 + lane] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_total = run;
        }
        __syncthreads();
        if (s_total > 3 || thr == 7) break;
        thr = 7;                 // cellKeyPoints.size() <= 3: clear and retry with the fixed fallback threshold
        __syncthreads();
    }
    // E
    uint32_t* out = d.cand + (size_t)f * d.cand_total + c.cand_off;
    for (int j = threadIdx.x; j < nwords; j += FAST_THREADS) {
        uint32_t w = bitmap[j];
        int pos = woff[j];
        while (w) {
            const int off = j * 32 + __ffs(w) - 1;
            w &= w - 1;
            const int y = off / pw, x = off - y * pw;
            if (pos < c.cand_cap) out[pos] = ((uint32_t)score[off + pw + 1] << 24) | ((uint32_t)(c.y0 + y) << 12) | (uint32_t)(c.x0 + x);
            else *d.err = 1;
            ++pos;
        }
    }
    if (threadIdx.x == 0) { hdr->n_base = s_total; hdr->n_a = s_total; hdr->n_b = s_total; }
}

// orb_fast_cells_tma with a leaner pass A (fast_screen.h, host-tested):
//   * one item = 8 pixels (two patch words): the rows y-3 / y+3 come in as one 64-bit shared load each, rows y-2 / y / y+2 as
//     32+64+32 bits - 11 shared loads per 8 pixels instead of 22 - and the address arithmetic, the survivor scan and the loop
//     control are paid once per 8 pixels;
//   * every warp compacts into its OWN segment of the candidate list (segments are sized by the items a warp owns), so the
//     shared atomic per 128 pixels and its vote/election code are gone; passes B and C walk the warp's own segment (the 32-item
//     chunks are dealt round-robin over the warps, so the segments are balanced). The list order never reaches the output:
//     keypoints are emitted in raster order from the bitmap (pass E), so the result is bit-identical to orb_fast_cells.
template <bool PDL>   // PDL: dependent launch behind the last resize (see orb_resize_w): cell geometry, barrier set-up and the smem pointers run in its tail
__global__ void __launch_bounds__(FAST_THREADS) orb_fast_cells_tma8(OrbDev d, const __grid_constant__ FastMaps maps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int s_total;
    __shared__ __align__(8) unsigned long long s_bar;
    const CellGeo c = d.cells[blockIdx.x];
    const int f = blockIdx.y + d.frame0;
    CellHdr* hdr = d.hdr + (size_t)f * d.n_cells + blockIdx.x;
    const int cw = c.x1 - c.x0, ch = c.y1 - c.y0;
    if (c.skipped || cw <= 0 || ch <= 0) {
        if (threadIdx.x == 0) { hdr->n_base = 0; hdr->n_a = 0; hdr->n_b = 0; }
        return;
    }
    const LevelGeo& L = d.levels[c.level];
    const int pw = L.fbw, pww = pw >> 2, bh = L.fbh;          // patch pitch = box width (multiple of 16 B: rows are 8 B aligned)
    uint8_t* smem = smem_raw + ((128u - (orb_smem_u32(smem_raw) & 127u)) & 127u);   // TMA destination: 128 B aligned
    const int nbits = pw * ch, nwords = (nbits + 31) >> 5;
    // the box starts at the 16 B aligned bordered-plane column ax0 <= x0-3 (TMA tile coordinates of 1-byte elements must be 16 B
    // aligned in the innermost dimension: an unaligned start raises an illegal-instruction fault, tools/tma_probe.cu)
    const int bx0 = c.x0 - 3 + EDGE, ax0 = bx0 & ~15, shift = bx0 - ax0;
    const int h0 = fastpx::first_pair(shift), G2 = fastpx::pairs_per_row(cw, shift), nitems = ch * G2;
    uint8_t* patch = smem;                                                                // [bh x pw], pixel (x,y) of the cell at (y+3)*pw + x+3+shift
    uint8_t* score = smem + ((pw * bh + 15) & ~15);                                       // [(ch+2) x pw], pixel (x,y) at (y+1)*pw + x+1
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(score + ((pw * (ch + 2) + 15) & ~15)); // [nwords], bit y*pw + x
    uint32_t* woff = bitmap + nwords;                                                     // [nwords] exclusive popcount scan
    uint16_t* list = reinterpret_cast<uint16_t*>(woff + nwords);                          // [8 * nitems] entries y*pw + x, one segment per warp
    if (threadIdx.x == 0) orb_mbar_init(&s_bar, 1);
    __syncthreads();
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");     // the pyramid is complete and visible from here on
    if (threadIdx.x == 0) orb_tma_box3(patch, &maps.m[c.level], ax0, c.y0 - 3 + EDGE, f, (unsigned)(pw * bh), &s_bar);
    constexpr int NW = FAST_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t* p0 = patch + 3 * pw + 3 + shift;
    uint16_t* seg = list + fastpx::seg_offset(wid, nitems, NW);
    fastpx::ItemWalk first;
    first.init((int)threadIdx.x, FAST_THREADS, G2);
    int thr = d.fast_th;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < (pw * (ch + 2) + 3) / 4; i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0u;
        for (int i = threadIdx.x; i < nwords; i += FAST_THREADS) bitmap[i] = 0u;
        if (pass == 0 && !orb_mbar_wait(&s_bar, 0)) { *d.err = 4; return; }   // the patch has landed (async-proxy writes are visible after the wait)
        __syncthreads();
        // A: one thread = 8 pixels = patch words 2h, 2h+1 of row y
        int nseg = 0;                                  // candidates in this warp's segment (warp-uniform)
        {
            const uint32_t* pwords = reinterpret_cast<const uint32_t*>(patch);
            const unsigned T1 = fastpx::screen_T1(thr), U1 = fastpx::screen_U1(thr);
            fastpx::ItemWalk it = first;
            for (int it0 = wid * 32; it0 < nitems; it0 += NW * 32) {
                unsigned m = 0;
                const int x0 = fastpx::pair_x0(it.h, shift);       // interior x of the item's first pixel
                if (it.y < ch) {                                   // <=> it0 + lane < nitems
                    const uint32_t* cp = pwords + (it.y + 3) * pww + 2 * (it.h + h0);    // even word index: 8 B aligned
                    const uint2 n3 = *reinterpret_cast<const uint2*>(cp - 3 * pww), s3 = *reinterpret_cast<const uint2*>(cp + 3 * pww);
                    const uint2 n2 = *reinterpret_cast<const uint2*>(cp - 2 * pww), s2 = *reinterpret_cast<const uint2*>(cp + 2 * pww);
                    const uint2 z = *reinterpret_cast<const uint2*>(cp);
                    const uint32_t n2l = cp[-2 * pww - 1], n2r = cp[-2 * pww + 2];
                    const uint32_t s2l = cp[2 * pww - 1], s2r = cp[2 * pww + 2];
                    const uint32_t zl = cp[-1], zr = cp[2];
                    m = fastpx::screen4(n3.x, s3.x, n2l, n2.x, n2.y, s2l, s2.x, s2.y, zl, z.x, z.y, T1, U1) |
                        fastpx::screen4(n3.y, s3.y, n2.x, n2.y, n2r, s2.x, s2.y, s2r, z.x, z.y, zr, T1, U1) << 4;
                    m &= fastpx::inside_mask8(x0, cw);             // pixels of the item outside the cell interior
                }
                const int cnt = __popc(m);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int tot = __shfl_sync(0xffffffffu, inc, 31);
                if (tot) {
                    int base = nseg + inc - cnt;
                    const int e0 = it.y * pw + x0;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        seg[base++] = (uint16_t)(e0 + j);
                    }
                    nseg += tot;
                }
                it.next();
            }
        }
        __syncwarp();
        // B: this warp's segment
        for (int i = lane; i < nseg; i += 32) {
            const int off = seg[i];
            const int m = fast_arc_score(p0 + off, pw);
            if (m > thr) score[off + pw + 1] = (uint8_t)(m - 1);
        }
        __syncthreads();
        // C
        for (int i = lane; i < nseg; i += 32) {
            const int off = seg[i];
            const uint8_t* q = score + off + pw + 1;
            const int sc = q[0];
            if (sc && sc > q[-pw - 1] && sc > q[-pw] && sc > q[-pw + 1] && sc > q[-1] && sc > q[1] && sc > q[pw - 1] && sc > q[pw] && sc > q[pw + 1])
                atomicOr(&bitmap[off >> 5], 1u << (off & 31));
        }
        __syncthreads();
        // D
        if (wid == 0) {
            int run = 0;
            for (int b0 = 0; b0 < nwords; b0 += 32) {
                const int v = (b0 + lane < nwords) ? __popc(bitmap[b0 + lane]) : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                if (b0 + lane < nwords) woff[b0 + lane] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_total = run;
        }
        __syncthreads();
        if (s_total > 3 || thr == 7) break;
        thr = 7;                 // cellKeyPoints.size() <= 3: clear and retry with the fixed fallback threshold
        __syncthreads();
    }
    // E
    uint32_t* out = d.cand + (size_t)f * d.cand_total + c.cand_off;
    for (int j = threadIdx.x; j < nwords; j += FAST_THREADS) {
        uint32_t w = bitmap[j];
        int pos = woff[j];
        while (w) {
            const int off = j * 32 + __ffs(w) - 1;
            w &= w - 1;
            const int y = off / pw, x = off - y * pw;
            if (pos < c.cand_cap) out[pos] = ((uint32_t)score[off + pw + 1] << 24) | ((uint32_t)(c.y0 + y) << 12) | (uint32_t)(c.x0 + x);
            else *d.err = 1;
            ++pos;
        }
    }
    if (threadIdx.x == 0) { hdr->n_base = s_total; hdr->n_a = s_total; hdr->n_b = s_total; }
}

// FAST-9-16 arc score M = max over the 16 nine-pixel arcs of min(+-(v - ring)); corner at t <=> M > t,
// cv::FAST's stored score == M-1 whatever t was. Returns 0 early when the pixel cannot be a corner at t:
// a 9-arc always contains one pixel of each opposite pair (k, k+8), so all 4 tested pairs must have a
// member beyond +-t of the centre.
__device__ __forceinline__ int fast_arc_score_qr(const uint8_t* __restrict__ p, int pw, int t) {
    const int v = p[0];
    const int c0 = v - p[3 * pw], c8 = v - p[-3 * pw];
    bool dk = (c0 > t) | (c8 > t), br = (c0 < -t) | (c8 < -t);
    if (!(dk | br)) return 0;
    const int c4 = v - p[3], c12 = v - p[-3];
    dk &= (c4 > t) | (c12 > t); br &= (c4 < -t) | (c12 < -t);
    if (!(dk | br)) return 0;
    const int c2 = v - p[2 * pw + 2], c10 = v - p[-2 * pw - 2];
    dk &= (c2 > t) | (c10 > t); br &= (c2 < -t) | (c10 < -t);
    if (!(dk | br)) return 0;
    const int c6 = v - p[-2 * pw + 2], c14 = v - p[2 * pw - 2];
    dk &= (c6 > t) | (c14 > t); br &= (c6 < -t) | (c14 < -t);
    if (!(dk | br)) return 0;
    // both polarities at once: each ring difference d is packed as the s16x2 pair (d, -d); an arc's score for the
    // "darker" / "brighter" test is the min over its 9 packed entries, computed with Blackwell's 3-input packed min
    // (VIMNMX3.S16x2): m3_k = min(q_k,q_k+1,q_k+2), m9_k = min(m3_k, m3_k+3, m3_k+6); M = max over k and both halves.
#define SE2_PK(dv) ((static_cast<unsigned>(dv) & 0xFFFFu) | (static_cast<unsigned>(-(dv)) << 16))
    unsigned q[16];
    q[0] = SE2_PK(c0);   q[1] = SE2_PK(v - p[3 * pw + 1]);    q[2] = SE2_PK(c2);    q[3] = SE2_PK(v - p[pw + 3]);
    q[4] = SE2_PK(c4);   q[5] = SE2_PK(v - p[-pw + 3]);       q[6] = SE2_PK(c6);    q[7] = SE2_PK(v - p[-3 * pw + 1]);
    q[8] = SE2_PK(c8);   q[9] = SE2_PK(v - p[-3 * pw - 1]);   q[10] = SE2_PK(c10);  q[11] = SE2_PK(v - p[-pw - 3]);
    q[12] = SE2_PK(c12); q[13] = SE2_PK(v - p[pw - 3]);       q[14] = SE2_PK(c14);  q[15] = SE2_PK(v - p[3 * pw - 1]);
#undef SE2_PK
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(q[k], q[(k + 1) & 15], q[(k + 2) & 15]);
    unsigned best = 0x80008000u;   // (-32768, -32768)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a9 = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned b9 = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a9, b9);
    }
    const int mdark = static_cast<short>(best & 0xFFFFu), mbright = static_cast<short>(best >> 16);
    return max(mdark, mbright);
}

// Fallback for frames whose grid cells are too large for orb_fast_cells' shared-memory candidate list (e.g. 1080p with
// 1000 features: 470 x 150 px cells): one thread per pixel, no compaction, 2.2 bytes of shared memory per pixel.
// one CTA per (cell, frame): cv::FAST(cell, fastTh, NMS) and, if that yields <= 3 keypoints, cv::FAST(cell, 7, NMS)
// (ORBextractor.cpp:616-623). The cell and its 3 px apron are staged once in shared memory with aligned 32-bit loads.
// Work unit = one 32-pixel row segment per warp iteration (no per-pixel division); segments are numbered in raster
// order, so one exclusive scan over the per-segment survivor counts gives every keypoint its raster-order slot.
__global__ void __launch_bounds__(FAST_THREADS) orb_fast_cells_big(OrbDev d) {
    extern __shared__ uint8_t smem[];
    __shared__ int s_total;
    const CellGeo c = d.cells[blockIdx.x];
    const int f = blockIdx.y + d.frame0;
    CellHdr* hdr = d.hdr + (size_t)f * d.n_cells + blockIdx.x;
    const int cw = c.x1 - c.x0, ch = c.y1 - c.y0;
    if (c.skipped || cw <= 0 || ch <= 0) {
        if (threadIdx.x == 0) { hdr->n_base = 0; hdr->n_a = 0; hdr->n_b = 0; }
        return;
    }
    const LevelGeo& L = d.levels[c.level];
    const uint8_t* plane = d.plain + f * d.frame_plane_bytes + L.plane_off;
    // patch columns start at the 4-aligned bordered-plane column ax0 <= x0-3 (plane base and pitch are 32 B aligned)
    const int bx0 = c.x0 - 3 + EDGE, by0 = c.y0 - 3 + EDGE;
    const int ax0 = bx0 & ~3, shift = bx0 - ax0;
    const int ph = ch + 6, sw = cw + 2, sh = ch + 2;
    const int pww = (shift + cw + 6 + 3) >> 2, pw = pww * 4;
    uint8_t* patch = smem;
    uint8_t* score = smem + ((pw * ph + 15) & ~15);
    const int nseg = (cw + 31) >> 5, nchunk = ch * nseg;
    int* cnt = reinterpret_cast<int*>(score + ((sw * sh + 15) & ~15));   // [nchunk] survivors per segment -> exclusive offsets
    for (int i = threadIdx.x; i < pww * ph; i += FAST_THREADS) {
        const int py = i / pww, pxw = i - py * pww;
        reinterpret_cast<uint32_t*>(patch)[i] = *reinterpret_cast<const uint32_t*>(plane + (size_t)(by0 + py) * L.pitch + ax0 + 4 * pxw);
    }
    constexpr int NW = FAST_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t* p0 = patch + 3 * pw + 3 + shift;
    int thr = d.fast_th;
    unsigned long long keepmask = 0;   // bit t: this lane's pixel of the warp's t-th segment survived NMS (first 64 segments)
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < (sw * sh + 3) / 4; i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0u;
        __syncthreads();
        for (int cidx = wid; cidx < nchunk; cidx += NW) {
            const int y = cidx / nseg, x = (cidx - y * nseg) * 32 + lane;
            if (x < cw) {
                const int m = fast_arc_score_qr(p0 + y * pw + x, pw, thr);
                if (m > thr) score[(y + 1) * sw + (x + 1)] = (uint8_t)(m - 1);
            }
        }
        __syncthreads();
        keepmask = 0;
        int t = 0;
        for (int cidx = wid; cidx < nchunk; cidx += NW, ++t) {
            const int y = cidx / nseg, x = (cidx - y * nseg) * 32 + lane;
            bool keep = false;
            if (x < cw) {
                const uint8_t* q = score + (y + 1) * sw + (x + 1);
                const int s = q[0];
                if (s) keep = s > q[-sw - 1] && s > q[-sw] && s > q[-sw + 1] && s > q[-1] && s > q[1] && s > q[sw - 1] && s > q[sw] && s > q[sw + 1];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) cnt[cidx] = __popc(bal);
            if (keep && t < 64) keepmask |= 1ull << t;
        }
        __syncthreads();
        if (wid == 0) {   // exclusive scan of the per-segment counts, in raster order
            int run = 0;
            for (int b0 = 0; b0 < nchunk; b0 += 32) {
                const int v = (b0 + lane < nchunk) ? cnt[b0 + lane] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                if (b0 + lane < nchunk) cnt[b0 + lane] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_total = run;
        }
        __syncthreads();
        if (s_total > 3 || thr == 7) break;
        thr = 7;                 // cellKeyPoints.size() <= 3: clear and retry with the fixed fallback threshold
        __syncthreads();
    }
    uint32_t* out = d.cand + (size_t)f * d.cand_total + c.cand_off;
    int t = 0;
    for (int cidx = wid; cidx < nchunk; cidx += NW, ++t) {
        const int y = cidx / nseg, x = (cidx - y * nseg) * 32 + lane;
        bool keep;
        int s = 0;
        if (t < 64) {
            keep = (keepmask >> t) & 1ull;
        } else {
            keep = false;
            if (x < cw) {
                const uint8_t* q = score + (y + 1) * sw + (x + 1);
                s = q[0];
                if (s) keep = s > q[-sw - 1] && s > q[-sw] && s > q[-sw + 1] && s > q[-1] && s > q[1] && s > q[sw - 1] && s > q[sw] && s > q[sw + 1];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            s = score[(y + 1) * sw + (x + 1)];
            const int pos = cnt[cidx] + __popc(bal & ((1u << lane) - 1));
            if (pos < c.cand_cap) out[pos] = ((uint32_t)s << 24) | ((uint32_t)(c.y0 + y) << 12) | (uint32_t)(c.x0 + x);
            else *d.err = 1;
        }
    }
    if (threadIdx.x == 0) { hdr->n_base = s_total; hdr->n_a = s_total; hdr->n_b = s_total; }
}

// one CTA per (level, frame): quota redistribution, retainBest per cell and per level. The cells' candidate lists
// are staged in shared memory when the level's total fits (SEL_STAGE entries), otherwise they are processed in place
// in global memory. retainBest is std::nth_element; one warp per cell runs the warp-cooperative introselect of
// introselect.h (same permutation as libstdc++'s, 32 elements per step), cells round-robin over the CTA's warps.
constexpr int SEL_STAGE = 12288;
constexpr int SEL_THREADS = 512;
__global__ void __launch_bounds__(SEL_THREADS) orb_select(OrbDev d) {
    extern __shared__ uint32_t sbuf[];   // [cap] level list | [4*nCells] ints | [SEL_STAGE] staged candidates
    __shared__ int wq[SEL_THREADS / 32][128];
    constexpr int NW = SEL_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int level = blockIdx.x, f = blockIdx.y + d.frame0;
    const LevelGeo& L = d.levels[level];
    const int nCells = L.nCells;
    const int cap = 2 * L.nDesired + 4 * nCells + 64;
    uint32_t* lbuf = sbuf;
    int* nTotal = (int*)(sbuf + cap);
    int* nToRetain = nTotal + nCells;
    int* kept = nToRetain + nCells;
    int* koff = kept + nCells;
    uint32_t* stage = (uint32_t*)(koff + nCells);
    __shared__ int s_total, s_staged;
    const CellHdr* hdr = d.hdr + (size_t)f * d.n_cells + L.cell_base;
    uint32_t* cand = d.cand + (size_t)f * d.cand_total;
    for (int c = threadIdx.x; c < nCells; c += SEL_THREADS) nTotal[c] = d.cells[L.cell_base + c].skipped ? 0 : hdr[c].n_base;
    __syncthreads();
    if (threadIdx.x == 0) {
        // :625-679 (skipped cells never enter the first pass: nToRetain/nTotal stay 0, bNoMore stays false)
        const int nfeaturesCell = L.nfeaturesCell;
        int nNoMore = 0, nToDistribute = 0, tot = 0;
        for (int c = 0; c < nCells; ++c) {
            const bool skipped = d.cells[L.cell_base + c].skipped;
            koff[c] = 0;  // bNoMore
            tot += nTotal[c];
            if (skipped) { nToRetain[c] = 0; continue; }
            if (nTotal[c] > nfeaturesCell) nToRetain[c] = nfeaturesCell;
            else { nToRetain[c] = nTotal[c]; nToDistribute += nfeaturesCell - nTotal[c]; koff[c] = 1; nNoMore++; }
        }
        while (nToDistribute > 0 && nNoMore < nCells) {
            const int nNew = (int)((float)nfeaturesCell + ceilf((float)nToDistribute / (float)(nCells - nNoMore)));
            nToDistribute = 0;
            for (int c = 0; c < nCells; ++c)
                if (!koff[c]) {
                    if (nTotal[c] > nNew) nToRetain[c] = nNew;
                    else { nToRetain[c] = nTotal[c]; nToDistribute += nNew - nTotal[c]; koff[c] = 1; nNoMore++; }
                }
        }
        s_staged = tot <= SEL_STAGE;
        int o = 0;
        for (int c = 0; c < nCells; ++c) { kept[c] = o; o += nTotal[c]; }   // staging offsets (reused below)
    }
    __syncthreads();
    const bool staged = s_staged;
    for (int c = wid; c < nCells; c += NW) {
        const int n = nToRetain[c], tot = nTotal[c];
        uint32_t* g = cand + d.cells[L.cell_base + c].cand_off;
        uint32_t* v = g;
        if (staged) {
            v = stage + kept[c];
            for (int i = lane; i < tot; i += 32) v[i] = g[i];
            __syncwarp();
        }
        if (lane == 0) koff[c] = (int)(v - (staged ? stage : cand));   // remember where the list lives
        if (tot > n && n > 0) se2gpu::kp_nth_element_warp(v, tot, n - 1, wq[wid]);   // KeyPointsFilter::retainBest + resize (:692-694)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int c = 0; c < nCells; ++c) { const int k = min(nTotal[c], nToRetain[c]); nTotal[c] = k; kept[c] = o; o += k; }
        if (o > cap) { *d.err = 2; o = cap; }
        s_total = o;
    }
    __syncthreads();
    for (int c = wid; c < nCells; c += NW) {
        const uint32_t* v = (staged ? stage : cand) + koff[c];
        const int o = kept[c];
        for (int i = lane; i < nTotal[c] && o + i < cap; i += 32) lbuf[o + i] = v[i];
    }
    __syncthreads();
    int total = s_total;
    if (total > L.nDesired) {  // :706-710
        if (wid == 0) se2gpu::kp_nth_element_warp(lbuf, total, L.nDesired - 1, wq[0]);
        total = L.nDesired;
        __syncthreads();
    }
    uint32_t* out = d.lkp + (size_t)f * d.lkp_total + L.kp_off;
    for (int i = threadIdx.x; i < total; i += SEL_THREADS) out[i] = lbuf[i];
    if (threadIdx.x == 0) d.lcount[f * d.nlevels + level] = total;
}

// test hook: one warp per list, lists in global memory
__global__ void __launch_bounds__(128) orb_debug_nth(uint32_t* v, const int* __restrict__ offs, const int* __restrict__ nth, int count) {
    __shared__ int wq[4][128];
    const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= count) return;
    se2gpu::kp_nth_element_warp(v + offs[k], offs[k + 1] - offs[k], nth[k], wq[threadIdx.x >> 5]);
}

// GaussianBlur 7x7 sigma 2 on the level ROI; the 16 px ring keeps its un-blurred reflect-101 copies.
// cv::GaussianBlur's float arithmetic: row pass = sequential fmaf over the 7 taps, column pass = centre tap then the
// three symmetric pairs, round-to-nearest-even, saturate. No shared memory: a tile is one warp = 128 columns x
// BLUR_TH output rows; a thread owns 4 adjacent columns (one output word) and walks down its strip with the last 7
// row-pass results in a register ring, so every source word is loaded once per row (3 coalesced words per thread:
// its own and both neighbours, the latter L1 hits) and every output row costs one 32-bit store.
__global__ void __launch_bounds__(256) orb_blur(OrbDev d, int tile0, int tile_end) {
    const int tile = tile0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= tile_end) return;
    const int lane = threadIdx.x & 31;
    const TileGeo t = d.tiles[tile];
    const int f = blockIdx.y + d.frame0;
    const int pitch = d.levels[t.level].pitch, Lw = d.levels[t.level].w, Lh = d.levels[t.level].h, H = Lh + 2 * EDGE;
    const int x4 = t.x0 + 4 * lane;
    if (x4 >= pitch) return;
    const size_t plane_off = f * d.frame_plane_bytes + d.levels[t.level].plane_off + x4;
    const uint8_t* __restrict__ src = d.plain + plane_off;
    uint8_t* __restrict__ dst = d.blurred + plane_off;
    const bool has_l = x4 >= 4, has_r = x4 + 8 <= pitch;
    uint32_t colmask = 0;          // 0xFF in the byte lanes of ROI columns
#pragma unroll
    for (int q = 0; q < 4; ++q) colmask |= (x4 + q >= EDGE && x4 + q < EDGE + Lw) ? (0xFFu << (8 * q)) : 0u;
    const float g0 = c_gauss[0], g1 = c_gauss[1], g2 = c_gauss[2], g3 = c_gauss[3], g4 = c_gauss[4], g5 = c_gauss[5], g6 = c_gauss[6];
    float hr[7][4];      // row-pass results of the last 7 source rows
    uint32_t cw[7];      // their centre words (ring pixels are copied through)
    // rows outside the plane only feed ring outputs (copies), so the row index is clamped instead of tested
    auto row_ptr = [&](int i) { return src + (size_t)min(max(t.y0 - 3 + i, 0), H - 1) * pitch; };
    uint32_t n0, n1, n2;  // words of the next source row (software prefetch, one row ahead)
    {
        const uint8_t* rp = row_ptr(0);
        n1 = __ldg(reinterpret_cast<const uint32_t*>(rp));
        n0 = has_l ? __ldg(reinterpret_cast<const uint32_t*>(rp - 4)) : 0u;
        n2 = has_r ? __ldg(reinterpret_cast<const uint32_t*>(rp + 4)) : 0u;
    }
    for (int i0 = 0; i0 < BLUR_TH + 6; i0 += 7) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int i = i0 + k;                       // source row t.y0 - 3 + i, kept in ring slot k
            const uint32_t w0 = n0, w1 = n1, w2 = n2;
            {
                const uint8_t* rp = row_ptr(i + 1);
                n1 = __ldg(reinterpret_cast<const uint32_t*>(rp));
                n0 = has_l ? __ldg(reinterpret_cast<const uint32_t*>(rp - 4)) : 0u;
                n2 = has_r ? __ldg(reinterpret_cast<const uint32_t*>(rp + 4)) : 0u;
            }
            float b[10];                                // bytes x4-3 .. x4+6
#pragma unroll
            for (int j = 0; j < 3; ++j) b[j] = (float)((w0 >> (8 * (j + 1))) & 255u);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[3 + j] = (float)((w1 >> (8 * j)) & 255u);
#pragma unroll
            for (int j = 0; j < 3; ++j) b[7 + j] = (float)((w2 >> (8 * j)) & 255u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float s = __fmul_rn(g0, b[q]);
                s = __fmaf_rn(g1, b[q + 1], s); s = __fmaf_rn(g2, b[q + 2], s); s = __fmaf_rn(g3, b[q + 3], s);
                s = __fmaf_rn(g4, b[q + 4], s); s = __fmaf_rn(g5, b[q + 5], s); s = __fmaf_rn(g6, b[q + 6], s);
                hr[k][q] = s;
            }
            cw[k] = w1;
            const int gy = t.y0 + i - 6;                // output row whose 7-row window ends at source row i
            if (i >= 6 && gy < H && gy < t.y0 + BLUR_TH) {
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float sacc = __fmul_rn(g3, hr[(k + 4) % 7][q]);
                    sacc = __fmaf_rn(g4, __fadd_rn(hr[(k + 5) % 7][q], hr[(k + 3) % 7][q]), sacc);
                    sacc = __fmaf_rn(g5, __fadd_rn(hr[(k + 6) % 7][q], hr[(k + 2) % 7][q]), sacc);
                    sacc = __fmaf_rn(g6, __fadd_rn(hr[k][q], hr[(k + 1) % 7][q]), sacc);
                    word |= (uint32_t)min(max(__float2int_rn(sacc), 0), 255) << (8 * q);
                }
                const uint32_t m = (gy >= EDGE && gy < EDGE + Lh) ? colmask : 0u;
                *reinterpret_cast<uint32_t*>(dst + (size_t)gy * pitch) = (word & m) | (cw[(k + 4) % 7] & ~m);   // centre word = source row i-3
            }
        }
    }
}

// cv::fastAtan2 (degrees), scalar float path without contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// one warp per output keypoint slot: orientation on the plain level, descriptor on the blurred level.
// BATCH: the disc column of a lane is read in three batches of 5 row pairs with all 10 loads of a batch issued before their
// sums (predicated off beyond the column's half-height) instead of one dependent +-v pair per loop trip: the kernel is bound by
// the latency of first-touch sectors (every plane byte comes from DRAM once), so the loads in flight per warp set its speed.
// Integer moments: any summation order gives the reference's m10, m01.
template <bool BATCH>
__global__ void __launch_bounds__(256) orb_orient_describe(OrbDev d, se2gpu_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                           int* __restrict__ counts) {
    __shared__ float4 patf[256];     // the 256 point pairs of the rBRIEF pattern as floats (x0, y0, x1, y1); test 8*lane+k at [k][lane]
    for (int i = threadIdx.x; i < 256; i += blockDim.x) patf[(i & 7) * 32 + (i >> 3)] = make_float4((float)d_pattern[4 * i], (float)d_pattern[4 * i + 1], (float)d_pattern[4 * i + 2], (float)d_pattern[4 * i + 3]);
    __syncthreads();
    const int f = blockIdx.y + d.frame0;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int* lc = d.lcount + f * d.nlevels;
    int level = -1, off = 0, total = 0;
    for (int l = 0; l < d.nlevels; ++l) {
        const int c = lc[l];
        if (level < 0 && slot < total + c) { level = l; off = total; }
        total += c;
    }
    if (slot == 0 && lane == 0) counts[f] = total;
    if (level < 0) return;
    const LevelGeo& L = d.levels[level];
    const uint32_t rec = d.lkp[(size_t)f * d.lkp_total + L.kp_off + (slot - off)];
    const int x = rec & 0xFFF, y = (rec >> 12) & 0xFFF, score = rec >> 24;
    const size_t base = f * d.frame_plane_bytes + L.plane_off + (size_t)(EDGE + y) * L.pitch + (EDGE + x);
    const uint8_t* center = d.plain + base;
    // IC_Angle: lane = column u = lane-15 of the 31x31 patch disc; it walks its rows in +-v pairs up to the column's
    // half-height (integer moments: any summation order gives the reference's m10, m01)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - HALF_PATCH, vm = c_vmax[abs(u)];
        const uint8_t* col = center + u;
        const int pitch = L.pitch;
        int colsum = col[0];
        if (BATCH) {
#pragma unroll
            for (int v0 = 1; v0 <= HALF_PATCH; v0 += 5) {
                int p[5], m[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {       // rows beyond the half-height re-read row vm (a cache hit) and are dropped below
                    const int off = min(v0 + k, vm) * pitch;
                    p[k] = col[off]; m[k] = col[-off];
                }
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (v0 + k <= vm) { colsum += p[k] + m[k]; m01 += (v0 + k) * (p[k] - m[k]); }
            }
        } else {
            for (int v = 1, off = pitch; v <= vm; ++v, off += pitch) {
                const int p = col[off], m = col[-off];
                colsum += p + m;
                m01 += v * (p - m);
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // steered BRIEF: a = (float)cos((double)angle_rad), b = (float)sin(...)
    const float factorPI = (float)(M_PI / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    double sd, cd;
    sincos((double)ang, &sd, &cd);
    const float a = (float)cd, b = (float)sd;
    const uint8_t* bc = d.blurred + base;
    unsigned val = 0;
    if (BATCH) {   // all 16 tap offsets first, then the 16 loads back to back, then the comparisons
        int o0[8], o1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 pt = patf[k * 32 + lane];
            const float x0 = pt.x, y0 = pt.y, x1 = pt.z, y1 = pt.w;
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            o0[k] = r0 * L.pitch + c0; o1[k] = r1 * L.pitch + c1;
        }
        int t0[8], t1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { t0[k] = bc[o0[k]]; t1[k] = bc[o1[k]]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) val |= (unsigned)(t0[k] < t1[k]) << k;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 pt = patf[k * 32 + lane];
            const float x0 = pt.x, y0 = pt.y, x1 = pt.z, y1 = pt.w;
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = bc[(ptrdiff_t)r0 * L.pitch + c0], t1 = bc[(ptrdiff_t)r1 * L.pitch + c1];
            val |= (unsigned)(t0 < t1) << k;
        }
    }
    // pack 4 bytes per lane group: lane 4q gets bytes 4q..4q+3
    unsigned w = val;
    w |= __shfl_down_sync(0xffffffffu, val, 1) << 8;
    w |= __shfl_down_sync(0xffffffffu, val, 2) << 16;
    w |= __shfl_down_sync(0xffffffffu, val, 3) << 24;
    const size_t oslot = (size_t)f * d.nfeatures + slot;
    if ((lane & 3) == 0) reinterpret_cast<unsigned*>(desc + oslot * 32)[lane >> 2] = w;
    if (lane == 0) {
        se2gpu_keypoint kp;
        kp.x = level ? __fmul_rn((float)x, L.scale) : (float)x;
        kp.y = level ? __fmul_rn((float)y, L.scale) : (float)y;
        kp.size = L.kp_size; kp.angle = angle; kp.response = (float)score; kp.octave = level; kp.class_id = -1;
        kps[oslot] = kp;
    }
}

inline int cv_round_f(float v) { return (int)lrintf(v); }

}  // namespace

// =================================================================================================
struct se2gpu_orb {
    int device = 0, nfeatures = 0, nlevels = 0, fast_th = 20, max_w = 0, max_h = 0, max_batch = 0;
    double scaleFactor = 1.2;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<int> mnFeaturesPerLevel;
    // geometry of the current frame size
    int cur_w = -1, cur_h = -1;
    std::vector<LevelGeo> levels;
    std::vector<CellGeo> cells;
    std::vector<TileGeo> tiles;
    size_t fast_smem = 0, select_smem = 0, resize_smem = 0, resize_w_smem = 0;
    int resize_rows = 0, resize_raw_pitch = 0;   // shared-memory box of orb_resize, sized from the scale factor
    bool fast_big = false;       // cells too large for the compacting FAST kernel: use orb_fast_cells_big
    int fast_tma = 0;            // cells staged by the TMA unit: 0 off, 1 orb_fast_cells_tma, 2 orb_fast_cells_tma8 (SE2GPU_ORB_FAST_TMA)
    size_t fast_tma_smem = 0;
    FastMaps fast_maps{};        // kernel parameter (__grid_constant__): one tensor map per level over d.plain
    // optional lens undistortion folded into level 0 (se2gpu_orb_set_undistort)
    bool und_on = false;
    float und_K[9] = {}, und_D[14] = {};
    int und_nd = 0, und_w = 0, und_h = 0;
    short2* d_und_m1 = nullptr; uint16_t* d_und_m2 = nullptr; size_t und_cap = 0;
    // capacities (computed for max_w x max_h)
    size_t cap_plane = 0, cap_cand = 0, cap_cells = 0, cap_tiles = 0, cap_tab = 0, cap_lkp = 0;
    OrbDev d{};
    LevelGeo* d_levels = nullptr; CellGeo* d_cells = nullptr; TileGeo* d_tiles = nullptr; int* d_itab = nullptr; short* d_stab = nullptr;

    uint8_t* d_in = nullptr; se2gpu_keypoint* d_kps = nullptr; uint8_t* d_desc = nullptr; int* d_counts = nullptr;
    std::vector<void*> bufs;
    int last_n = 0;
    se2gpu::Profiler prof;
    cudaStream_t side = nullptr;          // blur runs here, concurrently with FAST + selection
    cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr, ev_l1 = nullptr;
    // host-buffer path: two pipeline lanes (stream + side stream + events) so that the H2D of chunk c+1 and the D2H of
    // chunk c-1 overlap the kernels of chunk c
    cudaStream_t pipe[ORB_LANES] = {};   // host-path pipeline lanes (chunk k runs on lane k % lanes)
    // pinned staging for results when the caller's buffers are pageable (a D2H copy into pageable memory blocks the
    // host and would serialise the pipeline)
    int* pin_counts = nullptr; se2gpu_keypoint* pin_kps = nullptr; uint8_t* pin_desc = nullptr;
    // a host-buffer job that has been enqueued but not yet finished (se2gpu_orb_submit / _wait; se2gpu_orb_extract = both)
    struct Pending { bool on = false; int n = 0, lanes = 0; bool pipelined = false; se2gpu_keypoint* kps = nullptr; uint8_t* desc = nullptr; int* counts = nullptr;
                     se2gpu_keypoint* out_kps = nullptr; uint8_t* out_desc = nullptr; int* out_counts = nullptr; } pend;
    // asynchronous two-deep submission: a second, identical context so that two batches can be in flight (the copy engines move
    // batch k+1 in and batch k-1 out while the SMs work on batch k)
    se2gpu_orb* twin = nullptr;
    int submit_next = 0, wait_next = 0, in_flight = 0;
    int create_args[8] = {};
};

namespace {

// level geometry exactly as ORBextractor::ComputePyramid / ComputeKeyPoints derive it (float32 arithmetic)
int build_geometry(se2gpu_orb* h, int w, int hgt, bool dry, size_t* plane_bytes, size_t* cand_total, size_t* n_cells,
                   size_t* n_tiles, size_t* tab_total, size_t* max_fast_smem, size_t* max_select_smem,
                   std::vector<LevelGeo>* Lv, std::vector<CellGeo>* Cv, std::vector<TileGeo>* Tv, bool* fast_big_out = nullptr,
                   size_t* fast_tma_smem_out = nullptr) {
    (void)dry;
    const int nl = h->nlevels;
    std::vector<LevelGeo> L(nl);
    std::vector<CellGeo> C;
    std::vector<TileGeo> T;
    size_t poff = 0, coff = 0, toff = 0, fsm = 0, ssm = 0, fsm_big = 0, fsm_tma = 0;
    bool fast_big = false;   // some cell is too large for orb_fast_cells' shared-memory candidate list -> orb_fast_cells_big
    bool tma_ok = true;      // every level's cell patch fits a TMA box (<= 256 x 256) and 16-bit list offsets
    int kp_off = 0;
    const float imageRatio = (float)w / (float)hgt;   // mvImagePyramid[0].cols/rows (:538)
    for (int l = 0; l < nl; ++l) {
        LevelGeo& g = L[l];
        const float scale = h->mvInvScaleFactor[l];
        g.w = cv_round_f((float)w * scale); g.h = cv_round_f((float)hgt * scale);   // :794-795
        if (g.w < 2 * EDGE + 7 || g.h < 2 * EDGE + 7) return fail(SE2GPU_ERR_INVALID, "level %d of a %dx%d frame is %dx%d: too small for the 16 px border", l, w, hgt, g.w, g.h);
        if (g.w + 2 * EDGE > 4095 || g.h + 2 * EDGE > 4095) return fail(SE2GPU_ERR_CAPACITY, "frames wider/taller than 4063 px are not supported");
        g.pitch = (g.w + 2 * EDGE + 31) & ~31;
        g.plane_off = poff;
        poff += (size_t)g.pitch * (g.h + 2 * EDGE);
        poff = (poff + 255) & ~(size_t)255;
        g.nDesired = h->mnFeaturesPerLevel[l];
        g.cols = (int)sqrtf((float)g.nDesired / (5 * imageRatio));   // :542
        g.rows = (int)(imageRatio * g.cols);                          // :543
        if (g.cols < 1 || g.rows < 1) return fail(SE2GPU_ERR_INVALID, "level %d has a %dx%d cell grid (nfeatures too small): undefined in the reference", l, g.cols, g.rows);
        const int minB = EDGE, maxBX = g.w - EDGE, maxBY = g.h - EDGE;
        const int W = maxBX - minB, H = maxBY - minB;
        const int cellW = (int)ceilf((float)W / g.cols), cellH = (int)ceilf((float)H / g.rows);
        g.nCells = g.rows * g.cols;
        g.nfeaturesCell = (int)ceilf((float)g.nDesired / g.nCells);
        g.cell_base = (int)C.size();
        g.kp_off = kp_off; g.kp_cap = g.nDesired; kp_off += g.nDesired;
        g.scale = h->mvScaleFactor[l];
        g.kp_size = (float)(int)(PATCH * h->mvScaleFactor[l]);
        g.tab_off = (int)toff; toff += (size_t)g.w + g.h;
        std::vector<int> iniXCol(g.cols, 0);
        const size_t first_cell = C.size();
        int max_cw = 0, max_ch = 0;
        float hY = cellH + 6;
        for (int i = 0; i < g.rows; ++i) {
            const float iniY = minB + i * cellH - 3;
            bool rowSkipped = false;
            if (i == g.rows - 1) { hY = maxBY + 3 - iniY; if (hY <= 0) rowSkipped = true; }
            float hX = cellW + 6;
            for (int j = 0; j < g.cols; ++j) {
                CellGeo c{};
                c.level = l;
                if (rowSkipped) { c.skipped = 1; C.push_back(c); continue; }
                float iniX;
                if (i == 0) { iniX = minB + j * cellW - 3; iniXCol[j] = (int)iniX; } else iniX = iniXCol[j];
                if (j == g.cols - 1) { hX = maxBX + 3 - iniX; if (hX <= 0) { c.skipped = 1; C.push_back(c); continue; } }
                const int r0 = (int)iniY, r1 = (int)(iniY + hY), c0 = (int)iniX, c1 = (int)(iniX + hX);
                if (r1 > g.h || c1 > g.w || r0 < 0 || c0 < 0) return fail(SE2GPU_ERR_INVALID, "cell grid of level %d leaves the image (the reference asserts here)", l);
                c.x0 = c0 + 3; c.x1 = c1 - 3; c.y0 = r0 + 3; c.y1 = r1 - 3;
                const int cw = std::max(c.x1 - c.x0, 0), chh = std::max(c.y1 - c.y0, 0);
                c.cand_off = (int)coff;
                c.cand_cap = ((cw + 1) / 2) * ((chh + 1) / 2) + 8;   // strict 3x3 maxima: at most one per 2x2 block
                coff += c.cand_cap;
                if (cw > 0 && chh > 0) {
                    max_cw = std::max(max_cw, ((c.x0 - 3 + EDGE) & 15) + cw); max_ch = std::max(max_ch, chh);   // incl. the TMA alignment shift
                    const size_t pwb = (size_t)((cw + 6 + 3 + 3) / 4 + 1) * 4;   // worst-case alignment shift
                    const size_t nwords = (pwb * chh + 31) / 32;
                    if (pwb * (chh + 6) > 65535) fast_big = true;    // 16-bit patch offsets in the candidate list
                    fsm = std::max(fsm, ((pwb * (chh + 6) + 15) & ~(size_t)15) + ((pwb * (chh + 2) + 15) & ~(size_t)15) + nwords * 8 + (size_t)cw * chh * 2 + 64);
                    const size_t nchunk = (size_t)chh * ((cw + 31) / 32);
                    fsm_big = std::max(fsm_big, ((pwb * (chh + 6) + 15) & ~(size_t)15) + (((size_t)(cw + 2) * (chh + 2) + 15) & ~(size_t)15) + nchunk * 4 + 64);
                }
                C.push_back(c);
            }
        }
        // orb_fast_cells_tma: one TMA box per level = its widest x tallest cell patch (3 px apron), rows a multiple of 16 bytes
        g.fbw = (max_cw + 6 + 15) & ~15; g.fbh = max_ch + 6;
        if (max_cw > 0) {
            if (g.fbw > 256 || g.fbh > 256 || (size_t)g.fbw * g.fbh > 65535) tma_ok = false;
            for (size_t ci = first_cell; ci < C.size(); ++ci) {
                const int cw = C[ci].x1 - C[ci].x0, chh = C[ci].y1 - C[ci].y0;
                if (C[ci].skipped || cw <= 0 || chh <= 0) continue;
                const size_t pw = (size_t)g.fbw, nwords = (pw * chh + 31) / 32;
                // candidate list: 8 entries per (row, 8-pixel pair) item of orb_fast_cells_tma8 (>= cw*chh entries of orb_fast_cells_tma)
                const size_t list_bytes = (size_t)16 * chh * fastpx::pairs_per_row(cw, (C[ci].x0 - 3 + EDGE) & 15);
                fsm_tma = std::max(fsm_tma, (size_t)128 + ((pw * g.fbh + 15) & ~(size_t)15) + ((pw * (chh + 2) + 15) & ~(size_t)15) + nwords * 8 + list_bytes + 64);
            }
        }
        ssm = std::max(ssm, (size_t)(2 * g.nDesired + 4 * g.nCells + 64) * 4 + (size_t)g.nCells * 16 + (size_t)SEL_STAGE * 4);
        g.tile_base = (int)T.size();
        g.tiles_x = (g.w + 2 * EDGE + BLUR_TW - 1) / BLUR_TW; g.tiles_y = (g.h + 2 * EDGE + BLUR_TH - 1) / BLUR_TH;
        for (int ty = 0; ty < g.tiles_y; ++ty) for (int tx = 0; tx < g.tiles_x; ++tx) T.push_back(TileGeo{l, tx * BLUR_TW, ty * BLUR_TH});
    }
    *plane_bytes = poff; *cand_total = coff; *n_cells = C.size(); *n_tiles = T.size(); *tab_total = toff;
    if (fsm > 227 * 1024) fast_big = true;
    *max_fast_smem = fast_big ? fsm_big : fsm; *max_select_smem = ssm;
    if (fast_big_out) *fast_big_out = fast_big;
    if (fast_tma_smem_out) *fast_tma_smem_out = (tma_ok && !fast_big && fsm_tma <= 227 * 1024) ? fsm_tma : 0;
    if (Lv) *Lv = L;
    if (Cv) *Cv = C;
    if (Tv) *Tv = T;
    return SE2GPU_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (the library does not link libcuda, so that it also
// loads on a box without a driver, e.g. for the symbol check of the CPU test suite)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

constexpr bool ORIENT_BATCH_DEFAULT = true;    // measured: 0.0928 -> 0.0897 ms per 64 frames, bit-identical (profiles/r02b_orb_variants.jsonl)
constexpr int FAST_TMA_DEFAULT = 2;   // measured: orb_fast_cells 0.2976 -> 0.2543 ms per 64 frames, bit-identical (profiles/r02c_orb_variants.jsonl)
constexpr bool RESIZE_W_DEFAULT = true;     // measured: pyramid 0.1622 -> 0.155 ms per 64 frames, bit-identical (profiles/r02e_orb_variants.jsonl)
constexpr int PDL_DEFAULT = 2;               // measured: pyramid 0.154 -> 0.1346 ms, step 0.618 -> 0.597 ms per 64 frames, bit-identical (profiles/r02f_orb_variants.jsonl)
constexpr int BLUR_SPLIT_DEFAULT = 2;             // levels 0-1 behind the pyramid tail (round 1)
constexpr bool BLUR_B_AFTER_FAST_DEFAULT = true;   // no event between the last resize and FAST, so that FAST can be its programmatic dependent
constexpr int SUBMIT_CHUNKS_DEFAULT = 1;   // measured: 0.637 ms per 64-frame batch against 1.005 (4 chunks) / 0.857 (2) (profiles/r02d_orb_e2e_submit.jsonl)
int fast_tma_variant() {   // SE2GPU_ORB_FAST_TMA = 0: LDG/STS staging (orb_fast_cells), 1: orb_fast_cells_tma, 2: orb_fast_cells_tma8
    static const int v = [] { const char* e = getenv("SE2GPU_ORB_FAST_TMA"); const int x = e ? atoi(e) : FAST_TMA_DEFAULT; return x < 0 || x > 2 ? FAST_TMA_DEFAULT : x; }();
    return v;
}

// one 3-D u8 tensor per level over the batch of bordered planes: x = byte in the row (pitch), y = row, z = frame
bool encode_fast_maps(se2gpu_orb* h, size_t frame_plane_bytes) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || (frame_plane_bytes & 15)) return false;
    for (int l = 0; l < h->nlevels; ++l) {
        const LevelGeo& g = h->levels[l];
        if (g.fbw <= 6) { memset(&h->fast_maps.m[l], 0, sizeof(CUtensorMap)); continue; }   // no cell on this level ever launches a copy
        const cuuint64_t gdim[3] = {(cuuint64_t)g.pitch, (cuuint64_t)(g.h + 2 * EDGE), (cuuint64_t)h->max_batch};
        const cuuint64_t gstr[2] = {(cuuint64_t)g.pitch, (cuuint64_t)frame_plane_bytes};      // bytes, multiples of 16
        const cuuint32_t box[3] = {(cuuint32_t)g.fbw, (cuuint32_t)g.fbh, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        void* base = h->d.plain + g.plane_off;
        if (((uintptr_t)base & 15) || (g.pitch & 15)) return false;
        if (enc(&h->fast_maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    return true;
}

int set_geometry(se2gpu_orb* h, int w, int hgt, cudaStream_t s) {
    if (w == h->cur_w && hgt == h->cur_h) return SE2GPU_OK;
    size_t pb, ct, nc, nt, tt, fsm, ssm, fsm_tma = 0;
    bool big = false;
    int rc = build_geometry(h, w, hgt, false, &pb, &ct, &nc, &nt, &tt, &fsm, &ssm, &h->levels, &h->cells, &h->tiles, &big, &fsm_tma);
    if (rc != SE2GPU_OK) return rc;
    if (pb > h->cap_plane || ct > h->cap_cand || nc > h->cap_cells || nt > h->cap_tiles || tt > h->cap_tab)
        return fail(SE2GPU_ERR_CAPACITY, "frame %dx%d exceeds the capacity this extractor was created with (%dx%d)", w, hgt, h->max_w, h->max_h);
    if (fsm > 227 * 1024) return fail(SE2GPU_ERR_CAPACITY, "a FAST cell of a %dx%d frame needs %zu B of shared memory", w, hgt, fsm);
    // resize tables [upstream OpenCV resize.cpp: fixed-point bilinear coefficient tables]
    std::vector<int> itab(tt, 0);
    std::vector<short> stab(2 * tt, 0);
    for (int l = 1; l < h->nlevels; ++l) {
        const LevelGeo& g = h->levels[l];
        const LevelGeo& sg = h->levels[l - 1];
        const double scale_x = 1. / ((double)g.w / sg.w), scale_y = 1. / ((double)g.h / sg.h);
        int* xofs = itab.data() + g.tab_off; int* yofs = xofs + g.w;
        short* ialpha = stab.data() + 2 * (size_t)g.tab_off; short* ibeta = ialpha + 2 * g.w;
        for (int dx = 0; dx < g.w; ++dx) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= sg.w - 1) { fx = 0; sx = sg.w - 1; }
            xofs[dx] = sx;
            ialpha[2 * dx] = (short)std::min(std::max(cv_round_f((1.f - fx) * 2048.f), -32768), 32767);
            ialpha[2 * dx + 1] = (short)std::min(std::max(cv_round_f(fx * 2048.f), -32768), 32767);
        }
        for (int dy = 0; dy < g.h; ++dy) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            int sy = (int)floorf(fy);
            fy -= sy;
            yofs[dy] = sy;
            ibeta[2 * dy] = (short)cv_round_f((1.f - fy) * 2048.f);
            ibeta[2 * dy + 1] = (short)cv_round_f(fy * 2048.f);
        }
    }
    SE2_CUDA(cudaMemcpyAsync(h->d_levels, h->levels.data(), sizeof(LevelGeo) * h->levels.size(), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaMemcpyAsync(h->d_cells, h->cells.data(), sizeof(CellGeo) * h->cells.size(), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaMemcpyAsync(h->d_tiles, h->tiles.data(), sizeof(TileGeo) * h->tiles.size(), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaMemcpyAsync(h->d_itab, itab.data(), sizeof(int) * itab.size(), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaMemcpyAsync(h->d_stab, stab.data(), sizeof(short) * stab.size(), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaStreamSynchronize(s));
    h->fast_smem = fsm; h->select_smem = ssm; h->fast_big = big;
    {   // consecutive pyramid levels differ by mvScaleFactor[1] up to rounding of the level sizes
        double ratio = 1.0;
        for (size_t l = 1; l < h->levels.size(); ++l)
            ratio = std::max(ratio, std::max((double)h->levels[l - 1].w / h->levels[l].w, (double)h->levels[l - 1].h / h->levels[l].h));
        h->resize_rows = (int)ceil(RESIZE_TR * ratio) + 4;
        h->resize_raw_pitch = (((int)ceil(128 * ratio) + 4 + 15 + 15) / 16) * 16;
        h->resize_smem = (size_t)h->resize_rows * (128 * sizeof(int) + h->resize_raw_pitch);
        if (h->resize_smem > 200 * 1024) return fail(SE2GPU_ERR_CAPACITY, "scale factor %.3f needs %zu B of shared memory in orb_resize", ratio, h->resize_smem);
        SE2_CUDA(cudaFuncSetAttribute(orb_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->resize_smem));
        h->resize_w_smem = (size_t)h->resize_rows * (128 * sizeof(uint16_t) + h->resize_raw_pitch) + 16;   // 16-bit row-pass results, 16 B of slack behind the last staged row
        SE2_CUDA(cudaFuncSetAttribute(orb_resize_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->resize_w_smem));
        SE2_CUDA(cudaFuncSetAttribute(orb_resize_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->resize_w_smem));
    }
    SE2_CUDA(cudaFuncSetAttribute(orb_fast_cells_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fsm, 1024)));
    SE2_CUDA(cudaFuncSetAttribute(orb_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fsm, 1024)));
    // TMA-staged cells when every level's patch fits a box and the driver hands out the tensor-map encoder
    h->fast_tma = (fast_tma_variant() && fsm_tma > 0 && encode_fast_maps(h, pb)) ? fast_tma_variant() : 0;
    h->fast_tma_smem = fsm_tma;
    if (h->fast_tma == 1) SE2_CUDA(cudaFuncSetAttribute(orb_fast_cells_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fsm_tma, 1024)));
    if (h->fast_tma == 2) {
        SE2_CUDA(cudaFuncSetAttribute(orb_fast_cells_tma8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fsm_tma, 1024)));
        SE2_CUDA(cudaFuncSetAttribute(orb_fast_cells_tma8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fsm_tma, 1024)));
    }
    SE2_CUDA(cudaFuncSetAttribute(orb_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(ssm, 1024)));
    OrbDev& d = h->d;
    d.n_cells = (int)h->cells.size(); d.n_tiles = (int)h->tiles.size();
    d.frame_plane_bytes = pb; d.cand_total = ct;
    int lk = 0; for (auto& g : h->levels) lk += g.kp_cap;
    d.lkp_total = lk;
    h->cur_w = w; h->cur_h = hgt;
    return SE2GPU_OK;
}

// cv::undistort's map for a w x h frame [upstream OpenCV imgproc/undistort]: stripes of max(1, 4096/w) rows, for the stripe at
// row y0 the new camera matrix is A with cy - y0; ir = its LU inverse; rays accumulated column by column in double;
// radial/tangential/thin-prism model; CV_16SC2 + CV_16UC1 fixed point with 5 fraction bits (cvRound(u*32)).
int build_undistort_map(const float* K, const float* dist, int nd, int w, int hgt, std::vector<short2>& m1, std::vector<uint16_t>& m2) {
    double A[9], D[14] = {0};
    for (int i = 0; i < 9; ++i) A[i] = K[i];
    for (int i = 0; i < nd; ++i) D[i] = dist[i];
    m1.resize((size_t)w * hgt); m2.resize((size_t)w * hgt);
    const int stripe = std::min(std::max(1, 4096 / std::max(w, 1)), hgt);
    for (int y0 = 0; y0 < hgt; y0 += stripe) {
        double M[9], inv[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int i = 0; i < 9; ++i) M[i] = A[i];
        M[5] = A[5] - y0;
        for (int c = 0; c < 3; ++c) {                 // LU with partial pivoting, identity carried as right-hand side
            int piv = c;
            for (int r = c + 1; r < 3; ++r) if (std::fabs(M[r * 3 + c]) > std::fabs(M[piv * 3 + c])) piv = r;
            if (std::fabs(M[piv * 3 + c]) < 2.220446049250313e-14) return fail(SE2GPU_ERR_INVALID, "singular camera matrix");
            if (piv != c) for (int q = 0; q < 3; ++q) { std::swap(M[c * 3 + q], M[piv * 3 + q]); std::swap(inv[c * 3 + q], inv[piv * 3 + q]); }
            const double dneg = -1 / M[c * 3 + c];
            for (int r = c + 1; r < 3; ++r) {
                const double al = M[r * 3 + c] * dneg;
                for (int q = c + 1; q < 3; ++q) M[r * 3 + q] += al * M[c * 3 + q];
                for (int q = 0; q < 3; ++q) inv[r * 3 + q] += al * inv[c * 3 + q];
            }
        }
        for (int r = 2; r >= 0; --r)
            for (int q = 0; q < 3; ++q) {
                double acc = inv[r * 3 + q];
                for (int t = r + 1; t < 3; ++t) acc -= M[r * 3 + t] * inv[t * 3 + q];
                inv[r * 3 + q] = acc / M[r * 3 + r];
            }
        const int rows = std::min(stripe, hgt - y0);
        for (int i = 0; i < rows; ++i) {
            double X = i * inv[1] + inv[2], Y = i * inv[4] + inv[5], Wc = i * inv[7] + inv[8];
            short2* o1 = m1.data() + (size_t)(y0 + i) * w;
            uint16_t* o2 = m2.data() + (size_t)(y0 + i) * w;
            for (int jx = 0; jx < w; ++jx, X += inv[0], Y += inv[3], Wc += inv[6]) {
                const double iw = 1. / Wc, x = X * iw, y = Y * iw;
                const double x2 = x * x, y2 = y * y, r2 = x2 + y2, xy2 = 2 * x * y;
                const double kr = (1 + ((D[4] * r2 + D[1]) * r2 + D[0]) * r2) / (1 + ((D[7] * r2 + D[6]) * r2 + D[5]) * r2);
                const double xd = (x * kr + D[2] * xy2 + D[3] * (r2 + 2 * x2) + D[8] * r2 + D[9] * r2 * r2);
                const double yd = (y * kr + D[2] * (r2 + 2 * y2) + D[3] * xy2 + D[10] * r2 + D[11] * r2 * r2);
                const double su = (A[0] * xd + A[2]) * 32, sv = (A[4] * yd + A[5]) * 32;
                const int iu = su >= 2147483647.0 ? 2147483647 : su <= -2147483648.0 ? (int)-2147483648LL : (int)lrint(su);
                const int iv = sv >= 2147483647.0 ? 2147483647 : sv <= -2147483648.0 ? (int)-2147483648LL : (int)lrint(sv);
                o1[jx] = make_short2((short)(iu >> 5), (short)(iv >> 5));
                o2[jx] = (uint16_t)((iv & 31) * 32 + (iu & 31));
            }
        }
    }
    return SE2GPU_OK;
}

int ensure_undistort_map(se2gpu_orb* h, int w, int hgt, cudaStream_t s) {
    if (h->und_w == w && h->und_h == hgt) return SE2GPU_OK;
    std::vector<short2> m1; std::vector<uint16_t> m2;
    int rc = build_undistort_map(h->und_K, h->und_D, h->und_nd, w, hgt, m1, m2);
    if (rc != SE2GPU_OK) return rc;
    const size_t px = (size_t)w * hgt;
    if (px > h->und_cap) {
        if (h->d_und_m1) cudaFree(h->d_und_m1);
        if (h->d_und_m2) cudaFree(h->d_und_m2);
        h->d_und_m1 = nullptr; h->d_und_m2 = nullptr; h->und_cap = 0;
        SE2_CUDA(cudaMalloc((void**)&h->d_und_m1, px * sizeof(short2)));
        SE2_CUDA(cudaMalloc((void**)&h->d_und_m2, px * sizeof(uint16_t)));
        h->und_cap = px;
    }
    SE2_CUDA(cudaMemcpyAsync(h->d_und_m1, m1.data(), px * sizeof(short2), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaMemcpyAsync(h->d_und_m2, m2.data(), px * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    SE2_CUDA(cudaStreamSynchronize(s));      // the host vectors go out of scope
    h->und_w = w; h->und_h = hgt;
    return SE2GPU_OK;
}

int run_device(se2gpu_orb* h, const uint8_t* d_imgs, int n, int w, int hgt, int stride, size_t frame_stride,
               se2gpu_keypoint* d_kps, uint8_t* d_desc, int* d_counts, cudaStream_t s, int frame0 = 0, int lane = -1) {
    int rc = set_geometry(h, w, hgt, s);
    if (rc != SE2GPU_OK) return rc;
    if (h->und_on && (rc = ensure_undistort_map(h, w, hgt, s)) != SE2GPU_OK) return rc;
    OrbDev d = h->d;             // by-value copy carrying this launch group's frame offset
    d.frame0 = frame0;
    // pipeline lanes get their concurrency from each other, not from a blur side stream
    cudaStream_t side = lane < 0 ? h->side : nullptr;
    cudaEvent_t ev_pyr = h->ev_pyr, ev_blur = h->ev_blur;
    se2gpu::Profiler& pr = h->prof;
    SE2_NVTX("se2gpu.orb.run_device");
    nvtxRangePushA("se2gpu.orb.pyramid");
    pr.begin(0, s);
    {
        const LevelGeo& g = h->levels[0];
        if (h->und_on) {
            dim3 grid((g.pitch / 4 + 127) / 128, g.h + 2 * EDGE, n);
            SE2_LAUNCH(orb_pyr0_undistort, grid, 128, 0, s, d, g, d_imgs, stride, frame_stride, h->d_und_m1, h->d_und_m2);
        } else {
            const int aligned = (((uintptr_t)d_imgs | (uintptr_t)stride | (uintptr_t)frame_stride) & 15) == 0;
            const int nvec = aligned ? g.w / 16 : 0;      // aligned 16-byte interior vectors per row
            dim3 grid((nvec + 63) / 64 + 1, (g.h + 2 * EDGE + 4 * PYR0_ROWS - 1) / (4 * PYR0_ROWS), n);
            SE2_LAUNCH(orb_pyr0, grid, dim3(64, 4), 0, s, d, g, d_imgs, stride, frame_stride, nvec);
        }
    }
    // blur schedule on the side stream: group A = levels [0, splitA) starts as soon as level splitA-1 exists; group B = the rest starts
    // when the pyramid is complete (round-1 form) or, with SE2GPU_ORB_BLUR_B_AFTER_FAST=1, when FAST has been launched, i.e. it runs
    // next to the selection kernel, whose level-0 CTAs leave most SMs idle, instead of competing with the issue-bound FAST kernel
    // SE2GPU_ORB_PDL: 0 plain launches, 1 programmatic dependent launch along the resize chain, 2 also the FAST kernel behind the last resize
    static const int pdl_mode = [] { const char* e = getenv("SE2GPU_ORB_PDL"); const int v = e ? atoi(e) : PDL_DEFAULT; return v < 0 || v > 2 ? 0 : v; }();
    static const int env_split = [] { const char* e = getenv("SE2GPU_ORB_BLUR_SPLIT"); return e ? atoi(e) : BLUR_SPLIT_DEFAULT; }();
    static const bool b_after_fast = [] { const char* e = getenv("SE2GPU_ORB_BLUR_B_AFTER_FAST"); return e ? atoi(e) != 0 : BLUR_B_AFTER_FAST_DEFAULT; }();
    const int splitA = std::min(std::max(env_split, 1), h->nlevels);
    for (int l = 1; l < h->nlevels; ++l) {
        const LevelGeo& g = h->levels[l];
        dim3 grid((g.pitch + 127) / 128, (g.h + 2 * EDGE + RESIZE_TR - 1) / RESIZE_TR, n);
        static const bool resize_w = [] { const char* e = getenv("SE2GPU_ORB_RESIZE_W"); return e ? atoi(e) != 0 : RESIZE_W_DEFAULT; }();
        if (resize_w && pdl_mode >= 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(32, 8); cfg.dynamicSmemBytes = h->resize_w_smem; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            SE2_CUDA(cudaLaunchKernelEx(&cfg, orb_resize_w<true>, d, g, h->levels[l - 1], h->resize_rows, h->resize_raw_pitch));
            ::se2gpu::g_launches.fetch_add(1, std::memory_order_relaxed);
        } else if (resize_w) SE2_LAUNCH(orb_resize_w<false>, grid, dim3(32, 8), h->resize_w_smem, s, d, g, h->levels[l - 1], h->resize_rows, h->resize_raw_pitch);
        else SE2_LAUNCH(orb_resize, grid, dim3(32, 8), h->resize_smem, s, d, g, h->levels[l - 1], h->resize_rows, h->resize_raw_pitch);
        if (l == splitA - 1 && side && !pr.on) SE2_CUDA(cudaEventRecord(h->ev_l1, s));
    }
    if (splitA - 1 <= 0 && side && !pr.on) SE2_CUDA(cudaEventRecord(h->ev_l1, s));   // level 0 alone is group A (or there is only one level)
    pr.end(s);
    nvtxRangePop();
    // The blur of a level only needs that level's plane: on the side stream the blur of levels 0-1 (55 % of the pixels,
    // no shared memory, so it co-resides with the resize CTAs) starts as soon as level 1 exists and hides behind the
    // tail of the pyramid, a chain of small latency-bound launches; the remaining levels are blurred when the pyramid is
    // complete, next to FAST and the selection. The main stream joins before the descriptors are sampled. With the
    // profiler on everything stays on one stream so that the per-kernel event times are not polluted by the overlap.
    const bool overlap = !pr.on && side != nullptr;
    auto launch_blur = [&](cudaStream_t st, int t0, int t1) {
        if (t1 > t0) SE2_LAUNCH(orb_blur, dim3((t1 - t0 + 7) / 8, n), 256, 0, st, d, t0, t1);
    };
    const int tilesA = splitA < h->nlevels ? h->levels[splitA].tile_base : d.n_tiles;
    if (overlap) {
        SE2_CUDA(cudaStreamWaitEvent(side, h->ev_l1, 0));
        launch_blur(side, 0, tilesA);
        if (!b_after_fast) {
            SE2_CUDA(cudaEventRecord(ev_pyr, s));
            SE2_CUDA(cudaStreamWaitEvent(side, ev_pyr, 0));
            launch_blur(side, tilesA, d.n_tiles);
            SE2_CUDA(cudaEventRecord(ev_blur, side));
        }
    }
    SE2_NVTX("se2gpu.orb.fast_select_blur_describe");
    pr.begin(1, s);
    if (h->fast_big) SE2_LAUNCH(orb_fast_cells_big, dim3(d.n_cells, n), FAST_THREADS, h->fast_smem, s, d);
    else if (h->fast_tma == 2 && pdl_mode == 2 && !pr.on) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(d.n_cells, n); cfg.blockDim = dim3(FAST_THREADS); cfg.dynamicSmemBytes = h->fast_tma_smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SE2_CUDA(cudaLaunchKernelEx(&cfg, orb_fast_cells_tma8<true>, d, h->fast_maps));
        ::se2gpu::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    else if (h->fast_tma == 2) SE2_LAUNCH(orb_fast_cells_tma8<false>, dim3(d.n_cells, n), FAST_THREADS, h->fast_tma_smem, s, d, h->fast_maps);
    else if (h->fast_tma == 1) SE2_LAUNCH(orb_fast_cells_tma, dim3(d.n_cells, n), FAST_THREADS, h->fast_tma_smem, s, d, h->fast_maps);
    else SE2_LAUNCH(orb_fast_cells, dim3(d.n_cells, n), FAST_THREADS, h->fast_smem, s, d);
    pr.end(s);
    if (overlap && b_after_fast) {      // group B behind FAST in stream order, concurrent with the selection
        SE2_CUDA(cudaEventRecord(ev_pyr, s));
        SE2_CUDA(cudaStreamWaitEvent(side, ev_pyr, 0));
        launch_blur(side, tilesA, d.n_tiles);
        SE2_CUDA(cudaEventRecord(ev_blur, side));
    }
    pr.begin(2, s);
    SE2_LAUNCH(orb_select, dim3(h->nlevels, n), SEL_THREADS, h->select_smem, s, d);
    pr.end(s);
    if (overlap) {
        SE2_CUDA(cudaStreamWaitEvent(s, ev_blur, 0));
    } else {
        pr.begin(3, s);
        launch_blur(s, 0, d.n_tiles);
        pr.end(s);
    }
    const int warps = 8;
    pr.begin(4, s);
    // SE2GPU_ORB_ORIENT_BATCH=0: one dependent +-v load pair per loop trip (round-1 form)
    static const bool orient_batch = [] { const char* e = getenv("SE2GPU_ORB_ORIENT_BATCH"); return e ? atoi(e) != 0 : ORIENT_BATCH_DEFAULT; }();
    if (orient_batch) SE2_LAUNCH(orb_orient_describe<true>, dim3((h->nfeatures + warps - 1) / warps, n), warps * 32, 0, s, d, d_kps, d_desc, d_counts);
    else SE2_LAUNCH(orb_orient_describe<false>, dim3((h->nfeatures + warps - 1) / warps, n), warps * 32, 0, s, d, d_kps, d_desc, d_counts);
    pr.end(s);
    h->last_n = std::max(h->last_n * (frame0 > 0), frame0 + n);
    return SE2GPU_OK;
}

}  // namespace

extern "C" {

se2gpu_orb* se2gpu_orb_create(int nfeatures, float scale_factor, int nlevels, int fast_th, int max_w, int max_h,
                              int max_batch, int device) {
    if (nfeatures <= 0 || nlevels <= 0 || nlevels > MAX_LEVELS || !(scale_factor > 1.0f) || fast_th < 1 || fast_th > 254 ||
        max_w <= 0 || max_h <= 0 || max_batch <= 0) { fail(SE2GPU_ERR_INVALID, "bad ORB parameters"); return nullptr; }
    if (se2gpu::select_device(device) != SE2GPU_OK) return nullptr;
    se2gpu_orb* h = new se2gpu_orb;
    h->device = device; h->nfeatures = nfeatures; h->nlevels = nlevels; h->fast_th = fast_th;
    h->max_w = max_w; h->max_h = max_h; h->max_batch = max_batch;
    { int* a = h->create_args; a[0] = nfeatures; memcpy(&a[1], &scale_factor, sizeof(float)); a[2] = nlevels; a[3] = fast_th; a[4] = max_w; a[5] = max_h; a[6] = max_batch; a[7] = device; }
    h->scaleFactor = scale_factor;   // the reference keeps it in a double member (ORBextractor.h:67)
    // ORBextractor::ORBextractor, ORBextractor.cpp:463-520
    h->mvScaleFactor.resize(nlevels); h->mvInvScaleFactor.resize(nlevels); h->mnFeaturesPerLevel.resize(nlevels);
    h->mvScaleFactor[0] = 1;
    for (int i = 1; i < nlevels; i++) h->mvScaleFactor[i] = (float)(h->mvScaleFactor[i - 1] * h->scaleFactor);
    const float invScaleFactor = (float)(1.0f / h->scaleFactor);
    h->mvInvScaleFactor[0] = 1;
    for (int i = 1; i < nlevels; i++) h->mvInvScaleFactor[i] = h->mvInvScaleFactor[i - 1] * invScaleFactor;
    const float factor = (float)(1.0 / h->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { h->mnFeaturesPerLevel[l] = cv_round_f(nDesired); sum += h->mnFeaturesPerLevel[l]; nDesired *= factor; }
    h->mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    if (sum > nfeatures) { fail(SE2GPU_ERR_INVALID, "per-level quotas sum to %d > nfeatures %d for these parameters", sum, nfeatures); delete h; return nullptr; }
    int umax[16];
    {
        int v, v0, vmax = (int)floorf(HALF_PATCH * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = HALF_PATCH * HALF_PATCH;
        for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) { while (umax[v0] == umax[v0 + 1]) ++v0; umax[v] = v0; ++v0; }
    }
    float gk[7];
    { double g[7], s = 0; for (int i = 0; i < 7; ++i) { double x = i - 3; g[i] = std::exp(-0.5 * x * x / 4.0); s += g[i]; } for (int i = 0; i < 7; ++i) gk[i] = (float)(g[i] * (1. / s)); }
    size_t pb, ct, nc, nt, tt, fsm, ssm;
    int rc = build_geometry(h, max_w, max_h, true, &pb, &ct, &nc, &nt, &tt, &fsm, &ssm, nullptr, nullptr, nullptr);
    if (rc != SE2GPU_OK) { delete h; return nullptr; }
    // head-room so that smaller frames (different cell rounding) always fit
    h->cap_plane = pb + 4096; h->cap_cand = ct + ct / 8 + 4096; h->cap_cells = nc + 64; h->cap_tiles = nt + 64; h->cap_tab = tt + 64;
    h->cap_lkp = nfeatures + 64;
    const size_t B = max_batch;
    bool ok = true;
    auto A = [&](auto** p, size_t c) { if (ok && se2gpu::dev_alloc(p, c) == cudaSuccess) h->bufs.push_back(*p); else ok = false; };
    OrbDev& d = h->d;
    A(&h->d_levels, (size_t)nlevels); A(&h->d_cells, h->cap_cells); A(&h->d_tiles, h->cap_tiles); A(&h->d_itab, h->cap_tab); A(&h->d_stab, 2 * h->cap_tab);
    A(&d.plain, B * h->cap_plane); A(&d.blurred, B * h->cap_plane);
    A(&d.cand, B * h->cap_cand); A(&d.hdr, B * h->cap_cells); A(&d.lkp, B * h->cap_lkp); A(&d.lcount, B * nlevels); A(&d.err, 1);
    A(&h->d_in, B * (size_t)max_w * max_h); A(&h->d_kps, B * nfeatures); A(&h->d_desc, B * nfeatures * 32); A(&h->d_counts, B);
    if (!ok) { fail(SE2GPU_ERR_CUDA, "device allocation failed (%s)", cudaGetErrorString(cudaGetLastError())); se2gpu_orb_destroy(h); return nullptr; }
    cudaMemset(d.err, 0, sizeof(int));
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_pyr, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_blur, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_l1, cudaEventDisableTiming) != cudaSuccess) { h->side = nullptr; cudaGetLastError(); }
    for (int l = 0; l < ORB_LANES && h->side; ++l)
        if (cudaStreamCreateWithFlags(&h->pipe[l], cudaStreamNonBlocking) != cudaSuccess) { h->pipe[l] = nullptr; cudaGetLastError(); break; }
    cudaMemcpyToSymbol(c_umax, umax, sizeof umax);
    {
        int vmax[16];
        for (int u = 0; u <= HALF_PATCH; ++u) { vmax[u] = 0; for (int v = 0; v <= HALF_PATCH; ++v) if (umax[v] >= u) vmax[u] = v; }
        cudaMemcpyToSymbol(c_vmax, vmax, sizeof vmax);
    }
    cudaMemcpyToSymbol(c_gauss, gk, sizeof gk);
    d.nlevels = nlevels; d.nfeatures = nfeatures; d.fast_th = fast_th; d.t_lo = std::min(fast_th, 7);
    d.levels = h->d_levels; d.cells = h->d_cells; d.tiles = h->d_tiles; d.itab = h->d_itab; d.stab = h->d_stab;
    if (cudaDeviceSynchronize() != cudaSuccess) { fail(SE2GPU_ERR_CUDA, "init failed"); se2gpu_orb_destroy(h); return nullptr; }
    return h;
}

void se2gpu_orb_destroy(se2gpu_orb* h) {
    if (!h) return;
    if (h->twin) { se2gpu_orb_destroy(h->twin); h->twin = nullptr; }
    cudaSetDevice(h->device);
    for (void* p : h->bufs) cudaFree(p);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_pyr) cudaEventDestroy(h->ev_pyr);
    if (h->d_und_m1) cudaFree(h->d_und_m1);
    if (h->d_und_m2) cudaFree(h->d_und_m2);
    if (h->ev_blur) cudaEventDestroy(h->ev_blur);
    if (h->ev_l1) cudaEventDestroy(h->ev_l1);
    if (h->pin_counts) cudaFreeHost(h->pin_counts);
    if (h->pin_kps) cudaFreeHost(h->pin_kps);
    if (h->pin_desc) cudaFreeHost(h->pin_desc);
    for (int l = 0; l < ORB_LANES; ++l)
        if (h->pipe[l]) cudaStreamDestroy(h->pipe[l]);
    delete h;
}

int se2gpu_orb_extract_device(se2gpu_orb* h, const uint8_t* d_imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                              se2gpu_keypoint* d_kps, uint8_t* d_desc, int* d_counts, void* stream) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (n < 0 || n > h->max_batch) return fail(SE2GPU_ERR_CAPACITY, "batch %d exceeds max_batch %d", n, h->max_batch);
    SE2_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return SE2GPU_OK;
    if (!d_imgs || w <= 0 || hgt <= 0) { SE2_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * n, s)); return SE2GPU_OK; }
    if (stride < w) return fail(SE2GPU_ERR_INVALID, "stride < width");
    return run_device(h, d_imgs, n, w, hgt, stride, frame_stride, d_kps, d_desc, d_counts, s);
}

// enqueue one host-buffer batch on this context (copies in, kernels, copies out); nothing is waited for
static int orb_enqueue(se2gpu_orb* h, const uint8_t* imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                       se2gpu_keypoint* kps, uint8_t* desc, int* counts, bool submit_mode = false) {
    SE2_NVTX("se2gpu.orb.enqueue");
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (h->pend.on) return fail(SE2GPU_ERR_INVALID, "a submitted batch is still pending on this context: call se2gpu_orb_wait first");
    if (n < 0 || n > h->max_batch) return fail(SE2GPU_ERR_CAPACITY, "batch %d exceeds max_batch %d", n, h->max_batch);
    if (n == 0) return SE2GPU_OK;
    if (!imgs || w <= 0 || hgt <= 0) { for (int i = 0; i < n; ++i) counts[i] = 0; return SE2GPU_OK; }   // :730-731
    if (stride < w) return fail(SE2GPU_ERR_INVALID, "stride < width");
    if (w > h->max_w || hgt > h->max_h) return fail(SE2GPU_ERR_CAPACITY, "frame %dx%d exceeds %dx%d", w, hgt, h->max_w, h->max_h);
    SE2_CUDA(cudaSetDevice(h->device));
    int rc = set_geometry(h, w, hgt, nullptr);
    if (rc != SE2GPU_OK) return rc;
    // pipeline shape: `nchunks` chunks round-robin over `lanes` streams (H2D copy, kernels and D2H copies of a chunk are
    // stream-ordered; different lanes overlap), the first chunk `first_pct` percent of an even share so that less of
    // the initial H2D copy is exposed. SE2GPU_ORB_CHUNKS / _LANES / _FIRST override; 1 chunk = one synchronous pass.
    static const int env_chunks = [] { const char* e = getenv("SE2GPU_ORB_CHUNKS"); return e ? atoi(e) : 0; }();
    static const int env_first = [] { const char* e = getenv("SE2GPU_ORB_FIRST"); return e ? atoi(e) : 0; }();
    static const int env_lanes = [] { const char* e = getenv("SE2GPU_ORB_LANES"); return e ? atoi(e) : 0; }();
    int lanes = 0;
    while (lanes < ORB_LANES && h->pipe[lanes]) ++lanes;
    if (env_lanes > 0) lanes = std::min(lanes, env_lanes);
    // se2gpu_orb_submit keeps two batches in flight on twin contexts, so the copies of one batch already overlap the kernels of
    // the other: there the batch can go down in fewer, larger chunks (SE2GPU_ORB_SUBMIT_CHUNKS; 1 = the whole batch on one stream)
    static const int env_submit_chunks = [] { const char* e = getenv("SE2GPU_ORB_SUBMIT_CHUNKS"); return e ? atoi(e) : 0; }();
    const int nchunks = submit_mode ? (env_submit_chunks > 0 ? env_submit_chunks : SUBMIT_CHUNKS_DEFAULT) : (env_chunks > 0 ? env_chunks : 4);
    const bool pipelined = lanes >= 2 && !h->prof.on && n > 1 && (nchunks > 1 || submit_mode);
    const int chunk = pipelined ? std::max(1, (n + nchunks - 1) / nchunks) : n;
    const int first = (pipelined && nchunks > 1) ? std::max(1, std::min(n, chunk * (env_first > 0 ? env_first : 50) / 100)) : n;
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    se2gpu_keypoint* out_kps = kps; uint8_t* out_desc = desc; int* out_counts = counts;
    if (pipelined) {
        if (!h->pin_counts && cudaMallocHost((void**)&h->pin_counts, sizeof(int) * h->max_batch) != cudaSuccess) return fail(SE2GPU_ERR_CUDA, "cudaMallocHost failed");
        out_counts = h->pin_counts;
        if (!is_pinned(kps) || !is_pinned(desc)) {
            if (!h->pin_kps && (cudaMallocHost((void**)&h->pin_kps, sizeof(se2gpu_keypoint) * (size_t)h->max_batch * h->nfeatures) != cudaSuccess ||
                                cudaMallocHost((void**)&h->pin_desc, (size_t)32 * h->max_batch * h->nfeatures) != cudaSuccess))
                return fail(SE2GPU_ERR_CUDA, "cudaMallocHost failed");
            out_kps = h->pin_kps; out_desc = h->pin_desc;
        }
    }
    int lane = 0;
    for (int f0 = 0, m = 0; f0 < n; f0 += m, lane = (lane + 1) % std::max(lanes, 1)) {
        m = std::min(f0 == 0 ? first : std::max(chunk, (n - first + nchunks - 2) / std::max(1, nchunks - 1)), n - f0);
        cudaStream_t s = pipelined ? h->pipe[lane] : nullptr;
        // pack rows tightly on the device (pitch = w); frames of different chunks use disjoint device buffers
        if (stride == w && frame_stride == (size_t)w * hgt)
            SE2_CUDA(cudaMemcpyAsync(h->d_in + (size_t)f0 * w * hgt, imgs + (size_t)f0 * frame_stride, (size_t)m * w * hgt, cudaMemcpyHostToDevice, s));
        else
            for (int i = f0; i < f0 + m; ++i)
                SE2_CUDA(cudaMemcpy2DAsync(h->d_in + (size_t)i * w * hgt, w, imgs + i * frame_stride, stride, w, hgt, cudaMemcpyHostToDevice, s));
        // a single chunk keeps the blur on the context's side stream (like the device-resident entry); several chunks overlap each other
        rc = run_device(h, h->d_in, m, w, hgt, w, (size_t)w * hgt, h->d_kps, h->d_desc, h->d_counts, s, f0, (pipelined && nchunks > 1) ? lane : -1);
        if (rc != SE2GPU_OK) return rc;
        SE2_CUDA(cudaMemcpyAsync(out_counts + f0, h->d_counts + f0, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
        SE2_CUDA(cudaMemcpyAsync(out_kps + (size_t)f0 * h->nfeatures, h->d_kps + (size_t)f0 * h->nfeatures, sizeof(se2gpu_keypoint) * (size_t)m * h->nfeatures, cudaMemcpyDeviceToHost, s));
        SE2_CUDA(cudaMemcpyAsync(out_desc + (size_t)32 * f0 * h->nfeatures, h->d_desc + (size_t)32 * f0 * h->nfeatures, (size_t)32 * m * h->nfeatures, cudaMemcpyDeviceToHost, s));
    }
    h->pend.on = true; h->pend.n = n; h->pend.lanes = lanes; h->pend.pipelined = pipelined;
    h->pend.kps = kps; h->pend.desc = desc; h->pend.counts = counts; h->pend.out_kps = out_kps; h->pend.out_desc = out_desc; h->pend.out_counts = out_counts;
    return SE2GPU_OK;
}

// wait for the batch enqueued by orb_enqueue and hand the results to the caller's buffers
static int orb_finish(se2gpu_orb* h) {
    SE2_NVTX("se2gpu.orb.finish");
    if (!h->pend.on) return SE2GPU_OK;
    SE2_CUDA(cudaSetDevice(h->device));
    const se2gpu_orb::Pending p = h->pend;
    h->pend.on = false;
    if (p.pipelined) {
        for (int l = 0; l < p.lanes; ++l) SE2_CUDA(cudaStreamSynchronize(h->pipe[l]));
        memcpy(p.counts, p.out_counts, sizeof(int) * p.n);
        if (p.out_kps != p.kps) { memcpy(p.kps, p.out_kps, sizeof(se2gpu_keypoint) * (size_t)p.n * h->nfeatures); memcpy(p.desc, p.out_desc, (size_t)32 * p.n * h->nfeatures); }
    }
    int err = 0;
    SE2_CUDA(cudaMemcpyAsync(&err, h->d.err, sizeof(int), cudaMemcpyDeviceToHost, nullptr));
    SE2_CUDA(cudaStreamSynchronize(nullptr));
    h->last_n = p.n;
    if (err) {
        cudaMemset(h->d.err, 0, sizeof(int));
        if (err == 4) return fail(SE2GPU_ERR_CUDA, "the TMA copy of a FAST cell did not complete (tensor map / driver problem)");
        return fail(SE2GPU_ERR_CAPACITY, "internal candidate buffer overflow (code %d)", err);
    }
    return SE2GPU_OK;
}

int se2gpu_orb_extract(se2gpu_orb* h, const uint8_t* imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                       se2gpu_keypoint* kps, uint8_t* desc, int* counts) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (h->in_flight) return fail(SE2GPU_ERR_INVALID, "se2gpu_orb_extract while submitted batches are in flight: drain them with se2gpu_orb_wait");
    int rc = orb_enqueue(h, imgs, n, w, hgt, stride, frame_stride, kps, desc, counts);
    if (rc != SE2GPU_OK) return rc;
    return orb_finish(h);
}

int se2gpu_orb_submit(se2gpu_orb* h, const uint8_t* imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                      se2gpu_keypoint* kps, uint8_t* desc, int* counts) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (h->in_flight >= 2) return fail(SE2GPU_ERR_INVALID, "two batches are already in flight: call se2gpu_orb_wait first");
    if (!h->twin) {
        const int* a = h->create_args;
        float sf; memcpy(&sf, &a[1], sizeof sf);
        h->twin = se2gpu_orb_create(a[0], sf, a[2], a[3], a[4], a[5], a[6], a[7]);
        if (!h->twin) return SE2GPU_ERR_CUDA;
        if (h->und_on && se2gpu_orb_set_undistort(h->twin, h->und_K, h->und_nd ? h->und_D : nullptr, h->und_nd) != SE2GPU_OK) return SE2GPU_ERR_CUDA;
    }
    se2gpu_orb* ctx = (h->submit_next & 1) ? h->twin : h;
    int rc = orb_enqueue(ctx, imgs, n, w, hgt, stride, frame_stride, kps, desc, counts, true);
    if (rc != SE2GPU_OK) return rc;
    if (!ctx->pend.on) {          // empty image / n == 0: finished synchronously (outputs as se2gpu_orb_extract leaves them)
        return SE2GPU_OK;
    }
    h->submit_next ^= 1; h->in_flight++;
    return SE2GPU_OK;
}

int se2gpu_orb_wait(se2gpu_orb* h) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (!h->in_flight) return SE2GPU_OK;
    se2gpu_orb* ctx = (h->wait_next & 1) ? h->twin : h;
    h->wait_next ^= 1; h->in_flight--;
    return orb_finish(ctx);
}



int se2gpu_orb_debug_nth_element(uint32_t* values, const int* offsets, const int* nth, int count, int device) {
    if (count <= 0) return SE2GPU_OK;
    if (!values || !offsets || !nth) return fail(SE2GPU_ERR_INVALID, "null argument");
    if (se2gpu::select_device(device) != SE2GPU_OK) return SE2GPU_ERR_CUDA;
    const size_t total = (size_t)offsets[count];
    uint32_t* dv = nullptr; int* dofs = nullptr; int* dn = nullptr;
    SE2_CUDA(cudaMalloc((void**)&dv, std::max<size_t>(total, 1) * 4));
    SE2_CUDA(cudaMalloc((void**)&dofs, sizeof(int) * (count + 1)));
    SE2_CUDA(cudaMalloc((void**)&dn, sizeof(int) * count));
    SE2_CUDA(cudaMemcpy(dv, values, total * 4, cudaMemcpyHostToDevice));
    SE2_CUDA(cudaMemcpy(dofs, offsets, sizeof(int) * (count + 1), cudaMemcpyHostToDevice));
    SE2_CUDA(cudaMemcpy(dn, nth, sizeof(int) * count, cudaMemcpyHostToDevice));
    SE2_LAUNCH(orb_debug_nth, (count + 3) / 4, 128, 0, nullptr, dv, dofs, dn, count);
    SE2_CUDA(cudaGetLastError());
    SE2_CUDA(cudaMemcpy(values, dv, total * 4, cudaMemcpyDeviceToHost));
    cudaFree(dv); cudaFree(dofs); cudaFree(dn);
    return SE2GPU_OK;
}

int se2gpu_orb_set_undistort(se2gpu_orb* h, const float* K, const float* dist, int ndist) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    if (!K) { h->und_on = false; return SE2GPU_OK; }
    if (ndist != 0 && ndist != 4 && ndist != 5 && ndist != 8 && ndist != 12 && ndist != 14) return fail(SE2GPU_ERR_INVALID, "distortion vector must have 0, 4, 5, 8, 12 or 14 coefficients");
    if (ndist > 0 && !dist) return fail(SE2GPU_ERR_INVALID, "null distortion vector");
    if (ndist == 14 && (dist[12] != 0.f || dist[13] != 0.f)) return fail(SE2GPU_ERR_INVALID, "tilted-sensor coefficients (tauX, tauY) are not supported");
    memcpy(h->und_K, K, sizeof h->und_K);
    memset(h->und_D, 0, sizeof h->und_D);
    if (ndist > 0) memcpy(h->und_D, dist, sizeof(float) * ndist);
    h->und_nd = ndist; h->und_w = h->und_h = 0; h->und_on = true;
    return SE2GPU_OK;
}

int se2gpu_orb_debug_undistort_map(const float* K, const float* dist, int ndist, int w, int hgt, int16_t* m1, uint16_t* m2) {
    if (!K || !m1 || !m2 || w <= 0 || hgt <= 0 || ndist < 0 || ndist > 14 || (ndist > 0 && !dist)) return fail(SE2GPU_ERR_INVALID, "bad argument");
    std::vector<short2> a; std::vector<uint16_t> b;
    int rc = build_undistort_map(K, dist, ndist, w, hgt, a, b);
    if (rc != SE2GPU_OK) return rc;
    memcpy(m1, a.data(), a.size() * sizeof(short2));
    memcpy(m2, b.data(), b.size() * sizeof(uint16_t));
    return SE2GPU_OK;
}

int se2gpu_orb_profile(se2gpu_orb* h, int enable) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    SE2_CUDA(cudaSetDevice(h->device));
    h->prof.enable(enable != 0);
    return SE2GPU_OK;
}

int se2gpu_orb_profile_read(se2gpu_orb* h, double* ms, int* launches) {
    if (!h) return fail(SE2GPU_ERR_INVALID, "null handle");
    SE2_CUDA(cudaSetDevice(h->device));
    h->prof.flush();
    for (int g = 0; g < SE2GPU_ORB_PROFILE_GROUPS; ++g) { if (ms) ms[g] = h->prof.ms[g]; if (launches) launches[g] = h->prof.launches[g]; }
    return SE2GPU_OK;
}

int se2gpu_orb_level_dims(se2gpu_orb* h, int level, int* w, int* hgt, int* pitch) {
    if (!h || level < 0 || level >= h->nlevels || h->cur_w < 0) return fail(SE2GPU_ERR_INVALID, "no geometry");
    *w = h->levels[level].w; *hgt = h->levels[level].h; *pitch = h->levels[level].pitch;
    return SE2GPU_OK;
}

int se2gpu_orb_get_level(se2gpu_orb* h, int frame, int level, int blurred, uint8_t* out) {
    if (!h || level < 0 || level >= h->nlevels || h->cur_w < 0 || frame < 0 || frame >= h->last_n) return fail(SE2GPU_ERR_INVALID, "bad frame/level");
    SE2_CUDA(cudaSetDevice(h->device));
    const LevelGeo& g = h->levels[level];
    const uint8_t* src = (blurred ? h->d.blurred : h->d.plain) + (size_t)frame * h->d.frame_plane_bytes + g.plane_off;
    SE2_CUDA(cudaMemcpy(out, src, (size_t)g.pitch * (g.h + 2 * EDGE), cudaMemcpyDeviceToHost));
    return SE2GPU_OK;
}

}  // extern "C"

// Probe of cp.async.bulk.tensor.3d (UTMALDG) on a u8 tensor with the descriptor in different places:
//   mode 0: const __grid_constant__ CUtensorMap kernel parameter (the canonical form)
//   mode 1: array of maps inside a __grid_constant__ struct parameter, indexed with a value loaded from global memory
//   mode 2: descriptor in global memory (cudaMemcpy'd before the launch)
//   mode 3: like 1, but the index is a kernel argument (uniform)
// usage: tma_probe <mode> [box_w box_h x y z]   — prints OK / MISMATCH / the CUDA error. Build: nvcc -arch=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Maps { CUtensorMap m[16]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ void load_box(uint8_t* dst, const CUtensorMap* map, int x, int y, int z, unsigned bytes, unsigned long long* bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
    }
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
        if (done) break;
        if (clock64() - t0 > 400000000LL) { if (threadIdx.x == 0) printf("device: wait timed out\n"); break; }
    }
    __syncthreads();
}

extern __shared__ __align__(128) uint8_t dyn[];

__global__ void k_param(const __grid_constant__ CUtensorMap m, int x, int y, int z, int bytes, uint8_t* out) {
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* dst = dyn + ((128u - (smem_u32(dyn) & 127u)) & 127u);
    load_box(dst, &m, x, y, z, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = dst[i];
}
__global__ void k_array_dyn(const __grid_constant__ Maps ms, const int* idx, int x, int y, int z, int bytes, uint8_t* out) {
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* dst = dyn + ((128u - (smem_u32(dyn) & 127u)) & 127u);
    load_box(dst, &ms.m[idx[blockIdx.x]], x, y, z, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = dst[i];
}
__global__ void k_array_uni(const __grid_constant__ Maps ms, int idx, int x, int y, int z, int bytes, uint8_t* out) {
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* dst = dyn + ((128u - (smem_u32(dyn) & 127u)) & 127u);
    load_box(dst, &ms.m[idx], x, y, z, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = dst[i];
}
__global__ void k_global(const CUtensorMap* gm, int x, int y, int z, int bytes, uint8_t* out) {
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* dst = dyn + ((128u - (smem_u32(dyn) & 127u)) & 127u);
    load_box(dst, gm, x, y, z, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #e, __LINE__, cudaGetErrorString(_e)); return 2; } } while (0)

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int bw = argc > 2 ? atoi(argv[2]) : 128, bh = argc > 3 ? atoi(argv[3]) : 81;
    const int x = argc > 4 ? atoi(argv[4]) : 13, y = argc > 5 ? atoi(argv[5]) : 7, z = argc > 6 ? atoi(argv[6]) : 2;
    const int X = 672, Y = 512, Z = 4;
    const size_t plane = (size_t)X * Y + 4096 + 256;   // frame stride: a multiple of 16, not of the pitch (like the plane blocks)
    std::vector<uint8_t> h(plane * Z);
    for (int zz = 0; zz < Z; ++zz) for (int yy = 0; yy < Y; ++yy) for (int xx = 0; xx < X; ++xx) h[zz * plane + (size_t)yy * X + xx] = (uint8_t)(xx * 7 + yy * 13 + zz * 31);
    uint8_t* d = nullptr; CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) { printf("no encoder (%d)\n", (int)q); return 2; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    Maps ms; memset(&ms, 0, sizeof ms);
    const cuuint64_t gdim[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z}, gstr[2] = {(cuuint64_t)X, (cuuint64_t)plane};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, estr[3] = {1u, 1u, 1u};
    for (int l = 0; l < 16; ++l) {
        CUresult r = enc(&ms.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", (int)r); return 2; }
    }
    const int bytes = bw * bh;
    uint8_t* out = nullptr; CK(cudaMalloc(&out, bytes)); CK(cudaMemset(out, 0xEE, bytes));
    const size_t smem = bytes + 256;
    const int lvl = 5;
    if (mode == 0) { CK(cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_param<<<1, 256, smem>>>(ms.m[lvl], x, y, z, bytes, out); }
    else if (mode == 1) {
        int* didx; CK(cudaMalloc(&didx, 4)); CK(cudaMemcpy(didx, &lvl, 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_array_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_array_dyn<<<1, 256, smem>>>(ms, didx, x, y, z, bytes, out);
    } else if (mode == 3) { CK(cudaFuncSetAttribute(k_array_uni, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_array_uni<<<1, 256, smem>>>(ms, lvl, x, y, z, bytes, out); }
    else {
        CUtensorMap* gm; CK(cudaMalloc(&gm, sizeof(CUtensorMap))); CK(cudaMemcpy(gm, &ms.m[lvl], sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_global<<<1, 256, smem>>>(gm, x, y, z, bytes, out);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> o(bytes); CK(cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) {
        const int xx = x + c, yy = y + r;
        const uint8_t want = (xx < X && yy < Y && xx >= 0 && yy >= 0) ? (uint8_t)(xx * 7 + yy * 13 + z * 31) : 0;
        if (o[r * bw + c] != want) { if (bad < 4) printf("mismatch at row %d col %d: got %d want %d\n", r, c, o[r * bw + c], want); ++bad; }
    }
    printf("mode %d box %dx%d at (%d,%d,%d): %s\n", mode, bw, bh, x, y, z, bad ? "MISMATCH" : "OK");
    return bad ? 1 : 0;
}

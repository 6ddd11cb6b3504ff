// Micro-benchmark: write-only bandwidth of three ways to cover a 2 GiB buffer with constant 128-bit stores,
// persistent grid of 444 CTAs x 256 threads (the decode kernel's shape).
//   A  a CTA writes one contiguous 31 KB tile at a time (256 threads x 8 stores), tiles round robin over CTAs
//   B  a WARP writes 8 consecutive 3968-byte tiles (a 31 KB chunk) on its own, chunks round robin over warps
//   C  the 8 warps of a CTA write 8 adjacent 3968-byte tiles, then the next 8 (a warp's tiles are 31 KB apart)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_patterns write_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st4(uint4 *p, uint4 v) { asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
constexpr int TW4 = 248;   // uint4 per 3968-byte tile
__global__ void __launch_bounds__(256, 3) pat(uint4 *out, uint64_t n_tiles, int mode)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 v = make_uint4(mode, 0, 0, 0);
    if (mode == 0) {
        for (uint64_t c = blockIdx.x; c * 8 < n_tiles; c += gridDim.x) {
            uint4 *d = out + c * 8 * TW4;
            for (uint32_t i = threadIdx.x; i < 8 * TW4; i += 256) st4(d + i, v);
        }
    } else if (mode == 1) {
        const uint64_t gw = (uint64_t)blockIdx.x * 8 + warp, GW = (uint64_t)gridDim.x * 8;
        for (uint64_t c = gw; c * 8 < n_tiles; c += GW)
            for (int s = 0; s < 8; s++) {
                uint4 *d = out + (c * 8 + s) * TW4;
                for (uint32_t i = lane; i < TW4; i += 32) st4(d + i, v);
            }
    } else {
        for (uint64_t c = blockIdx.x; c * 64 < n_tiles; c += gridDim.x)
            for (int s = 0; s < 8; s++) {
                uint4 *d = out + (c * 64 + s * 8 + warp) * TW4;
                for (uint32_t i = lane; i < TW4; i += 32) st4(d + i, v);
            }
    }
}
int main()
{
    const uint64_t n_tiles = (1ull << 31) / 3968 / 64 * 64;
    uint4 *d;
    cudaMalloc(&d, n_tiles * 3968);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++)
        for (int mode = 0; mode < 3; mode++) {
            cudaEventRecord(a);
            pat<<<444, 256>>>(d, n_tiles, mode);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            printf("pattern %c: %.3f ms  %.0f GB/s\n", 'A' + mode, ms, n_tiles * 3968.0 / ms / 1e6);
        }
    return 0;
}

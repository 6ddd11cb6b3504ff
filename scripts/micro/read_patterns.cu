// Micro-benchmark: read-only bandwidth over a 2 GiB buffer, the compress kernel's shapes.
//   A  plain 128-bit loads, 444 CTAs x 256 threads, 8 loads in flight per thread
//   B  TMA bulk loads of 31744-byte tiles into a 3-stage shared-memory ring, 2 CTAs per SM (296 CTAs), the 8 consumer
//      warps release a stage as soon as it has landed (hold time 0: the ceiling of the staging scheme)
//   C  the same, each stage held for HOLD cycles after it has landed (what the worker warps do: ~1.1 us = 2100 cycles)
//   D  B with 6 stages of 15872 bytes
//   E  C with 6 stages of 15872 bytes, each held HOLD / 2 cycles
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o read_patterns read_patterns.cu && ./read_patterns [HOLD]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 3) plain(const uint4 *in, uint64_t n16, uint32_t *sink)
{
    uint32_t acc = 0;
    const uint64_t step = (uint64_t)gridDim.x * 256 * 8;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 * 8 + threadIdx.x; i + 7 * 256 < n16; i += step) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(in + i + k * 256));
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int STAGES, int BYTES>
__global__ void __launch_bounds__(288, 2) staged(const char *in, uint64_t n_tiles, uint32_t hold, uint32_t *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (warp == 8) {
        if (lane == 0)
            for (uint64_t i = 0; i < n_my; i++) {
                const uint32_t s = i % STAGES, use = i / STAGES;
                if (use > 0) mbar_wait(smem_u32(&empty[s]), (use - 1) & 1);
                mbar_expect(smem_u32(&full[s]), BYTES);
                bulk_g2s(smem_u32(smem + (size_t)s * BYTES), in + (blockIdx.x + i * gridDim.x) * (uint64_t)BYTES, BYTES, smem_u32(&full[s]));
            }
    } else {
        uint32_t acc = 0;
        for (uint64_t i = 0; i < n_my; i++) {
            const uint32_t s = i % STAGES, use = i / STAGES;
            mbar_wait(smem_u32(&full[s]), use & 1);
            acc += reinterpret_cast<const uint32_t *>(smem + (size_t)s * BYTES)[threadIdx.x];
            if (hold) {
                const long long t0 = clock64();
                while (clock64() - t0 < (long long)hold) {}
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

int main(int argc, char **argv)
{
    const uint32_t hold = argc > 1 ? atoi(argv[1]) : 2100;
    const uint64_t bytes = 1ull << 31;
    char *d;
    uint32_t *sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(d, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaFuncSetAttribute(staged<3, 31744>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 31744);
    cudaFuncSetAttribute(staged<6, 15872>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 15872);
    for (int rep = 0; rep < 3; rep++)
        for (int mode = 0; mode < 5; mode++) {
            cudaEventRecord(a);
            if (mode == 0) plain<<<444, 256>>>((const uint4 *)d, bytes / 16, sink);
            else if (mode == 1) staged<3, 31744><<<296, 288, 3 * 31744>>>(d, bytes / 31744, 0, sink);
            else if (mode == 2) staged<3, 31744><<<296, 288, 3 * 31744>>>(d, bytes / 31744, hold, sink);
            else if (mode == 3) staged<6, 15872><<<296, 288, 6 * 15872>>>(d, bytes / 15872, 0, sink);
            else staged<6, 15872><<<296, 288, 6 * 15872>>>(d, bytes / 15872, hold / 2, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            printf("pattern %c: %.3f ms  %.0f GB/s  (%s)\n", 'A' + mode, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

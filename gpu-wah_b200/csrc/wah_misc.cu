// wah_misc.cu -- small helper kernels: shard probe for the multi-GPU seam merge and the
// synthetic bitvector generators used by bench.py and the tests.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

// result[0] lead_groups, [1] lead_words, [2] lead_type, [3] trail_groups, [4] trail_type
__global__ void wah_shard_probe_kernel(const uint32_t *shard, uint64_t words, uint64_t *result)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t lead_groups = 0, lead_words = 0, lead_type = 0, trail_groups = 0, trail_type = 0;
    if (words > 0) {
        const uint32_t first = shard[0];
        if (is_fill(first)) {
            lead_type = (first >> 30) & 1u;
            // a run longer than the 30-bit counter is stored as several full words + the rest
            while (lead_words < words) {
                const uint32_t w = shard[lead_words];
                if (!is_fill(w) || ((w >> 30) & 1u) != lead_type) break;
                lead_groups += fill_count(w);
                lead_words++;
            }
        }
        const uint32_t last = shard[words - 1];
        if (is_fill(last)) {
            trail_groups = fill_count(last);
            trail_type = (last >> 30) & 1u;
        }
    }
    result[0] = lead_groups;
    result[1] = lead_words;
    result[2] = lead_type;
    result[3] = trail_groups;
    result[4] = trail_type;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// bit b of word i is set iff the 32-bit draw for (seed, 32 i + b) is below density * 2^32
__global__ void wah_gen_uniform_kernel(uint32_t *out, uint64_t n_words, uint64_t threshold, uint64_t seed)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        uint32_t w = 0;
#pragma unroll
        for (int h = 0; h < 16; h++) {
            const uint64_t r = splitmix64(seed ^ ((i << 4) + (uint64_t)h) * 0xD6E8FEB86659FD93ull);
            w |= ((uint64_t)(uint32_t)r < threshold ? 1u : 0u) << (2 * h);
            w |= ((r >> 32) < threshold ? 1u : 0u) << (2 * h + 1);
        }
        out[i] = w;
    }
}

__global__ void wah_gen_paint_runs_kernel(uint32_t *out, uint64_t n_words, const int64_t *start,
                                          const int64_t *len, uint64_t n_runs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_bits = n_words * 32ull;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += stride) {
        uint64_t b0 = (uint64_t)start[r];
        uint64_t b1 = b0 + (uint64_t)len[r];
        if (b0 >= n_bits) continue;
        if (b1 > n_bits) b1 = n_bits;
        if (b1 <= b0) continue;
        const uint64_t w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
        const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u);
        const uint32_t m1 = 0xFFFFFFFFu >> (31u - (uint32_t)((b1 - 1) & 31u));
        if (w0 == w1) {
            atomicOr(out + w0, m0 & m1);
        } else {
            atomicOr(out + w0, m0);
            for (uint64_t w = w0 + 1; w < w1; w++) out[w] = 0xFFFFFFFFu;
            atomicOr(out + w1, m1);
        }
    }
}

}  // namespace

// ---- query operators -------------------------------------------------------------------------

// set bits of the vector a stream stands for, straight from the stream: a literal contributes its popcount, a
// one-fill 31 bits per group (kernels.cu:337-348 is what a decoder would expand it to), a zero-fill nothing
__global__ void wah_popcount_kernel(const uint32_t *in, uint64_t c_words, unsigned long long *bits)
{
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c_words; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = in[i];
        mine += is_fill(w) ? ((w & BIT30) ? 31ull * fill_count(w) : 0ull) : (unsigned long long)__popc(w);
    }
    mine = warp_sum_u64(mine);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(bits, mine);
}

// a[i] = a[i] op b[i] on decoded vectors, 128 bits per thread and step (n4 = number of uint4)
__global__ void wah_logical_kernel(int op, uint4 *a, const uint4 *b, uint64_t n4)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 x = a[i];
        const uint4 y = b[i];
        switch (op) {
        case 0: x.x &= y.x; x.y &= y.y; x.z &= y.z; x.w &= y.w; break;
        case 1: x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w; break;
        case 2: x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w; break;
        default: x.x &= ~y.x; x.y &= ~y.y; x.z &= ~y.z; x.w &= ~y.w; break;
        }
        a[i] = x;
    }
}

// ---- sparse host transfers (wah_host.cu): bitvectors are mostly zero, so the host path moves only the 4 KiB blocks
//      that hold a set bit.  Upload: the copy threads pack the non-zero blocks of a chunk and their block numbers; this
//      kernel puts them in place in a buffer that was cleared.  Download: the non-zero blocks of the decoded vector are
//      packed per 16 MiB chunk (block test, scan of the flags, copy) and only those cross PCIe.
constexpr uint32_t XBLK = 4096;   // bytes per block

// block list[i] of `dst` (dst_bytes long) = packed[i]; one CTA of 256 threads per block
__global__ void __launch_bounds__(256) wah_scatter_blocks_kernel(char *dst, uint64_t dst_bytes, const char *packed, const uint32_t *list,
                                                                 uint32_t n)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t off = (uint64_t)list[i] * XBLK;
        const uint4 v = reinterpret_cast<const uint4 *>(packed + (uint64_t)i * XBLK)[threadIdx.x];
        if (off + (uint64_t)(threadIdx.x + 1u) * 16u <= dst_bytes) reinterpret_cast<uint4 *>(dst + off)[threadIdx.x] = v;
    }
}

// flags[b] = block b of src holds a non-zero byte; one warp per block (src is 16-byte aligned, bytes a multiple of 4)
__global__ void __launch_bounds__(256) wah_flag_blocks_kernel(const char *src, uint64_t bytes, uint8_t *flags, uint64_t n_blocks)
{
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t b = (uint64_t)blockIdx.x * 8u + (threadIdx.x >> 5); b < n_blocks; b += (uint64_t)gridDim.x * 8u) {
        const uint64_t lo = b * XBLK;
        uint32_t any = 0;
        if (lo + XBLK <= bytes) {
            const uint4 *q = reinterpret_cast<const uint4 *>(src + lo);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint4 v = ld_stream_v4(q + lane + 32 * k);
                any |= v.x | v.y | v.z | v.w;
            }
        } else {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(src + lo);
            for (uint64_t i = lane; lo + 4ull * i < bytes; i += 32) any |= q[i];
        }
        any = __any_sync(0xffffffffu, any != 0u);
        if (lane == 0) flags[b] = (uint8_t)any;
    }
}

// per chunk of `per_chunk` blocks (<= 4096): lists[chunk][i] = number (within the chunk) of its i-th non-zero block,
// counts[chunk] = how many; one CTA of 1024 threads per chunk, four flags per thread
__global__ void __launch_bounds__(1024) wah_list_blocks_kernel(const uint8_t *flags, uint64_t n_blocks, uint32_t per_chunk, uint32_t *lists,
                                                               uint32_t *counts)
{
    __shared__ uint32_t s_w[32];
    const uint64_t base = (uint64_t)blockIdx.x * per_chunk;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    uint32_t f[4], mine = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t k = 4u * t + i;
        f[i] = (k < per_chunk && base + k < n_blocks) ? flags[base + k] : 0u;
        mine += f[i];
    }
    const uint32_t incl = warp_incl_scan(mine);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t before = incl - mine, total = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const uint32_t v = s_w[k];
        if (k < (int)warp) before += v;
        total += v;
    }
    uint32_t *list = lists + (uint64_t)blockIdx.x * per_chunk;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (f[i]) list[before++] = 4u * t + i;
    if (t == 0) counts[blockIdx.x] = total;
}

// packed[chunk][i] = block lists[chunk][i] of the chunk; grid.y = chunk, CTAs of 256 threads walk the list
__global__ void __launch_bounds__(256) wah_pack_blocks_kernel(const char *src, uint64_t bytes, uint32_t per_chunk, const uint32_t *lists,
                                                              const uint32_t *counts, char *packed)
{
    const uint32_t chunk = blockIdx.y, n = counts[chunk];
    const uint32_t *list = lists + (uint64_t)chunk * per_chunk;
    const uint64_t cbase = (uint64_t)chunk * per_chunk * XBLK;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t off = cbase + (uint64_t)list[i] * XBLK + 16ull * threadIdx.x;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (off + 16u <= bytes) {
            v = ld_stream_v4(reinterpret_cast<const uint4 *>(src + off));
        } else if (off < bytes) {   // the vector's last, partial 16 bytes (bytes is a multiple of 4)
            const uint32_t *q = reinterpret_cast<const uint32_t *>(src + off);
            v.x = q[0];
            if (off + 8u <= bytes) v.y = q[1];
            if (off + 12u <= bytes) v.z = q[2];
        }
        reinterpret_cast<uint4 *>(packed + cbase + (uint64_t)i * XBLK)[threadIdx.x] = v;
    }
}

cudaError_t launch_scatter_blocks(void *d_dst, uint64_t dst_bytes, const void *d_packed, const uint32_t *d_list, uint32_t n,
                                  cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    wah_scatter_blocks_kernel<<<n < 4096u ? n : 4096u, 256, 0, stream>>>((char *)d_dst, dst_bytes, (const char *)d_packed, d_list, n);
    return cudaGetLastError();
}

cudaError_t launch_pack_nonzero_blocks(const void *d_src, uint64_t bytes, uint32_t blocks_per_chunk, uint8_t *d_flags, uint32_t *d_lists,
                                       uint32_t *d_counts, void *d_packed, cudaStream_t stream)
{
    const uint64_t n_blocks = (bytes + XBLK - 1) / XBLK;
    if (n_blocks == 0) return cudaSuccess;
    if (blocks_per_chunk == 0 || blocks_per_chunk > 4096u) return cudaErrorInvalidValue;
    const uint32_t n_chunks = (uint32_t)((n_blocks + blocks_per_chunk - 1) / blocks_per_chunk);
    const uint64_t g = (n_blocks + 7) / 8;
    wah_flag_blocks_kernel<<<(unsigned)(g < 148ull * 16 ? g : 148ull * 16), 256, 0, stream>>>((const char *)d_src, bytes, d_flags, n_blocks);
    wah_list_blocks_kernel<<<n_chunks, 1024, 0, stream>>>(d_flags, n_blocks, blocks_per_chunk, d_lists, d_counts);
    wah_pack_blocks_kernel<<<dim3(64, n_chunks), 256, 0, stream>>>((const char *)d_src, bytes, blocks_per_chunk, d_lists, d_counts, (char *)d_packed);
    return cudaGetLastError();
}

cudaError_t launch_shard_probe(const uint32_t *d_shard, uint64_t words, uint64_t *d_result, cudaStream_t stream)
{
    wah_shard_probe_kernel<<<1, 32, 0, stream>>>(d_shard, words, d_result);
    return cudaGetLastError();
}

cudaError_t launch_gen_uniform(uint32_t *d_out, uint64_t n_words, double density, uint64_t seed,
                               cudaStream_t stream)
{
    if (n_words == 0) return cudaSuccess;
    double t = density * 4294967296.0;
    if (t < 0) t = 0;
    if (t > 4294967296.0) t = 4294967296.0;
    const uint64_t threshold = (uint64_t)t;
    const uint64_t blocks = (n_words + 255) / 256;
    const int grid = (int)(blocks < 148ull * 16 ? blocks : 148ull * 16);
    wah_gen_uniform_kernel<<<grid, 256, 0, stream>>>(d_out, n_words, threshold, seed);
    return cudaGetLastError();
}

cudaError_t launch_gen_paint_runs(uint32_t *d_out, uint64_t n_words, const int64_t *d_start,
                                  const int64_t *d_len, uint64_t n_runs, cudaStream_t stream)
{
    if (n_runs == 0) return cudaSuccess;
    const uint64_t blocks = (n_runs + 255) / 256;
    const int grid = (int)(blocks < 148ull * 16 ? blocks : 148ull * 16);
    wah_gen_paint_runs_kernel<<<grid, 256, 0, stream>>>(d_out, n_words, d_start, d_len, n_runs);
    return cudaGetLastError();
}

cudaError_t launch_popcount(const uint32_t *d_in, uint64_t c_words, uint64_t *d_bits, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_bits, 0, sizeof(uint64_t), stream);
    if (e != cudaSuccess || c_words == 0) return e;
    const uint64_t want = (c_words + 1023) / 1024;
    wah_popcount_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, stream>>>(d_in, c_words, (unsigned long long *)d_bits);
    return cudaGetLastError();
}

cudaError_t launch_logical(int op, uint32_t *d_a, const uint32_t *d_b, uint64_t n_words, cudaStream_t stream)
{
    const uint64_t n4 = (n_words + 3) / 4;   // the buffers are padded to whole 16-byte units
    if (n4 == 0) return cudaSuccess;
    const uint64_t want = (n4 + 1023) / 1024;
    wah_logical_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, stream>>>(op, (uint4 *)d_a, (const uint4 *)d_b, n4);
    return cudaGetLastError();
}

}  // namespace wahb200

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

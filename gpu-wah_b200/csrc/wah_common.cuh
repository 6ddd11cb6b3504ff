// wah_common.cuh -- shared device helpers for the sm_100a WAH kernels.
//
// Format constants follow the reference's const.h:3-12 (ONES31, BIT31, BIT30).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wahb200 {

constexpr uint32_t ONES31   = 0x7FFFFFFFu;
constexpr uint32_t BIT31    = 0x80000000u;
constexpr uint32_t BIT30    = 0x40000000u;
constexpr uint32_t MAX_FILL = 0x3FFFFFFFu;   // 30-bit run counter (kernels.cu:300,334)

// One thread owns 32 consecutive 31-bit groups = 31 consecutive input words:
// the "row" of the reference (kernels.cu:68), kept in one thread instead of one warp.
constexpr int GROUPS_PER_THREAD = 32;
constexpr int WORDS_PER_THREAD  = 31;

// ---------------------------------------------------------------- memory ops

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming 128-bit load: read once, do not keep in L1
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// streaming stores: written once, never re-read by this kernel
__device__ __forceinline__ void st_stream_v4(uint4 *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The same with an L2 cache policy.  Output that is written once and not read again by the kernel is stored evict-first:
// measured on the decoder at 16 Gbit, 0.367 -> 0.336 ms for a stream of long fills (2 GB of stores that otherwise sit in
// the L2 until something pushes them out).
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_stream_v4_hint(uint4 *p, const uint4 &v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// The two big kernels of a compress / decompress pipeline follow each other on one stream.  Launched with
// programmatic stream serialisation, the next kernel's CTAs are scheduled as soon as the previous kernel's CTAs
// leave their SMs and wait here until that kernel has completed and flushed its memory: its launch latency and the
// previous kernel's drain overlap.  Nothing that another kernel may have written is touched before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- warp helpers

__device__ __forceinline__ uint32_t lane_id()
{
    return threadIdx.x & 31u;
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    // (the shuffle's own predicate says whether the source lane exists: two instructions per step instead of three)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile(
            "{\n\t"
            ".reg .u32 r;\n\t"
            ".reg .pred p;\n\t"
            "shfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff;\n\t"
            "@p add.u32 %0, %0, r;\n\t"
            "}"
            : "+r"(v)
            : "r"(d));
    return v;
}

__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v)
{
    const uint32_t lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += o;
    }
    return v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ---------------------------------------------------------------- WAH words

__device__ __forceinline__ uint32_t fill_word(uint32_t type, uint32_t count)
{
    // kernels.cu:244-248: BIT3130 | n for ones, BIT31 | n for zeros
    return BIT31 | (type << 30) | count;
}

__device__ __forceinline__ bool is_fill(uint32_t w) { return (w & BIT31) != 0u; }
__device__ __forceinline__ uint32_t fill_count(uint32_t w) { return w & MAX_FILL; }
__device__ __forceinline__ uint32_t word_groups(uint32_t w)
{
    // getCounts, kernels.cu:298-304
    return is_fill(w) ? fill_count(w) : 1u;
}

// 31-bit group g (0..) of a word array that starts on a group boundary:
// stream bits [31g, 31g+31), LSB first (kernels.cu:79).  Reads words (31g)>>5 and +1.
__device__ __forceinline__ uint32_t extract_group(const uint32_t *row, uint32_t g)
{
    const uint32_t bit = 31u * g;
    const uint32_t wi = bit >> 5, s = bit & 31u;
    return __funnelshift_r(row[wi], row[wi + 1], s) & ONES31;
}

}  // namespace wahb200

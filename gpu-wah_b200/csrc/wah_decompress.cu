// wah_decompress.cu -- WAH decompressor for sm_100a.
//
// Replaces the reference's getCounts + thrust::exclusive_scan + decompressWords +
// mergeWords (kernels.cu:291-385, decompress.cu:66-115), which materialise an
// 8-byte count per compressed word, a one-group-per-int intermediate array and
// expand every fill with a serial per-thread loop (kernels.cu:346-348).
//
// Here:
//   scan kernel   : one pass over the compressed words; per-tile group sums with a
//                   decoupled look-back give every tile its group offset, and each
//                   tile records, for every OUTPUT tile boundary (multiples of 8192
//                   groups) that falls into it, which compressed word covers it.
//   expand kernel : output-centric and therefore load balanced whatever the fill
//                   lengths are: a persistent grid walks output tiles of 8192 groups
//                   = 7936 words; one thread produces 32 groups = 31 output words
//                   (binary search for its first source word, then a short walk),
//                   31->32 repack with funnel shifts (kernels.cu:375), staged through
//                   shared memory and written as full 128-bit lines.
// Output tiles are aligned in group space to multiples of 32 groups = 31 words, so no
// output word is shared between threads or tiles and nothing needs atomics.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

constexpr uint64_t ST_EMPTY = 0, ST_AGG = 1, ST_INCL = 2;
constexpr uint64_t VALUE_MASK = (1ull << 62) - 1ull;

__device__ __forceinline__ uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ scan kernel

__global__ void __launch_bounds__(SCAN_THREADS) wah_scan_kernel(const ScanParams p)
{
    constexpr int NW = SCAN_THREADS / 32;
    constexpr uint64_t TG = EXPAND_TILE_GROUPS;
    __shared__ uint64_t s_wsum[NW];
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&p.hdr->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;

    // blocked arrangement: thread owns SCAN_ITEMS consecutive compressed words
    const uint64_t w_begin = (uint64_t)tile * SCAN_TILE_WORDS + (uint64_t)tid * SCAN_ITEMS;
    uint32_t w[SCAN_ITEMS];
    if (w_begin + SCAN_ITEMS <= p.c_words) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.in + w_begin);
#pragma unroll
        for (int v = 0; v < SCAN_ITEMS / 4; v++) {
            const uint4 x = ld_stream_v4(src + v);
            w[4 * v + 0] = x.x;
            w[4 * v + 1] = x.y;
            w[4 * v + 2] = x.z;
            w[4 * v + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++)
            w[i] = (w_begin + i < p.c_words) ? ld_stream_u32(p.in + w_begin + i) : BIT31;   // fill of 0 groups
    }

    uint32_t cnt[SCAN_ITEMS];
    uint64_t tsum = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        cnt[i] = word_groups(w[i]);   // getCounts, kernels.cu:298-304
        tsum += cnt[i];
        bad += (cnt[i] == 0u && w_begin + i < p.c_words) ? 1u : 0u;
    }
    bad = warp_sum(bad);
    if (lane == 0 && bad) atomicAdd(&p.hdr->bad_words, bad);

    const uint64_t incl = warp_incl_scan_u64(tsum);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();

    // ---- tile total + decoupled look-back (status:2 | value:62)
    if (warp == 0) {
        uint64_t tile_sum = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) tile_sum += s_wsum[k];
        uint64_t excl = 0;
        if (tile == 0u) {
            if (lane == 0) st_relaxed_u64(p.desc, (ST_INCL << 62) | tile_sum);
        } else {
            if (lane == 0) st_relaxed_u64(p.desc + tile, (ST_AGG << 62) | tile_sum);
            int64_t look = (int64_t)tile - 1 - (int64_t)lane;
            while (true) {
                uint64_t d;
                do {
                    d = look >= 0 ? ld_relaxed_u64(p.desc + look) : (ST_INCL << 62);
                } while (__any_sync(0xffffffffu, (d >> 62) == ST_EMPTY));
                const uint32_t incl_mask = __ballot_sync(0xffffffffu, (d >> 62) == ST_INCL);
                const uint32_t first_incl = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 32u;
                excl += warp_sum_u64(lane <= first_incl ? (d & VALUE_MASK) : 0ull);
                if (first_incl < 32u) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(p.desc + tile, (ST_INCL << 62) | (excl + tile_sum));
        }
        if (lane == 0) {
            s_base = excl;
            if (tile == p.n_tiles - 1u) {
                // decompress.cu:82-93: G = last offset + last count, realSize = ceil(31 G / 32)
                const uint64_t G = excl + tile_sum;
                const uint64_t words = (G >> 5) * 31ull + (((G & 31ull) * 31ull + 31ull) >> 5);
                p.hdr->groups = G;
                p.hdr->words = words;
                p.hdr->out_tiles = ceil_div_u64(G, TG);
                if (p.out_info) {
                    p.out_info[0] = words;
                    p.out_info[1] = G;
                }
            }
        }
    }
    __syncthreads();

    if (p.starts == nullptr) return;

    // ---- which compressed word covers each output-tile boundary k*TG ?
    uint64_t wprefix = 0;
#pragma unroll
    for (int k = 0; k < NW; k++)
        if (k < (int)warp) wprefix += s_wsum[k];
    uint64_t off = s_base + wprefix + incl - tsum;   // group offset of my first word
    const uint64_t k_limit = p.max_out_tiles + 1ull;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        // boundaries with off <= k*TG < off + cnt
        uint64_t k_first = ceil_div_u64(off, TG);
        uint64_t k_end = ceil_div_u64(off + cnt[i], TG);
        if (k_end > k_limit) k_end = k_limit;
        if (k_first > k_end) k_first = k_end;
        const bool heavy = (k_end - k_first) > 4ull;
        uint32_t hm = __ballot_sync(0xffffffffu, heavy);
        while (hm) {
            // a long fill spans many output tiles: the whole warp writes its boundaries
            const int srcl = __ffs(hm) - 1;
            hm &= hm - 1u;
            const uint64_t kf = __shfl_sync(0xffffffffu, k_first, srcl);
            const uint64_t ke = __shfl_sync(0xffffffffu, k_end, srcl);
            const uint64_t o = __shfl_sync(0xffffffffu, off, srcl);
            const uint64_t wi = __shfl_sync(0xffffffffu, w_begin, srcl) + (uint64_t)i;
            for (uint64_t k = kf + lane; k < ke; k += 32) p.starts[k] = make_ulonglong2(wi, o);
        }
        if (!heavy)
            for (uint64_t k = k_first; k < k_end; k++) p.starts[k] = make_ulonglong2(w_begin + i, off);
        off += cnt[i];
    }
}

// ---------------------------------------------------------------- expand kernel

__global__ void __launch_bounds__(EXPAND_THREADS, 3) wah_expand_kernel(const ExpandParams p)
{
    constexpr int NW = EXPAND_THREADS / 32;
    constexpr uint32_t CLAMP = 2u * EXPAND_TILE_GROUPS;   // any count >= tile span behaves the same
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_cw = smem;                         // EXPAND_MAX_CWORDS compressed words
    uint32_t *s_off = smem + EXPAND_MAX_CWORDS;    // their group offsets relative to the tile start
    uint32_t *s_stage = smem;                      // EXPAND_TILE_WORDS output words (aliases the above)
    __shared__ uint32_t s_wsum[NW];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t G = p.hdr->groups;
    uint64_t total_words = p.hdr->words;
    if (total_words > p.out_cap) total_words = p.out_cap;
    const uint64_t real_tiles = p.hdr->out_tiles;
    const uint64_t n_tiles = real_tiles < p.max_out_tiles ? real_tiles : p.max_out_tiles;

    for (uint64_t ot = blockIdx.x; ot < n_tiles; ot += gridDim.x) {
        const ulonglong2 st = p.starts[ot];
        const uint64_t ws = st.x;
        const uint64_t we = (ot + 1 < real_tiles) ? p.starts[ot + 1].x : p.c_words - 1;
        uint32_t nw = (uint32_t)(we - ws + 1);            // <= EXPAND_TILE_GROUPS + 1 for a well-formed stream
        if (nw > (uint32_t)EXPAND_MAX_CWORDS) nw = EXPAND_MAX_CWORDS;   // zero-length fills: flagged by the scan
        const uint64_t g_lo = ot * (uint64_t)EXPAND_TILE_GROUPS;
        const uint32_t skip = (uint32_t)(g_lo - st.y);    // groups of word ws that belong to earlier tiles

        for (uint32_t i = tid; i < nw; i += EXPAND_THREADS) s_cw[i] = ld_stream_u32(p.in + ws + i);
        __syncthreads();

        // ---- block scan of the (clamped) group counts -> s_off
        const uint32_t ipt = (nw + EXPAND_THREADS - 1) / EXPAND_THREADS;
        const uint32_t i0 = tid * ipt;
        uint32_t tsum = 0;
        for (uint32_t i = i0; i < i0 + ipt && i < nw; i++) {
            uint32_t c = word_groups(s_cw[i]);
            if (i == 0) c -= skip;
            tsum += c > CLAMP ? CLAMP : c;
        }
        const uint32_t incl = warp_incl_scan(tsum);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - tsum;
#pragma unroll
        for (int k = 0; k < NW; k++)
            if (k < (int)warp) run += s_wsum[k];
        for (uint32_t i = i0; i < i0 + ipt && i < nw; i++) {
            uint32_t c = word_groups(s_cw[i]);
            if (i == 0) c -= skip;
            s_off[i] = run;
            run += c > CLAMP ? CLAMP : c;
        }
        __syncthreads();

        // ---- my 32 groups -> 31 output words
        const uint32_t rel = 32u * tid;
        const bool active = g_lo + rel < G;
        uint32_t o[31];
        if (active) {
            // last word whose offset is <= rel
            uint32_t lo = 0, hi = nw;   // invariant: s_off[lo] <= rel, answer in [lo, hi)
            while (hi - lo > 1u) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_off[mid] <= rel) lo = mid;
                else hi = mid;
            }
            uint32_t idx = lo;
            uint32_t wv = s_cw[idx];
            uint32_t c = word_groups(wv);
            if (idx == 0) c -= skip;
            if (c > CLAMP) c = CLAMP;
            uint32_t rem = s_off[idx] + c - rel;                                   // groups of word idx left at rel
            uint32_t val = is_fill(wv) ? ((wv & BIT30) ? ONES31 : 0u) : wv;       // kernels.cu:337-354
            if (rem >= 32u) {
                // whole chunk inside one fill
                const uint32_t f = val ? 0xFFFFFFFFu : 0u;
#pragma unroll
                for (int j = 0; j < 31; j++) o[j] = f;
            } else {
                uint32_t pg = 0;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (rem == 0u) {
                        do {
                            idx++;
                            if (idx >= nw) {   // past the end of the stream: zero padding
                                val = 0u;
                                rem = 64u;
                                break;
                            }
                            wv = s_cw[idx];
                            rem = word_groups(wv);
                            val = is_fill(wv) ? ((wv & BIT30) ? ONES31 : 0u) : wv;
                        } while (rem == 0u);
                    }
                    // 31 -> 32 repack (mergeWords, kernels.cu:375):
                    // word j-1 = group[j-1] >> (j-1) | group[j] << (32-j)
                    if (j > 0) o[j - 1] = __funnelshift_r(pg << 1, val, j);
                    pg = val;
                    rem--;
                }
            }
        }
        __syncthreads();   // everyone is done with s_cw / s_off
        if (active) {
#pragma unroll
            for (int j = 0; j < 31; j++) s_stage[31u * tid + j] = o[j];
        }
        __syncthreads();

        // ---- coalesced write of the tile
        const uint64_t w_lo = ot * (uint64_t)EXPAND_TILE_WORDS;
        if (w_lo < total_words) {
            const uint64_t avail = total_words - w_lo;
            const uint32_t nout = avail < (uint64_t)EXPAND_TILE_WORDS ? (uint32_t)avail : (uint32_t)EXPAND_TILE_WORDS;
            uint32_t *dst = p.out + w_lo;
            const uint32_t nvec = nout >> 2;
            uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
            const uint4 *src4 = reinterpret_cast<const uint4 *>(s_stage);
            for (uint32_t i = tid; i < nvec; i += EXPAND_THREADS) st_stream_v4(dst4 + i, src4[i]);
            for (uint32_t i = (nvec << 2) + tid; i < nout; i += EXPAND_THREADS) dst[i] = s_stage[i];
        }
        __syncthreads();   // staging is reused as s_cw by the next tile
    }
}

}  // namespace

size_t expand_smem_bytes()
{
    return (size_t)(2 * EXPAND_MAX_CWORDS) * sizeof(uint32_t);
}

cudaError_t launch_scan(const ScanParams &p, cudaStream_t stream)
{
    wah_scan_kernel<<<p.n_tiles, SCAN_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandParams &p, int grid, cudaStream_t stream)
{
    const size_t smem = expand_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(wah_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wah_expand_kernel<<<grid, EXPAND_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace wahb200

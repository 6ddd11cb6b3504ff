// wah_compress.cu -- single-pass WAH compressor for sm_100a.
//
// Replaces the reference's compressData + thrust::exclusive_scan + moveData
// (kernels.cu:51-280, compress.cu:129-166) with ONE kernel: every input byte is
// read once and every output byte written once (4n + 4c bytes of HBM traffic).
//
// Work decomposition (B200-first, not the reference's):
//   thread  = 32 consecutive 31-bit groups = 31 input words (the reference gives
//             this "row" to a whole warp, kernels.cu:68)
//   warp    = 1024 groups = 992 words = one reference block, so in BLOCK1024 mode
//             (bit-exact to the reference: runs never cross a block, kernels.cu:256)
//             no run state ever leaves a warp
//   CTA     = tile of THREADS*31 words staged in shared memory with 16-byte
//             cp.async copies, classified from shared memory (stride 31 words per
//             thread = bank-conflict free)
//   grid    = tiles in ticket order; tile prefix (words emitted so far, length of
//             the fill run still open at the tile's end) by decoupled look-back over
//             one 64-bit descriptor per tile
//
// A run is emitted where it ENDS ("tail"): group k is a tail if it is a literal,
// or the next group has another type, or it is the last group of the stream /
// block.  The k-th tail is the k-th output word; a fill's length is the distance
// to the previous tail, which flows forward through the scans.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

// ---- tile descriptor: status:2 | no_tail:1 | open:30 | count:31 -------------------
constexpr uint64_t ST_EMPTY = 0, ST_AGG = 1, ST_INCL = 2;

__device__ __forceinline__ uint64_t desc_pack(uint64_t st, uint32_t no_tail, uint32_t open, uint32_t count)
{
    return (st << 62) | ((uint64_t)no_tail << 61) | ((uint64_t)open << 31) | (uint64_t)count;
}
__device__ __forceinline__ uint32_t desc_status(uint64_t d) { return (uint32_t)(d >> 62); }
__device__ __forceinline__ uint32_t desc_no_tail(uint64_t d) { return (uint32_t)(d >> 61) & 1u; }
__device__ __forceinline__ uint32_t desc_open(uint64_t d) { return (uint32_t)(d >> 31) & MAX_FILL; }
__device__ __forceinline__ uint32_t desc_count(uint64_t d) { return (uint32_t)d & 0x7FFFFFFFu; }

__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_addr, const void *gptr, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int THREADS, bool BLOCK_MODE>
__global__ void __launch_bounds__(THREADS, 3) wah_compress_kernel(const CompressParams p)
{
    constexpr int NW = THREADS / 32;
    constexpr int TILE_WORDS = THREADS * WORDS_PER_THREAD;
    constexpr int TILE_GROUPS = THREADS * GROUPS_PER_THREAD;
    constexpr int IN_WORDS = TILE_WORDS + 4;   // + look-ahead word, padded to 16 B
    constexpr int NVEC = IN_WORDS / 4;
    static_assert(TILE_WORDS % 4 == 0, "tile must be a whole number of 16-byte vectors");

    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_in = smem;               // IN_WORDS
    uint32_t *s_out = smem + IN_WORDS;   // TILE_GROUPS staged output words

    __shared__ uint32_t s_wcnt[NW];      // words emitted by each warp
    __shared__ uint32_t s_wopen[NW];     // groups after the warp's last tail (NW*32... if none: 1024)
    __shared__ uint32_t s_whas[NW];      // warp has at least one tail
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_excl;          // words emitted by earlier tiles of this launch
    __shared__ uint32_t s_carry;         // length of the run still open when this tile starts
    __shared__ uint32_t s_drop;          // 1 if the tile's first word was merged into out[base-1]
    __shared__ uint32_t s_tcnt;          // words emitted by this tile
    __shared__ uint32_t s_topen;         // groups after the tile's last tail

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // ticket order = start order, so every predecessor a tile waits on is resident or done
    if (tid == 0) {
        s_tile = atomicAdd(p.ticket, 1u);
        s_drop = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t col = tile / p.tiles_per_col;
    const uint32_t t = tile - col * p.tiles_per_col;           // tile index inside the column
    const uint64_t w0 = (uint64_t)t * TILE_WORDS;
    const uint32_t *src = p.in + (uint64_t)col * p.col_stride + w0;
    const uint64_t left = p.n_words - w0;                       // words from the tile start to the column end
    const uint32_t nload = left > (uint64_t)TILE_WORDS ? (uint32_t)TILE_WORDS + 1u : (uint32_t)left;

    // ---- stage the tile (+1 look-ahead word), zero filled past the end of the column
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_in);
#pragma unroll
        for (int k = 0; k < (NVEC + THREADS - 1) / THREADS; k++) {
            const uint32_t i = tid + k * THREADS;
            if (i < NVEC) {
                const uint32_t b = 4u * i;
                const uint32_t bytes = b >= nload ? 0u : (nload - b >= 4u ? 16u : (nload - b) * 4u);
                cp_async_16_zfill(s_base + 16u * i, bytes ? (const void *)(src + b) : (const void *)src, bytes);
            }
        }
        cp_async_wait_all();
    } else {
        for (uint32_t i = tid; i < IN_WORDS; i += THREADS) s_in[i] = i < nload ? ld_stream_u32(src + i) : 0u;
    }
    __syncthreads();

    // ---- classify my 32 groups (kernels.cu:79 regroup, :93-112 classification)
    const uint32_t *row = s_in + WORDS_PER_THREAD * tid;
    const uint64_t g_thread = (uint64_t)t * TILE_GROUPS + 32u * tid;   // first group, column relative
    uint32_t nvalid = 0;
    if (g_thread < p.groups) nvalid = (p.groups - g_thread) >= 32u ? 32u : (uint32_t)(p.groups - g_thread);
    const uint32_t vmask = nvalid == 32u ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);

    uint32_t Z = 0, O = 0;
    {
        uint32_t prev = row[0];
        uint32_t v = prev & ONES31;
        Z |= (v == 0u) ? 1u : 0u;
        O |= (v == ONES31) ? 1u : 0u;
#pragma unroll
        for (int j = 1; j < 31; j++) {
            const uint32_t cur = row[j];
            v = __funnelshift_r(prev, cur, 32 - j) & ONES31;
            Z |= (v == 0u) ? (1u << j) : 0u;
            O |= (v == ONES31) ? (1u << j) : 0u;
            prev = cur;
        }
        v = prev >> 1;
        Z |= (v == 0u) ? BIT31 : 0u;
        O |= (v == ONES31) ? BIT31 : 0u;
    }
    Z &= vmask;
    O &= vmask;
    const uint32_t F = Z | O;

    // type of the group after my chunk (first group of the next thread)
    uint32_t nz = 0, no = 0;
    if (!(BLOCK_MODE && lane == 31u) && g_thread + 32u < p.groups) {
        const uint32_t nx = row[31] & ONES31;
        nz = (nx == 0u) ? BIT31 : 0u;
        no = (nx == ONES31) ? BIT31 : 0u;
    }
    // tail = literal, or fill whose successor differs (run-end rule, kernels.cu:126-141);
    // the last group of the stream, and of every 1024-group block in BLOCK mode, has no successor
    const uint32_t T = ((~F) & vmask) | (Z & ~((Z >> 1) | nz)) | (O & ~((O >> 1) | no));
    const uint32_t cnt = __popc(T);
    const uint32_t my_open = T ? (uint32_t)__clz(T) : 32u;   // groups after my last tail

    // ---- warp scan: output offset and length of the run open at my chunk's start
    const uint32_t incl = warp_incl_scan(cnt);
    const uint32_t tb = __ballot_sync(0xffffffffu, T != 0u);
    const uint32_t below = tb & lanemask_lt();
    const uint32_t q = below ? 31u - (uint32_t)__clz(below) : 0u;
    const uint32_t open_q = __shfl_sync(0xffffffffu, my_open, q);
    uint32_t prev_open = below ? open_q + 32u * (lane - q - 1u) : 32u * lane;   // + warp carry if !below
    const uint32_t qlast = tb ? 31u - (uint32_t)__clz(tb) : 0u;
    const uint32_t open_last = __shfl_sync(0xffffffffu, my_open, qlast);
    const uint32_t wcnt = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
        s_wcnt[warp] = wcnt;
        s_whas[warp] = tb != 0u;
        s_wopen[warp] = tb ? open_last + 32u * (31u - qlast) : 1024u;
    }
    __syncthreads();

    // ---- tile aggregate + decoupled look-back (warp 0)
    if (warp == 0) {
        uint32_t tile_cnt = 0, tile_open = 0, tile_has = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            tile_cnt += s_wcnt[w];
            if (s_whas[w]) {
                tile_has = 1;
                tile_open = s_wopen[w];
            } else {
                tile_open += s_wopen[w];
            }
        }
        // the end of a column is always a tail (lanes past the end hold no groups)
        if (t == p.tiles_per_col - 1u) {
            tile_open = 0;
            tile_has = 1;
        }
        if (lane == 0) {
            s_tcnt = tile_cnt;
            s_topen = tile_open;
        }
        const bool defer_publish = (tile == 0u) && p.merge_prev;
        uint32_t excl = 0, carry = 0;
        if (tile == 0u) {
            if (lane == 0 && !defer_publish)
                st_relaxed_u64(p.desc, desc_pack(ST_INCL, 0, tile_open, tile_cnt));
        } else {
            if (lane == 0) st_relaxed_u64(p.desc + tile, desc_pack(ST_AGG, tile_has ^ 1u, tile_open, tile_cnt));
            // BLOCK mode never carries a run; CANONICAL restarts at every column
            bool open_done = BLOCK_MODE || (t == 0u);
            int64_t look = (int64_t)tile - 1 - (int64_t)lane;
            while (true) {
                uint64_t d;
                do {
                    d = look >= 0 ? ld_relaxed_u64(p.desc + look) : desc_pack(ST_INCL, 0, 0, 0);
                } while (__any_sync(0xffffffffu, desc_status(d) == ST_EMPTY));
                const uint32_t incl_mask = __ballot_sync(0xffffffffu, desc_status(d) == ST_INCL);
                const uint32_t first_incl = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 32u;
                const bool part = lane <= first_incl;     // first_incl == 32: every lane takes part
                excl += warp_sum(part ? desc_count(d) : 0u);
                if (!open_done) {
                    // walk towards older tiles until one ends a run (or carries a resolved value)
                    const uint32_t term_mask =
                        __ballot_sync(0xffffffffu, part && (desc_status(d) == ST_INCL || !desc_no_tail(d)));
                    const uint32_t first_term = term_mask ? (uint32_t)__ffs(term_mask) - 1u : 32u;
                    carry += warp_sum((part && lane <= first_term) ? desc_open(d) : 0u);
                    open_done = first_term < 32u;
                }
                if (first_incl < 32u) break;
                look -= 32;
            }
            const uint32_t incl_open = tile_has ? tile_open : carry + tile_open;
            if (lane == 0) st_relaxed_u64(p.desc + tile, desc_pack(ST_INCL, 0, incl_open, excl + tile_cnt));
        }
        if (lane == 0) {
            s_excl = excl;
            s_carry = (BLOCK_MODE || t == 0u) ? 0u : carry;
        }
    }
    __syncthreads();

    // ---- emit into the staging buffer
    uint32_t wprefix = 0, wcarry = 0;
    {
        bool found = false;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (w < (int)warp) {
                wprefix += s_wcnt[w];
                if (s_whas[w]) {
                    wcarry = s_wopen[w];
                    found = true;
                } else {
                    wcarry += s_wopen[w];
                }
            }
        }
        if (!found) wcarry += s_carry;
    }
    if (BLOCK_MODE) wcarry = 0;
    if (!below) prev_open += wcarry;
    const uint32_t my_off = wprefix + incl - cnt;

    const bool all_literal = __all_sync(0xffffffffu, T == 0xFFFFFFFFu && F == 0u);
    if (all_literal) {
        // every group of the warp is a literal: lane-per-output-word copy, conflict free
        const uint32_t *wrow = s_in + 992u * warp;
        uint32_t *wout = s_out + wprefix;
#pragma unroll 8
        for (uint32_t k = 0; k < 32u; k++) {
            const uint32_t g = 32u * k + lane;
            wout[g] = extract_group(wrow, g);
        }
    } else if (wcnt > 192u) {
        // dense warp: lanes cooperate on one thread-chunk at a time (conflict-free staging)
        const uint32_t *wrow = s_in + 992u * warp;
        for (uint32_t k = 0; k < 32u; k++) {
            const uint32_t Tk = __shfl_sync(0xffffffffu, T, k);
            if (Tk == 0u) continue;
            const uint32_t Fk = __shfl_sync(0xffffffffu, F, k);
            const uint32_t Ok = __shfl_sync(0xffffffffu, O, k);
            const uint32_t offk = __shfl_sync(0xffffffffu, my_off, k);
            const uint32_t pok = __shfl_sync(0xffffffffu, prev_open, k);
            const uint32_t bit = 1u << lane;
            if (Tk & bit) {
                const uint32_t lower = Tk & (bit - 1u);
                uint32_t w;
                if (Fk & bit) {
                    const uint32_t len = lower ? lane - (31u - (uint32_t)__clz(lower)) : lane + 1u + pok;
                    w = fill_word((Ok >> lane) & 1u, len);
                } else {
                    w = extract_group(wrow, 32u * k + lane);
                }
                s_out[offk + __popc(lower)] = w;
            }
        }
    } else {
        // sparse warp: every thread walks its own few tails
        uint32_t m = T, off = my_off;
        int prev = -1;
        uint32_t extra = prev_open;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1u;
            uint32_t w;
            if ((F >> j) & 1u) {
                w = fill_word((O >> j) & 1u, (uint32_t)(j - prev) + extra);
            } else {
                w = extract_group(row, (uint32_t)j);
            }
            s_out[off++] = w;
            prev = j;
            extra = 0;
        }
    }
    __syncthreads();

    // ---- tile totals, seam with an earlier launch, write out
    const uint32_t tile_cnt = s_tcnt;
    const uint64_t base = p.base_in ? *p.base_in : 0ull;

    if (tile == 0u && p.merge_prev) {
        // CANONICAL append: this launch continues a stream; fold my first run into its last word
        if (tid == 0) {
            uint32_t drop = 0;
            if (base > 0 && tile_cnt > 0 && base - 1 < p.out_cap) {
                const uint32_t pw = p.out[base - 1], fw = s_out[0];
                if (is_fill(pw) && is_fill(fw) && ((pw ^ fw) & BIT30) == 0u) {
                    const uint64_t total = (uint64_t)fill_count(pw) + fill_count(fw);
                    const uint32_t ty = (fw >> 30) & 1u;
                    if (total <= MAX_FILL) {
                        p.out[base - 1] = fill_word(ty, (uint32_t)total);
                        drop = 1;
                    } else {
                        p.out[base - 1] = fill_word(ty, MAX_FILL);
                        s_out[0] = fill_word(ty, (uint32_t)(total - MAX_FILL));
                    }
                }
            }
            s_drop = drop;
            st_relaxed_u64(p.desc, desc_pack(ST_INCL, 0, s_topen, tile_cnt - drop));
        }
        __syncthreads();
    }
    const uint32_t drop = s_drop;                 // only ever non-zero in tile 0
    const uint64_t dst0 = base + s_excl;          // s_excl already accounts for a dropped word
    if (tid == 0) {
        if (p.col_offsets && t == 0u) p.col_offsets[col] = dst0;
        if (tile == p.n_tiles - 1u) {
            const uint64_t total = dst0 + tile_cnt - drop;
            *p.total_out = total;
            if (p.col_offsets) p.col_offsets[p.n_cols] = total;
        }
    }
    for (uint32_t i = tid + drop; i < tile_cnt; i += THREADS) {
        const uint64_t idx = dst0 + i - drop;
        if (idx < p.out_cap) p.out[idx] = s_out[i];
    }
}

}  // namespace

size_t compress_smem_bytes()
{
    return (size_t)(COMPRESS_TILE_WORDS + 4 + COMPRESS_TILE_GROUPS) * sizeof(uint32_t);
}

cudaError_t launch_compress(const CompressParams &p, int mode, cudaStream_t stream)
{
    const size_t smem = compress_smem_bytes();
    cudaError_t e;
    if (mode == 0) {
        auto k = wah_compress_kernel<COMPRESS_THREADS, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<p.n_tiles, COMPRESS_THREADS, smem, stream>>>(p);
    } else {
        auto k = wah_compress_kernel<COMPRESS_THREADS, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<p.n_tiles, COMPRESS_THREADS, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace wahb200

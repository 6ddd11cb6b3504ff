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
//   CTA     = persistent, launched cooperatively (all CTAs co-resident), walks tiles
//             blockIdx + k * gridDim of THREADS*31 words.  The next tile's input is
//             already in flight (cp.async, triple-buffered shared memory) while the
//             current one is classified from shared memory (stride 31 words per
//             thread = bank-conflict free)
//   grid    = SMs x resident CTAs; tile prefix (words emitted so far, length of the
//             fill run still open at the tile's end) by decoupled look-back over one
//             64-bit descriptor per tile, with the WHOLE CTA reading a 512-tile
//             window per round (lock-stepped persistent CTAs would otherwise walk a
//             dozen 32-tile windows, one L2 round trip each).  The loop is software
//             pipelined: tile k+1 is classified and published before the look-back
//             of tile k, so a look-back never waits for a neighbour's classification.
//
// A run is emitted where it ENDS ("tail"): group k is a tail if it is a literal,
// or the next group has another type, or it is the last group of the stream /
// block.  The k-th tail is the k-th output word; a fill's length is the distance
// to the previous tail, which flows forward through the scans.  Output words go
// straight to global memory: a warp's words are one contiguous span.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

// ---- tile descriptor: status:2 | no_tail:1 | open:30 | count:31 -------------------
constexpr uint32_t ST_EMPTY = 0, ST_AGG = 1, ST_INCL = 2;

__device__ __forceinline__ uint64_t desc_pack(uint32_t st, uint32_t no_tail, uint32_t open, uint32_t count)
{
    return ((uint64_t)st << 62) | ((uint64_t)no_tail << 61) | ((uint64_t)open << 31) | (uint64_t)count;
}
__device__ __forceinline__ uint32_t desc_status(uint64_t d) { return (uint32_t)(d >> 62); }
__device__ __forceinline__ uint32_t desc_no_tail(uint64_t d) { return (uint32_t)(d >> 61) & 1u; }
__device__ __forceinline__ uint32_t desc_open(uint64_t d) { return (uint32_t)(d >> 31) & MAX_FILL; }
__device__ __forceinline__ uint32_t desc_count(uint64_t d) { return (uint32_t)d & 0x7FFFFFFFu; }

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_addr, const void *gptr, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int THREADS>
struct TileGeom {
    static constexpr int NW = THREADS / 32;
    static constexpr int TILE_WORDS = THREADS * WORDS_PER_THREAD;
    static constexpr int TILE_GROUPS = THREADS * GROUPS_PER_THREAD;
    static constexpr int BUF_WORDS = TILE_WORDS + 4;   // + look-ahead word, padded to 16 B
    static constexpr int NVEC = TILE_WORDS / 4;
    static_assert(TILE_WORDS % 4 == 0, "tile must be a whole number of 16-byte vectors");
};

// Start the asynchronous copy of one tile (+1 look-ahead word) into a shared-memory buffer,
// zero filled past the end of the column.  Always commits exactly one cp.async group.
template <int THREADS>
__device__ __forceinline__ void stage_tile(const CompressParams &p, uint32_t tile, uint32_t *buf, uint32_t tid)
{
    using G = TileGeom<THREADS>;
    if (tile < p.n_tiles) {
        const uint32_t col = tile / p.tiles_per_col;
        const uint32_t t = tile - col * p.tiles_per_col;
        const uint64_t w0 = (uint64_t)t * G::TILE_WORDS;
        const uint32_t *src = p.in + (uint64_t)col * p.col_stride + w0;
        const uint64_t left = p.n_words - w0;   // words from the tile start to the column end
        const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(buf);
        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
        if (aligned && left > (uint64_t)G::TILE_WORDS) {
            // full tile followed by at least one more word: plain 16-byte copies
            const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
            for (int k = 0; k < (G::NVEC + THREADS - 1) / THREADS; k++) {
                const uint32_t i = tid + k * THREADS;
                if (i < G::NVEC) cp_async_16(s_base + 16u * i, src4 + i);
            }
            if (tid == 0) cp_async_16_zfill(s_base + 16u * G::NVEC, src + G::TILE_WORDS, 4u);
        } else if (aligned) {
            const uint32_t nload = left > (uint64_t)G::TILE_WORDS ? (uint32_t)G::TILE_WORDS + 1u : (uint32_t)left;
            for (uint32_t i = tid; i < G::BUF_WORDS / 4; i += THREADS) {
                const uint32_t b = 4u * i;
                const uint32_t bytes = b >= nload ? 0u : (nload - b >= 4u ? 16u : (nload - b) * 4u);
                cp_async_16_zfill(s_base + 16u * i, bytes ? (const void *)(src + b) : (const void *)src, bytes);
            }
        } else {
            // column start not 16-byte aligned: scalar staging
            const uint32_t nload = left > (uint64_t)G::TILE_WORDS ? (uint32_t)G::TILE_WORDS + 1u : (uint32_t)left;
            for (uint32_t i = tid; i < G::BUF_WORDS; i += THREADS) buf[i] = i < nload ? ld_stream_u32(src + i) : 0u;
        }
    }
    cp_async_commit();
}

// Everything emission needs from classification, kept in registers while the NEXT tile is
// classified (the loop is software pipelined, see the kernel).
struct TileState {
    uint32_t T, F, O;       // tail / fill / one-fill masks of my 32 groups
    uint32_t my_off;        // tile-relative index of my first output word
    uint32_t prev_open;     // length of the run open at my chunk's start (without the tile's carry-in)
    uint32_t wcnt;          // words my warp emits
    uint32_t flags;         // bit0: a lower lane of my warp has a tail, bit1: a lower warp has a tail
    uint32_t tile_cnt, tile_open, tile_has;
};

template <int THREADS, bool BLOCK_MODE>
__global__ void __launch_bounds__(THREADS, 2) wah_compress_kernel(const CompressParams p)
{
    using G = TileGeom<THREADS>;
    constexpr int NW = G::NW;
    constexpr int LB = 2;   // descriptors each thread reads per look-back round (window = LB * THREADS tiles)

    extern __shared__ __align__(16) uint32_t smem[];   // three input buffers of BUF_WORDS

    __shared__ uint32_t s_wcnt[NW];      // words emitted by each warp
    __shared__ uint32_t s_wopen[NW];     // groups after the warp's last tail (1024 if it has none)
    __shared__ uint32_t s_whas[NW];      // warp has at least one tail
    __shared__ uint32_t s_lb_cnt[LB * NW];    // look-back partials of each warp's 32-tile windows
    __shared__ uint32_t s_lb_open[LB * NW];
    __shared__ uint32_t s_lb_flags[LB * NW];  // bit0: window holds an INCLUSIVE descriptor, bit1: open run resolved
    __shared__ uint32_t s_drop;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t stride = gridDim.x;
    if (tid == 0) s_drop = 0;
    const uint64_t base = p.base_in ? *p.base_in : 0ull;

    // classify tile `tile` from buffer s_in, publish its aggregate, return the per-thread state
    auto classify = [&](uint32_t tile, const uint32_t *s_in) -> TileState {
        const uint32_t col = tile / p.tiles_per_col;
        const uint32_t t = tile - col * p.tiles_per_col;   // tile index inside the column

        // ---- my 32 groups (kernels.cu:79 regroup, :93-112 classification)
        const uint32_t *row = s_in + WORDS_PER_THREAD * tid;
        const uint64_t g_thread = (uint64_t)t * G::TILE_GROUPS + 32u * tid;   // first group, column relative
        uint32_t nvalid = 0;
        if (g_thread < p.groups) nvalid = (p.groups - g_thread) >= 32u ? 32u : (uint32_t)(p.groups - g_thread);
        const uint32_t vmask = nvalid == 32u ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);

        uint32_t Z = 0, O = 0;
        {
            // u = group << 1 | (one junk bit): zero / all-ones tests on the top 31 bits
            uint32_t prev = row[0];
            uint32_t u = prev << 1;
            if ((u & 0xFFFFFFFEu) == 0u) Z |= 1u;
            if ((~u & 0xFFFFFFFEu) == 0u) O |= 1u;
#pragma unroll
            for (int j = 1; j < 31; j++) {
                const uint32_t cur_w = row[j];
                u = __funnelshift_r(prev, cur_w, 31 - j);
                if ((u & 0xFFFFFFFEu) == 0u) Z |= (1u << j);
                if ((~u & 0xFFFFFFFEu) == 0u) O |= (1u << j);
                prev = cur_w;
            }
            u = prev;
            if ((u & 0xFFFFFFFEu) == 0u) Z |= BIT31;
            if ((~u & 0xFFFFFFFEu) == 0u) O |= BIT31;
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
        uint32_t prev_open = below ? open_q + 32u * (lane - q - 1u) : 32u * lane;   // + carries if !below
        const uint32_t qlast = tb ? 31u - (uint32_t)__clz(tb) : 0u;
        const uint32_t open_last = __shfl_sync(0xffffffffu, my_open, qlast);
        const uint32_t wcnt = __shfl_sync(0xffffffffu, incl, 31);
        __syncthreads();   // the previous tile's warp aggregates have been consumed by everyone
        if (lane == 0) {
            s_wcnt[warp] = wcnt;
            s_whas[warp] = tb != 0u;
            s_wopen[warp] = tb ? open_last + 32u * (31u - qlast) : 1024u;
        }
        __syncthreads();

        // ---- tile aggregate (every thread; NW is small) and my warp's prefix
        TileState st;
        uint32_t tile_cnt = 0, tile_open = 0, tile_has = 0;
        uint32_t wprefix = 0, wcarry = 0;
        bool wfound = false;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t c = s_wcnt[w], h = s_whas[w], o = s_wopen[w];
            if (w < (int)warp) {
                wprefix += c;
                if (h) {
                    wcarry = o;
                    wfound = true;
                } else {
                    wcarry += o;
                }
            }
            tile_cnt += c;
            if (h) {
                tile_has = 1;
                tile_open = o;
            } else {
                tile_open += o;
            }
        }
        // the end of a column is always a tail (lanes past the end hold no groups)
        if (t == p.tiles_per_col - 1u) {
            tile_open = 0;
            tile_has = 1;
        }
        if (!BLOCK_MODE && !below) prev_open += wcarry;
        st.T = T;
        st.F = F;
        st.O = O;
        st.my_off = wprefix + incl - cnt;
        st.prev_open = prev_open;
        st.wcnt = wcnt;
        st.flags = (below ? 1u : 0u) | (wfound ? 2u : 0u);
        st.tile_cnt = tile_cnt;
        st.tile_open = tile_open;
        st.tile_has = tile_has;

        // publish: tile 0 has no predecessor (unless it still has to merge with an earlier launch)
        if (tid == 0) {
            if (tile == 0u) {
                if (!p.merge_prev) st_relaxed_u64(p.desc, desc_pack(ST_INCL, 0, tile_open, tile_cnt));
            } else {
                st_relaxed_u64(p.desc + tile, desc_pack(ST_AGG, tile_has ^ 1u, tile_open, tile_cnt));
            }
        }
        return st;
    };

    // ---- prologue: two tiles in flight, the first one classified and published
    uint32_t tile = blockIdx.x;
    stage_tile<THREADS>(p, tile, smem, tid);
    stage_tile<THREADS>(p, tile + stride, smem + G::BUF_WORDS, tid);
    TileState cur_st;
    if (tile < p.n_tiles) {
        cp_async_wait<1>();
        __syncthreads();
        cur_st = classify(tile, smem);
    }
    int cur = 0;   // buffer of `tile`

    // Software pipeline: tile k+1 is classified and its aggregate published BEFORE the look-back of
    // tile k.  CTAs of a round-robin grid run in step; without the skew every look-back would wait for
    // the slowest of its same-round predecessors to finish classifying.
    while (tile < p.n_tiles) {
        const uint32_t next = tile + stride;
        const int nxt = cur == 2 ? 0 : cur + 1, nn = nxt == 2 ? 0 : nxt + 1;
        const uint32_t *s_in = smem + cur * G::BUF_WORDS;
        // third buffer: last read by the emission of the tile before `tile`, which ended with barrier (C)
        stage_tile<THREADS>(p, next + stride, smem + nn * G::BUF_WORDS, tid);
        TileState next_st;
        if (next < p.n_tiles) {
            cp_async_wait<1>();
            __syncthreads();   // (A) `next` is staged
            next_st = classify(next, smem + nxt * G::BUF_WORDS);
        }

        const TileState st = cur_st;
        const uint32_t col = tile / p.tiles_per_col;
        const uint32_t t = tile - col * p.tiles_per_col;

        // ---- decoupled look-back, LB 32-tile windows per warp and round
        uint32_t excl = 0, carry = 0;
        if (tile != 0u) {
            // BLOCK mode never carries a run; CANONICAL restarts at every column
            bool open_done = BLOCK_MODE || (t == 0u);
            int64_t look = (int64_t)tile - 1 - (int64_t)tid;
            while (true) {
#pragma unroll
                for (int r = 0; r < LB; r++) {
                    const int64_t lk = look - (int64_t)r * THREADS;
                    uint64_t d;
                    if (lk >= 0) {
                        do {
                            d = ld_relaxed_u64(p.desc + lk);
                        } while (desc_status(d) == ST_EMPTY);
                    } else {
                        d = desc_pack(ST_INCL, 0, 0, 0);
                    }
                    const uint32_t incl_mask = __ballot_sync(0xffffffffu, desc_status(d) == ST_INCL);
                    const uint32_t first_incl = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 32u;
                    const bool part = lane <= first_incl;   // first_incl == 32: every lane takes part
                    const uint32_t csum = warp_sum(part ? desc_count(d) : 0u);
                    uint32_t osum = 0, flags = first_incl < 32u ? 1u : 0u;
                    if (!BLOCK_MODE) {
                        // towards older tiles until one ends a run (or carries a resolved value)
                        const uint32_t term_mask =
                            __ballot_sync(0xffffffffu, part && (desc_status(d) == ST_INCL || !desc_no_tail(d)));
                        const uint32_t first_term = term_mask ? (uint32_t)__ffs(term_mask) - 1u : 32u;
                        osum = warp_sum((part && lane <= first_term) ? desc_open(d) : 0u);
                        flags |= first_term < 32u ? 2u : 0u;
                    }
                    if (lane == 0) {
                        s_lb_cnt[r * NW + warp] = csum;
                        s_lb_open[r * NW + warp] = osum;
                        s_lb_flags[r * NW + warp] = flags;
                    }
                }
                __syncthreads();
                bool done = false;
#pragma unroll
                for (int w = 0; w < LB * NW; w++) {
                    if (!done) {
                        const uint32_t f = s_lb_flags[w];
                        excl += s_lb_cnt[w];
                        if (!open_done) {
                            carry += s_lb_open[w];
                            open_done = (f & 2u) != 0u;
                        }
                        done = (f & 1u) != 0u;
                    }
                }
                if (done) break;
                look -= LB * THREADS;
                __syncthreads();   // partials are rewritten in the next round
            }
            const uint32_t incl_open = st.tile_has ? st.tile_open : carry + st.tile_open;
            if (tid == 0) st_relaxed_u64(p.desc + tile, desc_pack(ST_INCL, 0, incl_open, excl + st.tile_cnt));
            if (BLOCK_MODE || t == 0u) carry = 0;
        }

        // ---- emit straight to global memory
        const uint32_t T = st.T, F = st.F, O = st.O;
        uint32_t prev_open = st.prev_open;
        if (!BLOCK_MODE && st.flags == 0u) prev_open += carry;   // no tail before me in this tile
        const uint32_t tile_cnt = st.tile_cnt;
        const uint64_t dst0 = base + excl;
        const uint32_t my_off = st.my_off;
        uint32_t *dst = p.out + dst0;
        // words of this tile the output buffer still has room for
        const uint32_t room = dst0 >= p.out_cap ? 0u
                              : (p.out_cap - dst0 >= (uint64_t)G::TILE_GROUPS ? (uint32_t)G::TILE_GROUPS
                                                                               : (uint32_t)(p.out_cap - dst0));
        const uint32_t *row = s_in + WORDS_PER_THREAD * tid;
        const uint32_t *wrow = s_in + 992u * warp;

        const bool all_literal = __all_sync(0xffffffffu, T == 0xFFFFFFFFu && F == 0u);
        if (all_literal) {
            // every group of the warp is a literal: lane-per-output-word, fully coalesced
            const uint32_t wprefix = __shfl_sync(0xffffffffu, my_off, 0);
#pragma unroll 8
            for (uint32_t k = 0; k < 32u; k++) {
                const uint32_t g = 32u * k + lane;
                const uint32_t pos = wprefix + g;
                if (pos < room) st_stream_u32(dst + pos, extract_group(wrow, g));
            }
        } else if (st.wcnt > 192u) {
            // dense warp: lanes cooperate on one thread-chunk at a time (contiguous stores)
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
                    const uint32_t pos = offk + __popc(lower);
                    if (pos < room) st_stream_u32(dst + pos, w);
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
                if (off < room) st_stream_u32(dst + off, w);
                off++;
                prev = j;
                extra = 0;
            }
        }

        // ---- seam with an earlier launch (CANONICAL append): fold my first run into its last word
        if (tile == 0u && p.merge_prev) {
            __syncthreads();   // the tile's words are in global memory, visible to the whole CTA
            if (tid == 0) {
                uint32_t drop = 0;
                if (base > 0 && tile_cnt > 0 && base < p.out_cap) {
                    const uint32_t pw = p.out[base - 1], fw = p.out[base];
                    if (is_fill(pw) && is_fill(fw) && ((pw ^ fw) & BIT30) == 0u) {
                        const uint64_t total = (uint64_t)fill_count(pw) + fill_count(fw);
                        const uint32_t ty = (fw >> 30) & 1u;
                        if (total <= MAX_FILL) {
                            p.out[base - 1] = fill_word(ty, (uint32_t)total);
                            drop = 1;
                        } else {
                            p.out[base - 1] = fill_word(ty, MAX_FILL);
                            p.out[base] = fill_word(ty, (uint32_t)(total - MAX_FILL));
                        }
                    }
                }
                s_drop = drop;
            }
            __syncthreads();
            if (s_drop) {
                // close the one-word gap: slide the tile's remaining words down (tile 0 of a continuation only)
                const uint32_t nmove = (tile_cnt < room ? tile_cnt : room);
                for (uint32_t i0 = 1; i0 < nmove; i0 += THREADS) {
                    const uint32_t i = i0 + tid;
                    uint32_t v = 0;
                    if (i < nmove) v = p.out[base + i];
                    __syncthreads();
                    if (i < nmove) p.out[base + i - 1] = v;
                    __syncthreads();
                }
            }
            if (tid == 0) st_relaxed_u64(p.desc, desc_pack(ST_INCL, 0, st.tile_open, tile_cnt - s_drop));
        }
        if (tid == 0) {
            const uint32_t drop = (tile == 0u) ? s_drop : 0u;
            if (p.col_offsets && t == 0u) p.col_offsets[col] = dst0;
            if (tile == p.n_tiles - 1u) {
                const uint64_t total = dst0 + tile_cnt - drop;
                *p.total_out = total;
                if (p.col_offsets) p.col_offsets[p.n_cols] = total;
            }
        }

        __syncthreads();   // (C) everyone is done with s_in and the look-back partials
        tile = next;
        cur = nxt;
        cur_st = next_st;
    }
    cp_async_wait<0>();
}

}  // namespace

size_t compress_smem_bytes()
{
    return (size_t)(3 * (COMPRESS_TILE_WORDS + 4)) * sizeof(uint32_t);
}

template <typename K>
static cudaError_t compress_grid(K kernel, size_t smem, int *grid)
{
    // all CTAs must be resident at once: the look-back spins on tiles owned by other CTAs
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, COMPRESS_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_compress(const CompressParams &p, int mode, cudaStream_t stream)
{
    const size_t smem = compress_smem_bytes();
    static int grids[2] = {0, 0};
    const int m = mode == 0 ? 0 : 1;
    const void *kernel = m == 0 ? (const void *)wah_compress_kernel<COMPRESS_THREADS, true>
                                : (const void *)wah_compress_kernel<COMPRESS_THREADS, false>;
    if (grids[m] == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int g = 0;
        e = m == 0 ? compress_grid(wah_compress_kernel<COMPRESS_THREADS, true>, smem, &g)
                   : compress_grid(wah_compress_kernel<COMPRESS_THREADS, false>, smem, &g);
        if (e != cudaSuccess) return e;
        grids[m] = g;
    }
    int grid = grids[m];
    if ((uint32_t)grid > p.n_tiles) grid = (int)p.n_tiles;
    CompressParams params = p;
    void *args[] = {&params};
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(COMPRESS_THREADS), args, smem, stream);
}

}  // namespace wahb200

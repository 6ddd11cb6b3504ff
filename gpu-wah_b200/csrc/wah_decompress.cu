// wah_decompress.cu -- WAH decompressor for sm_100a.
//
// Replaces the reference's getCounts + thrust::exclusive_scan + decompressWords +
// mergeWords (kernels.cu:291-385, decompress.cu:66-115), which materialise an
// 8-byte count per compressed word and a one-group-per-int intermediate array in
// HBM and expand every fill with a serial per-thread loop (kernels.cu:346-348).
//
// Here:
//   scan kernel   : one pass over the compressed words; per-tile group sums with a
//                   decoupled look-back (CTA-wide window) give every tile its group
//                   offset, and each tile records, for every OUTPUT tile boundary
//                   (multiples of 8192 groups) that falls into it, which compressed
//                   word covers it.
//   expand kernel : output-centric, hence load balanced whatever the fill lengths
//                   are.  A persistent grid walks output tiles of 8192 groups = 7936
//                   words.  Per tile the one-group-per-int array of the reference
//                   lives in SHARED memory: compressed words are read coalesced,
//                   scanned in registers and scattered (literal = one store, zero
//                   fill = nothing, one fill = a run of stores); then one thread
//                   repacks 32 groups into 31 words with compile-time funnel shifts
//                   (mergeWords, kernels.cu:375) and the tile leaves as 128-bit lines.
// Output tiles are aligned in group space to multiples of 32 groups = 31 words, so no
// output word is shared between threads or tiles and nothing needs atomics.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

constexpr uint64_t ST_EMPTY = 0, ST_AGG = 1, ST_INCL = 2;
constexpr uint64_t VALUE_MASK = (1ull << 62) - 1ull;
constexpr uint32_t TG_SHIFT = 13;
static_assert((1u << TG_SHIFT) == (uint32_t)EXPAND_TILE_GROUPS, "output tile must be 8192 groups");

// ------------------------------------------------------------------ scan kernel

// Group counts of one tile of compressed words, kept in registers while the NEXT tile is summed and
// published (software pipeline, as in the compressor).
struct ScanState {
    uint32_t cnt[SCAN_ITEMS];
    uint64_t first_off;   // group offset of my first word relative to the tile start
    uint64_t tile_sum;
};

__device__ __forceinline__ void scan_body(const ScanParams &p)
{
    constexpr int NW = SCAN_THREADS / 32;
    constexpr uint64_t TGM = (uint64_t)EXPAND_TILE_GROUPS - 1ull;
    __shared__ uint64_t s_wsum[NW];
    __shared__ uint64_t s_lb_sum[NW];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t stride = gridDim.x;

    // blocked arrangement: thread owns SCAN_ITEMS consecutive compressed words
    auto load = [&](uint32_t tile, uint32_t (&w)[SCAN_ITEMS]) {
        const uint64_t w_begin = (uint64_t)tile * SCAN_TILE_WORDS + (uint64_t)tid * SCAN_ITEMS;
        if (tile < p.n_tiles && w_begin + SCAN_ITEMS <= p.c_words) {
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
                w[i] = (tile < p.n_tiles && w_begin + i < p.c_words) ? ld_stream_u32(p.in + w_begin + i)
                                                                      : BIT31;   // fill of 0 groups
        }
    };

    // counts, block scan, publish the tile's aggregate
    auto summarize = [&](uint32_t tile, const uint32_t (&w)[SCAN_ITEMS]) -> ScanState {
        ScanState st;
        const uint64_t w_begin = (uint64_t)tile * SCAN_TILE_WORDS + (uint64_t)tid * SCAN_ITEMS;
        uint64_t tsum = 0;
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            st.cnt[i] = word_groups(w[i]);   // getCounts, kernels.cu:298-304
            tsum += st.cnt[i];
            bad += (st.cnt[i] == 0u && w_begin + i < p.c_words) ? 1u : 0u;
        }
        if (__any_sync(0xffffffffu, bad != 0u)) {
            bad = warp_sum(bad);
            if (lane == 0) atomicAdd(&p.hdr->bad_words, bad);
        }
        const uint64_t incl = warp_incl_scan_u64(tsum);
        __syncthreads();   // the previous tile's warp sums have been consumed
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint64_t tile_sum = 0, wprefix = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const uint64_t s = s_wsum[k];
            if (k < (int)warp) wprefix += s;
            tile_sum += s;
        }
        st.first_off = wprefix + incl - tsum;
        st.tile_sum = tile_sum;
        if (tid == 0) st_relaxed_u64(p.desc + tile, (ST_AGG << 62) | tile_sum);
        return st;
    };

    uint32_t tile = blockIdx.x;
    if (tile >= p.n_tiles) return;   // (uniform) more CTAs than scan tiles
    uint32_t raw[SCAN_ITEMS];
    load(tile, raw);
    ScanState cur = summarize(tile, raw);
    load(tile + stride, raw);
    uint64_t own_incl = 0;    // groups up to and including my previous tile
    bool first_tile = true;

    while (tile < p.n_tiles) {
        const uint32_t next = tile + stride;
        ScanState nxt;
        if (next < p.n_tiles) {
            nxt = summarize(next, raw);   // published before this tile's look-back (see wah_compress.cu)
            load(next + stride, raw);
        }

        // ---- group offset of the tile by a chained sum (see wah_compress.cu): the groups before my previous
        //      tile, that tile's, and the aggregates of the tiles in between -- all published by other CTAs
        //      as soon as they have counted the tile, never waiting for anybody's offset
        uint64_t excl;
        {
            const int64_t lo = first_tile ? 0 : (int64_t)tile - (int64_t)stride + 1;
            uint64_t acc = 0;
            for (int64_t lk = (int64_t)tile - 1 - (int64_t)tid; lk >= lo; lk -= SCAN_THREADS) {
                uint64_t d;
                do {
                    d = ld_relaxed_u64(p.desc + lk);
                } while ((d >> 62) == ST_EMPTY);
                acc += d & VALUE_MASK;
            }
            acc = warp_sum_u64(acc);
            if (lane == 0) s_lb_sum[warp] = acc;
            __syncthreads();
            uint64_t total = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) total += s_lb_sum[k];
            excl = own_incl + total;
            own_incl = excl + cur.tile_sum;
            first_tile = false;
        }
        if (tid == 0 && tile == p.n_tiles - 1u) {
            // decompress.cu:82-93: G = last offset + last count, realSize = ceil(31 G / 32)
            const uint64_t G = excl + cur.tile_sum;
            const uint64_t words = (G >> 5) * 31ull + (((G & 31ull) * 31ull + 31ull) >> 5);
            p.hdr->groups = G;
            p.hdr->words = words;
            p.hdr->out_tiles = (G + TGM) >> TG_SHIFT;
            if (p.out_info) {
                p.out_info[0] = words;
                p.out_info[1] = G;
            }
        }

        // ---- which compressed word covers each output-tile boundary k * 8192 ?
        if (p.starts != nullptr) {
            const uint64_t w_begin = (uint64_t)tile * SCAN_TILE_WORDS + (uint64_t)tid * SCAN_ITEMS;
            uint64_t off = excl + cur.first_off;   // group offset of my first word
            const uint64_t k_limit = p.max_out_tiles + 1ull;
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; i++) {
                // boundaries with off <= k * 8192 < off + cnt
                uint64_t k_first = (off + TGM) >> TG_SHIFT;
                uint64_t k_end = (off + cur.cnt[i] + TGM) >> TG_SHIFT;
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
                off += cur.cnt[i];
            }
        }
        __syncthreads();   // look-back partials are rewritten by the next tile; this tile's `starts` are written
        if (tid == 0) {
            __threadfence();
            atomicAdd(&p.hdr->scan_done, 1u);   // the expand phase starts when every scan tile has got here
        }
        tile = next;
        cur = nxt;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) wah_scan_kernel(const ScanParams p)
{
    scan_body(p);
}

// ---------------------------------------------------------------- expand kernel

constexpr int EXP_CHUNK = EXPAND_THREADS * 8;          // compressed words scanned per round
constexpr int GRP_WORDS = EXPAND_TILE_GROUPS + EXPAND_TILE_GROUPS / 32;   // rows of 32 groups padded to 33
constexpr int EXP_LIST = 512;                          // long one-fills deferred to a warp-wide store loop
constexpr uint32_t EXP_CLAMP = 2u * EXPAND_TILE_GROUPS;   // any count >= the tile span behaves the same

__device__ __forceinline__ uint32_t grp_pos(uint32_t g) { return g + (g >> 5); }

// ---- bulk (TMA) store of a finished tile, shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int SPARSE_MAX_WORDS = 4096;   // output tiles covered by at most this many compressed words take the bit-scatter path
constexpr int SP_ROUND = EXPAND_THREADS * 2;   // compressed words per bit-scatter round (two per thread)
constexpr int SP_LIST = 256;             // long one-fills of a round, finished warp-wide

// OR the stream bits [b0, b1) (tile relative) into the tile image; plain read-modify-write: the caller
// guarantees that no other thread touches the same words at the same time
__device__ __forceinline__ void set_bits(uint32_t *img, uint32_t b0, uint32_t b1)
{
    const uint32_t w0 = b0 >> 5, w1 = (b1 - 1u) >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u), m1 = 0xFFFFFFFFu >> (31u - ((b1 - 1u) & 31u));
    if (w0 == w1) {
        img[w0] |= m0 & m1;
    } else {
        img[w0] |= m0;
        for (uint32_t w = w0 + 1u; w < w1; w++) img[w] = 0xFFFFFFFFu;
        img[w1] |= m1;
    }
}

__device__ __forceinline__ void load8(const ExpandParams &p, uint64_t i0, uint32_t (&w)[8])
{
    if (i0 + 8 <= p.c_words) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.in + i0);
        const uint4 a = ld_stream_v4(src), b = ld_stream_v4(src + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = (i0 + i < p.c_words) ? ld_stream_u32(p.in + i0 + i) : BIT31;
    }
}

__device__ __forceinline__ void expand_body(const ExpandParams &p)
{
    constexpr int NW = EXPAND_THREADS / 32;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_grp = smem;                  // GRP_WORDS: one group per int, rows of 32 padded to 33
    uint32_t *s_stage = smem + GRP_WORDS;    // EXPAND_TILE_WORDS output words
    __shared__ uint32_t s_wsum[NW];
    __shared__ uint2 s_list[EXP_LIST];
    __shared__ uint32_t s_nlist;
    uint32_t n_sparse = 0;   // bit-scatter tiles so far: they alternate between the two tile images
    bool have_w = false;     // w[] holds the first words of the current tile (prefetched)

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t G = p.hdr->groups;
    uint64_t total_words = p.hdr->words;
    if (total_words > p.out_cap) total_words = p.out_cap;
    const uint64_t real_tiles = p.hdr->out_tiles;
    uint64_t n_tiles = real_tiles < p.max_out_tiles ? real_tiles : p.max_out_tiles;
    {
        const uint64_t need = (total_words + EXPAND_TILE_WORDS - 1) / EXPAND_TILE_WORDS;   // tiles with room
        if (n_tiles > need) n_tiles = need;
    }

    // where a tile's compressed words start / end (written by the scan kernel)
    auto tile_info = [&](uint64_t ot, uint64_t &ws, uint64_t &we, uint32_t &skip, uint32_t &first) {
        if (ot < n_tiles) {
            const ulonglong2 st = p.starts[ot];
            ws = st.x;
            we = (ot + 1 < real_tiles) ? p.starts[ot + 1].x : p.c_words - 1;
            skip = (uint32_t)((ot << TG_SHIFT) - st.y);   // groups of word ws that belong to earlier tiles
            first = (ws == we) ? p.in[ws] : 0u;           // a tile inside ONE word is written without decoding
        } else {
            ws = we = 0;
            skip = first = 0;
        }
    };

    uint64_t ot = blockIdx.x;
    uint64_t ws, we, ws_n, we_n;
    uint32_t skip, skip_n, first, first_n;
    uint32_t w[8];
    tile_info(ot, ws, we, skip, first);
    for (; ot < n_tiles; ot += gridDim.x) {
        // the next tile's bookkeeping is fetched now and its first words after the scatter below, so
        // neither load latency sits on the next iteration's critical path
        tile_info(ot + gridDim.x, ws_n, we_n, skip_n, first_n);

        const uint64_t g_lo = ot << TG_SHIFT;
        const uint64_t w_lo = ot * (uint64_t)EXPAND_TILE_WORDS;
        const uint64_t avail = total_words - w_lo;
        const uint32_t nout = avail < (uint64_t)EXPAND_TILE_WORDS ? (uint32_t)avail : (uint32_t)EXPAND_TILE_WORDS;
        uint32_t *dst = p.out + w_lo;
        uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
        const uint32_t nvec = nout >> 2;
        const uint64_t wa = ws & ~3ull;   // 16-byte aligned start; words before ws are ignored

        bool fast = false;
        if (ws == we && is_fill(first) && (ot + 1 < real_tiles || !(first & BIT30))) {
            // the whole tile lies inside one fill word (the stream's last tile may end in a partly
            // padded word: a one-fill there takes the general path)
            fast = true;
            const uint32_t f = (first & BIT30) ? 0xFFFFFFFFu : 0u;
            const uint4 v = make_uint4(f, f, f, f);
            for (uint32_t i = tid; i < nvec; i += EXPAND_THREADS) st_stream_v4(dst4 + i, v);
            for (uint32_t i = (nvec << 2) + tid; i < nout; i += EXPAND_THREADS) dst[i] = f;
        }
        if (fast) {
            ws = ws_n;
            we = we_n;
            skip = skip_n;
            first = first_n;
            have_w = false;
            continue;
        }

        const uint32_t nw_all = (uint32_t)(we - wa + 1);   // words wa .. we
        if (nw_all <= (uint32_t)SPARSE_MAX_WORDS && nout == (uint32_t)EXPAND_TILE_WORDS) {
            // ================= bit-scatter path (fill dominated data) =================
            // The tile image (7936 words) is cleared in shared memory, every literal ORs its 31 bits into
            // the one or two words it touches, every one-fill sets its bit range, zero fills cost nothing;
            // the finished image leaves through one bulk (TMA) store while the next tile is assembled in
            // the other image.  Words 2k and 2k+1 of the compressed stream are handled in separate phases:
            // two words handled at the same time are then at least two groups (62 bits) apart and never
            // touch the same 32-bit word, so plain read-modify-writes suffice.
            uint32_t *img = (n_sparse & 1u) ? s_grp : s_stage;
            n_sparse++;
            if (tid == 0) bulk_wait_read<1>();   // the store that last read this image is done
            __syncthreads();
            {
                uint4 *z = reinterpret_cast<uint4 *>(img);
                for (uint32_t i = tid; i < (uint32_t)EXPAND_TILE_WORDS / 4; i += EXPAND_THREADS) z[i] = make_uint4(0, 0, 0, 0);
            }
            int32_t running = 0;   // group offset (tile relative) of the round's first word
            for (uint32_t c0 = 0; c0 < nw_all; c0 += SP_ROUND) {
                const uint64_t i0 = wa + c0 + 2ull * tid;   // my two consecutive words
                uint32_t x[2];
                if (i0 + 2 <= p.c_words) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(p.in + i0);
                    x[0] = v.x;
                    x[1] = v.y;
                } else {
                    x[0] = i0 < p.c_words ? p.in[i0] : BIT31;
                    x[1] = BIT31;
                }
                uint32_t c[2];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint64_t gi = i0 + i;
                    uint32_t v = word_groups(x[i]);
                    if (gi < ws || gi > we) v = 0;   // outside this tile's word range
                    else if (gi == ws) v -= skip;    // part of the first word belongs to earlier tiles
                    c[i] = v > EXP_CLAMP ? EXP_CLAMP : v;
                }
                const uint32_t tsum = c[0] + c[1];
                const uint32_t incl = warp_incl_scan(tsum);
                __syncthreads();   // previous round's s_wsum / list consumed; first round: image cleared
                if (lane == 31) s_wsum[warp] = incl;
                if (tid == 0) s_nlist = 0;
                __syncthreads();
                int32_t off = running + (int32_t)(incl - tsum);
                uint32_t round_sum = 0;
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint32_t sv = s_wsum[k];
                    if (k < (int)warp) off += (int32_t)sv;
                    round_sum += sv;
                }
                running += (int32_t)round_sum;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (c[i] != 0u && off < EXPAND_TILE_GROUPS) {
                        const uint32_t wv = x[i];
                        const uint32_t g0 = (uint32_t)off;
                        if (!is_fill(wv)) {   // kernels.cu:351-354, packed at once (kernels.cu:375)
                            const uint32_t bit = 31u * g0, w = bit >> 5, sh = bit & 31u;
                            img[w] |= wv << sh;
                            if (sh > 1u) img[w + 1] |= wv >> (32u - sh);
                        } else if (wv & BIT30) {   // one-fill, kernels.cu:337-348
                            uint32_t g1 = g0 + c[i];
                            if (g1 > (uint32_t)EXPAND_TILE_GROUPS) g1 = EXPAND_TILE_GROUPS;
                            const uint32_t b0 = 31u * g0, b1 = 31u * g1;
                            if (b1 - b0 <= 32u * 48u) {
                                set_bits(img, b0, b1);
                            } else {
                                // long run: its two ragged ends now, the whole words in between warp-wide below
                                const uint32_t wlo = (b0 + 31u) >> 5, whi = b1 >> 5;
                                if (b0 & 31u) set_bits(img, b0, wlo << 5);
                                if (b1 & 31u) set_bits(img, whi << 5, b1);
                                const uint32_t e = atomicAdd(&s_nlist, 1u);
                                if (e < (uint32_t)SP_LIST) s_list[e] = make_uint2(wlo, whi);
                                else for (uint32_t w = wlo; w < whi; w++) img[w] = 0xFFFFFFFFu;
                            }
                        }
                    }
                    off += (int32_t)c[i];
                    if (i == 0) __syncthreads();   // even words done before odd words start
                }
                __syncthreads();
                const uint32_t nl = s_nlist < (uint32_t)SP_LIST ? s_nlist : (uint32_t)SP_LIST;
                for (uint32_t e = warp; e < nl; e += NW) {
                    const uint2 r = s_list[e];
                    for (uint32_t w = r.x + lane; w < r.y; w += 32) img[w] = 0xFFFFFFFFu;
                }
                if (running >= EXPAND_TILE_GROUPS) break;   // uniform: the tile is covered
            }
            // next tile's bookkeeping
            ws = ws_n;
            we = we_n;
            skip = skip_n;
            first = first_n;
            have_w = false;
            fence_async_smem();   // my writes to the image, visible to the bulk copy engine
            __syncthreads();
            if (tid == 0) bulk_s2g(dst, (uint32_t)__cvta_generic_to_shared(img), EXPAND_TILE_WORDS * 4u);
            continue;
        }
        // the general path below uses both images as scratch: no bulk store may still be reading them
        if (n_sparse) {
            if (tid == 0) bulk_wait_read<0>();
            __syncthreads();
        }

        // ---- 1. clear the group array
        {
            uint4 *z = reinterpret_cast<uint4 *>(s_grp);
            for (uint32_t i = tid; i < GRP_WORDS / 4; i += EXPAND_THREADS) z[i] = make_uint4(0, 0, 0, 0);
            if (tid == 0) s_nlist = 0;
        }
        __syncthreads();   // also: the previous tile's staging area has been written out

        // ---- 2. scan the compressed words of the tile in rounds of EXP_CHUNK and scatter them
        const uint32_t nw = (uint32_t)(we - wa + 1);    // words wa .. we
        int32_t running = 0;                            // group offset (tile relative) of the round's first word
        for (uint32_t c0 = 0; c0 < nw; c0 += EXP_CHUNK) {
            const uint64_t i0 = wa + c0 + 8ull * tid;   // my 8 consecutive words
            if (c0 != 0 || !have_w) load8(p, i0, w);
            uint32_t c[8];
            uint32_t tsum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint64_t gi = i0 + i;
                uint32_t x = word_groups(w[i]);
                if (gi < ws || gi > we) x = 0;           // outside this tile's word range
                else if (gi == ws) x -= skip;            // part of the first word belongs to earlier tiles
                c[i] = x > EXP_CLAMP ? EXP_CLAMP : x;
                tsum += c[i];
            }
            const uint32_t incl = warp_incl_scan(tsum);
            if (c0 != 0) __syncthreads();   // previous round's s_wsum consumed
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int32_t off = running + (int32_t)(incl - tsum);
            uint32_t round_sum = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const uint32_t sv = s_wsum[k];
                if (k < (int)warp) off += (int32_t)sv;
                round_sum += sv;
            }
            running += (int32_t)round_sum;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (c[i] != 0u && off < EXPAND_TILE_GROUPS) {
                    const uint32_t wv = w[i];
                    if (!is_fill(wv)) {
                        s_grp[grp_pos((uint32_t)off)] = wv;                      // kernels.cu:351-354
                    } else if (wv & BIT30) {                                      // one-fill, kernels.cu:337-348
                        const uint32_t lo = (uint32_t)off;
                        uint32_t hi = lo + c[i];
                        if (hi > (uint32_t)EXPAND_TILE_GROUPS) hi = EXPAND_TILE_GROUPS;
                        if (hi - lo <= 8u) {
                            for (uint32_t g = lo; g < hi; g++) s_grp[grp_pos(g)] = ONES31;
                        } else {
                            const uint32_t e = atomicAdd(&s_nlist, 1u);
                            if (e < EXP_LIST) s_list[e] = make_uint2(lo, hi);
                            else for (uint32_t g = lo; g < hi; g++) s_grp[grp_pos(g)] = ONES31;
                        }
                    }
                }
                off += (int32_t)c[i];
            }
            if (running >= EXPAND_TILE_GROUPS) break;   // uniform: the tile is covered
        }
        // start fetching the next tile's first words; they are consumed one iteration from now
        ws = ws_n;
        we = we_n;
        skip = skip_n;
        first = first_n;
        have_w = ot + gridDim.x < n_tiles;
        if (have_w) load8(p, (ws & ~3ull) + 8ull * tid, w);
        __syncthreads();

        // ---- 3. long one-fills: a warp per run
        {
            const uint32_t nl = s_nlist < (uint32_t)EXP_LIST ? s_nlist : (uint32_t)EXP_LIST;
            for (uint32_t e = warp; e < nl; e += NW) {
                const uint2 r = s_list[e];
                for (uint32_t g = r.x + lane; g < r.y; g += 32) s_grp[grp_pos(g)] = ONES31;
            }
            if (nl) __syncthreads();
        }

        // ---- 4. my 32 groups -> 31 output words (mergeWords, kernels.cu:375:
        //         word j = group[j] >> j | group[j+1] << (31-j)); rows padded to 33 = conflict free
        if (g_lo + 32ull * tid < G) {
            const uint32_t *r = s_grp + 33u * tid;
            uint32_t *o = s_stage + 31u * tid;
            uint32_t a = r[0];
#pragma unroll
            for (int j = 0; j < 31; j++) {
                const uint32_t b = r[j + 1];
                o[j] = __funnelshift_r(a << 1, b, j + 1);
                a = b;
            }
        }
        __syncthreads();

        // ---- 5. coalesced 128-bit write of the tile
        {
            const uint4 *src4 = reinterpret_cast<const uint4 *>(s_stage);
            for (uint32_t i = tid; i < nvec; i += EXPAND_THREADS) st_stream_v4(dst4 + i, src4[i]);
            for (uint32_t i = (nvec << 2) + tid; i < nout; i += EXPAND_THREADS) dst[i] = s_stage[i];
        }
        // no barrier here: the next tile's clear touches s_grp only, and its first barrier orders this
        // tile's staging reads before the next repack writes
    }
    if (tid == 0) bulk_wait_read<0>();   // shared memory must outlive the bulk stores that read it
}

__global__ void __launch_bounds__(EXPAND_THREADS, 3) wah_expand_kernel(const ExpandParams p)
{
    expand_body(p);
}

// Both phases in one persistent launch: every CTA first takes its share of the scan tiles, then -- once all
// scan tiles are done, i.e. the decoded size and every output tile's starting point are known -- its share of
// the output tiles.  Saves a launch and the idle tail / ramp between two kernels.
static_assert(SCAN_THREADS == EXPAND_THREADS, "the fused kernel runs both phases with one CTA shape");
__global__ void __launch_bounds__(EXPAND_THREADS, 3) wah_decode_kernel(const ScanParams sp, const ExpandParams ep)
{
    scan_body(sp);
    if (threadIdx.x == 0) {
        volatile uint32_t *done = &sp.hdr->scan_done;
        while (*done < sp.n_tiles) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
    expand_body(ep);
}

}  // namespace

size_t expand_smem_bytes()
{
    return (size_t)(GRP_WORDS + EXPAND_TILE_WORDS) * sizeof(uint32_t);
}

cudaError_t launch_scan(const ScanParams &p, cudaStream_t stream)
{
    // persistent + cooperative: the look-back spins on tiles owned by other CTAs, all must be resident
    static int max_grid = 0;
    if (max_grid == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wah_scan_kernel, SCAN_THREADS, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (per_sm > 4) per_sm = 4;
        max_grid = sms * per_sm;
    }
    int grid = max_grid;
    if ((uint32_t)grid > p.n_tiles) grid = (int)p.n_tiles;
    ScanParams params = p;
    void *args[] = {&params};
    return cudaLaunchCooperativeKernel((const void *)wah_scan_kernel, dim3(grid), dim3(SCAN_THREADS), args, 0, stream);
}

cudaError_t launch_decode(const ScanParams &sp, const ExpandParams &ep, cudaStream_t stream)
{
    // persistent, every CTA resident (both phases spin on results of other CTAs): SMs x occupancy CTAs
    const size_t smem = expand_smem_bytes();
    static int grid = 0;
    if (grid == 0) {
        cudaError_t e = cudaFuncSetAttribute(wah_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wah_decode_kernel, EXPAND_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        grid = sms * per_sm;
    }
    ScanParams a = sp;
    ExpandParams b = ep;
    void *args[] = {&a, &b};
    return cudaLaunchKernel((const void *)wah_decode_kernel, dim3(grid), dim3(EXPAND_THREADS), args, smem, stream);
}

cudaError_t launch_expand(const ExpandParams &p, int grid, cudaStream_t stream)
{
    const size_t smem = expand_smem_bytes();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wah_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    wah_expand_kernel<<<grid, EXPAND_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace wahb200

// wah_decompress.cu -- WAH decompressor for sm_100a.
//
// Replaces the reference's getCounts + thrust::exclusive_scan + decompressWords +
// mergeWords (kernels.cu:291-385, decompress.cu:66-115), which materialise an
// 8-byte count per compressed word and a one-group-per-int intermediate array in
// HBM and expand every fill with a serial per-thread loop (kernels.cu:346-348).
//
// Here: ONE persistent launch (wah_decode_kernel, 3 CTAs per SM), two phases per CTA:
//   scan phase    one pass over the compressed words, one tile per CTA: per-tile group sums, exchanged through a
//                 round aggregator, give every tile its group offset; each tile then records, for every OUTPUT
//                 tile boundary that falls into it, which compressed word covers it.
//   expand phase  output-centric, hence load balanced whatever the fill lengths are.  The grid walks output tiles
//                 of 8192 groups = 7936 words; a tile waits only for its own two boundary entries.  A tile is
//                 assembled as a bit image in SHARED memory -- a literal ORs its 31 bits in, a one-fill ORs its ends in
//                 and marks its whole 16-byte units in a coverage map, which a prefix XOR turns into one 128-bit
//                 store of ones per thread and eight units; zero fills cost nothing -- and leaves through a TMA bulk
//                 store.  All-literal tiles are repacked in registers with one shuffle per word;
//                 tiles with more than 4096 words go through the reference's one-group-per-int array
//                 (kernels.cu:321-359), kept in shared memory, and the 32 -> 31 repack of mergeWords (kernels.cu:375).
// Output tiles are aligned in group space to multiples of 32 groups = 31 words, so no output word is shared
// between threads or tiles and nothing in HBM needs atomics.  wah_scan_kernel is the scan phase alone (size query).
//
// A bitmap-index batch (n_cols streams back to back, each decoding to the same number of groups) is ONE launch as
// well: group offsets run over the concatenation, output tile k of column j starts at group j * col_groups + k * 8192
// and lands at out + j * col_stride + k * 7936.
#include "wah_common.cuh"
#include "wah_kernels.h"

namespace wahb200 {

namespace {

// Everything the CTAs exchange through the workspace is tagged with the launch's epoch (a number the host never
// repeats), so the workspace is not cleared between calls: what an earlier launch left behind reads as "not
// published yet".  A tile sum / tile offset is a 16-byte cell {value, epoch}, written and read as one access.
__device__ __forceinline__ void cell_store(ulonglong2 *c, uint64_t v, uint32_t epoch)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(c), "l"(v), "l"((uint64_t)epoch) : "memory");
}
__device__ __forceinline__ bool cell_load(const ulonglong2 *c, uint32_t epoch, uint64_t &v)
{
    uint64_t e;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(e) : "l"(c) : "memory");
    return e == (uint64_t)epoch;
}
constexpr uint64_t ENTRY_MASK = (1ull << 48) - 1ull;   // output-tile table: low 48 bits value, high 16 bits half of the epoch
constexpr uint32_t TG_SHIFT = 13;
static_assert((1u << TG_SHIFT) == (uint32_t)EXPAND_TILE_GROUPS, "output tile must be 8192 groups");

// A CTA that waits for another CTA of the grid polls politely and not for ever: after SPIN_LIMIT polls (about 2 s) it
// flags the launch as failed and carries on with whatever it has; the launch then ends with STATUS_TIMEOUT in
// d_out_info[2] instead of hanging the GPU.  Once one CTA has given up, the others notice within 1024 polls.
__device__ __forceinline__ bool spin_ok(uint32_t &left, DecodeHeader *hdr, uint32_t epoch)
{
    if (left != 0u) {
        left--;
        if ((left & 1023u) != 0u || *reinterpret_cast<volatile uint32_t *>(&hdr->error) != epoch) return true;
    }
    left = 0;
    *reinterpret_cast<volatile uint32_t *>(&hdr->error) = epoch;
    return false;
}

// ------------------------------------------------------------------ scan phase

// One scan tile = p.tile_words compressed words (a multiple of 1024, chosen by the host so that every stream is one
// tile per CTA; a tile longer than 8192 words is walked in sub-tiles).  A warp owns a contiguous eighth of a
// sub-tile, lane l fetching words 4l .. 4l+3 of every 128-word row into shared memory (cp.async):
//   pass 1  count the groups of my words (getCounts, kernels.cu:298-304), publish the tile sum;
//   offset  the last CTA of a round to publish scans the round's sums and hands every tile its offset;
//   pass 2  row by row, a warp scan gives every 4-word pack its group offset (done while the offset is in flight
//           when the tile is one sub-tile); record, for every output-tile boundary that falls into a word, which
//           word that is and where it starts.
constexpr int SCAN_MAXV = 8;   // 128-bit packs per lane and sub-tile

// An entry of the output-tile table is read by other CTAs while the scan is still running: x (word index + 1,
// 0 = not recorded) and y (the word's group offset) must appear together -- one 16-byte store, one 16-byte load.
__device__ __forceinline__ void store_entry(ulonglong2 *e, uint64_t x, uint64_t y, uint32_t epoch)
{
    x |= (uint64_t)(epoch & 0xFFFFu) << 48;
    y |= (uint64_t)(epoch >> 16) << 48;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"(x), "l"(y) : "memory");
}
// x (word index + 1) and y (group offset) of an entry, or x = 0 if this launch has not written it yet
__device__ __forceinline__ void load_entry(const ulonglong2 *e, uint32_t epoch, uint64_t &x, uint64_t &y)
{
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(e) : "memory");
    const bool ok = (uint32_t)(x >> 48) == (epoch & 0xFFFFu) && (uint32_t)(y >> 48) == (epoch >> 16);
    x = ok ? (x & ENTRY_MASK) : 0ull;
    y &= ENTRY_MASK;
}
constexpr int EXPAND_STATIC_ROUNDS = 4;   // output tiles of the first rounds are dealt round robin, later ones by ticket
constexpr int SCAN_HEAVY = 64;  // long fills of a tile queued for the CTA-wide boundary writer

// Where the output tiles start, in the group numbering of the whole stream.  A single stream: tile k at group
// k * 8192, table index k (cg = ~0, j stays 0).  A batch of columns of cg groups each: tile k of column j at group
// j * cg + k * 8192, table index j * tpc + k.  `base` / `j` are the column that holds the group offset last asked for.
struct ColumnCursor {
    uint64_t base;
    uint32_t j;
};
__device__ __forceinline__ void col_seek(ColumnCursor &c, uint64_t off, uint64_t cg)
{
    if (off - c.base >= cg) {   // (never for a single stream)
        if (off - c.base >= (cg << 2)) {
            c.j = (uint32_t)(off / cg);
            c.base = (uint64_t)c.j * cg;
        } else {
            do {
                c.base += cg;
                c.j++;
            } while (off - c.base >= cg);
        }
    }
}

struct BoundaryGeom {
    uint64_t cg;      // groups per column, ~0 for a single stream
    uint64_t k_lim;   // table entries per column that may be written (single stream: one more, the end of the last tile)
    uint32_t tpc;     // table indices per column
    uint32_t n_cols;
};

// Four consecutive compressed words, the first at group offset `off`: record every output-tile boundary that falls
// into one of them (off <= boundary < off + cnt).  A word that covers up to 4 boundaries records them itself; a long
// fill is queued for the whole CTA.  Out of line: this runs for under 1 % of the words and would otherwise be
// replicated 8 times in straight-line code that executes once per launch.
__device__ __noinline__ void record_boundaries(ulonglong2 *starts, uint32_t epoch, const BoundaryGeom g, ColumnCursor cur,
                                               uint64_t wi, uint64_t off, uint4 cnt, ulonglong4 *s_heavy, uint32_t *s_nheavy)
{
    constexpr uint64_t TGM = (uint64_t)EXPAND_TILE_GROUPS - 1ull;
    const uint32_t c[4] = {cnt.x, cnt.y, cnt.z, cnt.w};
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        if (c[j] != 0u) {
            col_seek(cur, off, g.cg);
            const uint64_t rel = off - cur.base;
            uint64_t k_first = (rel + TGM) >> TG_SHIFT;
            uint64_t k_end = (rel + c[j] + TGM) >> TG_SHIFT;
            // (a batch whose stream holds more columns than declared: the first boundary behind the last declared
            //  column is still recorded -- it is where that column's last tile ends)
            const uint64_t lim = cur.j < g.n_cols ? g.k_lim : ((cur.j == g.n_cols && g.tpc != 0u) ? 1ull : 0ull);
            if (k_end > lim) k_end = lim;
            if (k_first < k_end) {
                const uint64_t i0 = (uint64_t)cur.j * g.tpc;
                if (k_end - k_first > 4ull) {
                    const uint32_t e = atomicAdd(s_nheavy, 1u);
                    if (e < (uint32_t)SCAN_HEAVY) {
                        s_heavy[e] = make_ulonglong4(wi + j, off, i0 + k_first, i0 + k_end);
                        k_end = k_first;   // queued
                    }
                }
#pragma unroll 1
                for (uint64_t k = k_first; k < k_end; k++) store_entry(starts + i0 + k, wi + j + 1ull, off, epoch);
            }
        }
        off += c[j];
    }
}

// 16 bytes global -> shared without passing through registers (LDGSTS): `bytes` (0..16) are read, the rest of the
// 16 is zero filled.  The scan issues a whole sub-tile of these from a rolled loop -- every load is in flight before
// the first is used, and the code stays a few hundred bytes.  (It used to keep the words in registers, which needs
// every loop over them unrolled: 40 KB of straight-line code that runs once per launch, i.e. straight from DRAM --
// the instruction fetch, not the data, was what the scan phase waited for.)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int SCAN_SUB_WORDS = SCAN_MAXV * 4 * SCAN_THREADS;   // 8192 words = 32 KB: one sub-tile

__device__ __forceinline__ uint64_t pack_groups(const uint4 x)
{
    return (uint64_t)word_groups(x.x) + word_groups(x.y) + word_groups(x.z) + word_groups(x.w);
}
__device__ __forceinline__ uint32_t zero_fill(uint32_t x) { return (x & ~BIT30) == BIT31; }   // a fill of 0 groups

// `smem`: 2 * SCAN_SUB_WORDS words (the expand phase's shared memory, not in use yet).
__device__ __forceinline__ void scan_body(const ScanParams &p, uint32_t *smem)
{
    constexpr int NW = SCAN_THREADS / 32;
    constexpr uint64_t TGM = (uint64_t)EXPAND_TILE_GROUPS - 1ull;
    __shared__ uint64_t s_wsum[NW];
    __shared__ uint32_t s_wbad[NW];
    __shared__ uint64_t s_lb_sum[NW];
    __shared__ ulonglong4 s_heavy[SCAN_HEAVY];
    __shared__ uint32_t s_nheavy;
    __shared__ uint32_t s_flag;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t stride = gridDim.x;
    const bool batch = p.col_groups != ~0ull;
    BoundaryGeom geo;
    geo.cg = p.col_groups;
    geo.tpc = batch ? (uint32_t)p.max_out_tiles : 0u;
    geo.k_lim = batch ? p.max_out_tiles : p.max_out_tiles + 1ull;
    geo.n_cols = p.n_cols;
    // A tile is walked in sub-tiles of up to 8192 words (SCAN_MAXV rows of 128 words per warp), staged in shared
    // memory, two buffers.  A tile of ONE sub-tile -- every stream up to gridDim * 8192 words -- is still there in
    // pass 2; a longer tile is read a second time rather than split into several tiles, because every extra tile per
    // CTA is an extra round of the offset exchange below (measured: 10 us per round).
    const uint32_t rows_full = p.tile_words / (4u * SCAN_THREADS);   // rows of 128 words per warp and tile
    uint64_t *s_loc = reinterpret_cast<uint64_t *>(smem + SCAN_SUB_WORDS);   // one sub-tile: the second buffer is free
    uint32_t budget = SPIN_LIMIT;   // polls this thread may still spend waiting for other CTAs
#ifdef WAH_TRACE
    bool first_tile = true;
#endif
    // the counters of the NEXT launch on this device start from zero whatever an aborted launch left in them
    // (launches of one device are ordered by the library, so nobody is using that slot now)
    if (blockIdx.x == 0 && tid == 0 && p.next_ctr != nullptr && p.next_ctr != p.ctr) {
        p.next_ctr->agg_count = 0;
        p.next_ctr->bad_acc = 0;
        p.next_ctr->ticket = 0;
        p.next_ctr->done = 0;
    }

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += stride) {
        const uint64_t tile_begin = (uint64_t)tile * p.tile_words;
        const uint64_t w_first = tile_begin > (uint64_t)p.skip_words ? tile_begin : (uint64_t)p.skip_words;
        const uint64_t w_last = tile_begin + p.tile_words < p.c_words ? tile_begin + p.tile_words : p.c_words;
        // (the stream's last tile stops at the row that holds the last word: it is the tile everybody's offset waits for)
        const uint32_t rows = w_last - tile_begin < (uint64_t)p.tile_words
                                  ? (uint32_t)((w_last - tile_begin + 4u * SCAN_THREADS - 1u) / (4u * SCAN_THREADS))
                                  : rows_full;
        const uint32_t nsub = (rows + SCAN_MAXV - 1) / SCAN_MAXV;
        const uint32_t padding = rows * (4u * SCAN_THREADS) - (uint32_t)(w_last - w_first);
        const bool ragged = padding != 0u;   // words of my packs lie outside the stream
        uint32_t nv = 0;          // rows of the sub-tile at hand
        uint64_t seg_begin = 0;   // first word of my warp's part of it
        // A thread only ever reads back the 16-byte packs it fetched itself: no barrier between fetch and use.
        auto my_pack = [&](uint32_t buf, uint32_t v) -> uint4 * {
            return reinterpret_cast<uint4 *>(smem + buf * SCAN_SUB_WORDS) + (warp * SCAN_MAXV + v) * 32u + lane;
        };
        auto geometry = [&](uint32_t sub) {
            nv = rows - sub * SCAN_MAXV < (uint32_t)SCAN_MAXV ? rows - sub * SCAN_MAXV : (uint32_t)SCAN_MAXV;
            seg_begin = tile_begin + (uint64_t)sub * SCAN_SUB_WORDS + (uint64_t)warp * (nv * 128u);
        };
        auto fetch_sub = [&](uint32_t sub) {   // (leaves nv / seg_begin set for `sub`)
            geometry(sub);
#pragma unroll 1
            for (uint32_t v = 0; v < nv; v++) {
                const uint64_t i0 = seg_begin + (uint64_t)(v * 32u + lane) * 4u;
                const uint32_t bytes = i0 + 4 <= p.c_words ? 16u : (i0 < p.c_words ? (uint32_t)(p.c_words - i0) * 4u : 0u);
                cp_async16((uint32_t)__cvta_generic_to_shared(my_pack(sub & 1u, v)), p.in + (i0 < p.c_words ? i0 : 0), bytes);
            }
            cp_async_commit();
        };
        // Words of my packs that are not part of the stream (behind its end; up to 3 before its start when it does
        // not begin on a 16-byte boundary) become fills of 0 groups.  First and last tile only.
        auto patch_sub = [&](uint32_t sub) {
#pragma unroll 1
            for (uint32_t v = 0; v < nv; v++) {
                const uint64_t i0 = seg_begin + (uint64_t)(v * 32u + lane) * 4u;
                if (i0 + 4 > p.c_words || i0 < (uint64_t)p.skip_words) {
                    uint4 *pk = my_pack(sub & 1u, v);
                    uint4 x = *pk;
                    const uint64_t lo = p.skip_words, hi = p.c_words;
                    x.x = (i0 >= lo && i0 < hi) ? x.x : BIT31;
                    x.y = (i0 + 1 >= lo && i0 + 1 < hi) ? x.y : BIT31;
                    x.z = (i0 + 2 >= lo && i0 + 2 < hi) ? x.z : BIT31;
                    x.w = (i0 + 3 >= lo && i0 + 3 < hi) ? x.w : BIT31;
                    *pk = x;
                }
            }
        };

        // ---- pass 1: groups in the tile (getCounts, kernels.cu:298-304)
        uint64_t lsum = 0;
        uint32_t zc = 0;   // fills of 0 groups among my words: malformed -- unless they are my own padding
        fetch_sub(0);
#pragma unroll 1
        for (uint32_t sub = 0; sub < nsub; sub++) {
            if (sub + 1u < nsub) {
                fetch_sub(sub + 1u);   // into the other buffer (which this thread has finished reading)
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            geometry(sub);
            if (ragged) patch_sub(sub);
#pragma unroll 2
            for (uint32_t v = 0; v < nv; v++) {
                const uint4 x = *my_pack(sub & 1u, v);
                lsum += pack_groups(x);
                zc += zero_fill(x.x) + zero_fill(x.y) + zero_fill(x.z) + zero_fill(x.w);
            }
        }
        const uint64_t wtotal = warp_sum_u64(lsum);
        const uint32_t wbad = warp_sum(zc);
        __syncthreads();   // the previous tile's partial sums have been consumed
        if (lane == 0) {
            s_wsum[warp] = wtotal;
            s_wbad[warp] = wbad;
        }
        __syncthreads();
        uint64_t tile_sum = 0, wprefix1 = 0;   // wprefix1: groups in the lower warps' parts (meaningful if nsub == 1)
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const uint64_t sv = s_wsum[k];
            if (k < (int)warp) wprefix1 += sv;
            tile_sum += sv;
        }
        if (tid == 0) {
            uint32_t bad = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) bad += s_wbad[k];
            bad -= padding;
            if (bad) atomicAdd(&p.ctr->bad_acc, bad);
            cell_store(p.desc + tile, tile_sum, p.epoch);
        }
        // A tile whose every word is one group (literal-dense data: sum == number of words) needs no second look at
        // its words: boundary k * 8192 falls on word (k * 8192 - excl) of the tile.  (Single stream only.)
        const bool unit_tile = !batch && w_last > w_first && tile_sum == w_last - w_first;
        const bool pre_scan = p.starts != nullptr && !unit_tile && nsub == 1u;
#ifdef WAH_TRACE
        if (p.trace && tid == 0 && first_tile) {
            p.trace[(uint64_t)blockIdx.x * 64u + 48u] = ((uint64_t)unit_tile << 63) | ((uint64_t)nsub << 48) | tile_sum;
            p.trace[(uint64_t)blockIdx.x * 64u + 49u] = w_last - w_first;
        }
        if (p.trace && tid == 0 && first_tile) p.trace[(uint64_t)blockIdx.x * 64u + 4u] = (uint64_t)clock64();
#endif

        // ---- group offset of the tile.  The tiles of one round (tile / gridDim) are counted at about the same
        //      time by different CTAs; the LAST CTA to publish its sum scans the round's sums and hands every tile
        //      its offset.  (Letting every CTA add up the sums below it needs G^2 / 2 descriptor reads per round,
        //      all polling the same few cache lines: measured 6 us; this is one read per tile.)
        uint64_t excl;
        {
            const uint32_t round0 = tile - blockIdx.x;   // first tile of my round
            const uint32_t m = p.n_tiles - round0 < stride ? p.n_tiles - round0 : stride;   // tiles in it
            if (tid == 0) {
                __threadfence();   // my sum (and my malformed-word count) is visible before my arrival is
                // Arrivals are counted over the whole launch: a CTA arrives for round r + 1 only after it has read its
                // round-r offset, which exists only after ALL round-r arrivals -- so the count at the end of round r
                // is exactly round0 + m, and nothing has to be reset (or fenced) between rounds.
                s_flag = atomicAdd(&p.ctr->agg_count, 1u) == round0 + m - 1u;
                s_nheavy = 0;
            }
            __syncthreads();
            if (s_flag) {   // (uniform) I am the round's aggregator
                __threadfence();
                const uint32_t per = (m + SCAN_THREADS - 1) / SCAN_THREADS;   // consecutive tiles per thread
                uint64_t mine = 0;
#pragma unroll 1
                for (uint32_t j = 0; j < per; j++) {
                    const uint32_t k = tid * per + j;
                    if (k < m) {
                        uint64_t v;
                        cell_load(p.desc + round0 + k, p.epoch, v);   // published: its CTA has arrived
                        mine += v;
                    }
                }
                // groups before this round = offset + sum of the previous round's last tile (both published)
                uint64_t base0 = 0;
                if (round0 != 0u) {
                    uint64_t pe = 0, ps = 0;
                    while (!cell_load(p.excl + round0 - 1, p.epoch, pe)) {
                        if (!spin_ok(budget, p.hdr, p.epoch)) break;
                        __nanosleep(64);
                    }
                    cell_load(p.desc + round0 - 1, p.epoch, ps);
                    base0 = pe + ps;
                }
                const uint64_t incl = warp_incl_scan_u64(mine);
                if (lane == 31) s_lb_sum[warp] = incl;
                __syncthreads();
                uint64_t before = base0;
#pragma unroll
                for (int k = 0; k < NW; k++)
                    if (k < (int)warp) before += s_lb_sum[k];
                before += incl - mine;
#pragma unroll 1
                for (uint32_t j = 0; j < per; j++) {
                    const uint32_t k = tid * per + j;
                    if (k < m) {
                        cell_store(p.excl + round0 + k, before, p.epoch);
                        uint64_t v;
                        cell_load(p.desc + round0 + k, p.epoch, v);
                        before += v;
                    }
                }
                if (tid == 0 && round0 + m == p.n_tiles) {
                    // the last round: every CTA added its malformed-word count before it arrived for its last tile
                    const uint32_t bad = atomicAdd(&p.ctr->bad_acc, 0u);
                    p.hdr->bad_words = bad;
                    if (p.starts == nullptr) {
                        // size query: nobody arrives any more; leave the counters zeroed for the next launch and
                        // report the status (a full decode does both when its last CTA leaves the expand phase)
                        p.ctr->agg_count = 0;
                        p.ctr->bad_acc = 0;
                        if (p.out_info)
                            p.out_info[2] = (uint64_t)bad |
                                            (*reinterpret_cast<volatile uint32_t *>(&p.hdr->error) == p.epoch ? STATUS_TIMEOUT : 0ull);
                    }
                }
                __syncthreads();   // s_lb_sum is reused below
            }
            // While the offset is on its way: everything pass 2 needs that does not depend on it -- the group offset
            // of each of my 4-word packs relative to the tile (one 64-bit warp scan per row).
            if (pre_scan) {
                uint64_t row_base = wprefix1;
#pragma unroll 1
                for (uint32_t v = 0; v < nv; v++) {
                    const uint64_t sl = pack_groups(*my_pack(0, v));
                    const uint64_t incl = warp_incl_scan_u64(sl);
                    s_loc[v * SCAN_THREADS + tid] = row_base + incl - sl;
                    row_base += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            if (tid == 0) {
                uint64_t v = 0;
                while (!cell_load(p.excl + tile, p.epoch, v)) {
                    if (!spin_ok(budget, p.hdr, p.epoch)) break;
                    __nanosleep(64);
                }
                s_lb_sum[0] = v;
            }
            __syncthreads();
            excl = s_lb_sum[0];
#ifdef WAH_TRACE
            if (p.trace && tid == 0 && first_tile) p.trace[(uint64_t)blockIdx.x * 64u + 5u] = (uint64_t)clock64();
            first_tile = false;
#endif
        }
        if (tid == 0 && tile == p.n_tiles - 1u) {
            // decompress.cu:82-93: G = last offset + last count, realSize = ceil(31 G / 32)
            const uint64_t G = excl + tile_sum;
            const uint64_t words = (G >> 5) * 31ull + (((G & 31ull) * 31ull + 31ull) >> 5);
            p.hdr->groups = G;
            p.hdr->words = words;
            p.hdr->out_tiles = (G + TGM) >> TG_SHIFT;
            if (p.out_info) {
                // (a batch reports the words one column decodes to)
                p.out_info[0] = batch ? (p.col_groups >> 5) * 31ull + (((p.col_groups & 31ull) * 31ull + 31ull) >> 5) : words;
                p.out_info[1] = G;
            }
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(&p.hdr->valid) = p.epoch;   // the expand phase polls this
        }

        // ---- pass 2: which compressed word covers each output-tile boundary ?
        //      A word that covers up to 4 boundaries records them itself; a long fill (it may span a hundred
        //      thousand output tiles) is queued and written by the whole CTA afterwards.
        ColumnCursor cur;
        cur.base = 0;
        cur.j = 0;
        if (p.starts != nullptr && unit_tile) {
            uint64_t k_end = (excl + tile_sum + TGM) >> TG_SHIFT;
            if (k_end > geo.k_lim) k_end = geo.k_lim;
#pragma unroll 1
            for (uint64_t k = ((excl + TGM) >> TG_SHIFT) + tid; k < k_end; k += SCAN_THREADS)
                store_entry(p.starts + k, w_first + ((k << TG_SHIFT) - excl) + 1ull, k << TG_SHIFT, p.epoch);
        } else if (pre_scan) {
            // one sub-tile: the words are still in shared memory and their tile-relative offsets are known
#pragma unroll 1
            for (uint32_t v = 0; v < nv; v++) {
                const uint4 x = *my_pack(0, v);
                const uint64_t sl = pack_groups(x);
                const uint64_t off = excl + s_loc[v * SCAN_THREADS + tid];
                col_seek(cur, off, geo.cg);
                const uint64_t rel = off - cur.base;
                if (((rel + TGM) >> TG_SHIFT) != ((rel + sl + TGM) >> TG_SHIFT) || rel + sl > geo.cg)   // rare: a boundary in my 4 words
                    record_boundaries(p.starts, p.epoch, geo, cur, seg_begin + (uint64_t)(v * 32u + lane) * 4u, off,
                                      make_uint4(word_groups(x.x), word_groups(x.y), word_groups(x.z), word_groups(x.w)),
                                      s_heavy, &s_nheavy);
            }
#ifdef WAH_TRACE
            if (p.trace && lane == 0) p.trace[(uint64_t)blockIdx.x * 64u + 50u + warp] = (uint64_t)clock64();
#endif
        } else if (p.starts != nullptr) {
            uint64_t sub_base = excl;   // group offset of the sub-tile's first word
            fetch_sub(0);
#pragma unroll 1
            for (uint32_t sub = 0; sub < nsub; sub++) {
                if (sub + 1u < nsub) {
                    fetch_sub(sub + 1u);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                geometry(sub);
                if (ragged) patch_sub(sub);
                uint64_t lsub = 0;
#pragma unroll 2
                for (uint32_t v = 0; v < nv; v++) lsub += pack_groups(*my_pack(sub & 1u, v));
                const uint64_t wsub = warp_sum_u64(lsub);
                __syncthreads();   // s_wsum: the previous sub-tile's (or pass 1's) sums have been consumed
                if (lane == 0) s_wsum[warp] = wsub;
                __syncthreads();
                uint64_t row_base = sub_base;   // group offset of the row's first word
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint64_t sv = s_wsum[k];
                    if (k < (int)warp) row_base += sv;
                    sub_base += sv;
                }
#pragma unroll 1
                for (uint32_t v = 0; v < nv; v++) {
                    const uint4 x = *my_pack(sub & 1u, v);
                    const uint64_t sl = pack_groups(x);
                    const uint64_t incl = warp_incl_scan_u64(sl);
                    const uint64_t off = row_base + incl - sl;
                    col_seek(cur, off, geo.cg);
                    const uint64_t rel = off - cur.base;
                    if (((rel + TGM) >> TG_SHIFT) != ((rel + sl + TGM) >> TG_SHIFT) || rel + sl > geo.cg)   // rare: a boundary in my 4 words
                        record_boundaries(p.starts, p.epoch, geo, cur, seg_begin + (uint64_t)(v * 32u + lane) * 4u, off,
                                          make_uint4(word_groups(x.x), word_groups(x.y), word_groups(x.z), word_groups(x.w)),
                                          s_heavy, &s_nheavy);
                    row_base += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
        }
        if (p.starts != nullptr && !unit_tile) {
            __syncthreads();
            const uint32_t nh = s_nheavy < (uint32_t)SCAN_HEAVY ? s_nheavy : (uint32_t)SCAN_HEAVY;
#pragma unroll 1
            for (uint32_t e = 0; e < nh; e++) {
                const ulonglong4 h = s_heavy[e];
#pragma unroll 1
                for (uint64_t k = h.z + tid; k < h.w; k += SCAN_THREADS) store_entry(p.starts + k, h.x + 1ull, h.y, p.epoch);
            }
        }
        __syncthreads();   // partial sums and the heavy queue are rewritten by the next tile
    }
    // a size query that went wrong says so even if no aggregator of a last round ever gets to report
    if (tid == 0 && p.starts == nullptr && p.out_info != nullptr && budget == 0u) p.out_info[2] = STATUS_TIMEOUT;
}

__global__ void __launch_bounds__(SCAN_THREADS) wah_scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(16) uint32_t scan_smem[];
    scan_body(p, scan_smem);
}

// ---------------------------------------------------------------- expand phase

constexpr int EXP_CHUNK = EXPAND_THREADS * 8;          // compressed words scanned per round of the general path
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

// invariants of the expand phase, checked in -DWAH_TRACE builds only: the first violation's code and value end up in
// trace[62] (compute-sanitizer is not available on the GPU pool; this found what it would have)
#ifdef WAH_TRACE
#define DCHK(cond, code, val)                                                                     \
    do {                                                                                          \
        if (!(cond) && p.trace) atomicMax((unsigned long long *)&p.trace[62], ((unsigned long long)(code) << 48) | ((unsigned long long)(val) & 0xFFFFFFFFFFFFull)); \
    } while (0)
#else
#define DCHK(cond, code, val) \
    do {                      \
    } while (0)
#endif

constexpr int SPARSE_MAX_WORDS = 4096;          // output tiles covered by at most this many compressed words take the scatter path
constexpr int SC_ROUND = EXPAND_THREADS * 4;    // compressed words per scatter round (one 16-byte pack per thread)
constexpr int COV_WORDS = 64;                   // coverage map: one bit per 16-byte unit of the tile image (1984) and one for its end

// how a tile is expanded (decided by thread 0, which has the numbers in registers)
enum : uint32_t { PATH_STOP = 0, PATH_SKIP, PATH_CONST, PATH_UNIT, PATH_SCATTER, PATH_GENERAL };

// Thread 0 resolves a tile (it may have to wait for the scan) and publishes the result in shared memory, so that
// the whole CTA works from the same numbers -- every path below has CTA barriers.
struct TileRes {
    uint64_t ws;        // first compressed word of the tile
    uint64_t dst;       // word offset of the tile's first output word in p.out
    uint64_t next_ws;   // first compressed word of the CTA's next tile, ~0 = not known yet
    uint32_t nw;        // words (ws & ~3) .. we, the tile's last word
    uint32_t skip;      // groups of word ws that belong to earlier tiles
    uint32_t tg;        // groups in the tile: 8192, fewer at the end of a column / of the stream
    uint32_t nout;      // output words to write
    uint32_t path;
    uint32_t first;     // PATH_CONST: the fill word the tile lies in
};

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
    uint32_t *s_grp = smem;                  // GRP_WORDS: one group per int, rows of 32 padded to 33 / tile image 1
    uint32_t *s_stage = smem + GRP_WORDS;    // EXPAND_TILE_WORDS output words / tile image 0
    __shared__ uint32_t s_wsum[2][NW];
    __shared__ uint2 s_list[EXP_LIST];
    __shared__ __align__(16) uint32_t s_cov[COV_WORDS];
    __shared__ uint32_t s_nlist;
    __shared__ uint32_t s_marks;
    __shared__ TileRes s_res[2];
    uint32_t n_img = 0;   // scatter tiles so far: they alternate between the two tile images

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool batch = p.col_groups != ~0ull;
    const uint32_t tpc = (uint32_t)p.max_out_tiles;
    const uint64_t n_total = batch ? (uint64_t)p.n_cols * tpc : p.max_out_tiles;

    // An output tile can be expanded as soon as the scan has recorded where it starts and where the next one
    // starts (starts[k].x = word index + 1, 0 = not recorded yet) -- the offsets of the low tiles are known long
    // before a straggling scan tile at the far end of the stream is done.  The decoded size (the header) is only
    // needed to recognise the stream's last tile.
    struct Raw {
        uint64_t sx, sy, ex;
    };
    auto peek = [&](uint64_t ot_, Raw &r) {   // non-blocking
        uint64_t ey;
        load_entry(p.starts + ot_, p.epoch, r.sx, r.sy);
        load_entry(p.starts + ot_ + 1, p.epoch, r.ex, ey);
    };
    // The same two entries, requested but not looked at: thread 0 has them copied into shared memory (cp.async, no
    // register waits for them) two tiles ahead and decodes them when the tile comes up.  x and y each carry half of
    // the epoch, so an entry the scan had not written yet when it was fetched reads as unpublished and is fetched
    // again then.
    __shared__ __align__(16) ulonglong2 s_peek[4][2];
    auto request = [&](uint64_t ot_, uint32_t slot) {   // (thread 0)
        if (ot_ < n_total) {
            cp_async16((uint32_t)__cvta_generic_to_shared(&s_peek[slot][0]), p.starts + ot_, 16u);
            cp_async16((uint32_t)__cvta_generic_to_shared(&s_peek[slot][1]), p.starts + ot_ + 1, 16u);
        }
        cp_async_commit();   // (an empty group if there is no such tile: the groups are counted)
    };
    auto decode_peek = [&](uint32_t slot, Raw &o) {
        const uint32_t e_lo = p.epoch & 0xFFFFu, e_hi = p.epoch >> 16;
        ulonglong2 ea, eb;   // (written by the copy engine behind the compiler's back)
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(ea.x), "=l"(ea.y) : "r"((uint32_t)__cvta_generic_to_shared(&s_peek[slot][0])) : "memory");
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(eb.x), "=l"(eb.y) : "r"((uint32_t)__cvta_generic_to_shared(&s_peek[slot][1])) : "memory");
        const bool oka = (uint32_t)(ea.x >> 48) == e_lo && (uint32_t)(ea.y >> 48) == e_hi;
        const bool okb = (uint32_t)(eb.x >> 48) == e_lo && (uint32_t)(eb.y >> 48) == e_hi;
        o.sx = oka ? (ea.x & ENTRY_MASK) : 0ull;
        o.sy = ea.y & ENTRY_MASK;
        o.ex = okb ? (eb.x & ENTRY_MASK) : 0ull;
    };
    auto header_known = [&]() -> bool {
        if (*reinterpret_cast<const volatile uint32_t *>(&p.hdr->valid) != p.epoch) return false;
        __threadfence();
        return true;
    };
    // the 16-byte pack a thread handles in the first round of the scatter path
    auto pack_at = [&](uint64_t i0, uint4 &x) {
        if (i0 + 4 <= p.c_words) {
            x = *reinterpret_cast<const uint4 *>(p.in + i0);
        } else {
            x.x = i0 < p.c_words ? p.in[i0] : BIT31;
            x.y = i0 + 1 < p.c_words ? p.in[i0 + 1] : BIT31;
            x.z = i0 + 2 < p.c_words ? p.in[i0 + 2] : BIT31;
            x.w = BIT31;
        }
    };

    uint32_t w[8];
    uint4 xpre = make_uint4(0, 0, 0, 0), xpre_n = make_uint4(0, 0, 0, 0);
    uint64_t xpre_ws = ~0ull, xpre_n_ws = ~0ull;   // which tile start the prefetched words belong to
    // a finished tile image whose bulk store has not been issued yet: thread 0 issues it behind the NEXT tile's first
    // barrier, which is the one that orders every thread's writes to the image before it
    uint32_t pend_bytes = 0, pend_src = 0;
    uint32_t *pend_dst = nullptr;
    uint32_t budget = SPIN_LIMIT;
    // Which tiles a CTA expands: the first EXPAND_STATIC_ROUNDS rounds are dealt round robin (no communication, and
    // known before the scan phase is over); after that a CTA draws a ticket whenever it starts a tile -- the tile it
    // will expand four iterations later, so that the ticket, the tile's bookkeeping and its first words are all on
    // their way long before they are needed.  Tiles differ in cost and SMs in speed: with a fixed deal the slowest CTA
    // finished 5 us (of 36) after the median one.  All of this is thread 0's business.
    const uint64_t GD = gridDim.x;
    const bool tickets = p.dynamic_tiles != 0u;
    uint64_t ot = blockIdx.x, ot1 = ot + GD, ot2 = ot + 2ull * GD;   // (thread 0) this tile and the CTA's next two
    uint32_t tk_prev = 0, tk_new = 0;   // thread 0: tickets drawn one / zero iterations ago
    uint32_t it = 0;
    if (tid == 0) {
        request(ot, 0);
        request(ot1, 1);
    }
    for (;; it++, xpre = xpre_n, xpre_ws = xpre_n_ws) {
        {
            // Thread 0 draws a ticket.  (ptxas wraps an atom.add on a provably uniform address in its warp-aggregation
            // idiom, whose shuffle waits for the result on the spot, and a draw inside a branch is copied -- i.e.
            // waited for -- at the branch's end: hence an address the compiler cannot prove uniform (p.zero is 0) and a
            // predicated instruction.  The result is first used an iteration later.)
            tk_prev = tk_new;
            asm volatile(
                "{\n\t"
                ".reg .pred q;\n\t"
                "setp.ne.u32 q, %2, 0;\n\t"
                "@q atom.relaxed.gpu.global.add.u32 %0, [%1], 1;\n\t"
                "}"
                : "+r"(tk_new)
                : "l"(&p.ctr->ticket + (size_t)lane * p.zero), "r"((uint32_t)(tickets && tid == 0))
                : "memory");
        }

        // ---- resolve the tile (thread 0, blocking)
        if (tid == 0) {
            TileRes r;
            r.path = PATH_STOP;
            r.ws = 0;
            r.dst = 0;
            r.nw = r.skip = r.nout = r.first = 0;
            r.tg = EXPAND_TILE_GROUPS;
            r.next_ws = ~0ull;
            request(ot2, (it + 2u) & 3u);
            cp_async_wait<1>();   // the entries of this tile and the next have arrived (only the request above may be pending)
            const uint64_t ot3 = (!tickets || it == 0u) ? ot2 + GD : (uint64_t)EXPAND_STATIC_ROUNDS * GD + tk_prev;
            if (ot < n_total) {
                // where the tile starts in the stream's group numbering, and how many groups it holds
                uint32_t j = 0, k = (uint32_t)ot;
                uint64_t g_start = ot << TG_SHIFT;
                uint32_t tg = EXPAND_TILE_GROUPS;
                if (batch) {
                    j = (uint32_t)ot / tpc;
                    k = (uint32_t)ot - j * tpc;
                    const uint64_t in_col = (uint64_t)k << TG_SHIFT;
                    g_start = (uint64_t)j * p.col_groups + in_col;
                    if (p.col_groups - in_col < (uint64_t)EXPAND_TILE_GROUPS) tg = (uint32_t)(p.col_groups - in_col);
                }
                const uint64_t g_end = g_start + tg;   // where the next tile starts
                bool stop = false, last = false, hdr = false;
                Raw cur, nx1;
                decode_peek(it & 3u, cur);          // requested two tiles ago
                decode_peek((it + 1u) & 3u, nx1);   // requested one tile ago
                r.next_ws = nx1.sx != 0ull ? nx1.sx - 1ull : ~0ull;
                while (cur.sx == 0ull) {
                    if (!hdr) hdr = header_known();
                    if (hdr && g_start >= p.hdr->groups) {
                        stop = true;   // the stream ends before this tile
                        break;
                    }
                    if (!spin_ok(budget, p.hdr_rw, p.epoch)) {
                        stop = true;
                        break;
                    }
                    __nanosleep(128);   // polite polling, see scan_body
                    peek(ot, cur);
                }
                while (!stop && cur.ex == 0ull) {
                    if (!hdr) hdr = header_known();
                    if (hdr && g_end >= p.hdr->groups) {
                        last = true;   // the stream's last tile: it ends with the last compressed word
                        break;
                    }
                    if (!spin_ok(budget, p.hdr_rw, p.epoch)) {
                        stop = true;
                        break;
                    }
                    __nanosleep(128);
                    peek(ot, cur);
                }
                // room for the tile's words
                uint64_t w_lo = (uint64_t)k * EXPAND_TILE_WORDS;   // word offset in the column / the stream
                uint64_t total_words = p.out_cap;                  // single stream: capacity; batch: words per column
                if (last) {
                    const uint64_t G = p.hdr->groups;
                    if (G - g_start < (uint64_t)tg) tg = (uint32_t)(G - g_start);
                    if (!batch && p.hdr->words < total_words) total_words = p.hdr->words;
                }
                if (!stop) {
                    const uint64_t avail = w_lo < total_words ? total_words - w_lo : 0ull;
                    const uint32_t nout = avail < (uint64_t)EXPAND_TILE_WORDS ? (uint32_t)avail : (uint32_t)EXPAND_TILE_WORDS;
                    const uint64_t ws = cur.sx - 1ull;
                    const uint64_t we = last ? p.c_words - 1ull : cur.ex - 1ull;
                    const uint64_t span = we - (ws & ~3ull) + 1ull;
                    r.ws = ws;
                    r.dst = (uint64_t)j * p.col_stride + w_lo;
                    r.nw = span > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)span;
                    r.skip = (uint32_t)(g_start - cur.sy);   // groups of word ws that belong to earlier tiles
                    r.tg = tg;
                    r.nout = nout;
                    DCHK(ws < p.c_words && we < p.c_words && we >= ws, 4, ((uint64_t)(ws & 0xFFFFFF) << 24) | (we & 0xFFFFFF));
                    DCHK(g_start >= cur.sy, 5, ot);
                    if (nout == 0u) {
                        // no room: a single stream is cut short by the capacity here and for good, a column only here
                        r.path = batch ? PATH_SKIP : PATH_STOP;
                    } else if (ws == we) {
                        // the tile lies inside ONE word: written without decoding if that is a fill (the stream's
                        // last tile may end in a partly padded group: a one-fill there is expanded like any other tile)
                        r.first = p.in[ws];
                        r.path = (is_fill(r.first) && (!last || !(r.first & BIT30))) ? PATH_CONST : PATH_SCATTER;
                    } else if (!last && r.skip == 0u && we - ws == (uint64_t)EXPAND_TILE_GROUPS && tg == (uint32_t)EXPAND_TILE_GROUPS &&
                               nout == (uint32_t)EXPAND_TILE_WORDS) {
                        r.path = PATH_UNIT;   // 8192 groups from 8192 words: every word is one group
                    } else {
                        r.path = r.nw <= (uint32_t)SPARSE_MAX_WORDS ? PATH_SCATTER : PATH_GENERAL;
                    }
                }
            }
            if (r.path == PATH_SCATTER || r.path == PATH_GENERAL || r.path == PATH_STOP) bulk_wait_read<0>();   // the image this tile is built in is no longer being read
            s_res[it & 1u] = r;
            ot = ot1;
            ot1 = ot2;
            ot2 = ot3;
        }
        __syncthreads();
        if (tid == 0 && pend_bytes != 0u) {
            bulk_s2g(pend_dst, pend_src, pend_bytes);   // (every thread fenced its writes to the image before the barrier)
            pend_bytes = 0;
        }
        const TileRes &res = s_res[it & 1u];
        const uint32_t path = res.path;
        if (path == PATH_STOP) break;
        xpre_n_ws = res.next_ws;
        if (xpre_n_ws != ~0ull) pack_at((xpre_n_ws & ~3ull) + 4ull * tid, xpre_n);   // the next tile's first words, a tile ahead
        if (path == PATH_SKIP) continue;
        const uint64_t ws = res.ws;
        const uint32_t skip = res.skip, tg = res.tg, nout = res.nout;
        uint32_t *dst = p.out + res.dst;
        uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
        const uint32_t nvec = nout >> 2;
        const uint64_t wa = ws & ~3ull;   // 16-byte aligned start; words before ws are ignored
        const uint32_t nw = res.nw;       // words wa .. we
        const uint32_t w_end = nw - 1u;   // index of we relative to wa
        const uint32_t w_beg = (uint32_t)(ws - wa);
#ifdef WAH_TRACE
        if (p.trace && tid == 0) {
            const uint64_t k = it;
            if (k < 40) p.trace[(uint64_t)blockIdx.x * 64u + 8u + k] = (uint64_t)clock64();
            p.trace[(uint64_t)blockIdx.x * 64u + 59u] = ((uint64_t)path << 32) | (uint64_t)nw;
        }
#endif

        if (path == PATH_CONST) {
            const uint32_t f = (res.first & BIT30) ? 0xFFFFFFFFu : 0u;
            const uint4 v = make_uint4(f, f, f, f);
            for (uint32_t i = tid; i < nvec; i += EXPAND_THREADS) st_stream_v4(dst4 + i, v);
            for (uint32_t i = (nvec << 2) + tid; i < nout; i += EXPAND_THREADS) dst[i] = f;
            continue;
        }

        if (path == PATH_UNIT) {
            // ================= unit path (literal dominated data) =================
            // 8192 groups from 8192 words: every word is one group (a literal, or a fill of length 1).  A warp
            // takes 32 rows of 32 words with coalesced loads; output word j of a row needs groups j and j + 1
            // (kernels.cu:375), i.e. the neighbouring lane's word.  No shared memory, no barrier.
            const uint32_t *src = p.in + ws + 1024u * warp;
            uint32_t *o = dst + 992u * warp;
#pragma unroll 8
            for (uint32_t k = 0; k < 32u; k++) {
                const uint32_t x = ld_stream_u32(src + 32u * k + lane);
                const uint32_t v = is_fill(x) ? ((x & BIT30) ? ONES31 : 0u) : x;
                const uint32_t nx = __shfl_down_sync(0xffffffffu, v, 1);
                if (lane < 31u) st_stream_u32(o + 31u * k + lane, (v >> lane) | (nx << (31u - lane)));
            }
            continue;
        }

        if (path == PATH_SCATTER) {
            // ================= scatter path (up to 4096 compressed words in the tile) =================
            // The tile image (7936 words) is cleared in shared memory; the tile's words are scanned in rounds of 1024
            // (four per thread; warps beyond the last word idle).  A literal ORs its 31 bits into the one or two words
            // it touches.  A one-fill ORs the partial words at its two ends in, writes the up to three whole words
            // between each end and the next 16-byte boundary itself, and marks the whole 16-byte units in between in a
            // coverage map (bit u = unit u of the image): a toggle where they begin, a toggle where they end.  A prefix
            // XOR over the map (62 words: every warp does it for itself, two words per lane) then tells every unit
            // whether it lies inside a one-fill, and the CTA writes those units, one 128-bit store per thread and
            // eight units -- whatever the mix of runs, that part of the work is spread evenly over the threads (the
            // owner of a long one-fill used to write all of it himself while 255 threads waited at the next barrier).
            // Zero fills cost nothing.  Shared-memory atomics keep the scatter free of ordering between threads.
            uint32_t *img = (n_img & 1u) ? s_grp : s_stage;
            n_img++;
            {
                uint4 *z = reinterpret_cast<uint4 *>(img);
                for (uint32_t i = tid; i < (uint32_t)EXPAND_TILE_WORDS / 4; i += EXPAND_THREADS) z[i] = make_uint4(0, 0, 0, 0);
                if (tid < (uint32_t)COV_WORDS / 4) reinterpret_cast<uint4 *>(s_cov)[tid] = make_uint4(0, 0, 0, 0);
                if (tid == 0) s_marks = 0;
            }
            uint32_t running = 0;   // group offset (tile relative) of the round's first word
            uint4 xc = xpre;
            if (xpre_ws != ws && 128u * warp < nw) pack_at(wa + 4ull * tid, xc);
            uint4 xn = xc;
            uint32_t rnd = 0;
            for (uint32_t c0 = 0; c0 < nw; c0 += SC_ROUND, rnd++) {
                const bool active = c0 + 128u * warp < nw;   // words for my warp in this round
                uint32_t x[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
                uint32_t tsum = 0, incl = 0;
                if (active) {
                    const uint32_t r0 = c0 + 4u * tid;   // my four consecutive words, relative to wa
                    if (c0 + SC_ROUND + 128u * warp < nw) pack_at(wa + r0 + SC_ROUND, xn);   // the next round's words, a round ahead
                    x[0] = xc.x; x[1] = xc.y; x[2] = xc.z; x[3] = xc.w;
                    xc = xn;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t ri = r0 + i;
                        uint32_t v = word_groups(x[i]);
                        if (ri < w_beg || ri > w_end) v = 0;   // outside this tile's word range
                        else if (ri == w_beg) v -= skip;       // part of the first word belongs to earlier tiles
                        c[i] = v > EXP_CLAMP ? EXP_CLAMP : v;
                        tsum += c[i];
                    }
                    incl = warp_incl_scan(tsum);
                }
                if (lane == 31) s_wsum[rnd & 1u][warp] = incl;   // (0 from an idle warp)
                __syncthreads();   // the warp sums are there (first round: and the image is cleared)
                uint32_t off = running + (incl - tsum);
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint32_t sv = s_wsum[rnd & 1u][k];
                    if (k < (int)warp) off += sv;
                    running += sv;
                }
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (c[i] != 0u && off < tg) {
                            const uint32_t wv = x[i];
                            const uint32_t b0 = 31u * off, w0 = b0 >> 5, sh = b0 & 31u;
                            DCHK(w0 < (uint32_t)EXPAND_TILE_WORDS, 1, w0);
                            if (!is_fill(wv)) {   // kernels.cu:351-354, packed at once (kernels.cu:375)
                                atomicOr(img + w0, wv << sh);
                                if (sh > 1u) atomicOr(img + w0 + 1, wv >> (32u - sh));
                            } else if (wv & BIT30) {   // one-fill, kernels.cu:337-348
                                uint32_t g1 = off + c[i];
                                if (g1 > tg) g1 = tg;
                                const uint32_t b1 = 31u * g1;
                                DCHK(b1 > b0 && b1 <= 31u * EXPAND_TILE_GROUPS, 2, ((uint64_t)b0 << 24) | b1);
                                const uint32_t w1 = (b1 - 1u) >> 5;
                                const uint32_t m0 = 0xFFFFFFFFu << sh, m1 = 0xFFFFFFFFu >> (31u - ((b1 - 1u) & 31u));
                                if (w0 == w1) {
                                    atomicOr(img + w0, m0 & m1);
                                } else {
                                    atomicOr(img + w0, m0);
                                    atomicOr(img + w1, m1);
                                    // whole words a .. w1 - 1 are this run's alone; whole 16-byte units [ua, ub) of them
                                    const uint32_t a = w0 + 1u, ua = (a + 3u) >> 2, ub = w1 >> 2;
                                    const bool units = ub > ua;
                                    const uint32_t head_end = units ? 4u * ua : w1;   // words a .. head_end - 1: at most 3 (6 if no unit)
#pragma unroll
                                    for (uint32_t k = 0; k < 6u; k++)
                                        if (a + k < head_end) img[a + k] = 0xFFFFFFFFu;
                                    if (units) {
#pragma unroll
                                        for (uint32_t k = 0; k < 3u; k++)
                                            if (4u * ub + k < w1) img[4u * ub + k] = 0xFFFFFFFFu;
                                        atomicXor(s_cov + (ua >> 5), 1u << (ua & 31u));
                                        atomicXor(s_cov + (ub >> 5), 1u << (ub & 31u));
                                        s_marks = 1;
                                    }
                                }
                            }
                        }
                        off += c[i];
                    }
                }
                if (running >= tg) break;   // uniform: the tile is covered
            }
            __syncthreads();   // every mark is in the coverage map
            if (s_marks != 0u) {
                // prefix XOR over the 64 words of the map, two per lane
                const uint2 cv = reinterpret_cast<const uint2 *>(s_cov)[lane];
                uint32_t pa = cv.x, pb = cv.y;
                pa ^= pa << 1; pa ^= pa << 2; pa ^= pa << 4; pa ^= pa << 8; pa ^= pa << 16;
                pb ^= pb << 1; pb ^= pb << 2; pb ^= pb << 4; pb ^= pb << 8; pb ^= pb << 16;
                if (pa >> 31) pb = ~pb;
                const uint32_t tops = __ballot_sync(0xffffffffu, (pb >> 31) != 0u);
                if (__popc(tops & lanemask_lt()) & 1u) {
                    pa = ~pa;
                    pb = ~pb;
                }
                // unit u = 256 r + tid: bit `lane` of map word 8 r + warp, which lane (8 r + warp) / 2 holds
                const uint32_t mine = (warp & 1u) ? pb : pa;
                uint4 *img4 = reinterpret_cast<uint4 *>(img);
#pragma unroll
                for (uint32_t r8 = 0; r8 < 8u; r8++) {
                    const uint32_t v = __shfl_sync(0xffffffffu, mine, r8 * 4u + (warp >> 1));
                    const uint32_t u = r8 * 256u + tid;
                    if (((v >> lane) & 1u) != 0u && u < (uint32_t)EXPAND_TILE_WORDS / 4) img4[u] = make_uint4(~0u, ~0u, ~0u, ~0u);
                }
            }
            if (nout == (uint32_t)EXPAND_TILE_WORDS) {
                fence_async_smem();   // my writes to the image, visible to the bulk copy engine
                if (tid == 0) {       // issued behind the next tile's first barrier
                    pend_bytes = EXPAND_TILE_WORDS * 4u;
                    pend_src = (uint32_t)__cvta_generic_to_shared(img);
                    pend_dst = dst;
                }
            } else {
                // the last tile of a column or of the stream, or one cut short by the output capacity: the part that
                // exists, by hand
                __syncthreads();
                const uint4 *src4 = reinterpret_cast<const uint4 *>(img);
                for (uint32_t i = tid; i < nvec; i += EXPAND_THREADS) st_stream_v4(dst4 + i, src4[i]);
                for (uint32_t i = (nvec << 2) + tid; i < nout; i += EXPAND_THREADS) dst[i] = img[i];
            }
            continue;
        }

        // ================= general path (more than 4096 words in the tile) =================
        // uses both images as scratch: the bulk store issued above may still be reading one of them
        if (n_img != 0u) {
            if (tid == 0) bulk_wait_read<0>();
            __syncthreads();
        }

        // ---- 1. clear the group array
        {
            uint4 *z = reinterpret_cast<uint4 *>(s_grp);
            for (uint32_t i = tid; i < GRP_WORDS / 4; i += EXPAND_THREADS) z[i] = make_uint4(0, 0, 0, 0);
            if (tid == 0) s_nlist = 0;
        }
        __syncthreads();

        // ---- 2. scan the compressed words of the tile in rounds of EXP_CHUNK and scatter them
        uint32_t running = 0;                           // group offset (tile relative) of the round's first word
        for (uint32_t c0 = 0; c0 < nw; c0 += EXP_CHUNK) {
            const uint32_t r0 = c0 + 8u * tid;          // my 8 consecutive words, relative to wa
            load8(p, wa + r0, w);
            uint32_t c[8];
            uint32_t tsum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t ri = r0 + i;
                uint32_t x = word_groups(w[i]);
                if (ri < w_beg || ri > w_end) x = 0;     // outside this tile's word range
                else if (ri == w_beg) x -= skip;         // part of the first word belongs to earlier tiles
                c[i] = x > EXP_CLAMP ? EXP_CLAMP : x;
                tsum += c[i];
            }
            const uint32_t incl = warp_incl_scan(tsum);
            if (c0 != 0) __syncthreads();   // previous round's s_wsum consumed
            if (lane == 31) s_wsum[0][warp] = incl;
            __syncthreads();
            uint32_t off = running + (incl - tsum);
            uint32_t round_sum = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const uint32_t sv = s_wsum[0][k];
                if (k < (int)warp) off += sv;
                round_sum += sv;
            }
            running += round_sum;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (c[i] != 0u && off < tg) {
                    const uint32_t wv = w[i];
                    if (!is_fill(wv)) {
                        s_grp[grp_pos(off)] = wv;                                 // kernels.cu:351-354
                    } else if (wv & BIT30) {                                      // one-fill, kernels.cu:337-348
                        const uint32_t lo = off;
                        uint32_t hi = lo + c[i];
                        if (hi > tg) hi = tg;
                        if (hi - lo <= 8u) {
                            for (uint32_t g = lo; g < hi; g++) s_grp[grp_pos(g)] = ONES31;
                        } else {
                            const uint32_t e = atomicAdd(&s_nlist, 1u);
                            if (e < EXP_LIST) s_list[e] = make_uint2(lo, hi);
                            else for (uint32_t g = lo; g < hi; g++) s_grp[grp_pos(g)] = ONES31;
                        }
                    }
                }
                off += c[i];
            }
            if (running >= tg) break;   // uniform: the tile is covered
        }
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
        if (32u * tid < tg) {
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
        // no barrier here: whatever the next tile does to shared memory happens behind its first barrier
    }
#ifdef WAH_TRACE
    if (p.trace && tid == 0) p.trace[(uint64_t)blockIdx.x * 64u + 60u] = (uint64_t)clock64();
#endif
    if (tid == 0) bulk_wait_read<0>();   // shared memory must outlive the bulk stores that read it
#ifdef WAH_TRACE
    if (p.trace && tid == 0) p.trace[(uint64_t)blockIdx.x * 64u + 61u] = (uint64_t)clock64();
#endif
    if (p.ctr != nullptr && tid == 0) {
        // The last CTA to leave reports the launch's status and zeroes the counters for the next launch -- all of them,
        // so that a launch that went wrong (a CTA gave up waiting) still leaves its slot clean.  (My own draws are
        // performed before that: the fence orders them before my `done`.)
        // (a CTA that saw the launch fail says so itself: with poisoned counters there may be no "last CTA")
        if (p.out_info && *reinterpret_cast<volatile uint32_t *>(&p.hdr_rw->error) == p.epoch) p.out_info[2] = STATUS_TIMEOUT;
        __threadfence();
        if (atomicAdd(&p.ctr->done, 1u) == gridDim.x - 1u) {
            __threadfence();
            if (p.out_info) {
                uint64_t st = (uint64_t)p.hdr->bad_words;
                if (*reinterpret_cast<volatile uint32_t *>(&p.hdr_rw->error) == p.epoch) st |= STATUS_TIMEOUT;
                if (batch && p.hdr->groups != (uint64_t)p.n_cols * p.col_groups) st |= STATUS_BATCH_LENGTH;
                p.out_info[2] = st;
            }
            p.ctr->agg_count = 0;
            p.ctr->bad_acc = 0;
            p.ctr->ticket = 0;
            p.ctr->done = 0;
        }
    }
}

// Both phases in one persistent launch: every CTA first takes its share of the scan tiles, then its share of the
// output tiles, each of which waits only for its own two `starts` entries.  Saves a launch, the idle tail / ramp
// between two kernels, and the wait for the slowest scan tile.
static_assert(2 * SCAN_SUB_WORDS <= GRP_WORDS + EXPAND_TILE_WORDS, "the scan phase's two sub-tile buffers live in the expand phase's shared memory");
static_assert(SCAN_THREADS == EXPAND_THREADS, "the fused kernel runs both phases with one CTA shape");
#ifdef WAH_TRACE
#define DTRACE(slot, val)                                                                      \
    do {                                                                                       \
        if (ep.trace && threadIdx.x == 0) ep.trace[(uint64_t)blockIdx.x * 64u + (slot)] = (uint64_t)(val); \
    } while (0)
#else
#define DTRACE(slot, val) \
    do {                  \
    } while (0)
#endif
__global__ void __launch_bounds__(EXPAND_THREADS, 3) wah_decode_kernel(const ScanParams sp, const ExpandParams ep)
{
    pdl_launch_dependents();
    pdl_wait();   // the previous kernel on the stream (typically the compressor) is complete
    DTRACE(0, clock64());
#ifdef WAH_TRACE
    {
        uint64_t gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        DTRACE(6, gt);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        DTRACE(7, smid);
    }
#endif
    extern __shared__ __align__(16) uint32_t decode_smem[];   // the scan phase borrows the expand phase's tile images
    scan_body(sp, decode_smem);
    DTRACE(1, clock64());
    __syncthreads();
    DTRACE(2, clock64());
    expand_body(ep);
    DTRACE(3, clock64());
}

}  // namespace

size_t expand_smem_bytes()
{
    return (size_t)(GRP_WORDS + EXPAND_TILE_WORDS) * sizeof(uint32_t);
}

// ---- launch geometry, per device (a process may drive several devices: SM count, occupancy and the >48 KB dynamic
//      shared memory attribute belong to the device that is current at launch time)

namespace {
constexpr int MAX_DEVICES = 64;
struct DecodeLaunchState {
    int decode_grid = 0;   // SMs x occupancy of wah_decode_kernel
    int scan_grid = 0;     // ... of wah_scan_kernel
};
DecodeLaunchState g_dls[MAX_DEVICES];

cudaError_t current_device(int *dev)
{
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return e;
    if (*dev < 0 || *dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    return cudaSuccess;
}

cudaError_t decode_grid_for_device(int dev, int *grid)
{
    DecodeLaunchState &s = g_dls[dev];
    if (s.decode_grid == 0) {
        const size_t smem = expand_smem_bytes();
        cudaError_t e = cudaFuncSetAttribute(wah_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wah_decode_kernel, EXPAND_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        s.decode_grid = sms * per_sm;
    }
    *grid = s.decode_grid;
    return cudaSuccess;
}
}  // namespace

cudaError_t scan_tile_words(uint64_t c_words, uint32_t *tile_words)
{
    int dev = 0, grid = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    e = decode_grid_for_device(dev, &grid);
    if (e != cudaSuccess) return e;
    const uint64_t unit = 4ull * SCAN_THREADS;
    uint64_t tw = ((c_words + (uint64_t)grid - 1) / (uint64_t)grid + unit - 1) / unit * unit;
    if (tw < (uint64_t)SCAN_TILE_WORDS) tw = SCAN_TILE_WORDS;
    if (tw > 0xFFFFF000ull) tw = 0xFFFFF000ull;   // (a tile is walked in sub-tiles of 8192 words; one tile per CTA)
    *tile_words = (uint32_t)tw;
    return cudaSuccess;
}

cudaError_t launch_scan(const ScanParams &p, cudaStream_t stream)
{
    // persistent + cooperative: the offset exchange spins on tiles owned by other CTAs, all must be resident
    constexpr size_t smem = 2 * SCAN_SUB_WORDS * sizeof(uint32_t);
    int dev = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    DecodeLaunchState &s = g_dls[dev];
    if (s.scan_grid == 0) {
        e = cudaFuncSetAttribute(wah_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wah_scan_kernel, SCAN_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (per_sm > 4) per_sm = 4;
        s.scan_grid = sms * per_sm;
    }
    int grid = s.scan_grid;
    if ((uint32_t)grid > p.n_tiles) grid = (int)p.n_tiles;
    ScanParams params = p;
    void *args[] = {&params};
    return cudaLaunchCooperativeKernel((const void *)wah_scan_kernel, dim3(grid), dim3(SCAN_THREADS), args, smem, stream);
}

cudaError_t launch_decode(const ScanParams &sp, const ExpandParams &ep, cudaStream_t stream)
{
    // persistent, every CTA resident (both phases spin on results of other CTAs): SMs x occupancy CTAs
    int dev = 0, grid = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    e = decode_grid_for_device(dev, &grid);
    if (e != cudaSuccess) return e;
    ScanParams a = sp;
    ExpandParams b = ep;
    void *args[] = {&a, &b};
    return launch_pdl((const void *)wah_decode_kernel, grid, EXPAND_THREADS, args, expand_smem_bytes(), stream);
}

}  // namespace wahb200

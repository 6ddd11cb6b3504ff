// wah_decompress.cu -- WAH decompressor for sm_100a.
//
// Replaces the reference's getCounts + thrust::exclusive_scan + decompressWords +
// mergeWords (kernels.cu:291-385, decompress.cu:66-115), which materialise an
// 8-byte count per compressed word and a one-group-per-int intermediate array in
// HBM and expand every fill with a serial per-thread loop (kernels.cu:346-348).
//
// Here: ONE persistent launch (wah_decode_kernel, 3 CTAs of 8 warps per SM), two phases per CTA:
//   scan phase    one tile of the compressed stream per CTA: per-tile group sums, exchanged through a round aggregator,
//                 give every tile its group offset; each tile then records, for every OUTPUT tile boundary (a multiple
//                 of 1024 groups) that falls into it, which compressed word covers it.
//   expand phase  output centric and WARP autonomous, hence load balanced whatever the fill lengths are and free of CTA
//                 barriers.  An output tile is 1024 groups = 992 words; a warp waits only for the table entries of its
//                 own chunk of 8 tiles and expands each tile on its own: constant stores if the tile lies inside one
//                 fill; a shuffle repack if it is 1024 literals; otherwise the window path -- the tile's words are
//                 parked by rank in shared memory and flag a 1024-bit map where they start, then every lane walks the
//                 32 groups of its window and emits 31 output words (the 32 -> 31 repack of mergeWords,
//                 kernels.cu:375) into a tile image that leaves through a TMA bulk store.
// Output tiles are aligned in group space to multiples of 32 groups = 31 words, so no output word is shared
// between lanes or tiles and nothing in HBM needs atomics.  wah_scan_kernel is the scan phase alone (size query).
//
// A bitmap-index batch (n_cols streams back to back, each decoding to the same number of groups) is ONE launch as
// well: group offsets run over the concatenation, output tile k of column j starts at group j * col_groups + k * 1024
// and lands at out + j * col_stride + k * 992.
#include "wah_common.cuh"

#include <stdlib.h>
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
constexpr uint32_t TG_SHIFT = 10;
static_assert((1u << TG_SHIFT) == (uint32_t)EXPAND_TILE_GROUPS, "output tile must be 1024 groups");

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
    uint32_t ct;      // tiles per chunk of the expand phase (write_fill_entries)
};

// A long fill: table entries i_first .. i_end - 1 (column base i0) all name word wi at group offset off.  The expand phase
// asks for the entry of a chunk's first tile (a multiple of `ct` in its column) and, if that word turns out to cover the
// whole chunk, for nothing else -- so only those are written, and the nine at either end of the range (a chunk that the
// fill covers in part).  (A 16 Gbit vector of a few thousand fills is half a million entries from a handful of CTAs:
// writing them all took longer than the expand phase needed to start.)  Thread t of nt.
__device__ __forceinline__ void write_fill_entries(ulonglong2 *starts, uint32_t epoch, uint64_t wi, uint64_t off, uint64_t i_first,
                                                   uint64_t i_end, uint64_t i0, uint32_t ct_, uint32_t t, uint32_t nt)
{
    const uint64_t ct = ct_ ? ct_ : 1u;
    const uint64_t head_end = i_first + 9ull < i_end ? i_first + 9ull : i_end;
    const uint64_t tail_begin = i_end > head_end + 9ull ? i_end - 9ull : head_end;
    for (uint64_t i = i_first + t; i < head_end; i += nt) store_entry(starts + i, wi + 1ull, off, epoch);
    for (uint64_t i = tail_begin + t; i < i_end; i += nt) store_entry(starts + i, wi + 1ull, off, epoch);
    const uint64_t k_mid = (head_end - i0 + ct - 1ull) / ct * ct;   // first chunk start at or behind head_end
#pragma unroll 1
    for (uint64_t i = i0 + k_mid + (uint64_t)t * ct; i < tail_begin; i += (uint64_t)nt * ct) store_entry(starts + i, wi + 1ull, off, epoch);
}

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
                    if (e < (uint32_t)SCAN_HEAVY)
                        s_heavy[e] = make_ulonglong4(wi + j, off, (i0 + k_first) | ((i0 + k_end) << 32), i0);   // (table indices fit 32 bits)
                    else
                        write_fill_entries(starts, epoch, wi + j, off, i0 + k_first, i0 + k_end, i0, g.ct, 0u, 1u);   // (the queue is full)
                    k_end = k_first;
                }
#pragma unroll 1
                for (uint64_t k = k_first; k < k_end; k++) store_entry(starts + i0 + k, wi + j + 1ull, off, epoch);
            }
        }
        off += c[j];
    }
}

// The common case inline: a single stream and ONE boundary in the pack (with 1024-group output tiles every few packs of
// a fill-dominated stream hold one) -- find the word that covers it and record it; everything else out of line.
__device__ __forceinline__ void note_boundaries(ulonglong2 *starts, uint32_t epoch, const BoundaryGeom &g, const ColumnCursor &cur,
                                                bool batch, uint64_t kf, uint64_t ke, uint64_t wi, uint64_t off, const uint4 &x,
                                                ulonglong4 *s_heavy, uint32_t *s_nheavy)
{
    const uint32_t c0 = word_groups(x.x), c1 = word_groups(x.y), c2 = word_groups(x.z);
    if (!batch && ke - kf == 1ull) {
        if (kf < g.k_lim) {
            const uint32_t d = (uint32_t)((kf << TG_SHIFT) - off);   // the boundary's distance from the pack's first group (< 2^32)
            const uint32_t s1 = c0 + c1, s2 = s1 + c2;
            const uint32_t jj = (c0 <= d ? 1u : 0u) + (s1 <= d ? 1u : 0u) + (s2 <= d ? 1u : 0u);   // the word that covers it
            const uint32_t before = jj == 0u ? 0u : (jj == 1u ? c0 : (jj == 2u ? s1 : s2));
            store_entry(starts + kf, wi + jj + 1ull, off + before, epoch);
        }
        return;
    }
    record_boundaries(starts, epoch, g, cur, wi, off, make_uint4(c0, c1, c2, word_groups(x.w)), s_heavy, s_nheavy);
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
// The same with an L2 cache policy (l2_keep_policy): the compressed words are read up to three times (pass 1, pass 2,
// the expand phase) with 16 times their volume of output stores in between.
__device__ __forceinline__ void cp_async16_hint(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst_smem), "l"(src), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_keep_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_normal_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_v4_hint(const uint4 *p, uint64_t pol)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int SCAN_SUB_WORDS = SCAN_MAXV * 4 * SCAN_THREADS;   // 8192 words = 32 KB: one sub-tile
constexpr int SCAN_SUBSUMS = 32;   // sub-tiles of a tile whose per-warp sums pass 1 keeps for pass 2 (a longer tile adds them up again)

__device__ __forceinline__ uint64_t pack_groups(const uint4 x)
{
    return (uint64_t)word_groups(x.x) + word_groups(x.y) + word_groups(x.z) + word_groups(x.w);
}
__device__ __forceinline__ uint32_t pack_groups32(const uint4 x)   // (four counts below 2^30 each)
{
    return word_groups(x.x) + word_groups(x.y) + word_groups(x.z) + word_groups(x.w);
}
__device__ __forceinline__ uint32_t zero_fill(uint32_t x) { return (x & ~BIT30) == BIT31; }   // a fill of 0 groups
// groups of a pack (below 2^32) and how many of its four words are fills of 0 groups (the only words that hold no group)
__device__ __forceinline__ uint32_t pack_groups_zeros(const uint4 x, uint32_t &zeros)
{
    const uint32_t c0 = word_groups(x.x), c1 = word_groups(x.y), c2 = word_groups(x.z), c3 = word_groups(x.w);
    zeros += 4u - (min(c0, 1u) + min(c1, 1u) + min(c2, 1u) + min(c3, 1u));
    return c0 + c1 + c2 + c3;
}

// `smem`: 2 * SCAN_SUB_WORDS words (the expand phase's shared memory, not in use yet).
__device__ __forceinline__ void scan_body(const ScanParams &p, uint32_t *smem)
{
    constexpr int NW = SCAN_THREADS / 32;
    constexpr uint64_t TGM = (uint64_t)EXPAND_TILE_GROUPS - 1ull;
    __shared__ uint64_t s_wsum[NW];
    __shared__ uint32_t s_wbad[NW];
    __shared__ uint64_t s_lb_sum[NW];
    __shared__ uint64_t s_subsum[SCAN_SUBSUMS][NW];
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
    geo.ct = p.chunk_tiles;
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
        // (the scan is bound by its instruction count -- 24 warps per SM, about 450 instructions per thread and sub-tile in
        //  pass 1 before this was trimmed, against 4.3 TB/s -- so a tile that lies inside the stream, i.e. every tile but the
        //  first and the last, is fetched without the per-pack bounds arithmetic)
        auto fetch_sub = [&](uint32_t sub) {   // (leaves nv / seg_begin set for `sub`)
            geometry(sub);
            const uint64_t pol = p.l2_keep != 0u ? l2_keep_policy() : l2_normal_policy();
            if (!ragged) {
                const uint32_t *src = p.in + seg_begin + lane * 4u;
                uint32_t dst = (uint32_t)__cvta_generic_to_shared(my_pack(sub & 1u, 0));
#pragma unroll 1
                for (uint32_t v = 0; v < nv; v++, src += 128, dst += 512u) cp_async16_hint(dst, src, 16u, pol);
            } else {
#pragma unroll 1
                for (uint32_t v = 0; v < nv; v++) {
                    const uint64_t i0 = seg_begin + (uint64_t)(v * 32u + lane) * 4u;
                    const uint32_t bytes = i0 + 4 <= p.c_words ? 16u : (i0 < p.c_words ? (uint32_t)(p.c_words - i0) * 4u : 0u);
                    cp_async16_hint((uint32_t)__cvta_generic_to_shared(my_pack(sub & 1u, v)), p.in + (i0 < p.c_words ? i0 : 0), bytes, pol);
                }
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
        const bool keep_subsums = p.starts != nullptr && nsub > 1u && nsub <= (uint32_t)SCAN_SUBSUMS;
        uint64_t lsum = 0;
        uint32_t zc = 0;   // fills of 0 groups among my words: malformed -- unless they are my own padding
        if (nsub > 1u && !ragged) {
            // A tile of several sub-tiles is read again in pass 2 anyway: pass 1 takes its words straight into registers,
            // eight 16-byte loads per thread in flight.  (Measured on a 138 MB stream, pass 1 published by the slowest CTA
            // after: 41 us through the two shared-memory buffers, one sub-tile in flight; 36 us this way; slower again with
            // two sub-tiles in flight through the buffers, the packs copied to registers before the next request.)
            const uint64_t pol = p.l2_keep != 0u ? l2_keep_policy() : l2_normal_policy();
#pragma unroll 1
            for (uint32_t sub = 0; sub < nsub; sub++) {
                geometry(sub);
                const uint4 *src = reinterpret_cast<const uint4 *>(p.in + seg_begin) + lane;
                uint4 x[SCAN_MAXV];
#pragma unroll
                for (int v = 0; v < SCAN_MAXV; v++) x[v] = ld_v4_hint(src + ((uint32_t)v < nv ? v * 32 : 0), pol);   // (all in registers)
                uint64_t lsub = 0;
#pragma unroll
                for (int v = 0; v < SCAN_MAXV; v++)
                    if ((uint32_t)v < nv) lsub += pack_groups_zeros(x[v], zc);
                lsum += lsub;
                if (keep_subsums) {   // my warp's share of this sub-tile, for pass 2
                    const uint64_t wsub = warp_sum_u64(lsub);
                    if (lane == 0) s_subsum[sub][warp] = wsub;
                }
            }
        } else {
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
                uint64_t lsub = 0;
#pragma unroll 2
                for (uint32_t v = 0; v < nv; v++) lsub += pack_groups_zeros(*my_pack(sub & 1u, v), zc);
                lsum += lsub;
                if (keep_subsums) {   // my warp's share of this sub-tile, for pass 2
                    const uint64_t wsub = warp_sum_u64(lsub);
                    if (lane == 0) s_subsum[sub][warp] = wsub;
                }
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
        // a tile of several sub-tiles is read again in pass 2: its first sub-tile is requested now, while the offset is on its way
        const bool refetch = p.starts != nullptr && !unit_tile && nsub > 1u;
        if (refetch) fetch_sub(0);
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
                    if (p.starts == nullptr || p.scan_only) {
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
                const uint64_t kf = (rel + TGM) >> TG_SHIFT, ke = (rel + sl + TGM) >> TG_SHIFT;
                if (kf != ke || rel + sl > geo.cg)   // a boundary in my 4 words
                    note_boundaries(p.starts, p.epoch, geo, cur, batch, kf, ke, seg_begin + (uint64_t)(v * 32u + lane) * 4u, off, x,
                                    s_heavy, &s_nheavy);
            }
#ifdef WAH_TRACE
            if (p.trace && lane == 0) p.trace[(uint64_t)blockIdx.x * 64u + 50u + warp] = (uint64_t)clock64();
#endif
        } else if (p.starts != nullptr) {
            uint64_t sub_base = excl;   // group offset of the sub-tile's first word
            // words that hold under 8 groups on average: most 128-word rows hold no output-tile boundary and are skipped
            const bool skip_rows = tile_sum < 8ull * (w_last - w_first);
            ColumnCursor rcur;          // (uniform) the column that holds the row at hand
            rcur.base = 0;
            rcur.j = 0;
            // (sub-tile 0 was requested before the offset exchange)
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
                if (!skip_rows) {
                    // (most rows hold a boundary: every row is scanned -- in 32-bit arithmetic relative to the sub-tile unless
                    //  the warp's eight rows hold 2^31 groups or more)
                    uint64_t wsub;
                    if (keep_subsums) {
                        wsub = s_subsum[sub][warp];   // (pass 1 left them; the barrier behind pass 1 has been passed)
                    } else {
                        uint64_t lsub = 0;
#pragma unroll 2
                        for (uint32_t v = 0; v < nv; v++) lsub += pack_groups32(*my_pack(sub & 1u, v));
                        wsub = warp_sum_u64(lsub);
                        __syncthreads();   // s_wsum: the previous sub-tile's (or pass 1's) sums have been consumed
                        if (lane == 0) s_wsum[warp] = wsub;
                        __syncthreads();
                    }
                    const bool narrow = wsub < (1ull << 31);
                    uint64_t row_base = sub_base;   // group offset of the row's first word
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        const uint64_t sv = keep_subsums ? s_subsum[sub][k] : s_wsum[k];
                        if (k < (int)warp) row_base += sv;
                        sub_base += sv;
                    }
                    if (narrow && !batch) {
                        // a single stream, the warp's rows below 2^31 groups: offsets relative to the output-tile boundary
                        // at or below the warp's first row, in 32 bits; 64 bits only for the pack that holds a boundary
                        const uint64_t q0 = row_base & ~TGM;
                        uint32_t rb = (uint32_t)(row_base - q0);
#pragma unroll 1
                        for (uint32_t v = 0; v < nv; v++) {
                            const uint4 x = *my_pack(sub & 1u, v);
                            const uint32_t s32 = pack_groups32(x);
                            const uint32_t i32 = warp_incl_scan(s32);
                            const uint32_t e = rb + (i32 - s32);
                            const uint32_t kf = (e + (uint32_t)TGM) >> TG_SHIFT, ke = (e + s32 + (uint32_t)TGM) >> TG_SHIFT;
                            if (kf != ke)   // a boundary in my 4 words
                                note_boundaries(p.starts, p.epoch, geo, cur, false, (q0 >> TG_SHIFT) + kf, (q0 >> TG_SHIFT) + ke,
                                                seg_begin + (uint64_t)(v * 32u + lane) * 4u, q0 + e, x, s_heavy, &s_nheavy);
                            rb += __shfl_sync(0xffffffffu, i32, 31);
                        }
                        continue;
                    }
#pragma unroll 1
                    for (uint32_t v = 0; v < nv; v++) {
                        const uint4 x = *my_pack(sub & 1u, v);
                        uint64_t sl, off, rsum;
                        if (narrow) {
                            const uint32_t s32 = pack_groups32(x);
                            const uint32_t i32 = warp_incl_scan(s32);
                            sl = s32;
                            off = row_base + (uint64_t)(i32 - s32);
                            rsum = __shfl_sync(0xffffffffu, i32, 31);
                        } else {
                            sl = pack_groups(x);
                            const uint64_t incl = warp_incl_scan_u64(sl);
                            off = row_base + incl - sl;
                            rsum = __shfl_sync(0xffffffffu, incl, 31);
                        }
                        col_seek(cur, off, geo.cg);
                        const uint64_t rel = off - cur.base;
                        const uint64_t kf = (rel + TGM) >> TG_SHIFT, ke = (rel + sl + TGM) >> TG_SHIFT;
                        if (kf != ke || rel + sl > geo.cg)   // a boundary in my 4 words
                            note_boundaries(p.starts, p.epoch, geo, cur, batch, kf, ke, seg_begin + (uint64_t)(v * 32u + lane) * 4u, off, x,
                                            s_heavy, &s_nheavy);
                        row_base += rsum;
                    }
                    continue;
                }
                // the groups of each of my warp's rows (128 words): lane v ends up with row v's.  A pack's four counts fit
                // 32 bits; two 16-bit halves summed over the warp by redux.sync are exact.
                uint64_t myrow = 0;
#pragma unroll 2
                for (uint32_t v = 0; v < nv; v++) {
                    const uint32_t s32 = pack_groups32(*my_pack(sub & 1u, v));
                    const uint32_t lo = __reduce_add_sync(0xffffffffu, s32 & 0xFFFFu), hi = __reduce_add_sync(0xffffffffu, s32 >> 16);
                    if (lane == v) myrow = ((uint64_t)hi << 16) + lo;
                }
                uint64_t rincl = myrow;   // prefix over the rows (lanes 0 .. SCAN_MAXV - 1)
#pragma unroll
                for (int dd = 1; dd < SCAN_MAXV; dd <<= 1) {
                    const uint64_t o = __shfl_up_sync(0xffffffffu, rincl, dd);
                    if ((int)lane >= dd) rincl += o;
                }
                const uint64_t wsub = __shfl_sync(0xffffffffu, rincl, SCAN_MAXV - 1);
                __syncthreads();   // s_wsum: the previous sub-tile's (or pass 1's) sums have been consumed
                if (lane == 0) s_wsum[warp] = wsub;
                __syncthreads();
                uint64_t row_base = sub_base;   // group offset of my warp's first row
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint64_t sv = s_wsum[k];
                    if (k < (int)warp) row_base += sv;
                    sub_base += sv;
                }
                // Only a row that holds an output-tile boundary is looked at word by word (a literal-dense stream has
                // one in eight rows, a stream of long fills one in hundreds), in 32-bit arithmetic relative to the row
                // unless the row holds 2^32 groups or more.
#pragma unroll 1
                for (uint32_t v = 0; v < nv; v++) {
                    const uint64_t rsum = __shfl_sync(0xffffffffu, myrow, v);
                    const uint64_t roff = row_base + __shfl_sync(0xffffffffu, rincl - myrow, v);
                    col_seek(rcur, roff, geo.cg);
                    const uint64_t rel0 = roff - rcur.base;
                    if (((rel0 + TGM) >> TG_SHIFT) == ((rel0 + rsum + TGM) >> TG_SHIFT) && rel0 + rsum <= geo.cg) continue;
                    const uint4 x = *my_pack(sub & 1u, v);
                    uint64_t sl, off;
                    if (rsum < (1ull << 32)) {
                        const uint32_t s32 = pack_groups32(x);
                        sl = s32;
                        off = roff + (uint64_t)(warp_incl_scan(s32) - s32);
                    } else {
                        sl = pack_groups(x);
                        off = roff + warp_incl_scan_u64(sl) - sl;
                    }
                    ColumnCursor cur = rcur;
                    col_seek(cur, off, geo.cg);
                    const uint64_t rel = off - cur.base;
                    const uint64_t kf = (rel + TGM) >> TG_SHIFT, ke = (rel + sl + TGM) >> TG_SHIFT;
                    if (kf != ke || rel + sl > geo.cg)   // a boundary in my 4 words
                        note_boundaries(p.starts, p.epoch, geo, cur, batch, kf, ke, seg_begin + (uint64_t)(v * 32u + lane) * 4u, off, x,
                                        s_heavy, &s_nheavy);
                }
            }
        }
        if (p.starts != nullptr && !unit_tile) {
            __syncthreads();
            const uint32_t nh = s_nheavy < (uint32_t)SCAN_HEAVY ? s_nheavy : (uint32_t)SCAN_HEAVY;
#pragma unroll 1
            for (uint32_t e = 0; e < nh; e++) {
                const ulonglong4 h = s_heavy[e];
                write_fill_entries(p.starts, p.epoch, h.x, h.y, h.z & 0xFFFFFFFFull, h.z >> 32, h.w, p.chunk_tiles, tid, SCAN_THREADS);
            }
        }
        __syncthreads();   // partial sums and the heavy queue are rewritten by the next tile
    }
    // a size query that went wrong says so even if no aggregator of a last round ever gets to report
    if (tid == 0 && (p.starts == nullptr || p.scan_only) && p.out_info != nullptr && budget == 0u) p.out_info[2] = STATUS_TIMEOUT;
}

__global__ void __launch_bounds__(SCAN_THREADS) wah_scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(16) uint32_t scan_smem[];
    scan_body(p, scan_smem);
}

// ---------------------------------------------------------------- expand phase
//
// WARP autonomous: an output tile is 1024 groups = 992 words (3968 bytes, a multiple of 128), one warp expands it from
// start to finish -- its own slice of shared memory, __syncwarp() only, no CTA barrier anywhere in the phase.  (The
// first version of this phase gave a tile of 8192 groups to the CTA: thread 0 resolved the tile while 255 threads
// waited, the low warps scanned the tile's few words while the high warps waited, ...: ncu put 20 - 40 % of all issue
// slots into barrier stalls.)  Tiles are dealt in chunks of 8 consecutive tiles (31 KB of output), the first rounds
// round robin, later chunks by ticket.

constexpr int MAX_CHUNK_TILES = 8;                                            // tiles per chunk (the unit of work distribution): 8, fewer for short streams
constexpr int CW_WORDS = EXPAND_TILE_GROUPS + EXPAND_TILE_GROUPS / 32 + 8;   // a tile's words by rank (0 .. 1025), rows of 32 padded to 33
constexpr int FLAG_WORDS = EXPAND_TILE_GROUPS / 32 + 4;                       // bit g: a word starts at group g of the tile (+ a word for g = 1024)
constexpr int WARP_SMEM_WORDS = CW_WORDS + EXPAND_TILE_WORDS + FLAG_WORDS;    // 8368 bytes per warp
static_assert(EXPAND_TILE_GROUPS + 1 + ((EXPAND_TILE_GROUPS + 1) >> 5) < CW_WORDS, "s_cw too small");
static_assert((CW_WORDS * 4) % 16 == 0 && (WARP_SMEM_WORDS * 4) % 16 == 0, "the tile image must be 16-byte aligned (bulk store)");
constexpr uint32_t EXP_CLAMP = 2u * EXPAND_TILE_GROUPS;   // any count >= the tile span behaves the same
// the packed warp scan of the window path: groups in the low 20 bits (a round's 128 words x EXP_CLAMP fit), words above
constexpr uint32_t RANK_SHIFT = 20, RANK_ONE = 1u << RANK_SHIFT, GROUP_MASK = RANK_ONE - 1u;
static_assert(128u * EXP_CLAMP <= GROUP_MASK, "packed warp scan overflows");

// where the tile's word of rank r is parked: neighbouring lanes of the window walk read words about 32 ranks apart
// when the tile is literal dense -- the padding keeps those reads on different banks
__device__ __forceinline__ uint32_t cw_pos(uint32_t r) { return r + (r >> 5); }
// the 31 bits of a group that word x holds (kernels.cu:337-354)
__device__ __forceinline__ uint32_t group_bits(uint32_t x)
{
    const uint32_t f = (uint32_t)((int32_t)(x << 1) >> 31) & ONES31;   // fill: all ones or all zeros
    return is_fill(x) ? f : x;
}

// how a tile is expanded (decided once per chunk by the lane that holds the tile's table entry)
enum : uint32_t { PATH_STOP = 0, PATH_SKIP, PATH_CONST, PATH_UNIT, PATH_WINDOW };

// W consecutive compressed words for a lane, the first at src[i0]; `room` = words from src to the end of the stream
// (words behind it read as fills of 0 groups).  src is 16-byte aligned and i0 a multiple of W.
template <int W>
__device__ __forceinline__ void load_words(const uint32_t *src, uint32_t i0, uint64_t room, uint32_t (&x)[W])
{
    if ((uint64_t)i0 + W <= room) {
        if (W == 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + i0);
            x[0] = v.x; x[1 % W] = v.y; x[2 % W] = v.z; x[3 % W] = v.w;
        } else if (W == 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(src + i0);
            x[0] = v.x; x[1 % W] = v.y;
        } else {
            x[0] = src[i0];
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; i++) x[i] = (uint64_t)i0 + i < room ? src[i0 + i] : BIT31;
    }
}

// Word phase of the window path, one round: my W consecutive words x[], the first at index r0 from the tile's aligned
// start.  w_beg = index of the tile's first word there, w_span = index of its last word among the tile's words.
template <int W, bool PAD>
__device__ __forceinline__ void park_words(const uint32_t (&x)[W], uint32_t r0, uint32_t w_beg, uint32_t w_span, uint32_t skip,
                                           uint32_t tg, uint32_t &running, uint32_t &rk_run, uint32_t *s_cw, uint32_t *s_flag)
{
    uint32_t c[W];
    uint32_t tsum = 0;   // groups in the low 20 bits, words that hold a group above
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t rel = r0 + i - w_beg;   // index among the tile's words (wraps for the up to 3 words before it)
        uint32_t v = word_groups(x[i]);
        if (rel > w_span) v = 0;               // outside this tile's word range
        if (rel == 0u) v -= skip;              // part of the first word belongs to earlier tiles
        v = v > EXP_CLAMP ? EXP_CLAMP : v;
        c[i] = v;
        tsum += v + (v != 0u ? RANK_ONE : 0u);
    }
    const uint32_t incl = warp_incl_scan(tsum);
    const uint32_t excl = incl - tsum;
    uint32_t off = running + (excl & GROUP_MASK);
    uint32_t rk = rk_run + (excl >> RANK_SHIFT);
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    running += tot & GROUP_MASK;
    rk_run += tot >> RANK_SHIFT;
#pragma unroll
    for (int i = 0; i < W; i++) {
        // (the tile's last word may start exactly where the tile ends: it is parked like the others -- the ranks stay
        //  consecutive -- but reads as zeros: behind a short tile lies padding)
        if (c[i] != 0u && off <= tg) {
            const uint32_t gb = off < tg ? group_bits(x[i]) : 0u;
            s_cw[PAD ? cw_pos(rk) : rk] = gb;
            atomicOr(s_flag + (off >> 5), 1u << (off & 31u));
        }
        rk += c[i] != 0u ? 1u : 0u;
        off += c[i];
    }
}

// The walk of the window path: group JJ of the lane's window.  Where the flag map says a word starts, the address steps
// to it and the word's group bits are loaded -- both under the flag's predicate (the flag bits go into predicate registers
// seven at a time, R2P): a lane inside a fill issues no load at all, so a shared-memory request serves the two or three
// lanes that change word at this step instead of all 32 (measured: unconditional loads kept the shared-memory pipe 70 %
// busy).  One shift and one funnel shift make output word JJ - 1 = group JJ-1 >> (JJ-1) | group JJ << (32-JJ)
// (kernels.cu:375).  Five instructions per group.  (Parking both operand forms, {bits << 1, bits}, and loading 8 bytes
// saved the shift but cost two register moves per step: a predicated load of a register pair.)
// OP < 0: the output word is stored; OP = 0..3 (AND, OR, XOR, ANDNOT): it is combined into the word already there
// (the logical operators: the image holds the first operand's tile).
template <int OP>
__device__ __forceinline__ uint32_t combine(uint32_t x, uint32_t y)
{
    return OP == 0 ? (x & y) : (OP == 1 ? (x | y) : (OP == 2 ? (x ^ y) : (x & ~y)));
}
template <int JJ, int OP = -1>
__device__ __forceinline__ void walk_from(uint32_t a, uint32_t F, uint32_t v, uint32_t *o)
{
    if constexpr (JJ < 32) {
        const uint32_t lo = v << 1;
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            ".reg .b32 t;\n\t"
            "and.b32 t, %2, %3;\n\t"
            "setp.ne.u32 q, t, 0;\n\t"
            "@q add.u32 %0, %0, 4;\n\t"
            "@q ld.shared.u32 %1, [%0];\n\t"
            "}"
            : "+r"(a), "+r"(v)
            : "r"(F), "n"(1u << JJ));
        if constexpr (OP < 0)
            o[JJ - 1] = __funnelshift_r(lo, v, JJ);
        else
            o[JJ - 1] = combine<OP>(o[JJ - 1], __funnelshift_r(lo, v, JJ));
        walk_from<JJ + 1, OP>(a, F, v, o);
    }
}

// ---- bulk (TMA) store of a finished tile, shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes, bool evict_first)
{
    if (evict_first) {
        const uint64_t pol = l2_evict_first_policy();
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem),
                     "r"(bytes), "l"(pol)
                     : "memory");
    } else {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                     : "memory");
    }
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

__device__ __forceinline__ void expand_body(const ExpandParams &p)
{
    constexpr int NW = EXPAND_THREADS / 32;
    constexpr uint32_t TG = (uint32_t)EXPAND_TILE_GROUPS, TW = (uint32_t)EXPAND_TILE_WORDS;
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t *s_cw = smem + warp * WARP_SMEM_WORDS;   // the tile's words that hold a group, by rank: what a group of theirs decodes to
    uint32_t *s_stage = s_cw + CW_WORDS;              // the tile image: 992 output words
    uint32_t *s_flag = s_stage + TW;                  // bit g: a word starts at group g of the tile
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(s_stage);

    const bool batch = p.col_groups != ~0ull;
    const uint32_t tpc = (uint32_t)p.max_out_tiles;                   // tiles per column (single stream: tiles the capacity has room for)
    // tiles per chunk: 8 (31 KB of output per draw), fewer when the stream is so short that 8 would leave warps without
    // work or with twice the work of their neighbours (launch_decode picks it from the number of tiles and warps)
    const uint32_t CT = p.chunk_tiles;
    const uint32_t cpc = (tpc + CT - 1u) / CT;                        // chunks per column
    const uint32_t n_chunks = (batch ? p.n_cols : 1u) * cpc;   // (the table has fewer than 2^32 entries: wah_capi.cu)
    const uint32_t GW = gridDim.x * NW, gw = blockIdx.x * NW + warp;
    const bool tickets = p.dynamic_tiles != 0u;

    // Chunk c = tiles k0 .. k0 + 7 of column j (a single stream is one column).  Lane l <= 8 holds the table entry
    // of tile k0 + l: {index + 1 of the compressed word that covers the tile's first group (0 = not recorded), that
    // word's group offset}; the entry behind a column's last tile is the next column's first.
    auto where = [&](uint32_t c, uint32_t &j, uint32_t &k0) {
        j = batch ? (uint32_t)(c / cpc) : 0u;
        k0 = (uint32_t)(c - (uint64_t)j * cpc) * CT;
    };
    auto fetch = [&](uint32_t c, uint64_t &x, uint64_t &y) {
        x = 0;
        y = 0;
        if (c < n_chunks && lane <= CT) {
            uint32_t j, k0;
            where(c, j, k0);
            if (k0 + lane <= tpc) load_entry(p.starts + (uint64_t)j * tpc + k0 + lane, p.epoch, x, y);
        }
    };
    uint32_t budget = SPIN_LIMIT;
    bool hdr = false;                  // the decoded size is known (only needed where the stream ends: read there, not kept)
    auto total_groups = [&]() -> uint64_t { return *reinterpret_cast<const volatile uint64_t *>(&p.hdr->groups); };
    // Software pipeline over a warp's chunks -- nothing a tile needs is asked for when it is needed:
    //   chunk i + 2   its table entries are requested                      (fetch)
    //   chunk i + 1   the first word of each of its tiles is requested     (needs the entries)
    //   chunk i       is resolved (needs the first words) and expanded; tile s + 1's words are requested before
    //                 tile s is expanded.
    // The first EXPAND_STATIC_ROUNDS rounds of chunks are dealt round robin (known before the scan phase is over);
    // after that a warp draws a ticket three chunks ahead of the chunk it is for.
    static_assert(EXPAND_STATIC_ROUNDS >= 3, "the first three chunks of a warp are dealt statically");
    uint32_t c0 = gw, c1 = gw + GW, c2 = gw + 2u * GW;
    uint64_t e0x, e0y, e1x, e1y, e2x, e2y;
    fetch(c0, e0x, e0y);
    fetch(c1, e1x, e1y);
    auto first_word = [&](uint64_t x) -> uint32_t { return (lane <= CT && x != 0ull) ? ld_stream_u32(p.in + (x - 1ull)) : 0u; };
    uint32_t first0 = first_word(e0x), first1;
    bool stop = false;
    for (uint32_t it = 0; !stop && c0 < n_chunks; it++) {
        // Lane 0 draws a ticket.  (ptxas wraps an atom.add on a provably uniform address in its warp-aggregation idiom,
        // whose shuffle waits for the result on the spot -- an L2 round trip per chunk, 7 % of the stall samples of a
        // 1 Gbit decode: hence an address the compiler cannot prove uniform (p.zero is 0) and a predicated instruction.
        // The result is first used at the end of the iteration.)
        uint32_t tk = 0;
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "setp.ne.u32 q, %2, 0;\n\t"
            "@q atom.relaxed.gpu.global.add.u32 %0, [%1], 1;\n\t"
            "}"
            : "+r"(tk)
            : "l"(&p.ctr->ticket + (size_t)lane * p.zero), "r"((uint32_t)(tickets && it + 3u >= (uint32_t)EXPAND_STATIC_ROUNDS && lane == 0u))
            : "memory");
        fetch(c2, e2x, e2y);
        first1 = first_word(e1x);

        uint32_t j, k0;
        where(c0, j, k0);
        const uint32_t nt = tpc - k0 < CT ? tpc - k0 : CT;   // tiles in the chunk
        const uint64_t col_g0 = batch ? (uint64_t)j * p.col_groups : 0ull;
        // the group where my entry's tile starts (the entry behind a column's last tile: where the column ends)
        const uint64_t my_g = (batch && k0 + lane >= tpc) ? col_g0 + p.col_groups : col_g0 + ((uint64_t)(k0 + lane) << TG_SHIFT);

        // ---- the chunk's entries must have been recorded by the scan -- or lie behind the end of the stream
        if (__any_sync(0xffffffffu, lane <= nt && e0x == 0ull)) {
            const bool had = e0x != 0ull;
            const uint64_t end_g = __shfl_sync(0xffffffffu, my_g, nt);   // the group where the chunk ends
            for (;;) {
                // A chunk inside one long fill: the scan records the entry of the chunk's first tile only (write_fill_entries).
                // If the word that entry names covers the whole chunk, every tile of the chunk starts in it.
                const uint64_t x0 = __shfl_sync(0xffffffffu, e0x, 0), y0 = __shfl_sync(0xffffffffu, e0y, 0);
                if (x0 != 0ull) {
                    const uint32_t f0 = ld_stream_u32(p.in + (x0 - 1ull));
                    if (is_fill(f0) && y0 + fill_count(f0) > end_g) {
                        e0x = x0;
                        e0y = y0;
                        break;
                    }
                }
                const bool missing = lane <= nt && e0x == 0ull && !(hdr && my_g >= total_groups());
                if (!__any_sync(0xffffffffu, missing)) break;
                if (!hdr) {
                    if (*reinterpret_cast<const volatile uint32_t *>(&p.hdr->valid) == p.epoch) {
                        __threadfence();
                        hdr = true;
                        continue;
                    }
                }
                bool ok = true;
                if (lane == 0) ok = spin_ok(budget, p.hdr_rw, p.epoch);
                if (!__shfl_sync(0xffffffffu, ok, 0)) {
                    stop = true;
                    break;
                }
                __nanosleep(256);   // polite polling, see scan_body
                if (missing) load_entry(p.starts + (uint64_t)j * tpc + k0 + lane, p.epoch, e0x, e0y);
            }
            if (stop) break;
            if (!had) first0 = first_word(e0x);
        }

        // ---- resolve the chunk's tiles, lane l its tile l (the arithmetic once per chunk and in parallel instead of
        //      once per tile and uniform); what the tile loop needs comes back by shuffle, 32 bits at a time
        uint64_t my_ws = 0;
        uint32_t my_nw = 0, my_skip = 0, my_meta = PATH_STOP;
        {
            const uint64_t nxt = __shfl_down_sync(0xffffffffu, e0x, 1);
            if (lane < nt && e0x != 0ull) {
                const uint32_t k = k0 + lane;
                const uint64_t in_col = (uint64_t)k << TG_SHIFT;
                const uint64_t g_start = col_g0 + in_col;
                uint32_t tg = TG;   // groups in the tile: fewer at the end of a column / of the stream
                if (batch && p.col_groups - in_col < (uint64_t)TG) tg = (uint32_t)(p.col_groups - in_col);
                const bool last = nxt == 0ull;   // the stream's last tile: it ends with the last compressed word
                const uint64_t w_lo = (uint64_t)k * TW;   // word offset in the column / the stream
                uint64_t total_words = p.out_cap;         // single stream: capacity; batch: words per column
                if (last) {   // (the header is known: that is how the missing entry was told from a late one)
                    const uint64_t G = total_groups(), Gwords = *reinterpret_cast<const volatile uint64_t *>(&p.hdr->words);
                    if (G - g_start < (uint64_t)tg) tg = (uint32_t)(G - g_start);
                    if (!batch && Gwords < total_words) total_words = Gwords;
                }
                const uint64_t avail = w_lo < total_words ? total_words - w_lo : 0ull;
                const uint32_t nout = avail < (uint64_t)TW ? (uint32_t)avail : TW;
                const uint64_t ws = e0x - 1ull;
                const uint64_t we = last ? p.c_words - 1ull : nxt - 1ull;
                const uint64_t span = we - (ws & ~3ull) + 1ull;
                DCHK(ws < p.c_words && we < p.c_words && we >= ws, 4, ((uint64_t)(ws & 0xFFFFFF) << 24) | (we & 0xFFFFFF));
                DCHK(g_start >= e0y, 5, c0);
                my_ws = ws;
                my_nw = span > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)span;   // words (ws & ~3) .. we
                my_skip = (uint32_t)(g_start - e0y);                           // groups of word ws that belong to earlier tiles
                uint32_t path;
                if (nout == 0u)
                    path = batch ? PATH_SKIP : PATH_STOP;   // no room: a single stream is cut short by the capacity here and for good, a column only here
                else if (ws == we && is_fill(first0) && (!last || !(first0 & BIT30)))
                    path = PATH_CONST;    // inside ONE fill word (the stream's last tile may end in a partly padded group: a one-fill there is expanded like any other tile)
                else if (!last && my_skip == 0u && we - ws == (uint64_t)TG && tg == TG && nout == TW)
                    path = PATH_UNIT;     // 1024 groups from 1024 words: every word is one group
                else
                    path = PATH_WINDOW;
                my_meta = path | (tg << 4) | (nout << 16);
            }
        }
        uint32_t *dst_chunk = p.out + (uint64_t)j * p.col_stride + (uint64_t)k0 * TW;

        // what the tile at hand needs (uniform), and the first 128 words of a window-path tile, four per lane
        uint32_t meta = __shfl_sync(0xffffffffu, my_meta, 0);
        uint64_t ws = __shfl_sync(0xffffffffu, my_ws, 0);
        uint32_t nw = __shfl_sync(0xffffffffu, my_nw, 0);
        uint32_t xp[4] = {BIT31, BIT31, BIT31, BIT31};
        if ((meta & 15u) == PATH_WINDOW && 4u * lane < nw) load_words<4>(p.in + (ws & ~3ull), 4u * lane, p.c_words - (ws & ~3ull), xp);

        for (uint32_t s = 0; s < nt; s++) {
            const uint32_t path = meta & 15u, tg = (meta >> 4) & 0xFFFu, nout = meta >> 16;
            const uint64_t ws_t = ws;
            const uint32_t nw_t = nw;
            uint32_t xc[4] = {xp[0], xp[1], xp[2], xp[3]};
            // the next tile's numbers and first words: on their way while this tile is expanded
            meta = __shfl_sync(0xffffffffu, my_meta, s + 1u);
            if (s + 1u >= nt) meta = PATH_STOP;
            if ((meta & 15u) >= PATH_UNIT) {
                ws = __shfl_sync(0xffffffffu, my_ws, s + 1u);
                nw = __shfl_sync(0xffffffffu, my_nw, s + 1u);
                if ((meta & 15u) == PATH_WINDOW && 4u * lane < nw) load_words<4>(p.in + (ws & ~3ull), 4u * lane, p.c_words - (ws & ~3ull), xp);
            }
            if (path == PATH_STOP) {   // the stream (or the room for it) ends before this tile, and before every later one
                stop = true;
                break;
            }
            if (path == PATH_SKIP) continue;
            uint32_t *dst = dst_chunk + s * TW;
            uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
            const uint32_t nvec = nout >> 2;

            if (path == PATH_CONST) {
                // ================= the tile lies inside ONE fill word: written without decoding =================
                const uint32_t first = __shfl_sync(0xffffffffu, first0, s);
                const uint32_t f = (first & BIT30) ? 0xFFFFFFFFu : 0u;
                const uint4 v = make_uint4(f, f, f, f);
                if (p.l2_stream_out) {
                    const uint64_t pol = l2_evict_first_policy();
                    for (uint32_t i = lane; i < nvec; i += 32u) st_stream_v4_hint(dst4 + i, v, pol);
                } else {
                    for (uint32_t i = lane; i < nvec; i += 32u) st_stream_v4(dst4 + i, v);
                }
                for (uint32_t i = (nvec << 2) + lane; i < nout; i += 32u) dst[i] = f;
                continue;
            }

            if (path == PATH_UNIT) {
                // ================= unit path (literal dominated data) =================
                // 1024 groups from 1024 words: every word is one group (a literal, or a fill of length 1).  32 rows of 32
                // words with coalesced loads; output word j of a row needs groups j and j + 1 (kernels.cu:375), i.e.
                // the neighbouring lane's word.  No shared memory.
                // (all of a half tile's loads are in flight before the first is used: the path is bound by the bytes a warp
                //  keeps in flight, not by its instructions)
                const uint32_t *src = p.in + ws_t + lane;
                uint32_t *d = dst + lane;
#pragma unroll 1
                for (uint32_t h = 0; h < 2u; h++, src += 512, d += 496) {
                    uint32_t x[16];
#pragma unroll
                    for (int r = 0; r < 16; r++) x[r] = ld_stream_u32(src + 32 * r);
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        const uint32_t v = group_bits(x[r]);
                        const uint32_t nv = __shfl_down_sync(0xffffffffu, v, 1);
                        if (lane < 31u) st_stream_u32(d + 31 * r, (v >> lane) | (nv << (31u - lane)));
                    }
                }
                continue;
            }

            // ================= window path (every other tile) =================
            // Step 1, word centric: the tile's compressed words are taken 32, 64 or (in rounds) 128 at a time, one, two or
            //   four per lane.  One packed warp scan gives every word its tile-relative group offset AND its rank among
            //   the words that hold at least one group; the 31 bits its groups decode to (the literal; all ones or all
            //   zeros for a fill, kernels.cu:337-354) are parked at s_cw[rank], and bit `offset` of a 1024-bit flag map is
            //   set ("a word starts at this group").  Nothing else is done per word -- no branch on what kind it is.
            // Step 2, output centric: lane t owns window t = groups 32 t .. 32 t + 31 = output words 31 t .. 31 t + 30
            //   (no word is shared between lanes).  The number of flags below the window is the rank of the word that
            //   covers its first group; walking the 32 flag bits, the lane steps to the next word wherever a flag is set
            //   and emits output word j = group j >> j | group j+1 << (31 - j) (mergeWords, kernels.cu:375) into the
            //   tile image.  Straight-line code, the same five or six instructions per group whatever the mix of fills and
            //   literals (walk_from); if all 32 windows lie inside fills the constants are written without the walk.
            // The image leaves through one TMA bulk store.
            const uint32_t skip = __shfl_sync(0xffffffffu, my_skip, s);
            const uint32_t w_beg = (uint32_t)ws_t & 3u;                   // words before ws are ignored
            const uint32_t w_span = nw_t - 1u - w_beg;                    // index of we among the tile's words
            bool image_done = false;
            if (nw_t <= 128u) {
                // A tile of zero fills and literals only (uniformly sparse data): no walk.  The image is cleared and every
                // literal ORs its 31 bits into the one or two output words they fall into; zero fills cost nothing.
                // (a one-fill among the up to three words of the first / last pack that belong to the neighbouring tiles
                //  sends the tile down the window path for nothing: not worth a range check per word)
                const bool one_fill = (xc[0] & xc[0] << 1 | xc[1] & xc[1] << 1 | xc[2] & xc[2] << 1 | xc[3] & xc[3] << 1) >> 31;
                if (!__any_sync(0xffffffffu, one_fill)) {
                    if (lane == 0) bulk_wait_read<0>();   // the previous tile's bulk store has read the image
                    __syncwarp();
                    uint4 *z = reinterpret_cast<uint4 *>(s_stage);
#pragma unroll
                    for (uint32_t i = 0; i < (TW / 4u + 31u) / 32u; i++)
                        if (lane + 32u * i < TW / 4u) z[lane + 32u * i] = make_uint4(0u, 0u, 0u, 0u);
                    uint32_t c[4], tsum = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t rel = 4u * lane + i - w_beg;
                        uint32_t v = word_groups(xc[i]);
                        if (rel > w_span) v = 0;
                        if (rel == 0u) v -= skip;
                        c[i] = v > EXP_CLAMP ? EXP_CLAMP : v;
                        tsum += c[i];
                    }
                    uint32_t off = warp_incl_scan(tsum) - tsum;
                    __syncwarp();   // the image is cleared
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (c[i] != 0u && off < tg && !is_fill(xc[i])) {   // kernels.cu:351-354, packed at once (kernels.cu:375)
                            const uint32_t b0 = 31u * off, w0 = b0 >> 5, sh = b0 & 31u;
                            atomicOr(s_stage + w0, xc[i] << sh);
                            if (sh > 1u) atomicOr(s_stage + w0 + 1u, xc[i] >> (32u - sh));
                        }
                        off += c[i];
                    }
                    image_done = true;
                }
            }
            if (!image_done) {
                s_flag[lane] = 0;
                if (lane == 0) s_flag[32] = 0;
                __syncwarp();
                // a tile of many words is literal dense: neighbouring lanes of the walk then read words about 32 ranks apart, and
                // the words are parked in rows of 32 padded to 33 (cw_pos) to keep those reads on different banks
                const bool pad = nw_t > 256u;
                uint32_t running = 0;   // group offset (tile relative) of the round's first word
                uint32_t rk_run = 0;    // words of earlier rounds that hold at least one group
                if (nw_t <= 32u) {
                    // word `lane`: component lane & 3 of lane (lane >> 2)'s pack
                    uint32_t x[1];
                    const uint32_t a = __shfl_sync(0xffffffffu, xc[0], lane >> 2), b = __shfl_sync(0xffffffffu, xc[1], lane >> 2);
                    const uint32_t c = __shfl_sync(0xffffffffu, xc[2], lane >> 2), d = __shfl_sync(0xffffffffu, xc[3], lane >> 2);
                    x[0] = (lane & 2u) ? ((lane & 1u) ? d : c) : ((lane & 1u) ? b : a);
                    park_words<1, false>(x, lane, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
                } else if (nw_t <= 64u) {
                    // words 2 lane, 2 lane + 1: a half of lane (lane >> 1)'s pack
                    uint32_t x[2];
                    const uint32_t a = __shfl_sync(0xffffffffu, xc[0], lane >> 1), b = __shfl_sync(0xffffffffu, xc[1], lane >> 1);
                    const uint32_t c = __shfl_sync(0xffffffffu, xc[2], lane >> 1), d = __shfl_sync(0xffffffffu, xc[3], lane >> 1);
                    x[0] = (lane & 1u) ? c : a;
                    x[1] = (lane & 1u) ? d : b;
                    park_words<2, false>(x, 2u * lane, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
                } else {
                    const uint32_t *src = p.in + (ws_t & ~3ull);              // 16-byte aligned start
                    const uint64_t room = p.c_words - (ws_t & ~3ull);         // words from there to the end of the stream
                    uint32_t xn[4] = {BIT31, BIT31, BIT31, BIT31};
#pragma unroll 1
                    for (uint32_t r0 = 4u * lane;; r0 += 128u) {   // my four consecutive words, relative to the aligned start
                        const bool more = r0 - 4u * lane + 128u < nw_t;
                        if (more && r0 + 128u < nw_t) load_words<4>(src, r0 + 128u, room, xn);   // the next round's words, a round ahead
                        if (pad)
                            park_words<4, true>(xc, r0, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
                        else
                            park_words<4, false>(xc, r0, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
#pragma unroll
                        for (int i = 0; i < 4; i++) xc[i] = xn[i];
                        if (!more || running >= tg) break;   // uniform: the tile is covered
                    }
                }
                if (lane == 0) {
                    if (tg < TG) {
                        // a short tile (the end of a column / of the stream): what lies behind it reads as a zero fill
                        const uint32_t re = rk_run <= TG ? rk_run : TG + 1u;
                        s_cw[pad ? cw_pos(re) : re] = 0u;
                        atomicOr(s_flag + (tg >> 5), 1u << (tg & 31u));
                    }
                    bulk_wait_read<0>();   // the previous tile's bulk store has read the image
                }
                __syncwarp();   // words parked, flags set, image free

                const uint32_t F = s_flag[lane];
                const uint32_t pc = __popc(F);
                uint32_t r = warp_incl_scan(pc) - pc + (F & 1u) - 1u;   // rank of the word that covers my first group
                DCHK(r <= TG + 1u, 2, r);
                uint32_t *o = s_stage + 31u * lane;
                uint32_t v = s_cw[pad ? cw_pos(r) : r];
                if (__any_sync(0xffffffffu, (F >> 1) != 0u)) {
                    if (!pad) {
                        walk_from<1>((uint32_t)__cvta_generic_to_shared(s_cw) + 4u * r, F, v, o);
                    } else {
#pragma unroll
                        for (int jj = 1; jj < 32; jj++) {
                            r += (F >> jj) & 1u;
                            const uint32_t nv = s_cw[cw_pos(r)];
                            o[jj - 1] = __funnelshift_r(v << 1, nv, jj);
                            v = nv;
                        }
                    }
                } else {
                    // every window lies inside one word (a fill, or the zeros behind a short tile)
                    const uint32_t v1 = v << 1;
#pragma unroll
                    for (int jj = 1; jj < 32; jj++) o[jj - 1] = __funnelshift_r(v1, v, jj);
                }
            }
            if (nout == TW) {
                fence_async_smem();   // my writes to the image, visible to the bulk copy engine
                __syncwarp();
                if (lane == 0) bulk_s2g(dst, stage_addr, TW * 4u, p.l2_stream_out != 0u);
            } else {
                // the last tile of a column or of the stream, or one cut short by the output capacity: the part that
                // exists, by hand
                __syncwarp();
                const uint4 *src4 = reinterpret_cast<const uint4 *>(s_stage);
                for (uint32_t i = lane; i < nvec; i += 32u) st_stream_v4(dst4 + i, src4[i]);
                for (uint32_t i = (nvec << 2) + lane; i < nout; i += 32u) dst[i] = s_stage[i];
                __syncwarp();
            }
        }

        // ---- shift the pipeline
        c0 = c1;
        e0x = e1x;
        e0y = e1y;
        first0 = first1;
        c1 = c2;
        e1x = e2x;
        e1y = e2y;
        c2 = it + 3u < (uint32_t)EXPAND_STATIC_ROUNDS || !tickets ? gw + (it + 3u) * GW
                                                                  : (uint32_t)EXPAND_STATIC_ROUNDS * GW + __shfl_sync(0xffffffffu, tk, 0);
        if (c2 < c1) c2 = n_chunks;   // (wrapped: behind the last chunk)
    }
    if (lane == 0) bulk_wait_read<0>();   // shared memory must outlive the bulk stores that read it
    __syncthreads();
    if (p.ctr != nullptr && tid == 0) {
        // The last CTA to leave reports the launch's status and zeroes the counters for the next launch -- all of them,
        // so that a launch that went wrong (a CTA gave up waiting) still leaves its slot clean.  (My own draws are
        // performed before that: the barrier and the fence order them before my `done`.)
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

// ---------------------------------------------------------------- logical operators on two compressed vectors
//
// result = compress(decode(a) op decode(b)) in BLOCK1024 mode without either vector ever being decoded into HBM
// (the specification is wah_oracle_logical, oracle/wah_oracle.c).  Both streams are scanned (the scan
// phase above, table entries for every tile); then a warp takes a tile of 1024 groups -- one block of the reference
// encoder (kernels.cu:256,273-280) -- at a time: a tile that lies inside one fill of each operand becomes one fill word
// without a group being looked at; otherwise the first operand's tile is expanded into an image in shared memory (the
// window path of the expand phase), the second operand's walk combines its output words into that image, and the image
// is encoded the way the compressor's warp encodes a block (classify, run-end rule, warp scan, compaction:
// kernels.cu:79,93-141,244-248).  The words of tile t go to slot t of a scratch array; a prefix sum over the tiles' word
// counts and a gather make the stream.  Traffic: the two streams, the result twice.

constexpr int LOG_IMG_WORDS = EXPAND_TILE_WORDS + 8;                         // the image + the word the last group's extraction touches
constexpr int LOG_WARP_SMEM_WORDS = CW_WORDS + FLAG_WORDS + LOG_IMG_WORDS;   // 8.4 KB per warp
constexpr uint32_t LOG_CONST0 = 1u, LOG_CONST1 = 2u;

struct LogicalParams {
    const uint32_t *a, *b;
    uint64_t ca, cb;
    const ulonglong2 *starts_a, *starts_b;   // one entry per tile (+ 1), written by the two scans
    const DecodeHeader *hdr_a, *hdr_b;
    uint32_t epoch_a, epoch_b;
    uint64_t groups;                          // groups of the vectors the streams stand for
    uint32_t n_tiles;
    int op;
    uint32_t *slots;                          // n_tiles x 1024 words
    uint32_t *counts;                         // n_tiles
    unsigned long long *partials;             // words of every 256 tiles (zero at launch)
};

// Expands the part of tile t (groups g_start .. g_start + tg) that stream `in` covers into `img`; what lies behind the
// stream's end reads as zeros.  Returns LOG_CONST0 / LOG_CONST1 if the whole tile is zeros / ones without an image.
template <int OP>
__device__ __forceinline__ uint32_t expand_operand(const uint32_t *in, uint64_t c_words, const ulonglong2 *starts, uint32_t epoch,
                                                   uint64_t G, uint32_t t, uint32_t tg_tile, uint32_t *s_cw, uint32_t *s_flag,
                                                   uint32_t *img, uint32_t lane)
{
    constexpr uint32_t TG = (uint32_t)EXPAND_TILE_GROUPS;
    const uint64_t g_start = (uint64_t)t << TG_SHIFT;
    if (c_words == 0 || g_start >= G) return LOG_CONST0;
    const uint32_t tg = G - g_start < (uint64_t)tg_tile ? (uint32_t)(G - g_start) : tg_tile;
    uint64_t sx, sy, nx, ny;
    load_entry(starts + t, epoch, sx, sy);
    load_entry(starts + t + 1, epoch, nx, ny);
    if (sx == 0ull) return LOG_CONST0;   // (cannot happen for g_start < G)
    const bool last = nx == 0ull;        // the stream ends in this tile
    const uint64_t ws = sx - 1ull, we = last ? c_words - 1ull : nx - 1ull;
    const uint32_t skip = (uint32_t)(g_start - sy);
    if (ws == we) {
        const uint32_t first = in[ws];
        if (is_fill(first)) {
            if (!(first & BIT30)) return LOG_CONST0;
            if (tg == tg_tile) return LOG_CONST1;
        }
    }
    const uint64_t wa = ws & ~3ull;
    const uint64_t span = we - wa + 1ull;
    const uint32_t nw = span > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)span;
    const uint32_t w_beg = (uint32_t)ws & 3u, w_span = nw - 1u - w_beg;
    const uint32_t *src = in + wa;
    const uint64_t room = c_words - wa;
    const bool pad = nw > 256u;
    s_flag[lane] = 0;
    if (lane == 0) s_flag[32] = 0;
    __syncwarp();
    uint32_t running = 0, rk_run = 0;
    if (nw <= 32u) {
        uint32_t x[1];
        load_words<1>(src, lane, room, x);
        park_words<1, false>(x, lane, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
    } else if (nw <= 64u) {
        uint32_t x[2];
        load_words<2>(src, 2u * lane, room, x);
        park_words<2, false>(x, 2u * lane, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
    } else {
        uint32_t xc[4], xn[4] = {BIT31, BIT31, BIT31, BIT31};
        load_words<4>(src, 4u * lane, room, xc);
#pragma unroll 1
        for (uint32_t r0 = 4u * lane;; r0 += 128u) {
            const bool more = r0 - 4u * lane + 128u < nw;
            if (more) load_words<4>(src, r0 + 128u, room, xn);
            if (pad)
                park_words<4, true>(xc, r0, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
            else
                park_words<4, false>(xc, r0, w_beg, w_span, skip, tg, running, rk_run, s_cw, s_flag);
#pragma unroll
            for (int i = 0; i < 4; i++) xc[i] = xn[i];
            if (!more || running >= tg) break;
        }
    }
    if (lane == 0 && tg < TG) {
        const uint32_t re = rk_run <= TG ? rk_run : TG + 1u;
        s_cw[pad ? cw_pos(re) : re] = 0u;
        atomicOr(s_flag + (tg >> 5), 1u << (tg & 31u));
    }
    __syncwarp();
    const uint32_t F = s_flag[lane];
    const uint32_t pc = __popc(F);
    uint32_t r = warp_incl_scan(pc) - pc + (F & 1u) - 1u;
    uint32_t *o = img + 31u * lane;
    uint32_t v = s_cw[pad ? cw_pos(r) : r];
    if (!pad) {
        walk_from<1, OP>((uint32_t)__cvta_generic_to_shared(s_cw) + 4u * r, F, v, o);
    } else {
#pragma unroll
        for (int jj = 1; jj < 32; jj++) {
            r += (F >> jj) & 1u;
            const uint32_t nv = s_cw[cw_pos(r)];
            const uint32_t w = __funnelshift_r(v << 1, nv, jj);
            if constexpr (OP < 0)
                o[jj - 1] = w;
            else
                o[jj - 1] = combine<OP>(o[jj - 1], w);
            v = nv;
        }
    }
    __syncwarp();   // image complete; s_cw and the flag map are free again
    return 0u;
}

__device__ __forceinline__ uint32_t apply_op(int op, uint32_t x, uint32_t y)
{
    return op == 0 ? (x & y) : (op == 1 ? (x | y) : (op == 2 ? (x ^ y) : (x & ~y)));
}

__global__ void __launch_bounds__(EXPAND_THREADS, 3) wah_logical_tiles_kernel(const LogicalParams p)
{
    constexpr uint32_t TG = (uint32_t)EXPAND_TILE_GROUPS, TW = (uint32_t)EXPAND_TILE_WORDS;
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *s_cw = smem + warp * LOG_WARP_SMEM_WORDS;
    uint32_t *s_flag = s_cw + CW_WORDS;
    uint32_t *img = s_flag + FLAG_WORDS;
    const uint64_t Ga = p.ca ? p.hdr_a->groups : 0ull, Gb = p.cb ? p.hdr_b->groups : 0ull;
    const uint32_t GW = gridDim.x * (EXPAND_THREADS / 32), gw = blockIdx.x * (EXPAND_THREADS / 32) + warp;
    if (lane < 8u) img[TW + lane] = 0;   // the word the extraction of a row's last group reads behind the image
    for (uint32_t t = gw; t < p.n_tiles; t += GW) {
        const uint64_t g_start = (uint64_t)t << TG_SHIFT;
        const uint32_t tg = p.groups - g_start < (uint64_t)TG ? (uint32_t)(p.groups - g_start) : TG;
        uint32_t *slot = p.slots + (uint64_t)t * TG;
        uint32_t *row_rw = img + 31u * lane;
        // ---- the first operand's tile into the image, the second combined into it by its walk (a constant operand by a
        //      loop over the lane's 31 words, unless the operator leaves the other operand as it is)
        const uint32_t ka = expand_operand<-1>(p.a, p.ca, p.starts_a, p.epoch_a, Ga, t, tg, s_cw, s_flag, img, lane);
        if (ka != 0u) {
            const uint32_t kb = expand_operand<-1>(p.b, p.cb, p.starts_b, p.epoch_b, Gb, t, tg, s_cw, s_flag, img, lane);
            if (kb != 0u) {
                // both operands constant over the tile: one fill word, no group looked at
                const uint32_t r = apply_op(p.op, ka == LOG_CONST1 ? 1u : 0u, kb == LOG_CONST1 ? 1u : 0u) & 1u;
                if (lane == 0) {
                    slot[0] = fill_word(r, tg);
                    p.counts[t] = 1u;
                    atomicAdd(p.partials + (t >> 8), 1ull);
                }
                continue;
            }
            const uint32_t fa = ka == LOG_CONST1 ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int k = 0; k < 31; k++) row_rw[k] = apply_op(p.op, fa, row_rw[k]);
        } else {
            uint32_t kb;
            switch (p.op) {
            case 0: kb = expand_operand<0>(p.b, p.cb, p.starts_b, p.epoch_b, Gb, t, tg, s_cw, s_flag, img, lane); break;
            case 1: kb = expand_operand<1>(p.b, p.cb, p.starts_b, p.epoch_b, Gb, t, tg, s_cw, s_flag, img, lane); break;
            case 2: kb = expand_operand<2>(p.b, p.cb, p.starts_b, p.epoch_b, Gb, t, tg, s_cw, s_flag, img, lane); break;
            default: kb = expand_operand<3>(p.b, p.cb, p.starts_b, p.epoch_b, Gb, t, tg, s_cw, s_flag, img, lane); break;
            }
            if (kb != 0u) {
                const uint32_t fb = kb == LOG_CONST1 ? 0xFFFFFFFFu : 0u;
                // (x op 0 = x for OR, XOR, ANDNOT; x AND ones = x)
                const bool same = fb == 0u ? p.op != 0 : p.op == 0;
                if (!same) {
#pragma unroll
                    for (int k = 0; k < 31; k++) row_rw[k] = apply_op(p.op, row_rw[k], fb);
                }
            }
        }
        __syncwarp();
        // ---- encode the block (the compressor's warp, BLOCK1024 mode)
        const uint32_t *row = img + 31u * lane;
        uint32_t nvalid = tg > 32u * lane ? tg - 32u * lane : 0u;
        if (nvalid > 32u) nvalid = 32u;
        const uint32_t vmask = nvalid == 32u ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);
        uint32_t Z = 0, O = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t g = extract_group(row, j);
            Z |= (g == 0u ? 1u : 0u) << j;
            O |= (g == ONES31 ? 1u : 0u) << j;
        }
        Z &= vmask;
        O &= vmask;
        const uint32_t F = Z | O;
        // type of the group after my 32 (the next lane's first); the block's last group has no successor
        const uint32_t nzb = __shfl_down_sync(0xffffffffu, Z & 1u, 1), nob = __shfl_down_sync(0xffffffffu, O & 1u, 1);
        const uint32_t nz = (lane != 31u && nzb) ? BIT31 : 0u, no = (lane != 31u && nob) ? BIT31 : 0u;
        // tail = literal, or fill whose successor differs (run-end rule, kernels.cu:126-141)
        const uint32_t T = ((~F) & vmask) | (Z & ~((Z >> 1) | nz)) | (O & ~((O >> 1) | no));
        const uint32_t cnt = __popc(T);
        const uint32_t my_open = T ? (uint32_t)__clz(T) : 32u;   // groups after my last tail
        const uint32_t incl = warp_incl_scan(cnt);
        const uint32_t tb = __ballot_sync(0xffffffffu, T != 0u);
        const uint32_t below = tb & lanemask_lt();
        const uint32_t qb = below ? 31u - (uint32_t)__clz(below) : 0u;
        const uint32_t open_q = __shfl_sync(0xffffffffu, my_open, qb);
        const uint32_t prev_open = below ? open_q + 32u * (lane - qb - 1u) : 32u * lane;   // run open where my groups start
        const uint32_t off = incl - cnt;
        const uint32_t wcnt = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t m = T & ~F;   // literals: the group itself (kernels.cu:107-112,256)
        while (m) {
            const uint32_t j = 31u - (uint32_t)__clz(m);
            m ^= 1u << j;
            s_cw[off + __popc(T & ((1u << j) - 1u))] = extract_group(row, j);
        }
        m = T & F;             // fills: BIT31 | type << 30 | length (kernels.cu:244-248)
        while (m) {
            const uint32_t j = 31u - (uint32_t)__clz(m);
            m ^= 1u << j;
            const uint32_t lower = T & ((1u << j) - 1u);
            const uint32_t len = lower ? j - (31u - (uint32_t)__clz(lower)) : j + 1u + prev_open;
            s_cw[off + __popc(lower)] = fill_word((O >> j) & 1u, len);
        }
        __syncwarp();
        for (uint32_t i = lane; i < wcnt; i += 32u) slot[i] = s_cw[i];
        if (lane == 0) {
            p.counts[t] = wcnt;
            atomicAdd(p.partials + (t >> 8), (unsigned long long)wcnt);
        }
        __syncwarp();
    }
}

// exclusive prefix over the words of every 256 tiles, in place (one CTA: a 16 Gbit vector has 2114 of them); total into *total
__global__ void __launch_bounds__(1024) wah_logical_offsets_kernel(unsigned long long *partials, uint32_t n, uint64_t *total)
{
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_base;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 1024u) {
        const uint64_t mine = i0 + t < n ? partials[i0 + t] : 0ull;
        const uint64_t incl = warp_incl_scan_u64(mine);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint64_t before = s_base + incl - mine, all = 0;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const uint64_t v = s_w[k];
            if (k < (int)warp) before += v;
            all += v;
        }
        if (i0 + t < n) partials[i0 + t] = before;
        __syncthreads();
        if (t == 0) s_base += all;
        __syncthreads();
    }
    if (t == 0) *total = s_base;
}

// the stream: CTA b takes tiles 256 b .. 256 b + 255 -- their offsets by a scan of their counts on top of the group's --
// and copies tile t's words from slot t to their place, a warp per tile
__global__ void __launch_bounds__(256) wah_logical_gather_kernel(const uint32_t *slots, const uint32_t *counts, const unsigned long long *partials,
                                                                 uint32_t n_tiles, uint32_t *out, uint64_t out_cap)
{
    __shared__ uint32_t s_w[8];
    __shared__ uint64_t s_off[256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t t0 = blockIdx.x * 256u;
    const uint32_t mine = t0 + tid < n_tiles ? counts[t0 + tid] : 0u;
    const uint32_t incl = warp_incl_scan(mine);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint64_t before = partials[blockIdx.x] + (incl - mine);
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (k < (int)warp) before += s_w[k];
    s_off[tid] = before;
    __syncthreads();
    // (blockIdx.y: which eighth of the group's tiles this CTA copies -- every CTA of a group does the group's scan)
    for (uint32_t k = 32u * blockIdx.y + warp; k < 32u * blockIdx.y + 32u && t0 + k < n_tiles; k += 8u) {
        const uint32_t c = counts[t0 + k];
        const uint64_t o = s_off[k];
        const uint32_t *s = slots + (uint64_t)(t0 + k) * EXPAND_TILE_GROUPS;
        for (uint32_t i = lane; i < c; i += 32u)
            if (o + i < out_cap) out[o + i] = s[i];
    }
}

// Both phases in one persistent launch: every CTA first takes its share of the scan tiles, then its share of the
// output tiles, each of which waits only for its own two `starts` entries.  Saves a launch, the idle tail / ramp
// between two kernels, and the wait for the slowest scan tile.
static_assert(2 * SCAN_SUB_WORDS <= (EXPAND_THREADS / 32) * WARP_SMEM_WORDS, "the scan phase's two sub-tile buffers live in the expand phase's shared memory");
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
    return (size_t)(EXPAND_THREADS / 32) * WARP_SMEM_WORDS * sizeof(uint32_t);
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
    // every warp should get several chunks (the first EXPAND_STATIC_ROUNDS are dealt round robin): 8 tiles per chunk for
    // long streams, fewer for short ones
    const uint64_t tiles = ep.max_out_tiles * (uint64_t)(ep.n_cols ? ep.n_cols : 1u), warps = (uint64_t)grid * (EXPAND_THREADS / 32);
    b.chunk_tiles = tiles >= 64ull * warps ? 8u : (tiles >= 24ull * warps ? 4u : 2u);   // (measured: 1 never pays -- the per-chunk work is not amortised)
    static const int forced = [] {
        const char *e = getenv("WAH_B200_CHUNK_TILES");   // (experiments)
        const int v = e ? atoi(e) : 0;
        return v == 1 || v == 2 || v == 4 || v == 8 ? v : 0;
    }();
    if (forced) b.chunk_tiles = (uint32_t)forced;
    a.chunk_tiles = b.chunk_tiles;   // (the scan phase records the entries of long fills at chunk starts only)
    // The compressed words are read by pass 1, by pass 2 and by the expand phase, whose output stores -- 16 to 1000 times
    // the volume -- flow through the same L2: the scan's loads ask the L2 to keep the stream, the output of the fill and
    // window paths is stored evict-first.  (Measured at 16 Gbit: 0.367 -> 0.336 ms at d = 0.0001, 0.381 -> 0.344 ms at
    // d = 0.01, 0.425 -> 0.403 ms at d = 0.1, 0.531 -> 0.524 ms at d = 0.5, whose 133 MB stream does not fit; keeping only
    // half or a quarter of such a stream was no better.  The same hints on the unit path's stores, on the compressor's
    // output and on its input loads measured as no gain and are not there.)
    a.l2_keep = l2_hints() ? 1u : 0u;
    b.l2_stream_out = a.l2_keep;
    void *args[] = {&a, &b};
    return launch_pdl((const void *)wah_decode_kernel, grid, EXPAND_THREADS, args, expand_smem_bytes(), stream);
}

cudaError_t launch_logical_compressed(const LogicalJob &job, cudaStream_t stream)
{
    constexpr size_t smem = (size_t)(EXPAND_THREADS / 32) * LOG_WARP_SMEM_WORDS * sizeof(uint32_t);
    static int grid_of[MAX_DEVICES] = {};
    int dev = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    if (grid_of[dev] == 0) {
        e = cudaFuncSetAttribute(wah_logical_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wah_logical_tiles_kernel, EXPAND_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        grid_of[dev] = sms * per_sm;
    }
    if (job.n_tiles == 0) return cudaMemsetAsync(job.total, 0, sizeof(uint64_t), stream);
    LogicalParams p;
    p.a = job.a;
    p.b = job.b;
    p.ca = job.ca;
    p.cb = job.cb;
    p.starts_a = job.starts_a;
    p.starts_b = job.starts_b;
    p.hdr_a = job.hdr_a;
    p.hdr_b = job.hdr_b;
    p.epoch_a = job.epoch_a;
    p.epoch_b = job.epoch_b;
    p.groups = job.groups;
    p.n_tiles = job.n_tiles;
    p.op = job.op;
    p.slots = job.slots;
    p.counts = job.counts;
    p.partials = reinterpret_cast<unsigned long long *>(job.offsets);
    const uint32_t n_groups = (job.n_tiles + 255u) / 256u;   // groups of 256 tiles
    e = cudaMemsetAsync(job.offsets, 0, (size_t)n_groups * 8, stream);
    if (e != cudaSuccess) return e;
    const uint32_t want = (job.n_tiles + 7u) / 8u;
    const int grid = (int)(want < (uint32_t)grid_of[dev] ? want : (uint32_t)grid_of[dev]);
    wah_logical_tiles_kernel<<<grid, EXPAND_THREADS, smem, stream>>>(p);
    wah_logical_offsets_kernel<<<1, 1024, 0, stream>>>(p.partials, n_groups, job.total);
    wah_logical_gather_kernel<<<dim3(n_groups, 8), 256, 0, stream>>>(job.slots, job.counts, p.partials, job.n_tiles, job.out, job.out_cap);
    return cudaGetLastError();
}

}  // namespace wahb200

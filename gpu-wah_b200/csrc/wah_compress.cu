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
//   tile    = NWORK reference blocks, one per WORKER warp
//   CTA     = persistent (every CTA of the grid resident: SMs x occupancy, or a cooperative launch) and warp
//             specialised:
//               NWORK worker warps   classify a tile from shared memory, publish its aggregate (one 64-bit
//                                    descriptor per tile) to the other CTAs, compact its words into a ring
//                                    in shared memory -- or, for a tile of more words than the ring takes,
//                                    write them once the offset is known
//               2 control warps      taking the CTA's tiles in turn: tile offset = the CTA's previous tile's
//                                    end + the descriptors of the tiles in between (a chained sum: no warp
//                                    ever waits for another CTA's control warp)
//               1 writer warp        copies a tile's words from the ring to their place in the output
//               1 producer warp      cp.async.bulk (TMA) of tile blockIdx + k * gridDim into a ring of
//                                    STAGES shared-memory buffers, mbarrier signalled
//             The roles talk through mbarriers only; there is no __syncthreads in the loop.  The workers run
//             up to QDEPTH tiles ahead of the control warps and the writer, so the latency of a tile's offset
//             is covered by the classification of the tiles behind it.
//
// A run is emitted where it ENDS ("tail"): group k is a tail if it is a literal,
// or the next group has another type, or it is the last group of the stream /
// block.  The k-th tail is the k-th output word; a fill's length is the distance
// to the previous tail, which flows forward through the scans.
#include "wah_common.cuh"
#include "wah_kernels.h"

#include <cstdlib>

namespace wahb200 {

namespace {

// ---- tile descriptor: epoch:32 | (unused) | no_tail:1 | open:14 | count:14 ----------------------------
//   count    words the tile emits (0 .. 8192)
//   no_tail  the tile holds no run end: its groups all belong to one run that is still open
//   open     groups after the tile's last run end (the whole tile, 8192, if no_tail)
//   epoch    the launch that wrote it.  Every launch gets a fresh number from the host, so whatever an earlier
//            launch (or nobody) left in the workspace reads as "not published yet": no memset per call.
// Written once, by the tile's workers, as soon as the tile is classified.

__device__ __forceinline__ uint64_t desc_pack(uint32_t epoch, uint32_t no_tail, uint32_t open, uint32_t count)
{
    return ((uint64_t)epoch << 32) | (no_tail << 28) | (open << 14) | count;
}
__device__ __forceinline__ bool desc_empty(uint64_t d, uint32_t epoch) { return (uint32_t)(d >> 32) != epoch; }
__device__ __forceinline__ uint32_t desc_no_tail(uint64_t d) { return ((uint32_t)d >> 28) & 1u; }
__device__ __forceinline__ uint32_t desc_open(uint64_t d) { return ((uint32_t)d >> 14) & 0x3FFFu; }
__device__ __forceinline__ uint32_t desc_count(uint64_t d) { return (uint32_t)d & 0x3FFFu; }

// ---- mbarrier / bulk-copy primitives (PTX) -----------------------------------------

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// The same for a warp that is in no hurry (producer: three stages ahead; writer: behind by design): it sleeps between
// tries instead of re-issuing try_wait back to back on a scheduler the worker warps need every issue slot of.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok) __nanosleep(128);
    } while (!ok);
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- optional phase trace (compile with -DWAH_TRACE; p.trace = [n_ctas][iters][8] u64) ----------
#ifdef WAH_TRACE
#define TRACE(i, slot, val)                                                                                  \
    do {                                                                                                     \
        if (p.trace && (i) < 64u) p.trace[((uint64_t)blockIdx.x * 64u + (i)) * 8u + (slot)] = (uint64_t)(val); \
    } while (0)
__device__ __forceinline__ uint64_t gtime()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else
#define TRACE(i, slot, val) \
    do {                    \
    } while (0)
#endif

// ---- kernel geometry ------------------------------------------------------------------

constexpr int WARP_RING = 512;    // staged output words per worker warp (the CTA's ring is NWORK times this)
#ifndef WAH_QDEPTH
#define WAH_QDEPTH 8
#endif
constexpr int QDEPTH = WAH_QDEPTH;   // tiles a CTA may have between classification and copy-out
constexpr uint32_t MODE_RING = 0, MODE_DIRECT = 1;

template <int NWORK, int STAGES>
struct Geom {
    static constexpr int TILE_WORDS = NWORK * 992;
    static constexpr int TILE_GROUPS = NWORK * 1024;
    static constexpr int PAD_FRONT = 4;                       // row[-1] of the tile's first thread
    static constexpr int STAGE_WORDS = TILE_WORDS + 8;        // + 4 in front, + look-ahead word (16 B) behind
    static constexpr int THREADS = (NWORK + 4) * 32;          // workers, two control warps, producer, writer
    static constexpr int RING_WORDS = NWORK * WARP_RING;      // power of two for NWORK = 4, 8
    static_assert((RING_WORDS & (RING_WORDS - 1)) == 0, "ring size must be a power of two");
    static_assert(TILE_GROUPS <= 8192, "descriptor fields are 14 bits");
    static_assert(TILE_WORDS == COMPRESS_TILE_WORDS, "the C ABI sizes its descriptor array from COMPRESS_TILE_WORDS");
};

// producer -> workers, one per input stage
struct StageInfo {
    uint32_t gvalid;     // groups of the tile that exist (0 .. TILE_GROUPS)
    uint32_t has_next;   // a group of the same column follows the tile (CANONICAL look-ahead)
    uint32_t tile;       // global tile index
    uint32_t pad;
};

// warp aggregates of a tile, exchanged at the workers' barrier and read by the control / writer warps
// (slot = CTA-local tile index % QDEPTH)
template <int NWORK>
struct WarpAgg {
    uint32_t wcnt[NWORK];      // words emitted by each warp
    uint32_t wopen[NWORK];     // groups after the warp's last tail (1024 if it has none)
    uint32_t whas[NWORK];      // warp has at least one tail
};

// one per tile in flight between the workers and the control warps (slot = CTA-local tile index % QDEPTH)
template <int NWORK>
struct TileMeta {
    // workers -> control
    uint32_t tile_cnt, tile_open, tile_has;
    uint32_t ring_base;        // ring position of the tile's first staged word
    uint32_t mode;             // MODE_RING: words staged in the ring, MODE_DIRECT: the workers write them
    // control -> workers in MODE_DIRECT
    int32_t lead_adjust;       // added to the launch's very first word (seam with an earlier launch)
    uint64_t dst;              // output word index of the tile's first word
    uint32_t wcarry[NWORK];    // groups of a run still open where the warp starts (CANONICAL)
    // control warp of tile i -> control warp of tile i + 1 (the CTA's chain)
    uint32_t cnt_end;          // words of this launch up to the end of the tile
    uint32_t open_end;         // groups of the run still open at the end of the tile
    uint32_t failed;           // a control warp of this CTA has given up waiting for descriptors
};

template <int NWORK, int STAGES>
struct Smem {
    using G = Geom<NWORK, STAGES>;
    uint32_t stage[STAGES][G::STAGE_WORDS];
    uint32_t ring[G::RING_WORDS];
    TileMeta<NWORK> meta[QDEPTH];
    WarpAgg<NWORK> wagg[QDEPTH];
    StageInfo info[STAGES];
    uint64_t full[STAGES], empty[STAGES];              // input ring: producer <-> workers
    uint64_t agg[QDEPTH], pref[QDEPTH], done[QDEPTH];  // tile queue: workers <-> control
    uint64_t chain[QDEPTH];                            // control warp of a tile -> control warp of the next tile
};

// literal group j (0..31) of the row that starts at `row`: stream bits [31 j, 31 j + 31), LSB first
// (kernels.cu:79).  Reads row[j-1] and row[j]; for j = 0 the clamped shift by 32 drops row[-1].
__device__ __forceinline__ uint32_t row_group(const uint32_t *row, uint32_t j)
{
    const uint32_t *q = row + j;
    return __funnelshift_rc(q[-1], q[0], 32u - j) & ONES31;
}

// Set bit `bit` of Z / O if the group held in the top 31 bits of u is all zeros / all ones.  The bit is
// added with a predicated multiply-add (`one` is a register holding 1): that is an IMAD, issued to the FMA
// pipe, which leaves the ALU pipe -- the kernel's bottleneck -- the funnel shift and the two compares.
__device__ __forceinline__ void classify_group_rt(uint32_t u, uint32_t one, uint32_t bit, uint32_t &Z, uint32_t &O)
{
    asm("{\n\t"
        ".reg .pred pz, po;\n\t"
        "setp.lt.u32 pz, %2, 2;\n\t"
        "setp.gt.u32 po, %2, 0xFFFFFFFD;\n\t"
        "@pz mad.lo.u32 %0, %3, %4, %0;\n\t"
        "@po mad.lo.u32 %1, %3, %4, %1;\n\t"
        "}"
        : "+r"(Z), "+r"(O)
        : "r"(u), "r"(one), "r"(bit));
}
template <int J>
__device__ __forceinline__ void classify_group(uint32_t u, uint32_t one, uint32_t &Z, uint32_t &O)
{
    classify_group_rt(u, one, 1u << J, Z, O);
}

template <int NWORK, int STAGES, bool BLOCK_MODE>
__global__ void __launch_bounds__((NWORK + 4) * 32, 2) wah_compress_kernel(const CompressParams p)
{
    using G = Geom<NWORK, STAGES>;
    using SM = Smem<NWORK, STAGES>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t stride = gridDim.x;
    const uint32_t n_my = blockIdx.x < p.n_tiles ? (p.n_tiles - blockIdx.x + stride - 1u) / stride : 0u;

    pdl_launch_dependents();   // the next kernel on the stream may take the SMs as this one leaves them
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), NWORK);
        }
        for (int q = 0; q < QDEPTH; q++) {
            mbar_init(smem_u32(&sm.agg[q]), NWORK);
            mbar_init(smem_u32(&sm.pref[q]), 1);
            mbar_init(smem_u32(&sm.done[q]), 1);
            mbar_init(smem_u32(&sm.chain[q]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();   // the previous kernel on the stream is complete: its output (our input, the workspace) is visible
    __syncthreads();

    if (warp == NWORK + 1) {
        // =========================================================== producer warp
        for (uint32_t i = 0; i < n_my; i++) {
            const uint32_t s = i % STAGES, use = i / STAGES;
            const uint32_t tile = blockIdx.x + i * stride;
            const uint32_t col = tile / p.tiles_per_col;
            const uint32_t t = tile - col * p.tiles_per_col;
            const uint64_t w0 = (uint64_t)t * G::TILE_WORDS;
            const uint32_t *src = p.in + (uint64_t)col * p.col_stride + w0;
            const uint64_t left = p.n_words - w0;                  // words from the tile start to the column end
            const uint64_t g0 = (uint64_t)t * G::TILE_GROUPS;
            const uint64_t gleft = p.groups - g0;
            uint32_t *buf = &sm.stage[s][G::PAD_FRONT];
            const uint32_t bar = smem_u32(&sm.full[s]);
            // words to stage: the tile and, if the column goes on, one look-ahead word
            const uint32_t nload = left > (uint64_t)G::TILE_WORDS ? (uint32_t)G::TILE_WORDS + 1u : (uint32_t)left;
            const uint32_t gvalid = gleft >= (uint64_t)G::TILE_GROUPS ? (uint32_t)G::TILE_GROUPS : (uint32_t)gleft;
            const uint32_t has_next = gleft > (uint64_t)G::TILE_GROUPS ? 1u : 0u;
            const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
            // whole 16-byte units by TMA, the ragged end (and zero words behind it) by hand
            const uint32_t nbulk = left >= (uint64_t)G::TILE_WORDS + 4u ? (uint32_t)G::TILE_WORDS + 4u : (nload & ~3u);
            // Everything above is worked out before the wait, and the wait does not sleep: the stage is refilled as soon as
            // the last worker lets go of it.  (scripts/micro/read_patterns.cu: the read rate of this staging scheme falls
            // from 7.1 to 5.8 TB/s when a stage is held 2700 instead of 2100 cycles, and the workers need 2200.)
            if (use > 0) mbar_wait(smem_u32(&sm.empty[s]), (use - 1u) & 1u);
            if (lane == 0) {
                sm.info[s].gvalid = gvalid;
                sm.info[s].has_next = has_next;
                sm.info[s].tile = tile;
            }
            if (aligned) {
                if (nbulk < nload + 2u && lane < 8u) {
                    const uint32_t i0 = nbulk + lane;
                    if (i0 < (uint32_t)G::TILE_WORDS + 4u) buf[i0] = i0 < nload ? ld_stream_u32(src + i0) : 0u;
                }
                __syncwarp();
                if (lane == 0) {
                    TRACE(i, 0, clock64());
                    if (nbulk) {
                        mbar_arrive_expect_tx(bar, nbulk * 4u);
                        bulk_g2s(smem_u32(buf), src, nbulk * 4u, bar);
                    } else {
                        mbar_arrive(bar);
                    }
                }
            } else {
                // column start not 16-byte aligned: plain staging by the producer warp
                for (uint32_t i0 = lane; i0 < (uint32_t)G::TILE_WORDS + 4u; i0 += 32u)
                    buf[i0] = i0 < nload ? ld_stream_u32(src + i0) : 0u;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);
            }
        }
    } else if (warp == NWORK || warp == NWORK + 3) {
        // ============================================================ control warps
        // Two of them, taking the CTA's tiles in turn.  (Phase trace of the single control warp: 1.5 us per tile in
        // BLOCK1024 mode, 2.0 us in CANONICAL mode, with no descriptor missing -- the workers need 1.1 us and ran the
        // full QDEPTH tiles ahead: that warp set the pace of the CTA.)  The sum over the other CTAs' descriptors, which
        // is what takes the time, does not depend on this CTA's own running totals; those pass from the warp of tile i
        // to the warp of tile i + 1 through shared memory (sm.chain), two additions per tile.
        const uint32_t cw = warp == NWORK ? 0u : 1u;
        // Offsets by a chained sum instead of a status-polling look-back: this CTA owns tiles b, b + G,
        // b + 2G, ... (G = gridDim), so the words before tile k are the words before its previous tile
        // k - G, plus that tile's, plus the aggregates of the G - 1 tiles in between.  Those are published
        // by the workers of the other CTAs (which run ahead of their control warps), so a control warp
        // never waits for another control warp.
        constexpr int LBN = 10;   // descriptors per lane and round (32 * LBN = 320 >= 2 CTAs x 148 SMs)
        const uint64_t base = p.base_in ? *p.base_in : 0ull;
        const int32_t lead_adjust = p.lead_adjust ? *p.lead_adjust : 0;
        // Descriptor fetches this warp may still spend waiting for other CTAs (about 2 s).  Every CTA of the grid has
        // to be resident for the aggregates to appear; if that ever fails the launch ends with *total_out = ~0
        // (WAH_ERR_CUDA at the host entry points) instead of hanging the GPU.
        uint32_t budget = COMPRESS_SPIN_LIMIT;
        bool failed = base == ~0ull;   // an earlier launch of a chained stream failed
        uint64_t d[LBN];               // first window of my next tile, requested while the other warp's tile is at hand
        auto request = [&](int64_t hi, int64_t lo) {
#pragma unroll
            for (int r = 0; r < LBN; r++) {
                const int64_t lk = hi - (int64_t)lane - 32 * r;
                d[r] = lk >= lo ? ld_relaxed_u64(p.desc + lk) : ~((uint64_t)p.epoch << 32);
            }
        };
        if (cw == 0u && n_my > 0) request((int64_t)blockIdx.x - 1, 0);
        if (cw == 1u && n_my > 1) request((int64_t)(blockIdx.x + stride) - 1, (int64_t)blockIdx.x + 1);
        for (uint32_t i = cw; i < n_my; i += 2u) {
            const uint32_t q = i % QDEPTH, use = i / QDEPTH;
            const uint32_t tile = blockIdx.x + i * stride;
            const uint32_t col = tile / p.tiles_per_col;
            const uint32_t t = tile - col * p.tiles_per_col;
            TileMeta<NWORK> &mt = sm.meta[q];

            // my own CTA's workers first: the other CTAs' run at the same pace, so their descriptors are (nearly
            // always) there by now and the sum below does not poll the L2 for tiles that are still being classified
            mbar_wait(smem_u32(&sm.agg[q]), use & 1u);
            if (lane == 0) TRACE(i, 3, clock64());

            // ---- sum the aggregates of tiles [lo, hi], nearest first
            uint32_t between = 0, open_between = 0;   // words of / run open across the tiles between my CTA's previous tile and this one
            bool open_done;
            {
                const int64_t lo = i == 0 ? 0 : (int64_t)tile - (int64_t)stride + 1;
                int64_t hi = (int64_t)tile - 1;
                open_done = BLOCK_MODE || (t == 0u);   // BLOCK mode never carries; CANONICAL restarts per column
                uint32_t csum = 0, osum = 0;
                bool fresh = true;
#ifdef WAH_TRACE
                uint32_t polls = 0;
#endif
                while (hi >= lo) {
                    if (!fresh) request(hi, lo);
                    fresh = false;
#pragma unroll
                    for (int r = 0; r < LBN; r++) {
                        const int64_t lk = hi - (int64_t)lane - 32 * r;
                        const bool in = lk >= lo;
                        while (__any_sync(0xffffffffu, in && desc_empty(d[r], p.epoch))) {
                            if (budget == 0u) {   // (uniform) give up: the tile counts as empty, the launch as failed
                                if (in && desc_empty(d[r], p.epoch)) d[r] = (uint64_t)p.epoch << 32;
                                failed = true;
                                break;
                            }
                            budget--;
                            if (in && desc_empty(d[r], p.epoch)) d[r] = ld_relaxed_u64(p.desc + lk);
#ifdef WAH_TRACE
                            polls++;
#endif
                        }
                        if (in) csum += desc_count(d[r]);
                        if (!open_done) {
                            // towards older tiles until one ends a run
                            const uint32_t term_mask = __ballot_sync(0xffffffffu, in && !desc_no_tail(d[r]));
                            const uint32_t first_term = term_mask ? (uint32_t)__ffs(term_mask) - 1u : 32u;
                            if (in && lane <= first_term) osum += desc_open(d[r]);
                            open_done = first_term < 32u;
                        }
                    }
                    hi -= 32 * LBN;
                }
#ifdef WAH_TRACE
                if (lane == 0) TRACE(i, 7, polls);
#endif
                between = warp_sum(csum);
                if (!BLOCK_MODE) open_between = warp_sum(osum);
            }
            // request my next tile's window now: it is in flight for the whole of the other warp's tile
            if (i + 2u < n_my) request((int64_t)(tile + 2u * stride) - 1, (int64_t)(tile + stride) + 1);

            // ---- the CTA's chain: where its previous tile ended (the other control warp's tile)
            uint32_t own_cnt = 0, own_open = 0;
            if (i > 0u) {
                const uint32_t qp = (i - 1u) % QDEPTH;
                mbar_wait(smem_u32(&sm.chain[qp]), ((i - 1u) / QDEPTH) & 1u);
                own_cnt = sm.meta[qp].cnt_end;
                own_open = sm.meta[qp].open_end;
                failed = failed || sm.meta[qp].failed != 0u;
            }
            const uint32_t excl = own_cnt + between;
            // no run end between my CTA's previous tile and this one: the run open at its end goes on
            uint32_t carry = open_between + (open_done ? 0u : own_open);
            if (BLOCK_MODE || t == 0u) carry = 0;

            // ---- tile aggregate (computed and published to the other CTAs by the workers)
            const uint32_t tile_cnt = mt.tile_cnt, tile_open = mt.tile_open, tile_has = mt.tile_has;
            // the CTA's running totals, for its next tile
            if (lane == 0) {
                mt.cnt_end = excl + tile_cnt;
                mt.open_end = tile_has ? tile_open : carry + tile_open;
                mt.failed = failed ? 1u : 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.chain[q]));

            // ---- hand the offsets to the writer warp (and to the workers, if they write this tile themselves)
            const uint64_t dst0 = base + excl;
            // run still open where each warp starts (CANONICAL): lane w looks at the warps below it
            uint32_t run = carry;
            if (!BLOCK_MODE) {
#pragma unroll
                for (int w = 0; w < NWORK - 1; w++) {
                    const uint32_t h = sm.wagg[q].whas[w], o = sm.wagg[q].wopen[w];
                    if ((int)lane > w) run = h ? o : run + o;
                }
            }
            if (lane < NWORK) mt.wcarry[lane] = run;
            if (lane == 0) {
                mt.dst = dst0;
                mt.lead_adjust = (excl == 0u) ? lead_adjust : 0;   // no word of this launch precedes the tile
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&sm.pref[q]));
                TRACE(i, 4, clock64());
            }
            if (lane == 0) {
                if (p.col_offsets && t == 0u) p.col_offsets[col] = dst0;
                if (tile == p.n_tiles - 1u) {
                    // (the last tile's offset depends on every other CTA's aggregates: it is the one that notices)
                    const uint64_t total = failed ? ~0ull : dst0 + tile_cnt;
                    *p.total_out = total;
                    if (p.col_offsets) p.col_offsets[p.n_cols] = total;
                }
            }
        }
    } else if (warp == NWORK + 2) {
        // ============================================================= writer warp
        // copies the words the workers staged in the ring to their place in the output, coalesced
        for (uint32_t i = 0; i < n_my; i++) {
            const uint32_t q = i % QDEPTH;
            TileMeta<NWORK> &mt = sm.meta[q];
            mbar_wait_relaxed(smem_u32(&sm.pref[q]), (i / QDEPTH) & 1u);
            if (mt.mode == MODE_RING) {
                const uint32_t tile_cnt = mt.tile_cnt;
                const uint64_t dst0 = mt.dst;
                const uint32_t rb = mt.ring_base;
                // the first word of each warp may close a run that started before the warp
                {
                    const uint32_t a_cnt = lane < NWORK ? sm.wagg[q].wcnt[lane] : 0u;
                    uint32_t a_scan = a_cnt;
#pragma unroll
                    for (int dd = 1; dd < NWORK; dd <<= 1) {
                        const uint32_t o = __shfl_up_sync(0xffffffffu, a_scan, dd);
                        if ((int)lane >= dd) a_scan += o;
                    }
                    const uint32_t pre = a_scan - a_cnt;
                    uint32_t add = (BLOCK_MODE || lane >= NWORK) ? 0u : mt.wcarry[lane < NWORK ? lane : 0];
                    if (pre == 0u) add += (uint32_t)mt.lead_adjust;   // the launch's first word
                    if (a_cnt != 0u && add != 0u) sm.ring[(rb + pre) & (G::RING_WORDS - 1)] += add;
                }
                __syncwarp();
                const uint32_t room =
                    dst0 >= p.out_cap ? 0u
                                      : (p.out_cap - dst0 >= (uint64_t)G::TILE_GROUPS ? (uint32_t)G::TILE_GROUPS
                                                                                       : (uint32_t)(p.out_cap - dst0));
                const uint32_t ncopy = tile_cnt < room ? tile_cnt : room;
                uint32_t *dst = p.out + dst0;
#pragma unroll 4
                for (uint32_t k = lane; k < ncopy; k += 32u)
                    st_stream_u32(dst + k, sm.ring[(rb + k) & (G::RING_WORDS - 1)]);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&sm.done[q]));
                TRACE(i, 5, clock64());
            }
        }
    } else {
        // ============================================================= worker warps
        const uint32_t one = p.one;   // 1, opaque to the compiler (see classify_group_rt)
        uint32_t head = 0;   // words this CTA has staged so far (ring position, monotonic, same in every warp)
        uint32_t jd = 0;     // oldest tile of this CTA not yet known to be copied out
        // An all-literal warp of a dense tile writes its 1024 words straight from the input stage once the tile's
        // offset is known.  It does not wait for that: the write is deferred until the NEXT tile has been
        // classified, so the offset's latency is covered (nothing but the stage has to be kept for it).
        bool pend = false;
        uint32_t pend_s = 0, pend_q = 0, pend_par = 0;
        auto flush_pending = [&]() {
            if (!pend) return;
            mbar_wait(smem_u32(&sm.pref[pend_q]), pend_par);
            const uint64_t dstw = sm.meta[pend_q].dst + 1024u * warp;   // every warp of such a tile emits 1024 words... (checked below)
            const uint32_t room =
                dstw >= p.out_cap ? 0u : (p.out_cap - dstw >= 1024ull ? 1024u : (uint32_t)(p.out_cap - dstw));
            uint32_t *dst = p.out + dstw;
            const uint32_t *wrow = &sm.stage[pend_s][G::PAD_FRONT] + 992u * warp;
#pragma unroll 8
            for (uint32_t k = 0; k < 32u; k++) {
                const uint32_t g = 32u * k + lane;
                if (g < room) st_stream_u32(dst + g, extract_group(wrow, g));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.empty[pend_s]));
            pend = false;
        };

        for (uint32_t i = 0; i < n_my; i++) {
            const uint32_t s = i % STAGES, q = i % QDEPTH;
            TileMeta<NWORK> &mt = sm.meta[q];
            // the queue slot is free again once tile i - QDEPTH has been copied out
            while (jd + QDEPTH <= i) {
                mbar_wait(smem_u32(&sm.done[jd % QDEPTH]), (jd / QDEPTH) & 1u);
                jd++;
            }
            mbar_wait(smem_u32(&sm.full[s]), (i / STAGES) & 1u);
            if (tid == 0) TRACE(i, 1, clock64());

            const uint32_t *stage = &sm.stage[s][G::PAD_FRONT];
            const uint32_t *row = stage + WORDS_PER_THREAD * tid;
            const uint32_t gvalid = sm.info[s].gvalid;
            const uint32_t has_next = sm.info[s].has_next;
            const uint32_t g_thread = 32u * tid;
            uint32_t nvalid = gvalid > g_thread ? gvalid - g_thread : 0u;
            if (nvalid > 32u) nvalid = 32u;
            const uint32_t vmask = nvalid == 32u ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);

            // ---- regroup 32 -> 31 bit (kernels.cu:79) and classify (kernels.cu:93-112)
            uint32_t Z = 0, O = 0;
            {
                // u = group << 1 | (one junk bit): group == 0 <=> u < 2, group == ONES31 <=> u > 0xFFFFFFFD
                uint32_t prev = row[0];
                classify_group<0>(prev << 1, one, Z, O);
#pragma unroll
                for (int j = 1; j < 31; j++) {
                    const uint32_t cur_w = row[j];
                    classify_group_rt(__funnelshift_r(prev, cur_w, 31 - j), one, 1u << j, Z, O);
                    prev = cur_w;
                }
                classify_group_rt(prev, one, BIT31, Z, O);
            }
            Z &= vmask;
            O &= vmask;
            const uint32_t F = Z | O;

            // type of the group after my chunk (first group of the next thread / tile)
            uint32_t nz = 0, no = 0;
            const bool succ = BLOCK_MODE ? (lane != 31u && g_thread + 32u < gvalid)
                                         : (g_thread + 32u < gvalid || (tid == NWORK * 32u - 1u && has_next));
            if (succ) {
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
            const uint32_t qb = below ? 31u - (uint32_t)__clz(below) : 0u;
            const uint32_t open_q = __shfl_sync(0xffffffffu, my_open, qb);
            // (the run open where the WARP starts is added to the warp's first word at copy-out)
            const uint32_t prev_open = below ? open_q + 32u * (lane - qb - 1u) : 32u * lane;
            const uint32_t off = incl - cnt;
            const uint32_t wcnt = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t qlast = tb ? 31u - (uint32_t)__clz(tb) : 0u;
            const uint32_t open_last = __shfl_sync(0xffffffffu, my_open, qlast);
            const bool all_literal = __all_sync(0xffffffffu, T == 0xFFFFFFFFu && F == 0u);

            // ---- exchange the warp aggregates (workers' named barrier); lane w then holds warp w's
            WarpAgg<NWORK> &wa = sm.wagg[q];
            if (lane == 31u) {
                wa.wcnt[warp] = wcnt;
                wa.whas[warp] = tb != 0u;
                wa.wopen[warp] = tb ? open_last + 32u * (31u - qlast) : 1024u;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            const uint32_t a_cnt = lane < NWORK ? wa.wcnt[lane] : 0u;
            uint32_t a_scan = a_cnt;
#pragma unroll
            for (int dd = 1; dd < NWORK; dd <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, a_scan, dd);
                if ((int)lane >= dd) a_scan += o;
            }
            const uint32_t tile_cnt = __shfl_sync(0xffffffffu, a_scan, NWORK - 1);
            const uint32_t wprefix = __shfl_sync(0xffffffffu, a_scan - a_cnt, warp);
            const bool ring_mode = tile_cnt <= (uint32_t)G::RING_WORDS;
            if (warp == 0) {
                // publish the aggregate to the other CTAs at once: their offset sums never wait for this
                // CTA's control warp
                uint32_t tile_open = 0, tile_has = 0;
#pragma unroll
                for (int w = 0; w < NWORK; w++) {
                    const uint32_t h = wa.whas[w], o = wa.wopen[w];
                    if (h) {
                        tile_has = 1;
                        tile_open = o;
                    } else {
                        tile_open += o;
                    }
                }
                if (BLOCK_MODE || has_next == 0u) {   // the end of a column is always a tail
                    tile_open = 0;
                    tile_has = 1;
                }
                if (lane == 0) {
                    st_relaxed_u64(p.desc + sm.info[s].tile, desc_pack(p.epoch, tile_has ^ 1u, tile_open, tile_cnt));
                    mt.tile_cnt = tile_cnt;
                    mt.tile_open = tile_open;
                    mt.tile_has = tile_has;
                    mt.ring_base = head;
                    mt.mode = ring_mode ? MODE_RING : MODE_DIRECT;
                }
            }
            if (tid == 0) TRACE(i, 2, clock64());
            flush_pending();   // the previous tile's deferred words

            // compaction of my words into the ring at `origin`: those with warp-relative index in
            // [lo, lo + WARP_RING) when `windowed`, else all of them
            auto compact = [&](uint32_t origin, uint32_t lo, bool windowed) {
                // (a round of the windowed form concerns the threads whose words [off, off + cnt) reach into it)
                if (windowed && (off + cnt <= lo || off >= lo + (uint32_t)WARP_RING)) return;
                // literals: the group itself (kernels.cu:107-112,256)
                uint32_t m = T & ~F;
                while (m) {
                    const uint32_t j = 31u - (uint32_t)__clz(m);
                    m ^= 1u << j;
                    const uint32_t idx = off + __popc(T & ((1u << j) - 1u)) - lo;
                    if (!windowed || idx < (uint32_t)WARP_RING)
                        sm.ring[(origin + idx) & (G::RING_WORDS - 1)] = row_group(row, j);
                }
                // fills: BIT31 | type << 30 | length (kernels.cu:244-248)
                m = T & F;
                while (m) {
                    const uint32_t j = 31u - (uint32_t)__clz(m);
                    m ^= 1u << j;
                    const uint32_t lower = T & ((1u << j) - 1u);
                    const uint32_t len = lower ? j - (31u - (uint32_t)__clz(lower)) : j + 1u + prev_open;
                    const uint32_t idx = off + __popc(lower) - lo;
                    if (!windowed || idx < (uint32_t)WARP_RING)
                        sm.ring[(origin + idx) & (G::RING_WORDS - 1)] = fill_word((O >> j) & 1u, len);
                }
            };

            if (ring_mode) {
                // ---- stage the tile's words back to back in the ring; the writer warp copies them out
                //      once the tile's offset is known.  Wait until the ring has room for them.
                while (jd < i && head - sm.meta[jd % QDEPTH].ring_base + tile_cnt > (uint32_t)G::RING_WORDS) {
                    mbar_wait(smem_u32(&sm.done[jd % QDEPTH]), (jd / QDEPTH) & 1u);
                    jd++;
                }
                compact(head + wprefix, 0u, false);
                head += tile_cnt;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&sm.empty[s]));
                    mbar_arrive(smem_u32(&sm.agg[q]));
                }
            } else {
                // ---- too many words for the ring (dense data): wait for the tile's offset and write them here
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sm.agg[q]));
                if (all_literal && wprefix == 1024u * warp) {
                    pend = true;
                    pend_s = s;
                    pend_q = q;
                    pend_par = (i / QDEPTH) & 1u;
                    if (tid == 0) TRACE(i, 6, clock64());
                    continue;
                }
                while (jd < i) {   // ring drained: my slice of it serves as staging area below
                    mbar_wait(smem_u32(&sm.done[jd % QDEPTH]), (jd / QDEPTH) & 1u);
                    jd++;
                }
                mbar_wait(smem_u32(&sm.pref[q]), (i / QDEPTH) & 1u);
                const uint64_t dstw = mt.dst + wprefix;
                const uint32_t room =
                    dstw >= p.out_cap ? 0u : (p.out_cap - dstw >= 1024ull ? 1024u : (uint32_t)(p.out_cap - dstw));
                uint32_t *dst = p.out + dstw;
                if (all_literal) {
                    // every group of the warp is a literal: lane-per-output-word, fully coalesced
                    const uint32_t *wrow = stage + 992u * warp;
#pragma unroll 8
                    for (uint32_t k = 0; k < 32u; k++) {
                        const uint32_t g = 32u * k + lane;
                        if (g < room) st_stream_u32(dst + g, extract_group(wrow, g));
                    }
                } else {
                    uint32_t first_add = BLOCK_MODE ? 0u : mt.wcarry[warp];
                    if (wprefix == 0u) first_add += (uint32_t)mt.lead_adjust;   // the launch's first word
                    const uint32_t mine = warp * WARP_RING;
                    for (uint32_t lo = 0; lo < wcnt; lo += WARP_RING) {
                        compact(mine, lo, true);
                        __syncwarp();
                        const uint32_t nround = wcnt - lo < (uint32_t)WARP_RING ? wcnt - lo : (uint32_t)WARP_RING;
                        for (uint32_t k = lane; k < nround; k += 32u) {
                            uint32_t v = sm.ring[mine + k];
                            if (lo + k == 0u) v += first_add;
                            if (lo + k < room) st_stream_u32(dst + lo + k, v);
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));
            }
            if (tid == 0) TRACE(i, 6, clock64());
        }
        flush_pending();
    }
}

// ---- seam between two launches of one CANONICAL stream (inputs beyond MAX_LAUNCH_GROUPS) ----------
//
// The next segment may start with a fill run that continues the last word written so far.  The
// probe measures that leading run straight from the input; the seam kernel then either takes the
// previous word back (the next launch re-emits it, lengthened by lead_adjust) or, if the sum does
// not fit the 30-bit counter, saturates the previous word and shortens the next one.

// result[0] = leading bits of the segment that equal its first bit (atomicMin over CTAs)
__global__ void wah_lead_probe_kernel(const uint32_t *in, uint64_t n_words, unsigned long long *result)
{
    constexpr uint64_t CHUNK = 8192;   // words per CTA step
    __shared__ unsigned long long s_best;
    const uint32_t pattern = (in[0] & 1u) ? 0xFFFFFFFFu : 0u;
    for (uint64_t c0 = (uint64_t)blockIdx.x * CHUNK; c0 < n_words; c0 += (uint64_t)gridDim.x * CHUNK) {
        if (threadIdx.x == 0) s_best = *((volatile unsigned long long *)result);
        __syncthreads();
        const unsigned long long best = s_best;
        __syncthreads();
        if (best < c0 * 32ull) return;   // an earlier chunk already ended the run
        unsigned long long mine = ~0ull;
        for (uint64_t i = c0 + threadIdx.x; i < c0 + CHUNK && i < n_words; i += blockDim.x) {
            const uint32_t x = in[i] ^ pattern;
            if (x) {
                mine = i * 32ull + (uint64_t)(__ffs(x) - 1);
                break;
            }
        }
        if (mine != ~0ull) atomicMin(result, mine);
    }
}

// slot: words written so far (rewritten if the last word is taken back); adjust: for the next launch
__global__ void wah_seam_kernel(const uint32_t *in, uint64_t n_words, uint64_t groups, uint32_t *out, uint64_t out_cap,
                                uint64_t *slot, const unsigned long long *lead_bits, int32_t *adjust)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t base = *slot;
    int32_t adj = 0;
    if (base > 0 && base <= out_cap) {
        const uint32_t type = in[0] & 1u;
        unsigned long long bits = *lead_bits;
        if (bits > n_words * 32ull) bits = n_words * 32ull;
        // the zero padded last group continues a zero run; a one run ends at the last whole group
        uint64_t lead = (bits == n_words * 32ull && type == 0u) ? groups : bits / 31ull;
        const uint32_t pw = out[base - 1];
        if (lead > 0 && is_fill(pw) && ((pw >> 30) & 1u) == type) {
            const uint64_t c1 = fill_count(pw);
            if (c1 + lead <= (uint64_t)MAX_FILL) {
                *slot = base - 1;          // the next launch overwrites the word with the merged run
                adj = (int32_t)c1;
            } else {
                out[base - 1] = fill_word(type, MAX_FILL);
                adj = (int32_t)c1 - (int32_t)MAX_FILL;
            }
        }
    }
    *adjust = adj;
}

template <int NWORK, int STAGES>
size_t smem_bytes_t()
{
    return sizeof(Smem<NWORK, STAGES>) + 128;
}

constexpr int CFG_NWORK = COMPRESS_TILE_WORDS / 992;
constexpr int CFG_STAGES = 3;
// two CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per CTA
static_assert(sizeof(Smem<CFG_NWORK, CFG_STAGES>) + 128 <= (233472 - 2048) / 2, "two CTAs per SM must fit");

}  // namespace

size_t compress_smem_bytes()
{
    return smem_bytes_t<CFG_NWORK, CFG_STAGES>();
}

cudaError_t launch_seam(const uint32_t *d_in, uint64_t n_words, uint64_t groups, uint32_t *d_out, uint64_t out_cap,
                        uint64_t *d_slot, unsigned long long *d_lead_bits, int32_t *d_adjust, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_lead_bits, 0xFF, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    const uint64_t chunks = (n_words + 8191) / 8192;
    const int grid = (int)(chunks < 592 ? chunks : 592);
    wah_lead_probe_kernel<<<grid, 256, 0, stream>>>(d_in, n_words, d_lead_bits);
    wah_seam_kernel<<<1, 32, 0, stream>>>(d_in, n_words, groups, d_out, out_cap, d_slot, d_lead_bits, d_adjust);
    return cudaGetLastError();
}

cudaError_t launch_compress(const CompressParams &p, int mode, cudaStream_t stream)
{
    // A control warp spins on aggregates published by the workers of OTHER CTAs, so every CTA must get
    // an SM slot without waiting for a CTA of this grid to exit: the grid is capped at SMs x occupancy.
    // (Tiles are owned statically, tile k only ever waits for tiles < k, so CTAs that are delayed by
    // foreign work on the GPU merely delay the others.)  WAH_B200_COOPERATIVE=1 makes the launch
    // cooperative, which has the driver verify co-residency at the price of a slower launch.
    constexpr int THREADS = Geom<CFG_NWORK, CFG_STAGES>::THREADS;
    constexpr int MAX_DEV = 64;
    const size_t smem = compress_smem_bytes();
    // per device: the dynamic shared memory attribute and SMs x occupancy belong to the device current at launch time
    static int grids[MAX_DEV][2] = {};
    const int m = mode == 0 ? 0 : 1;
    const void *kernel = m == 0 ? (const void *)wah_compress_kernel<CFG_NWORK, CFG_STAGES, true>
                                : (const void *)wah_compress_kernel<CFG_NWORK, CFG_STAGES, false>;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    if (grids[dev][m] == 0) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        grids[dev][m] = sms * per_sm;
    }
    int grid = grids[dev][m];
    if ((uint32_t)grid > p.n_tiles) grid = (int)p.n_tiles;
    CompressParams params = p;
    void *args[] = {&params};
    static const bool cooperative = [] {
        const char *e = getenv("WAH_B200_COOPERATIVE");
        return e && e[0] == '1';
    }();
    if (cooperative) return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(THREADS), args, smem, stream);
    return launch_pdl(kernel, grid, THREADS, args, smem, stream);
}

}  // namespace wahb200

// wah_kernels.h -- internal interface between the C ABI (wah_capi.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>

namespace wahb200 {

// ------------------------------------------------------------------ compress

// One launch covers at most MAX_LAUNCH_GROUPS groups: the length of a run that is still open is carried from tile to
// tile in 32-bit arithmetic and ends up in a 30-bit fill counter.  (A tile descriptor itself is
// epoch:32 | no_tail:1 | open:14 | count:14, see wah_compress.cu.)
constexpr uint64_t MAX_LAUNCH_GROUPS = 0x3FFFFFFFull;

constexpr uint32_t COMPRESS_SPIN_LIMIT = 1u << 21;            // descriptor polls (an L2 round trip each) before a control warp gives up
constexpr int COMPRESS_THREADS = 256;                         // 8 warps = 8 reference blocks per tile
constexpr int COMPRESS_TILE_WORDS = COMPRESS_THREADS * 31;    // 7936 words  (31 744 B)
constexpr int COMPRESS_TILE_GROUPS = COMPRESS_THREADS * 32;   // 8192 groups

struct CompressParams {
    const uint32_t *in;      // first column of this launch
    uint64_t n_words;        // words per column
    uint64_t groups;         // groups per column = ceil(32 n / 31)
    uint64_t col_stride;     // words between column starts
    uint32_t tiles_per_col;
    uint32_t n_tiles;        // tiles_per_col * n_cols
    uint32_t n_cols;
    const int32_t *lead_adjust;  // nullptr or device int: added to the launch's first word (launch_seam)
    uint32_t *out;
    uint64_t out_cap;
    uint64_t *desc;          // [n_tiles] tile descriptors; need not be cleared (tagged with `epoch`)
    uint32_t epoch;          // unique per launch (never repeated within 2^32 launches of this process)
    const uint64_t *base_in; // words already in `out` (nullptr = 0)
    uint64_t *total_out;     // receives base + words emitted by this launch
    uint64_t *col_offsets;   // nullptr or [n_cols + 1] for this launch's columns
    uint32_t one;            // always 1 (an opaque multiplier that keeps bit accumulation on the FMA pipe)
    uint64_t *trace;         // nullptr; phase timestamps in -DWAH_TRACE builds (scripts/trace_compress.py)
};

size_t compress_smem_bytes();
cudaError_t launch_compress(const CompressParams &p, int mode, cudaStream_t stream);
// CANONICAL stream continued by another launch: decide how the next segment's leading run joins
// the last word written so far (*d_slot words); writes *d_slot and *d_adjust for that launch
cudaError_t launch_seam(const uint32_t *d_in, uint64_t n_words, uint64_t groups, uint32_t *d_out, uint64_t out_cap,
                        uint64_t *d_slot, unsigned long long *d_lead_bits, int32_t *d_adjust, cudaStream_t stream);

// ---------------------------------------------------------------- decompress

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE_WORDS = SCAN_THREADS * 8;              // 2048: smallest scan tile, the unit the workspace is sized by

constexpr int EXPAND_THREADS = 256;
constexpr int EXPAND_TILE_GROUPS = 1024;                       // groups per output tile: a warp's tile, and the grain of the boundary table
constexpr int EXPAND_TILE_WORDS = 992;                         // = 31 x 32 output words (3968 bytes)

// header written by the scan kernel, read by the expand kernel and the host
struct DecodeHeader {
    uint64_t groups;        // G
    uint64_t words;         // ceil(31 G / 32)
    uint64_t out_tiles;     // ceil(G / EXPAND_TILE_GROUPS)
    uint32_t bad_words;     // zero-length fills seen (written in the last round)
    uint32_t valid;         // == the launch's epoch once groups / words / out_tiles are final
    uint32_t error;         // == the launch's epoch if a CTA gave up waiting for another CTA (WAH_STATUS_TIMEOUT)
    uint32_t pad0;
    uint64_t pad[3];
};

// status word of a decode launch: d_out_info[2] (include/wah_b200.h)
constexpr uint64_t STATUS_BAD_MASK = 0xFFFFFFFFull;      // number of zero-length fill words in the stream
constexpr uint64_t STATUS_TIMEOUT = 1ull << 32;          // a CTA gave up waiting for another CTA of the grid
constexpr uint64_t STATUS_BATCH_LENGTH = 1ull << 33;     // batch: the stream does not decode to n_cols * groups_per_col groups
// polls (an L2 round trip and a __nanosleep of >= 64 ns each) after which a CTA that waits for another CTA gives up: about 2 s.
// The kernels need every CTA of the grid resident at the same time; if that ever fails (a foreign context holding SMs,
// a poisoned counter slot) they end with an error instead of hanging the GPU.
constexpr uint32_t SPIN_LIMIT = 1u << 21;

// Counters that must be 0 when a launch starts.  They live in a slot of a small library-owned array (never in the
// caller's workspace, whose content is arbitrary) and every launch leaves its slot zeroed again.
struct DecodeCounters {
    uint32_t agg_count;     // scan tiles that have published their sum so far
    uint32_t bad_acc;       // zero-length fills seen so far
    uint32_t ticket;        // tiles handed out beyond the static rounds (decode: output tiles of the expand phase)
    uint32_t done;          // CTAs that draw no more tickets (the last one zeroes ticket and done)
    uint64_t pad[2];
};

// L2 cache-policy hints of the decoder (launch_decode); WAH_B200_L2_HINTS=0 (experiments, read once) switches them off
inline bool l2_hints()
{
    static const bool on = [] {
        const char *e = getenv("WAH_B200_L2_HINTS");
        return !(e && e[0] == '0');
    }();
    return on;
}

struct ScanParams {
    const uint32_t *in;
    uint64_t c_words;        // words from `in` to the end of the stream
    uint32_t skip_words;     // 0..3: words at `in` that precede the stream (in is 16-byte aligned, the stream need not be)
    uint32_t n_tiles;        // ceil(c_words / tile_words)
    uint32_t tile_words;     // words per scan tile: a multiple of 4 * SCAN_THREADS (scan_tile_words())
    ulonglong2 *desc;        // [n_tiles] tile sums {value, epoch}
    ulonglong2 *excl;        // [n_tiles] tile offsets {value, epoch}, written by each round's aggregator
    uint32_t epoch;          // unique per launch: whatever else is in the workspace reads as unpublished
    DecodeHeader *hdr;       // written by the launch, never read before that
    DecodeCounters *ctr;     // zero at launch, zero again when the launch is over
    ulonglong2 *starts;      // nullptr (size query) or [n_cols * tiles_per_col + 2]: {compressed word index + 1, its group offset}
    uint64_t max_out_tiles;  // output tiles per column the table has room for (single stream: from the capacity)
    uint64_t *out_info;      // nullptr or device u64[3] {words, groups, status}
    // bitmap-index batch: the stream is n_cols columns back to back, each decoding to col_groups groups; output tile
    // (j, k) = tile k of column j has table index j * max_out_tiles + k and starts at group j * col_groups + k * 1024.
    // Single stream: n_cols = 1, col_groups = ~0.
    uint64_t col_groups;
    uint32_t n_cols;
    DecodeCounters *next_ctr;   // the slot the NEXT launch will use: zeroed by this launch's first CTA
    uint32_t chunk_tiles;    // == ExpandParams::chunk_tiles (0 for a size query)
    uint32_t scan_only;      // the scan phase alone (wah_scan_kernel): its last round leaves the counters zeroed and reports the status
    uint32_t l2_keep;        // 1: the scan's loads ask the L2 to keep the stream for the reads that follow (pass 2, the expand
                             // phase); set by launch_decode
    uint64_t *trace;         // nullptr; -DWAH_TRACE builds only
};

struct ExpandParams {
    const uint32_t *in;
    uint64_t c_words;
    const DecodeHeader *hdr;
    uint32_t epoch;
    const ulonglong2 *starts;
    uint64_t max_out_tiles;  // output tiles per column (single stream: tiles the capacity has room for)
    uint32_t *out;
    uint64_t out_cap;        // single stream: capacity of out; batch: words to write per column (<= decoded words of a column)
    uint64_t col_groups;     // batch: groups per column; single stream: ~0
    uint64_t col_stride;     // batch: words between the columns' first output words
    uint32_t n_cols;         // 1 = single stream
    DecodeHeader *hdr_rw;    // == hdr (the expand phase reports a timeout through it)
    uint64_t *out_info;      // nullptr or device u64[3]: the last CTA to leave writes the status into [2]
    uint32_t dynamic_tiles;  // 0: tiles are dealt round robin; else: by ticket after EXPAND_STATIC_ROUNDS rounds
    uint32_t chunk_tiles;    // tiles per chunk of work: 8, 4, 2 or 1 (set by launch_decode)
    uint32_t zero;           // 0 (opaque to the compiler, see the ticket draw in expand_body)
    uint32_t l2_stream_out;  // 1: the output is stored with an L2 evict-first hint (it is never read again; the stream is)
    DecodeCounters *ctr;     // zero at launch; the last CTA to leave zeroes it again
    uint64_t *trace;         // nullptr; phase timestamps in -DWAH_TRACE builds (scripts/trace_decode.py)
};

size_t expand_smem_bytes();
// scan tile size for a stream of c_words: one tile per CTA of the decode grid while that keeps a tile between
// at least SCAN_TILE_WORDS (the unit the workspace is sized by)
cudaError_t scan_tile_words(uint64_t c_words, uint32_t *tile_words);
cudaError_t launch_scan(const ScanParams &p, cudaStream_t stream);
cudaError_t launch_decode(const ScanParams &sp, const ExpandParams &ep, cudaStream_t stream);   // scan + expand, one launch

// a op b on two compressed vectors, BLOCK1024 result (wah_decompress.cu); both streams scanned with a table entry per tile
struct LogicalJob {
    const uint32_t *a, *b;
    uint64_t ca, cb;
    const ulonglong2 *starts_a, *starts_b;
    const DecodeHeader *hdr_a, *hdr_b;
    uint32_t epoch_a, epoch_b;
    uint64_t groups;
    uint32_t n_tiles;
    int op;
    uint32_t *slots;       // n_tiles x 1024 words of scratch
    uint32_t *counts;      // n_tiles
    uint64_t *offsets;     // ceil(n_tiles / 256): words before every 256 tiles
    uint32_t *out;
    uint64_t out_cap;
    uint64_t *total;       // device u64: words of the result
};
cudaError_t launch_logical_compressed(const LogicalJob &job, cudaStream_t stream);

// launch with programmatic stream serialisation (see pdl_wait in wah_common.cuh)
inline cudaError_t launch_pdl(const void *kernel, int grid, int threads, void **args, size_t smem, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, kernel, args);
}

// ----------------------------------------------------------- per-device launch state (wah_device.cu)

// Held around every launch of a persistent kernel: launches of one device are ordered (see wah_device.cu).
class LaunchOrder {
   public:
    explicit LaunchOrder(cudaStream_t stream);
    ~LaunchOrder();
    LaunchOrder(const LaunchOrder &) = delete;
    LaunchOrder &operator=(const LaunchOrder &) = delete;
    cudaError_t status() const { return err_; }
    int device() const { return dev_; }
    // the decode kernel's counter slot for this launch, and the slot the next launch will get
    cudaError_t counter_slots(DecodeCounters **mine, DecodeCounters **next);

   private:
    cudaStream_t stream_;
    int dev_ = -1;
    cudaError_t err_ = cudaSuccess;
};
void forget_stream(cudaStream_t stream);
cudaError_t poison_counter_slots();

// --------------------------------------------------------------------- misc

cudaError_t launch_shard_probe(const uint32_t *d_shard, uint64_t words, uint64_t *d_result /*[5]*/,
                               cudaStream_t stream);
// sparse host transfers (wah_host.cu): 4 KiB blocks, lists of block numbers
cudaError_t launch_scatter_blocks(void *d_dst, uint64_t dst_bytes, const void *d_packed, const uint32_t *d_list, uint32_t n,
                                  cudaStream_t stream);
cudaError_t launch_pack_nonzero_blocks(const void *d_src, uint64_t bytes, uint32_t blocks_per_chunk, uint8_t *d_flags, uint32_t *d_lists,
                                       uint32_t *d_counts, void *d_packed, cudaStream_t stream);
cudaError_t launch_popcount(const uint32_t *d_in, uint64_t c_words, uint64_t *d_bits, cudaStream_t stream);
cudaError_t launch_logical(int op, uint32_t *d_a, const uint32_t *d_b, uint64_t n_words, cudaStream_t stream);
cudaError_t launch_gen_uniform(uint32_t *d_out, uint64_t n_words, double density, uint64_t seed,
                               cudaStream_t stream);
cudaError_t launch_gen_paint_runs(uint32_t *d_out, uint64_t n_words, const int64_t *d_start,
                                  const int64_t *d_len, uint64_t n_runs, cudaStream_t stream);

}  // namespace wahb200

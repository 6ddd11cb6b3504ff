// wah_kernels.h -- internal interface between the C ABI (wah_capi.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wahb200 {

// ------------------------------------------------------------------ compress

// One launch covers at most MAX_LAUNCH_GROUPS groups so that a tile descriptor
// (status:2 | no_tail:1 | open:30 | count:31) fits one 64-bit word.
constexpr uint64_t MAX_LAUNCH_GROUPS = 0x3FFFFFFFull;

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
constexpr int SCAN_ITEMS = 8;                                  // compressed words per thread
constexpr int SCAN_TILE_WORDS = SCAN_THREADS * SCAN_ITEMS;     // 2048

constexpr int EXPAND_THREADS = 256;
constexpr int EXPAND_TILE_GROUPS = EXPAND_THREADS * 32;        // 8192 groups per output tile
constexpr int EXPAND_TILE_WORDS = EXPAND_THREADS * 31;         // 7936 output words

// header written by the scan kernel, read by the expand kernel and the host
struct DecodeHeader {
    uint64_t groups;        // G
    uint64_t words;         // ceil(31 G / 32)
    uint64_t out_tiles;     // ceil(G / EXPAND_TILE_GROUPS)
    uint32_t bad_words;     // zero-length fills seen (written in the last round)
    uint32_t valid;         // == the launch's epoch once groups / words / out_tiles are final
    uint64_t pad[4];
};

// Counters that must be 0 when a launch starts.  They live in a slot of a small library-owned array (never in the
// caller's workspace, whose content is arbitrary) and every launch leaves its slot zeroed again.
struct DecodeCounters {
    uint32_t agg_count;     // scan tiles that have published their sum so far
    uint32_t bad_acc;       // zero-length fills seen so far
    uint32_t ticket;        // tiles handed out beyond the static rounds (decode: output tiles of the expand phase)
    uint32_t done;          // CTAs that draw no more tickets (the last one zeroes ticket and done)
    uint64_t pad[2];
};

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
    ulonglong2 *starts;      // nullptr (size query) or [max_out_tiles + 1]: {compressed word index, its group offset}
    uint64_t max_out_tiles;
    uint64_t *out_info;      // nullptr or device u64[2] {words, groups}
    uint64_t *trace;         // nullptr; -DWAH_TRACE builds only
};

struct ExpandParams {
    const uint32_t *in;
    uint64_t c_words;
    const DecodeHeader *hdr;
    uint32_t epoch;
    const ulonglong2 *starts;
    uint64_t max_out_tiles;
    uint32_t *out;
    uint64_t out_cap;
    uint32_t zero;           // 0 (opaque to the compiler, see the ticket draw in expand_body)
    DecodeCounters *ctr;     // nullptr: tiles are dealt round robin; else: dynamically after EXPAND_STATIC_ROUNDS rounds
    uint64_t *trace;         // nullptr; phase timestamps in -DWAH_TRACE builds (scripts/trace_decode.py)
};

size_t expand_smem_bytes();
// scan tile size for a stream of c_words: one tile per CTA of the decode grid while that keeps a tile between
// at least SCAN_TILE_WORDS (the unit the workspace is sized by)
uint32_t scan_tile_words(uint64_t c_words);
cudaError_t launch_scan(const ScanParams &p, cudaStream_t stream);
cudaError_t launch_decode(const ScanParams &sp, const ExpandParams &ep, cudaStream_t stream);   // scan + expand, one launch

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

// --------------------------------------------------------------------- misc

cudaError_t launch_shard_probe(const uint32_t *d_shard, uint64_t words, uint64_t *d_result /*[6]*/,
                               cudaStream_t stream);
cudaError_t launch_popcount(const uint32_t *d_in, uint64_t c_words, uint64_t *d_bits, cudaStream_t stream);
cudaError_t launch_logical(int op, uint32_t *d_a, const uint32_t *d_b, uint64_t n_words, cudaStream_t stream);
cudaError_t launch_gen_uniform(uint32_t *d_out, uint64_t n_words, double density, uint64_t seed,
                               cudaStream_t stream);
cudaError_t launch_gen_paint_runs(uint32_t *d_out, uint64_t n_words, const int64_t *d_start,
                                  const int64_t *d_len, uint64_t n_runs, cudaStream_t stream);

}  // namespace wahb200

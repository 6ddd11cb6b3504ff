// wah_capi.cu -- the extern "C" boundary of libwah_b200.so (include/wah_b200.h) and the
// host orchestration behind it.  Mirrors what the reference's host functions do around
// their kernels (compress.cu:41-209, decompress.cu:18-141) without the per-call
// cudaMalloc/cudaFree, Thrust scans and 8-byte synchronous copies inside the compute step.
#include "../../include/wah_b200.h"
#include "wah_kernels.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace wahb200;

// ------------------------------------------------------------------------ errors

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// shared with wah_host.cu
int wah_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? WAH_ERR_NOMEM : WAH_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(e__));                                           \
    } while (0)

extern "C" const char *wah_last_error_string(void) { return g_err; }
extern "C" int wah_version(void) { return WAH_B200_VERSION; }

// ------------------------------------------------------------------------- sizes

extern "C" uint64_t wah_num_groups(uint64_t n_words)
{
    // compress.cu:74-81, in 64-bit safe form: 32 n = 31 q + r
    return (n_words / 31ull) * 32ull + ((n_words % 31ull) * 32ull + 30ull) / 31ull;
}
extern "C" uint64_t wah_max_compressed_words(uint64_t n_words) { return wah_num_groups(n_words); }
extern "C" uint64_t wah_decoded_words(uint64_t groups)
{
    // decompress.cu:84-92
    return (groups / 32ull) * 31ull + ((groups % 32ull) * 31ull + 31ull) / 32ull;
}

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------- compress

// workspace: [0,64) two u64 ping-pong slots chaining launches | [64,128) seam scratch (lead bits u64,
// adjust i32) | descriptors
static constexpr size_t WS_SLOTS = 0, WS_SEAM = 64, WS_DESC = 128;
// tiles one launch may cover (descriptor fields are 30/31 bit, wah_kernels.h)
static constexpr uint64_t HW_MAX_LAUNCH_TILES = MAX_LAUNCH_GROUPS / COMPRESS_TILE_GROUPS;
static uint64_t g_max_launch_tiles = HW_MAX_LAUNCH_TILES;
#define MAX_LAUNCH_TILES g_max_launch_tiles

// test hook: shrink the segment one launch covers so that the multi-launch path (and its CANONICAL
// seam) can be exercised with small inputs; 0 restores the default
static uint64_t *g_trace = nullptr;
extern "C" void wah_test_set_trace(uint64_t *d_trace) { g_trace = d_trace; }

extern "C" void wah_test_set_max_launch_tiles(uint64_t tiles)
{
    g_max_launch_tiles = (tiles == 0 || tiles > HW_MAX_LAUNCH_TILES) ? HW_MAX_LAUNCH_TILES : tiles;
}

static uint64_t compress_tiles(uint64_t n_words) { return ceil_div(n_words, COMPRESS_TILE_WORDS); }

extern "C" size_t wah_compress_workspace_bytes(uint64_t n_words)
{
    uint64_t tiles = compress_tiles(n_words);
    if (tiles > MAX_LAUNCH_TILES) tiles = MAX_LAUNCH_TILES;
    return WS_DESC + (size_t)(tiles + 1) * sizeof(uint64_t);
}

extern "C" size_t wah_compress_batch_workspace_bytes(uint64_t n_cols, uint64_t words_per_col)
{
    uint64_t tiles = compress_tiles(words_per_col) * n_cols;
    if (tiles > MAX_LAUNCH_TILES) tiles = MAX_LAUNCH_TILES;
    return WS_DESC + (size_t)(tiles + 1) * sizeof(uint64_t);
}

// Tile descriptors carry the number of the launch that wrote them, so the descriptor array is never cleared
// between calls.  A workspace seen for the first time is cleared once: its content is arbitrary.
#include <atomic>
#include <mutex>
static uint32_t next_epoch()
{
    static std::atomic<uint32_t> e{0x5EED0001u};
    uint32_t v = e.fetch_add(1u);
    if (v == 0u) v = e.fetch_add(1u);
    return v;
}
// WAH_B200_STATIC_TILES=1: deal the decoder's output tiles round robin (the scheme before tickets; for A/B timing)
static bool static_tiles()
{
    static const bool v = [] {
        const char *e = getenv("WAH_B200_STATIC_TILES");
        return e && e[0] == '1';
    }();
    return v;
}

static int check_mode(int mode)
{
    if (mode != WAH_BLOCK1024 && mode != WAH_CANONICAL) return fail(WAH_ERR_INVALID, "unknown mode %d", mode);
    return WAH_OK;
}

extern "C" int wah_compress_batch_device(const uint32_t *d_in, uint64_t n_cols, uint64_t words_per_col,
                                         uint64_t col_stride_words, int mode, uint32_t *d_out,
                                         uint64_t out_capacity_words, uint64_t *d_col_offsets,
                                         void *d_workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (check_mode(mode)) return WAH_ERR_INVALID;
    if (!d_col_offsets) return fail(WAH_ERR_INVALID, "d_col_offsets is null");
    if (n_cols == 0 || words_per_col == 0) {
        CUDA_TRY(cudaMemsetAsync(d_col_offsets, 0, (size_t)(n_cols + 1) * sizeof(uint64_t), stream));
        return WAH_OK;
    }
    if (!d_in || !d_out || !d_workspace) return fail(WAH_ERR_INVALID, "null device pointer");
    if (!aligned16(d_workspace)) return fail(WAH_ERR_INVALID, "workspace must be 16-byte aligned");
    if (n_cols > 1 && col_stride_words < words_per_col)
        return fail(WAH_ERR_INVALID, "col_stride_words < words_per_col");
    if (workspace_bytes < wah_compress_batch_workspace_bytes(n_cols, words_per_col))
        return fail(WAH_ERR_CAPACITY, "workspace too small: %zu < %zu", workspace_bytes,
                    wah_compress_batch_workspace_bytes(n_cols, words_per_col));
    const uint64_t tiles_per_col = compress_tiles(words_per_col);
    if (tiles_per_col > MAX_LAUNCH_TILES)
        return fail(WAH_ERR_INVALID, "a batch column may hold at most %llu words",
                    (unsigned long long)(MAX_LAUNCH_TILES * COMPRESS_TILE_WORDS));
    const uint64_t cols_per_launch = MAX_LAUNCH_TILES / tiles_per_col;

    char *ws = static_cast<char *>(d_workspace);
    uint64_t *slots = reinterpret_cast<uint64_t *>(ws + WS_SLOTS);
    int launch = 0;
    for (uint64_t c0 = 0; c0 < n_cols; c0 += cols_per_launch, launch++) {
        const uint64_t nc = (n_cols - c0) < cols_per_launch ? (n_cols - c0) : cols_per_launch;
        CompressParams p;
        memset(&p, 0, sizeof(p));
        p.in = d_in + c0 * col_stride_words;
        p.n_words = words_per_col;
        p.groups = wah_num_groups(words_per_col);
        p.col_stride = col_stride_words;
        p.tiles_per_col = (uint32_t)tiles_per_col;
        p.n_tiles = (uint32_t)(tiles_per_col * nc);
        p.n_cols = (uint32_t)nc;
        p.lead_adjust = nullptr;
        p.one = 1;
        p.out = d_out;
        p.out_cap = out_capacity_words;
        p.desc = reinterpret_cast<uint64_t *>(ws + WS_DESC);
        p.base_in = launch == 0 ? nullptr : slots + (launch & 1);
        p.total_out = slots + ((launch + 1) & 1);
        p.col_offsets = d_col_offsets + c0;
        p.epoch = next_epoch();
        LaunchOrder order(stream);
        CUDA_TRY(order.status());
        CUDA_TRY(launch_compress(p, mode, stream));
    }
    return WAH_OK;
}

extern "C" int wah_compress_device(const uint32_t *d_in, uint64_t n_words, int mode, uint32_t *d_out,
                                   uint64_t out_capacity_words, uint64_t *d_out_words, void *d_workspace,
                                   size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (check_mode(mode)) return WAH_ERR_INVALID;
    if (!d_out_words) return fail(WAH_ERR_INVALID, "d_out_words is null");
    if (n_words == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out_words, 0, sizeof(uint64_t), stream));
        return WAH_OK;
    }
    if (!d_in || !d_out || !d_workspace) return fail(WAH_ERR_INVALID, "null device pointer");
    if (!aligned16(d_workspace)) return fail(WAH_ERR_INVALID, "workspace must be 16-byte aligned");
    if (workspace_bytes < wah_compress_workspace_bytes(n_words))
        return fail(WAH_ERR_CAPACITY, "workspace too small: %zu < %zu", workspace_bytes,
                    wah_compress_workspace_bytes(n_words));

    // A stream longer than one launch can describe is cut into segments at tile boundaries
    // (multiples of 992 words); each launch appends to the output of the previous one and, in
    // CANONICAL mode, its leading run is joined with the previous launch's last word (launch_seam).
    const uint64_t seg_words = MAX_LAUNCH_TILES * COMPRESS_TILE_WORDS;
    char *ws = static_cast<char *>(d_workspace);
    uint64_t *slots = reinterpret_cast<uint64_t *>(ws + WS_SLOTS);
    int launch = 0;
    for (uint64_t w0 = 0; w0 < n_words; w0 += seg_words, launch++) {
        const uint64_t nw = (n_words - w0) < seg_words ? (n_words - w0) : seg_words;
        const bool last = w0 + nw >= n_words;
        CompressParams p;
        memset(&p, 0, sizeof(p));
        p.in = d_in + w0;
        p.n_words = nw;
        p.groups = wah_num_groups(nw);
        p.col_stride = 0;
        p.tiles_per_col = (uint32_t)compress_tiles(nw);
        p.n_tiles = p.tiles_per_col;
        p.n_cols = 1;
        p.lead_adjust = nullptr;
        p.one = 1;
        if (launch > 0 && mode == WAH_CANONICAL) {
            // the segment's leading run may continue the last word written so far
            p.lead_adjust = reinterpret_cast<int32_t *>(ws + WS_SEAM + 8);
            CUDA_TRY(launch_seam(p.in, nw, p.groups, d_out, out_capacity_words, slots + (launch & 1),
                                 reinterpret_cast<unsigned long long *>(ws + WS_SEAM),
                                 reinterpret_cast<int32_t *>(ws + WS_SEAM + 8), stream));
        }
        p.out = d_out;
        p.out_cap = out_capacity_words;
        p.desc = reinterpret_cast<uint64_t *>(ws + WS_DESC);
        p.base_in = launch == 0 ? nullptr : slots + (launch & 1);
        p.total_out = last ? d_out_words : slots + ((launch + 1) & 1);
        p.col_offsets = nullptr;
        p.trace = g_trace;
        p.epoch = next_epoch();
        LaunchOrder order(stream);
        CUDA_TRY(order.status());
        CUDA_TRY(launch_compress(p, mode, stream));
    }
    return WAH_OK;
}

// -------------------------------------------------------------------- decompress

static uint64_t scan_tiles(uint64_t c_words) { return ceil_div(c_words, SCAN_TILE_WORDS); }
static uint64_t max_out_tiles(uint64_t out_capacity_words) { return ceil_div(out_capacity_words, EXPAND_TILE_WORDS); }
static size_t ws_desc_off() { return sizeof(DecodeHeader); }
static size_t ws_starts_off(uint64_t c_words)
{
    size_t o = ws_desc_off() + 2 * (size_t)scan_tiles(c_words) * sizeof(ulonglong2);   // tile sums + tile offsets
    return (o + 15) & ~(size_t)15;
}
// workspace: header | one tile sum and one tile offset per 2048-word scan tile | one table entry per output tile
static size_t decode_ws_bytes(uint64_t c_words, uint64_t table_entries)
{
    return ws_starts_off(c_words) + (size_t)(table_entries + 2) * sizeof(ulonglong2);
}

extern "C" size_t wah_decompress_workspace_bytes(uint64_t c_words, uint64_t out_capacity_words)
{
    return decode_ws_bytes(c_words, max_out_tiles(out_capacity_words));
}

// One decode launch.  Single stream: n_cols = 1, col_groups = ~0, out_cap = capacity of d_out.  Batch: n_cols columns
// of col_groups groups each, out_cap = words written per column, column j at d_out + j * col_stride.
struct DecodeJob {
    const uint32_t *d_in;
    uint64_t c_words;
    uint32_t skip_words;
    uint32_t *d_out;
    uint64_t out_cap;
    uint64_t n_cols, col_groups, col_stride;
    bool expand;
    bool table_only = false;   // the scan phase alone, but with the output-tile table (out_cap sizes it): the logical operators
    ScanParams *scan_out = nullptr;   // receives the launch's parameters (epoch, header, table)
};

static int decode_launch(const DecodeJob &job, uint64_t *d_out_info, void *d_workspace, size_t workspace_bytes,
                         cudaStream_t stream)
{
    const bool batch = job.col_groups != ~0ull;
    const bool table = job.expand || job.table_only;
    const uint64_t tpc = batch ? ceil_div(job.col_groups, EXPAND_TILE_GROUPS) : (table ? max_out_tiles(job.out_cap) : 0);
    if (tpc > 0xFFFFFFF0ull || job.n_cols > 0xFFFFFFF0ull || job.n_cols * tpc > 0xFFFFFFF0ull)
        return fail(WAH_ERR_INVALID, "too many output tiles for one launch");
    const uint64_t entries = table ? job.n_cols * tpc : 0;
    const size_t need = decode_ws_bytes(job.c_words, entries);
    if (workspace_bytes < need) return fail(WAH_ERR_CAPACITY, "workspace too small: %zu < %zu", workspace_bytes, need);
    if (scan_tiles(job.c_words) > 0x7FFFFFFFull) return fail(WAH_ERR_INVALID, "compressed stream too long");

    char *ws = static_cast<char *>(d_workspace);
    ScanParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.in = job.d_in;
    sp.c_words = job.c_words;
    sp.skip_words = job.skip_words;
    CUDA_TRY(scan_tile_words(job.c_words, &sp.tile_words));
    sp.n_tiles = (uint32_t)ceil_div(job.c_words, sp.tile_words);
    sp.hdr = reinterpret_cast<DecodeHeader *>(ws);
    sp.desc = reinterpret_cast<ulonglong2 *>(ws + ws_desc_off());
    sp.excl = sp.desc + scan_tiles(job.c_words);
    sp.epoch = next_epoch();
    sp.starts = table ? reinterpret_cast<ulonglong2 *>(ws + ws_starts_off(job.c_words)) : nullptr;
    sp.max_out_tiles = tpc;
    sp.out_info = d_out_info;
    sp.col_groups = job.col_groups;
    sp.n_cols = (uint32_t)job.n_cols;
    sp.trace = g_trace;
    LaunchOrder order(stream);
    CUDA_TRY(order.status());
    CUDA_TRY(order.counter_slots(&sp.ctr, &sp.next_ctr));
    if (!job.expand) {
        sp.scan_only = 1;
        sp.chunk_tiles = 1;   // (a table entry for every tile, long fills included)
        if (job.scan_out) *job.scan_out = sp;
        CUDA_TRY(launch_scan(sp, stream));
        return WAH_OK;
    }
    ExpandParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.in = job.d_in;
    ep.c_words = job.c_words;
    ep.hdr = sp.hdr;
    ep.hdr_rw = sp.hdr;
    ep.epoch = sp.epoch;
    ep.starts = sp.starts;
    ep.max_out_tiles = tpc;
    ep.out = job.d_out;
    ep.out_cap = job.out_cap;
    ep.col_groups = job.col_groups;
    ep.col_stride = job.col_stride;
    ep.n_cols = (uint32_t)job.n_cols;
    ep.out_info = d_out_info;
    ep.dynamic_tiles = static_tiles() ? 0u : 1u;
    ep.ctr = sp.ctr;
    ep.trace = g_trace;
    CUDA_TRY(launch_decode(sp, ep, stream));
    return WAH_OK;
}

static int decompress_common(const uint32_t *d_in, uint64_t c_words, uint32_t *d_out, uint64_t out_cap,
                             uint64_t *d_out_info, void *d_workspace, size_t workspace_bytes, bool expand,
                             cudaStream_t stream)
{
    if (!d_out_info) return fail(WAH_ERR_INVALID, "d_out_info is null");
    if (c_words == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out_info, 0, 3 * sizeof(uint64_t), stream));
        return WAH_OK;
    }
    if (!d_in || !d_workspace || (expand && !d_out)) return fail(WAH_ERR_INVALID, "null device pointer");
    if (!aligned16(d_in) || !aligned16(d_workspace) || (expand && !aligned16(d_out)))
        return fail(WAH_ERR_INVALID, "device buffers must be 16-byte aligned");
    DecodeJob job;
    job.d_in = d_in;
    job.c_words = c_words;
    job.skip_words = 0;
    job.d_out = d_out;
    job.out_cap = expand ? out_cap : 0;
    job.n_cols = 1;
    job.col_groups = ~0ull;
    job.col_stride = 0;
    job.expand = expand;
    return decode_launch(job, d_out_info, d_workspace, workspace_bytes, stream);
}

extern "C" int wah_decompress_device(const uint32_t *d_in, uint64_t c_words, uint32_t *d_out,
                                     uint64_t out_capacity_words, uint64_t *d_out_info, void *d_workspace,
                                     size_t workspace_bytes, void *stream)
{
    return decompress_common(d_in, c_words, d_out, out_capacity_words, d_out_info, d_workspace, workspace_bytes,
                             true, (cudaStream_t)stream);
}

extern "C" int wah_decoded_size_device(const uint32_t *d_in, uint64_t c_words, uint64_t *d_out_info,
                                       void *d_workspace, size_t workspace_bytes, void *stream)
{
    return decompress_common(d_in, c_words, nullptr, 0, d_out_info, d_workspace, workspace_bytes, false,
                             (cudaStream_t)stream);
}

// bitmap-index batch: n_cols compressed columns back to back (the layout wah_compress_batch_device writes), ONE launch.
// The reference decodes one vector per call (decompress.cu:61-115: three kernels, a Thrust scan and four
// cudaMalloc / cudaFree pairs each); its callers would loop over the columns.
extern "C" size_t wah_decompress_batch_workspace_bytes(uint64_t n_cols, uint64_t c_total_words, uint64_t words_per_col)
{
    return decode_ws_bytes(c_total_words, n_cols * ceil_div(wah_num_groups(words_per_col), EXPAND_TILE_GROUPS));
}

extern "C" int wah_decompress_batch_device(const uint32_t *d_in, uint64_t c_total_words, uint64_t n_cols,
                                           uint64_t words_per_col, uint32_t *d_out, uint64_t out_col_stride_words,
                                           uint64_t out_col_words, uint64_t *d_out_info, void *d_workspace,
                                           size_t workspace_bytes, void *stream)
{
    if (!d_out_info) return fail(WAH_ERR_INVALID, "d_out_info is null");
    if (n_cols == 0 || words_per_col == 0 || c_total_words == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out_info, 0, 3 * sizeof(uint64_t), (cudaStream_t)stream));
        if (n_cols != 0 && words_per_col != 0) return fail(WAH_ERR_FORMAT, "an empty stream cannot hold %llu columns", (unsigned long long)n_cols);
        return WAH_OK;
    }
    if (!d_in || !d_out || !d_workspace) return fail(WAH_ERR_INVALID, "null device pointer");
    if (!aligned16(d_in) || !aligned16(d_out) || !aligned16(d_workspace))
        return fail(WAH_ERR_INVALID, "device buffers must be 16-byte aligned");
    if (out_col_stride_words % 4 != 0) return fail(WAH_ERR_INVALID, "out_col_stride_words must be a multiple of 4");
    if (out_col_words > out_col_stride_words && n_cols > 1)
        return fail(WAH_ERR_INVALID, "out_col_words > out_col_stride_words");
    const uint64_t groups = wah_num_groups(words_per_col);
    if (out_col_words > wah_decoded_words(groups)) out_col_words = wah_decoded_words(groups);
    DecodeJob job;
    job.d_in = d_in;
    job.c_words = c_total_words;
    job.skip_words = 0;
    job.d_out = d_out;
    job.out_cap = out_col_words;
    job.n_cols = n_cols;
    job.col_groups = groups;
    job.col_stride = out_col_stride_words;
    job.expand = true;
    return decode_launch(job, d_out_info, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int wah_test_poison_counter_slots(void)
{
    CUDA_TRY(poison_counter_slots());
    return WAH_OK;
}

// ----------------------------------------------------------------- range sharding

extern "C" int wah_shard_record_device(const uint32_t *d_shard, uint64_t words, uint64_t groups,
                                       wah_shard_record *h_record, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h_record) return fail(WAH_ERR_INVALID, "h_record is null");
    memset(h_record, 0, sizeof(*h_record));
    h_record->words = words;
    h_record->groups = groups;
    if (words == 0) return WAH_OK;
    if (!d_shard) return fail(WAH_ERR_INVALID, "d_shard is null");
    // 64 bytes of device scratch per device, allocated once (a cudaMalloc / cudaFree pair per call costs more than
    // the probe); calls on one device are serialised by the lock
    static std::mutex mu;
    static void *scratch[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(WAH_ERR_INVALID, "device ordinal %d not supported", dev);
    std::lock_guard<std::mutex> g(mu);
    if (!scratch[dev]) CUDA_TRY(cudaMalloc(&scratch[dev], 8 * sizeof(uint64_t)));
    CUDA_TRY(launch_shard_probe(d_shard, words, (uint64_t *)scratch[dev], stream));
    uint64_t r[5];
    CUDA_TRY(cudaMemcpyAsync(r, scratch[dev], sizeof(r), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    h_record->lead_groups = r[0];
    h_record->lead_words = r[1];
    h_record->lead_type = (uint32_t)r[2];
    h_record->trail_groups = r[3];
    h_record->trail_type = (uint32_t)r[4];
    return WAH_OK;
}

extern "C" int wah_stitch_plan(const wah_shard_record *rec, int n_shards, int mode, uint64_t *skip_words,
                               uint64_t *dst_offset, uint64_t *seam_offset, uint32_t *seam_count,
                               uint32_t *seam_words, uint64_t *total_words)
{
    if (check_mode(mode)) return WAH_ERR_INVALID;
    if (n_shards < 0 || (n_shards && (!rec || !skip_words || !dst_offset || !seam_offset || !seam_count || !seam_words)))
        return fail(WAH_ERR_INVALID, "null argument");
    const uint64_t M = 0x3FFFFFFFull;
    uint64_t total = 0;
    bool open = false;        // the stream so far ends in a fill word ...
    uint32_t open_type = 0;   // ... of this type
    uint64_t open_count = 0;  // ... and this length
    for (int r = 0; r < n_shards; r++) {
        skip_words[r] = 0;
        seam_count[r] = 0;
        seam_offset[r] = total;
        dst_offset[r] = total;
        const wah_shard_record &s = rec[r];
        if (s.words == 0) continue;
        if (mode == WAH_CANONICAL && open && s.lead_groups > 0 && s.lead_type == open_type) {
            // the run that ends the stream so far continues into this shard: re-split the sum
            uint64_t sum = open_count + s.lead_groups;
            const uint64_t full = sum / M, rest = sum % M;
            const uint64_t k = full + (rest ? 1 : 0);
            if (k > WAH_MAX_SEAM_WORDS) return fail(WAH_ERR_INVALID, "seam run too long (%llu words)", (unsigned long long)k);
            uint32_t *sw = seam_words + (size_t)r * WAH_MAX_SEAM_WORDS;
            for (uint64_t i = 0; i < full; i++) sw[i] = 0x80000000u | (open_type << 30) | (uint32_t)M;
            if (rest) sw[full] = 0x80000000u | (open_type << 30) | (uint32_t)rest;
            seam_count[r] = (uint32_t)k;
            seam_offset[r] = total - 1;
            total = total - 1 + k;
            skip_words[r] = s.lead_words;
            dst_offset[r] = total;
            const uint64_t body = s.words - s.lead_words;
            total += body;
            if (body > 0) {
                open = s.trail_groups > 0;
                open_type = s.trail_type;
                open_count = s.trail_groups;
            } else {
                open = true;   // the whole shard was that run
                open_count = rest ? rest : M;
            }
        } else {
            total += s.words;
            open = s.trail_groups > 0;
            open_type = s.trail_type;
            open_count = s.trail_groups;
        }
    }
    if (total_words) *total_words = total;
    return WAH_OK;
}

// --------------------------------------------------------------------- query operators

extern "C" int wah_popcount_device(const uint32_t *d_in, uint64_t c_words, uint64_t *d_bits, void *stream)
{
    if (!d_bits || (c_words && !d_in)) return fail(WAH_ERR_INVALID, "null device pointer");
    CUDA_TRY(launch_popcount(d_in, c_words, d_bits, (cudaStream_t)stream));
    return WAH_OK;
}

// workspace of wah_logical_device: two decoded operands, the decoder's and the encoder's scratch, 64 B of scalars
static size_t logical_operand_bytes(uint64_t n_words) { return (size_t)((n_words + 8 + 3) / 4 * 4) * 4; }
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// workspace of the compressed-domain path: the two scans' workspaces (header, tile cells, a table entry per tile), the
// tiles' word counts and offsets, a slot of 1024 words per tile
struct LogicalLayout {
    uint64_t n_tiles;
    size_t ws_a, ws_b, off_b, off_counts, off_offsets, off_info, off_slots, total;
};
static LogicalLayout logical_layout(uint64_t n_words, uint64_t ca_words, uint64_t cb_words)
{
    LogicalLayout l;
    l.n_tiles = ceil_div(wah_num_groups(n_words), EXPAND_TILE_GROUPS);
    l.ws_a = align256(decode_ws_bytes(ca_words, l.n_tiles));
    l.ws_b = align256(decode_ws_bytes(cb_words, l.n_tiles));
    l.off_b = l.ws_a;
    l.off_counts = l.off_b + l.ws_b;
    l.off_offsets = l.off_counts + align256(l.n_tiles * 4);
    l.off_info = l.off_offsets + align256(l.n_tiles * 8);
    l.off_slots = l.off_info + 256;
    l.total = l.off_slots + align256(l.n_tiles * (size_t)EXPAND_TILE_GROUPS * 4);
    return l;
}

static bool logical_plain()
{
    static const bool v = [] {
        const char *e = getenv("WAH_B200_LOGICAL_PLAIN");
        return e && e[0] == '1';
    }();
    return v;
}

extern "C" size_t wah_logical_workspace_bytes(uint64_t n_words, uint64_t ca_words, uint64_t cb_words)
{
    const uint64_t cmax = ca_words > cb_words ? ca_words : cb_words;
    const size_t plain = 2 * align256(logical_operand_bytes(n_words)) + 256 + align256(wah_decompress_workspace_bytes(cmax, n_words + 8)) +
                         align256(wah_compress_workspace_bytes(n_words));
    const size_t compressed = logical_layout(n_words, ca_words, cb_words).total;
    return plain > compressed ? plain : compressed;
}

// BLOCK1024 result straight from the two streams (wah_decompress.cu, wah_logical_tiles_kernel): neither operand is decoded
// into HBM; cost proportional to the compressed sizes plus a few instructions per 1024-group tile
static int logical_compressed(int op, const uint32_t *d_a, uint64_t ca_words, const uint32_t *d_b, uint64_t cb_words, uint64_t n_words,
                              uint32_t *d_out, uint64_t out_capacity_words, uint64_t *d_out_words, char *ws, cudaStream_t stream)
{
    const LogicalLayout l = logical_layout(n_words, ca_words, cb_words);
    if (l.n_tiles > 0xFFFFFFF0ull) return fail(WAH_ERR_INVALID, "too many tiles for one launch");
    uint64_t *info = reinterpret_cast<uint64_t *>(ws + l.off_info);
    ScanParams spa, spb;
    memset(&spa, 0, sizeof(spa));
    memset(&spb, 0, sizeof(spb));
    const uint32_t *ins[2] = {d_a, d_b};
    const uint64_t cs[2] = {ca_words, cb_words};
    ScanParams *sps[2] = {&spa, &spb};
    char *wss[2] = {ws, ws + l.off_b};
    const size_t wsb[2] = {l.ws_a, l.ws_b};
    for (int i = 0; i < 2; i++) {
        if (cs[i] == 0) continue;   // an empty stream: zeros
        if (!ins[i] || !aligned16(ins[i])) return fail(WAH_ERR_INVALID, "operands must be 16-byte aligned device buffers");
        DecodeJob job;
        job.d_in = ins[i];
        job.c_words = cs[i];
        job.skip_words = 0;
        job.d_out = nullptr;
        job.out_cap = l.n_tiles * (uint64_t)EXPAND_TILE_WORDS;
        job.n_cols = 1;
        job.col_groups = ~0ull;
        job.col_stride = 0;
        job.expand = false;
        job.table_only = true;
        job.scan_out = sps[i];
        if (int rc = decode_launch(job, info + 4 * i, wss[i], wsb[i], stream)) return rc;
    }
    LogicalJob lj;
    lj.a = d_a;
    lj.b = d_b;
    lj.ca = ca_words;
    lj.cb = cb_words;
    lj.starts_a = spa.starts;
    lj.starts_b = spb.starts;
    lj.hdr_a = spa.hdr;
    lj.hdr_b = spb.hdr;
    lj.epoch_a = spa.epoch;
    lj.epoch_b = spb.epoch;
    lj.groups = wah_num_groups(n_words);
    lj.n_tiles = (uint32_t)l.n_tiles;
    lj.op = op;
    lj.slots = reinterpret_cast<uint32_t *>(ws + l.off_slots);
    lj.counts = reinterpret_cast<uint32_t *>(ws + l.off_counts);
    lj.offsets = reinterpret_cast<uint64_t *>(ws + l.off_offsets);
    lj.out = d_out;
    lj.out_cap = out_capacity_words;
    lj.total = d_out_words;
    CUDA_TRY(launch_logical_compressed(lj, stream));
    return WAH_OK;
}

extern "C" int wah_logical_device(int op, const uint32_t *d_a, uint64_t ca_words, const uint32_t *d_b, uint64_t cb_words,
                                  uint64_t n_words, int mode, uint32_t *d_out, uint64_t out_capacity_words,
                                  uint64_t *d_out_words, void *d_workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (op < WAH_OP_AND || op > WAH_OP_ANDNOT) return fail(WAH_ERR_INVALID, "unknown operator %d", op);
    if (check_mode(mode)) return WAH_ERR_INVALID;
    if (!d_workspace || !aligned16(d_workspace)) return fail(WAH_ERR_INVALID, "workspace must be a 16-byte aligned device buffer");
    const size_t need = wah_logical_workspace_bytes(n_words, ca_words, cb_words);
    if (workspace_bytes < need) return fail(WAH_ERR_CAPACITY, "workspace too small: %zu < %zu", workspace_bytes, need);
    char *ws = static_cast<char *>(d_workspace);
    if (mode == WAH_BLOCK1024 && !logical_plain()) {
        if (!d_out_words || (out_capacity_words && !d_out)) return fail(WAH_ERR_INVALID, "null output");
        return logical_compressed(op, d_a, ca_words, d_b, cb_words, n_words, d_out, out_capacity_words, d_out_words, ws, stream);
    }
    // CANONICAL results (and WAH_B200_LOGICAL_PLAIN=1): both operands are expanded into scratch, combined word by word,
    // and the result is compressed again -- three validated kernels and one trivial one, about 6 x 4n bytes of traffic.
    const size_t ob = align256(logical_operand_bytes(n_words));
    uint32_t *buf_a = reinterpret_cast<uint32_t *>(ws), *buf_b = reinterpret_cast<uint32_t *>(ws + ob);
    uint64_t *info = reinterpret_cast<uint64_t *>(ws + 2 * ob);
    const uint64_t cmax = ca_words > cb_words ? ca_words : cb_words;
    const size_t dws_bytes = align256(wah_decompress_workspace_bytes(cmax, n_words + 8));
    void *dws = ws + 2 * ob + 256;
    void *cws = ws + 2 * ob + 256 + dws_bytes;
    // a stream that decodes to fewer than n_words words counts as zero-extended
    CUDA_TRY(cudaMemsetAsync(buf_a, 0, 2 * ob, stream));
    if (int rc = wah_decompress_device(d_a, ca_words, buf_a, n_words + 8, info, dws, dws_bytes, stream)) return rc;
    if (int rc = wah_decompress_device(d_b, cb_words, buf_b, n_words + 8, info + 4, dws, dws_bytes, stream)) return rc;
    CUDA_TRY(launch_logical(op, buf_a, buf_b, n_words, stream));
    return wah_compress_device(buf_a, n_words, mode, d_out, out_capacity_words, d_out_words, cws,
                               wah_compress_workspace_bytes(n_words), stream);
}

// --------------------------------------------------------------------- generators

extern "C" int wah_gen_uniform_device(uint32_t *d_out, uint64_t n_words, double density, uint64_t seed,
                                      void *stream)
{
    if (n_words && !d_out) return fail(WAH_ERR_INVALID, "d_out is null");
    CUDA_TRY(launch_gen_uniform(d_out, n_words, density, seed, (cudaStream_t)stream));
    return WAH_OK;
}

extern "C" int wah_gen_paint_runs_device(uint32_t *d_out, uint64_t n_words, const int64_t *d_start_bits,
                                         const int64_t *d_len_bits, uint64_t n_runs, void *stream)
{
    if (n_runs && (!d_out || !d_start_bits || !d_len_bits)) return fail(WAH_ERR_INVALID, "null device pointer");
    CUDA_TRY(launch_gen_paint_runs(d_out, n_words, d_start_bits, d_len_bits, n_runs, (cudaStream_t)stream));
    return WAH_OK;
}

// wah_container.cpp -- a self-describing container around raw WAH word streams (host only, no CUDA).
//
// The reference keeps a compressed vector as a bare uint32 array plus a length in a local variable; the only
// trace of a file format is a commented-out dump of the words (tests.cpp:278-281).  A stream that leaves the
// process (disk, wire, another rank) needs what those locals held: the encoder mode, the uncompressed length,
// and -- for a bitmap index or a range-sharded vector -- where each column / shard starts.  Layout, little endian:
//
//   offset  0  char[8]  "WAHB200\0"
//           8  u32      container version (1)
//          12  u32      encoder mode (WAH_BLOCK1024 / WAH_CANONICAL)
//          16  u64      n_streams      columns of a bitmap index, shards of one vector, or 1
//          24  u64      words_per_stream   uncompressed 32-bit words each stream decodes to
//          32  u64      total_words    compressed words in the payload
//          40  u64      checksum       sum of the payload words, each multiplied by (its index | 1), plus the stream
//                                  offsets, each multiplied by an odd constant times (its index + 1), mod 2^64
//          48  u64      header_bytes   offset of the payload from the start (a multiple of 16)
//          56  u64      reserved (0)
//          64  u64[n_streams + 1]  word offset of each stream in the payload, the last = total_words
//          header_bytes            u32[total_words]   the streams, back to back
#include "../../include/wah_b200.h"

#include <cstring>

int wah_set_error(int code, const char *fmt, ...);   // wah_capi.cu

namespace {

constexpr char MAGIC[8] = {'W', 'A', 'H', 'B', '2', '0', '0', '\0'};
constexpr uint32_t VERSION = 1;
constexpr size_t FIXED = 64;

struct Header {
    char magic[8];
    uint32_t version, mode;
    uint64_t n_streams, words_per_stream, total_words, checksum, header_bytes, reserved;
};
static_assert(sizeof(Header) == FIXED, "container header layout");

uint64_t checksum(const uint32_t *w, uint64_t n)
{
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) s += (uint64_t)w[i] * (i | 1ull);
    return s;
}

uint64_t table_checksum(const uint64_t *o, uint64_t n)
{
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) s += o[i] * (0x9E3779B97F4A7C15ull * (i + 1));
    return s;
}

uint64_t header_bytes_for(uint64_t n_streams) { return (FIXED + 8 * (n_streams + 1) + 15) & ~(uint64_t)15; }

}  // namespace

extern "C" uint64_t wah_container_bytes(uint64_t n_streams, uint64_t total_words)
{
    return header_bytes_for(n_streams) + 4 * total_words;
}

extern "C" int wah_container_pack(void *dst, uint64_t dst_bytes, int mode, uint64_t n_streams,
                                  uint64_t words_per_stream, const uint64_t *stream_offsets, const uint32_t *words)
{
    if (mode != WAH_BLOCK1024 && mode != WAH_CANONICAL) return wah_set_error(WAH_ERR_INVALID, "unknown mode %d", mode);
    if (!dst || !stream_offsets || n_streams == 0) return wah_set_error(WAH_ERR_INVALID, "null argument / no stream");
    const uint64_t total = stream_offsets[n_streams];
    if (total && !words) return wah_set_error(WAH_ERR_INVALID, "words is null");
    if (stream_offsets[0] != 0) return wah_set_error(WAH_ERR_INVALID, "the first stream must start at word 0");
    for (uint64_t i = 0; i < n_streams; i++)
        if (stream_offsets[i] > stream_offsets[i + 1]) return wah_set_error(WAH_ERR_INVALID, "stream offsets must not decrease");
    const uint64_t hb = header_bytes_for(n_streams);
    if (dst_bytes < hb + 4 * total)
        return wah_set_error(WAH_ERR_CAPACITY, "container needs %llu bytes, buffer has %llu", (unsigned long long)(hb + 4 * total),
                             (unsigned long long)dst_bytes);
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, MAGIC, 8);
    h.version = VERSION;
    h.mode = (uint32_t)mode;
    h.n_streams = n_streams;
    h.words_per_stream = words_per_stream;
    h.total_words = total;
    h.checksum = checksum(words, total) + table_checksum(stream_offsets, n_streams + 1);
    h.header_bytes = hb;
    char *d = static_cast<char *>(dst);
    memset(d, 0, hb);
    memcpy(d, &h, sizeof(h));
    memcpy(d + FIXED, stream_offsets, 8 * (n_streams + 1));
    if (total) memcpy(d + hb, words, 4 * total);
    return WAH_OK;
}

extern "C" int wah_container_unpack(const void *src, uint64_t src_bytes, int *mode, uint64_t *n_streams,
                                    uint64_t *words_per_stream, const uint64_t **stream_offsets,
                                    const uint32_t **words, uint64_t *total_words)
{
    if (!src) return wah_set_error(WAH_ERR_INVALID, "src is null");
    if (src_bytes < FIXED) return wah_set_error(WAH_ERR_FORMAT, "container shorter than its header");
    Header h;
    memcpy(&h, src, sizeof(h));
    if (memcmp(h.magic, MAGIC, 8) != 0) return wah_set_error(WAH_ERR_FORMAT, "not a WAH container (bad magic)");
    if (h.version != VERSION) return wah_set_error(WAH_ERR_FORMAT, "container version %u not supported", h.version);
    if (h.mode != WAH_BLOCK1024 && h.mode != WAH_CANONICAL) return wah_set_error(WAH_ERR_FORMAT, "unknown mode %u", h.mode);
    if (h.n_streams == 0 || h.n_streams > (src_bytes - FIXED) / 8 || h.header_bytes != header_bytes_for(h.n_streams) ||
        h.header_bytes > src_bytes || h.total_words > (src_bytes - h.header_bytes) / 4)
        return wah_set_error(WAH_ERR_FORMAT, "container sizes do not fit the buffer (truncated?)");
    const char *s = static_cast<const char *>(src);
    if ((reinterpret_cast<uintptr_t>(s) & 7u) != 0) return wah_set_error(WAH_ERR_INVALID, "container must be 8-byte aligned");
    const uint64_t *offs = reinterpret_cast<const uint64_t *>(s + FIXED);
    if (offs[0] != 0 || offs[h.n_streams] != h.total_words) return wah_set_error(WAH_ERR_FORMAT, "stream offset table is inconsistent");
    for (uint64_t i = 0; i < h.n_streams; i++)
        if (offs[i] > offs[i + 1]) return wah_set_error(WAH_ERR_FORMAT, "stream offset table is inconsistent");
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s + h.header_bytes);
    if (checksum(w, h.total_words) + table_checksum(offs, h.n_streams + 1) != h.checksum)
        return wah_set_error(WAH_ERR_FORMAT, "checksum mismatch (payload or offset table damaged)");
    if (mode) *mode = (int)h.mode;
    if (n_streams) *n_streams = h.n_streams;
    if (words_per_stream) *words_per_stream = h.words_per_stream;
    if (stream_offsets) *stream_offsets = offs;
    if (words) *words = w;
    if (total_words) *total_words = h.total_words;
    return WAH_OK;
}

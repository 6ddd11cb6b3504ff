/*
 * wah_b200.h -- C ABI of libwah_b200.so: Word-Aligned-Hybrid (WAH) compress /
 * decompress of 32-bit-word bitvectors on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of holgus103/GPU-WAH.  The
 * reference is a C++/CUDA program whose whole public surface is two host
 * functions, compress() (compress.h:12-18) and decompress() (decompress.h:11-17);
 * include/compress.h and include/decompress.h re-declare exactly those and
 * libwah_b200.so exports them with the same C++ mangling, implemented on top of
 * the functions below.  Everything here is plain C: pointers, sizes, no CUDA or
 * torch types (a stream is passed as void*, 0 = the legacy default stream).
 *
 * Data formats (identical to the reference):
 *   uncompressed : n 32-bit words = 32 n bits, LSB first (kernels.cu:79)
 *   group k      : stream bits [31k, 31k+31), zero padded past the end
 *   literal word : bit31 = 0, bits 30..0 = the group (never all 0 / all 1)
 *   fill word    : bit31 = 1, bit30 = fill bit, bits 29..0 = run length in groups
 *                  (const.h:3-12, kernels.cu:244-248, 298-304, 332-354)
 *
 * Every function returns WAH_OK (0) or a WAH_ERR_* code; wah_last_error_string()
 * describes the last failure on the calling thread.  There is no CPU fallback:
 * without a usable CUDA device every compute entry point fails with WAH_ERR_CUDA.
 */
#ifndef WAH_B200_H_
#define WAH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAH_B200_VERSION 200

/* encoder modes */
#define WAH_BLOCK1024 0 /* bit-exact to the reference encoder: runs never cross a block of
                           1024 groups = 992 input words (kernels.cu:256,273-280; compress.cu:146,166) */
#define WAH_CANONICAL 1 /* maximal runs over the whole stream (split only at the 30-bit counter) */

/* status codes */
#define WAH_OK            0
#define WAH_ERR_INVALID   1 /* bad argument (null pointer, misaligned buffer, unknown mode)  */
#define WAH_ERR_CUDA      2 /* a CUDA runtime call failed / no device                        */
#define WAH_ERR_NOMEM     3 /* host or device allocation failed                               */
#define WAH_ERR_CAPACITY  4 /* output or workspace buffer too small                           */
#define WAH_ERR_FORMAT    5 /* compressed stream is malformed (zero-length fill)              */

const char *wah_last_error_string(void);
int wah_version(void);

/* ---- sizes (pure host arithmetic) ------------------------------------------------ */

/* ceil(32 n / 31): number of 31-bit groups, = worst-case compressed words (compress.cu:74-81) */
uint64_t wah_num_groups(uint64_t n_words);
uint64_t wah_max_compressed_words(uint64_t n_words);
/* ceil(31 G / 32): words a stream of G groups decodes to (decompress.cu:82-93) */
uint64_t wah_decoded_words(uint64_t groups);

/* ---- device-resident API (buffers in HBM, asynchronous on `stream`) ----------------
 *
 * Both big kernels are persistent grids whose CTAs exchange results with each other, so two of them must not run on
 * one device at the same time: the library orders its own launches per device (a launch on another stream than the
 * previous one waits for an event behind it; launches on one stream pay nothing).  A CTA that waits for another CTA
 * gives up after about two seconds: compress then reports *d_out_words = UINT64_MAX, decompress
 * WAH_STATUS_TIMEOUT in d_out_info[2] -- an error instead of a hung GPU (e.g. when a foreign context holds SMs).   */

/* replaces compressData + thrust::exclusive_scan + moveData (kernels.cu:51-280, compress.cu:129-166).
 *   d_in            n_words input words, 16-byte aligned
 *   d_out           receives the compressed words; writes past out_capacity_words are dropped
 *   d_out_words     device u64, receives c (the true length even if it exceeds the capacity); UINT64_MAX if the
 *                   launch failed (see above)
 *   d_workspace     wah_compress_workspace_bytes(n_words) bytes, 16-byte aligned.  Scratch: its content
 *                   need not be initialised or preserved (what the kernels exchange through it is tagged
 *                   with a per-launch number), but it must not be shared by launches that may overlap.  */
size_t wah_compress_workspace_bytes(uint64_t n_words);
int wah_compress_device(const uint32_t *d_in, uint64_t n_words, int mode,
                        uint32_t *d_out, uint64_t out_capacity_words, uint64_t *d_out_words,
                        void *d_workspace, size_t workspace_bytes, void *stream);

/* bitmap-index batch: n_cols independent streams of words_per_col words, column j starting at
 * d_in + j * col_stride_words.  Compressed columns are written back to back into d_out;
 * d_col_offsets (device, n_cols+1 u64) receives the word offset of every column.           */
size_t wah_compress_batch_workspace_bytes(uint64_t n_cols, uint64_t words_per_col);
int wah_compress_batch_device(const uint32_t *d_in, uint64_t n_cols, uint64_t words_per_col,
                              uint64_t col_stride_words, int mode,
                              uint32_t *d_out, uint64_t out_capacity_words, uint64_t *d_col_offsets,
                              void *d_workspace, size_t workspace_bytes, void *stream);

/* replaces getCounts + thrust::exclusive_scan + decompressWords + mergeWords
 * (kernels.cu:291-385, decompress.cu:66-115).
 *   d_in            c_words compressed words, 16-byte aligned
 *   d_out           receives ceil(31 G / 32) words (16-byte aligned); words past the capacity are dropped
 *   d_out_info      device u64[3]: [0] = decoded words, [1] = decoded groups G, [2] = status of the launch:
 *                   0 = fine, else WAH_STATUS_* bits (the reference checks nothing in decompress(): decompress.cu:48-52)
 *   d_workspace     wah_decompress_workspace_bytes(c_words, out_capacity_words) bytes; scratch as above   */
#define WAH_STATUS_BAD_WORDS_MASK 0xFFFFFFFFull  /* number of malformed words (fills of 0 groups) in the stream        */
#define WAH_STATUS_TIMEOUT        (1ull << 32)   /* the kernel gave up waiting for part of its own grid (see below)     */
#define WAH_STATUS_BATCH_LENGTH   (1ull << 33)   /* batch: the stream does not decode to n_cols equal columns           */
size_t wah_decompress_workspace_bytes(uint64_t c_words, uint64_t out_capacity_words);
int wah_decompress_device(const uint32_t *d_in, uint64_t c_words,
                          uint32_t *d_out, uint64_t out_capacity_words, uint64_t *d_out_info,
                          void *d_workspace, size_t workspace_bytes, void *stream);
/* bitmap-index batch, ONE launch: d_in holds n_cols compressed columns back to back (the layout
 * wah_compress_batch_device writes; c_total_words = its d_col_offsets[n_cols]), every column the compressed form
 * of words_per_col input words, i.e. of wah_num_groups(words_per_col) groups.  Column j is decoded to
 * d_out + j * out_col_stride_words (a multiple of 4); out_col_words (<= the stride) words are written per column.
 * d_out_info: device u64[3]: [0] = words one column decodes to, [1] = groups in the whole stream, [2] = status.   */
size_t wah_decompress_batch_workspace_bytes(uint64_t n_cols, uint64_t c_total_words, uint64_t words_per_col);
int wah_decompress_batch_device(const uint32_t *d_in, uint64_t c_total_words, uint64_t n_cols, uint64_t words_per_col,
                                uint32_t *d_out, uint64_t out_col_stride_words, uint64_t out_col_words,
                                uint64_t *d_out_info, void *d_workspace, size_t workspace_bytes, void *stream);
/* size query only (the scan half of the above); d_out_info as above */
int wah_decoded_size_device(const uint32_t *d_in, uint64_t c_words, uint64_t *d_out_info,
                            void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- host-buffer API (what the reference's compress()/decompress() do) ------------- */

/* H2D copy, kernels, D2H copy on the current device; *h_out is malloc()ed, release it with
 * wah_free() (or free()).  The three optional floats receive milliseconds for
 * H2D(+allocation) / compute / D2H(+release), like the reference's out-params
 * (compress.cu:205-207, decompress.cu:136-138).  Page-locked inputs are DMAed directly; pageable
 * memory moves through a ring of pinned bounce buffers filled by a few copy threads
 * (WAH_B200_COPY_THREADS, default min(16, cores)).  Transfers of 8 MiB and more are sparse: 4 KiB
 * blocks that are all zero do not cross PCIe (the device input is cleared first, the malloc()ed
 * result comes from calloc()); WAH_B200_SPARSE_COPY=0 moves every byte.  Calls are serialised by a
 * process-wide lock.                                                                              */
int wah_compress_host(const uint32_t *h_in, uint64_t n_words, int mode,
                      uint32_t **h_out, uint64_t *out_words,
                      float *ms_h2d, float *ms_compute, float *ms_d2h);
int wah_decompress_host(const uint32_t *h_in, uint64_t c_words,
                        uint32_t **h_out, uint64_t *out_words,
                        float *ms_h2d, float *ms_compute, float *ms_d2h);
void wah_free(void *p);

/* Same work with caller-provided result buffers (no malloc; a page-locked buffer is written by DMA
 * directly).  *out_words receives the true length; WAH_ERR_CAPACITY if it exceeds the capacity.   */
int wah_compress_host_into(const uint32_t *h_in, uint64_t n_words, int mode,
                           uint32_t *h_out, uint64_t out_capacity_words, uint64_t *out_words);
int wah_decompress_host_into(const uint32_t *h_in, uint64_t c_words,
                             uint32_t *h_out, uint64_t out_capacity_words, uint64_t *out_words);

/* The host entry points keep their device buffers, pinned bounce buffers and copy threads between
 * calls (the reference allocates and frees everything per call); this releases them.            */
void wah_host_release(void);
/* bytes the last host entry point of this process moved over PCIe, each way */
void wah_host_last_transfer_bytes(uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* ---- range sharding of one vector across GPUs ------------------------ */

/* One record per shard, exchanged with an all-gather.  A shard is the compressed form of a
 * contiguous range of the vector that starts at a multiple of 992 words.                   */
typedef struct wah_shard_record {
    uint64_t words;        /* compressed words in the shard                                   */
    uint64_t groups;       /* groups the shard decodes to                                     */
    uint64_t lead_groups;  /* total length of the leading fill run (0 if it starts literal)   */
    uint64_t lead_words;   /* compressed words that leading run occupies                      */
    uint64_t trail_groups; /* length of the last word if it is a fill, else 0                 */
    uint32_t lead_type;    /* fill bit of the leading run                                     */
    uint32_t trail_type;   /* fill bit of the trailing fill                                   */
} wah_shard_record;

/* fill a record from a shard resident on the device (tiny kernel + 48-byte copy, synchronous) */
int wah_shard_record_device(const uint32_t *d_shard, uint64_t words, uint64_t groups,
                            wah_shard_record *h_record, void *stream);

/* Where every shard lands in the concatenated stream.  For shard r:
 *   skip_words[r]   leading words of shard r that are absorbed into the seam before it
 *   dst_offset[r]   word offset in the global stream of shard r's word skip_words[r]
 *   seam_offset[r]  word offset of the seam words written in front of shard r (r >= 1)
 *   seam_count[r]   0..k seam words, seam_words[r*WAH_MAX_SEAM_WORDS + i]
 * The seam words replace the last word of the previous non-empty data and the absorbed
 * leading words.  WAH_BLOCK1024: no merging, plain concatenation.  Returns total words.    */
#define WAH_MAX_SEAM_WORDS 8
int wah_stitch_plan(const wah_shard_record *records, int n_shards, int mode,
                    uint64_t *skip_words, uint64_t *dst_offset,
                    uint64_t *seam_offset, uint32_t *seam_count, uint32_t *seam_words,
                    uint64_t *total_words);

/* ---- query operators on compressed vectors (not in the reference) --------- */

/* *d_bits = number of set bits of the vector the stream stands for, computed from the stream alone */
int wah_popcount_device(const uint32_t *d_in, uint64_t c_words, uint64_t *d_bits, void *stream);

#define WAH_OP_AND    0
#define WAH_OP_OR     1
#define WAH_OP_XOR    2
#define WAH_OP_ANDNOT 3 /* a & ~b */
/* out = compress(decompress(a) op decompress(b)) for two streams that stand for vectors of n_words words each (a
 * stream that decodes to fewer words counts as zero-extended), in either encoder mode; bit-identical to what
 * wah_compress_device produces for the combined vector.  A WAH_BLOCK1024 result is made from the two streams
 * directly -- neither operand is decoded into HBM: tiles of 1024 groups that lie inside fills of both operands become
 * one fill word each, the others are expanded, combined and encoded in shared memory; a WAH_CANONICAL result (or
 * WAH_B200_LOGICAL_PLAIN=1) expands both operands into the workspace, combines them and compresses the result.
 * d_out / d_out_words / capacity as for wah_compress_device.  Asynchronous on `stream`.                       */
size_t wah_logical_workspace_bytes(uint64_t n_words, uint64_t ca_words, uint64_t cb_words);
int wah_logical_device(int op, const uint32_t *d_a, uint64_t ca_words, const uint32_t *d_b, uint64_t cb_words,
                       uint64_t n_words, int mode, uint32_t *d_out, uint64_t out_capacity_words,
                       uint64_t *d_out_words, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- container: compressed streams that leave the process (host only) ---------------- */

/* The reference keeps a compressed vector as a bare word array plus a length in a local variable (its only trace of
 * a file format is a commented-out dump, tests.cpp:278-281).  The container adds what those locals held: a 64-byte
 * header (magic "WAHB200", version, mode, n_streams, words_per_stream, total_words, checksum, header_bytes), the
 * word offset of every stream (columns of a bitmap index as wah_compress_batch_device lays them out, shards of one
 * vector, or a single stream), then the words.  Little endian; layout in gpu-wah_b200/csrc/wah_container.cpp.   */
uint64_t wah_container_bytes(uint64_t n_streams, uint64_t total_words);
/* stream_offsets: n_streams + 1 word offsets, first 0, last = total words */
int wah_container_pack(void *dst, uint64_t dst_bytes, int mode, uint64_t n_streams, uint64_t words_per_stream,
                       const uint64_t *stream_offsets, const uint32_t *words);
/* Validates magic, version, sizes against src_bytes, the offset table and the checksum (WAH_ERR_FORMAT otherwise);
 * the out-pointers (each may be NULL) receive views INTO src, which must be 8-byte aligned.                      */
int wah_container_unpack(const void *src, uint64_t src_bytes, int *mode, uint64_t *n_streams,
                         uint64_t *words_per_stream, const uint64_t **stream_offsets, const uint32_t **words,
                         uint64_t *total_words);

/* ---- synthetic bitvector generators used by bench.py / the tests (device) ---------- */

/* i.i.d. Bernoulli(density) bits, counter-based (splitmix64), reproducible for (seed, word index) */
int wah_gen_uniform_device(uint32_t *d_out, uint64_t n_words, double density, uint64_t seed, void *stream);
/* set stream bits [start[i], start[i]+len[i]) for n_runs runs (d_out pre-zeroed; runs disjoint) */
int wah_gen_paint_runs_device(uint32_t *d_out, uint64_t n_words, const int64_t *d_start_bits,
                              const int64_t *d_len_bits, uint64_t n_runs, void *stream);

/* ---- test hook ---------------------------------------------------------------------- */

/* A stream longer than one kernel launch can describe (2^30 - 1 groups) is compressed as several
 * chained launches.  This shrinks the per-launch segment to `tiles` tiles of 7936 words so that
 * the tests can drive that path with small inputs; 0 restores the default.  Not thread safe.   */
void wah_test_set_max_launch_tiles(uint64_t tiles);
/* Overwrites the decode kernel's library-owned counter slots of the current device with garbage -- what a launch that
 * was killed half way would leave behind.  The next decode launch then times out (WAH_STATUS_TIMEOUT), the one after
 * it works again: every launch zeroes its successor's slot.  Synchronises the device.                             */
int wah_test_poison_counter_slots(void);

#ifdef __cplusplus
}
#endif
#endif /* WAH_B200_H_ */

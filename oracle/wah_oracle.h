/*
 * wah_oracle.h -- CPU oracle for the WAH compress/decompress hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpu-wah_b200/) may
 * include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * The reference (holgus103/GPU-WAH) has no CPU encoder/decoder; this file is a
 * sequential restatement of what its CUDA kernels compute (kernels.cu), pinned
 * by the reference's own golden vectors (tests.cpp:146,162-163,169,183-184,
 * 197-198,209-210) through the unmodified tests.cpp (see oracle/Makefile,
 * target ref_tests_oracle) and, on a GPU box, against the shim-built reference
 * kernels themselves (oracle/_ref/libgpuwah_ref.so).
 */
#ifndef WAH_ORACLE_H_
#define WAH_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Encoder modes.
 * WAH_ORACLE_BLOCK1024: what the reference encoder emits -- canonical WAH inside
 *   every block of 1024 31-bit groups (= 992 input words), blocks concatenated,
 *   fills never merged across a block boundary (kernels.cu:256, 273-280;
 *   compress.cu:146,166; tests.cpp:166-172 pins one block of zeros -> 0x80000400).
 * WAH_ORACLE_CANONICAL: maximal runs over the whole stream; a run longer than
 *   the 30-bit counter is split into full 0x3FFFFFFF chunks followed by the rest.
 */
enum { WAH_ORACLE_BLOCK1024 = 0, WAH_ORACLE_CANONICAL = 1 };

#define WAH_ORACLE_MAX_FILL 0x3FFFFFFFu

/* number of 31-bit groups of an n-word input: ceil(32 n / 31)  (compress.cu:74-81) */
uint64_t wah_oracle_num_groups(uint64_t n_words);

/* 31-bit group k of the LSB-first bit stream, zero padded past the end
 * (kernels.cu:79, tests.cpp:94-97). */
uint32_t wah_oracle_group(const uint32_t *in, uint64_t n_words, uint64_t k);

/* Encode.  out must hold wah_oracle_num_groups(n) words.  Returns c. */
uint64_t wah_oracle_compress(const uint32_t *in, uint64_t n_words, int mode, uint32_t *out);

/* Total groups a compressed stream expands to (getCounts + scan, kernels.cu:291-309,
 * decompress.cu:74-82). */
uint64_t wah_oracle_decoded_groups(const uint32_t *cw, uint64_t c_words);

/* ceil(31 G / 32): the reference's *outSize (decompress.cu:82-93). */
uint64_t wah_oracle_decoded_words(uint64_t groups);

/* Decode.  out must hold wah_oracle_decoded_words(G) words.  Returns that size. */
uint64_t wah_oracle_decompress(const uint32_t *cw, uint64_t c_words, uint32_t *out);

/* Merge adjacent same-type fills of any valid stream into the CANONICAL form.
 * out must hold c_words words.  Returns the new length. */
uint64_t wah_oracle_canonicalize(const uint32_t *cw, uint64_t c_words, uint32_t *out);

/* Multi-threaded (OpenMP) encoder/decoder used as the timed CPU baseline.
 * Same output as the sequential functions.  nthreads <= 0: all cores. */
uint64_t wah_oracle_compress_mt(const uint32_t *in, uint64_t n_words, int mode, uint32_t *out, int nthreads);
uint64_t wah_oracle_decompress_mt(const uint32_t *cw, uint64_t c_words, uint32_t *out, int nthreads);
int wah_oracle_max_threads(void);

/* Batch of equally sized columns, each an independent stream (bitmap index).
 * offsets[n_cols+1] receives the word offset of every column inside out. */
uint64_t wah_oracle_compress_batch(const uint32_t *in, uint64_t n_cols, uint64_t words_per_col,
                                   int mode, uint32_t *out, uint64_t *offsets);

/* bench helper: compress + decompress of every column, columns dealt to the threads (see wah_oracle.c) */
uint64_t wah_oracle_roundtrip_columns_mt(const uint32_t *in, uint64_t n_cols, uint64_t words_per_col, int mode,
                                         int nthreads, int verify, uint64_t *mismatches);

/* out = a op b (0 AND, 1 OR, 2 XOR, 3 ANDNOT) computed on the runs of two streams that stand for vectors of
 * `groups` groups each (a shorter stream counts as zero-extended); equals
 * compress(decompress(a) op decompress(b)) word for word.  out must hold `groups` words.  Returns the length. */
uint64_t wah_oracle_logical(int op, const uint32_t *a, uint64_t ca_words, const uint32_t *b, uint64_t cb_words,
                            uint64_t groups, int mode, uint32_t *out);

/* Host-side synthetic inputs for the CPU arms of bench.py (wah_datagen.c): run clustered (two-state Markov chain,
 * 1-runs of mean mean_run_bits) and i.i.d. Bernoulli(density) bits, LSB first, written as packed words. */
void wah_oracle_gen_clustered(uint32_t *out, uint64_t n_words, double density, double mean_run_bits, uint64_t seed);
void wah_oracle_gen_uniform(uint32_t *out, uint64_t n_words, double density, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* WAH_ORACLE_H_ */

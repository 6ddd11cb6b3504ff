/*
 * wah_oracle.c -- sequential + OpenMP CPU restatement of the GPU-WAH hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see wah_oracle.h).  Every function cites the
 * reference lines (under /root/reference) whose behaviour it restates.
 *
 * Format (const.h:3-12): literal = 31-bit group with bit31 clear;
 * fill = BIT31 | (type << 30) | count-in-groups (30 bit).
 */
#include "wah_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BIT31  0x80000000u
#define BIT30  0x40000000u
#define ONES31 0x7FFFFFFFu
#define MAXFILL WAH_ORACLE_MAX_FILL

uint64_t wah_oracle_num_groups(uint64_t n)
{
    /* compress.cu:74-81: maxExpectedSize = ceil(32 n / 31) */
    unsigned __int128 bits = (unsigned __int128)n * 32u;
    return (uint64_t)((bits + 30u) / 31u);
}

uint64_t wah_oracle_decoded_words(uint64_t groups)
{
    /* decompress.cu:84-92: realSize = ceil(31 G / 32) */
    unsigned __int128 bits = (unsigned __int128)groups * 31u;
    return (uint64_t)((bits + 31u) / 32u);
}

uint32_t wah_oracle_group(const uint32_t *in, uint64_t n, uint64_t k)
{
    /* kernels.cu:79: lane id of a 31-word row sees
     *   ONES31 & ((word[id-1] >> (32-id)) | (word[id] << id))
     * i.e. stream bits [31k, 31k+31), LSB first (same formula tests.cpp:94-97). */
    unsigned __int128 bit = (unsigned __int128)k * 31u;
    uint64_t w = (uint64_t)(bit >> 5);
    unsigned s = (unsigned)(bit & 31u);
    uint64_t lo = w < n ? in[w] : 0;
    uint64_t hi = (w + 1) < n ? in[w + 1] : 0;
    return (uint32_t)(((lo | (hi << 32)) >> s) & ONES31);
}

/* ------------------------------------------------------------------ encoder */

typedef struct {
    uint32_t *out;
    uint64_t c;
    int type;        /* -1: no open run, 0: zero fill, 1: one fill */
    uint32_t count;
} enc_t;

static inline void enc_flush(enc_t *e)
{
    /* kernels.cu:244-248: BIT3130 | count for ones, BIT31 | count for zeros */
    if (e->type >= 0) {
        e->out[e->c++] = BIT31 | ((uint32_t)e->type << 30) | e->count;
        e->type = -1;
        e->count = 0;
    }
}

static inline void enc_fill(enc_t *e, int t, uint32_t cnt)
{
    /* run-end rule kernels.cu:126-141 (a run ends where the next group differs) */
    if (e->type != t) {
        enc_flush(e);
        e->type = t;
    }
    /* 30-bit counter: only reachable in CANONICAL mode */
    while (cnt) {
        uint32_t room = MAXFILL - e->count;
        if (room == 0) {
            enc_flush(e);
            e->type = t;
            room = MAXFILL;
        }
        uint32_t take = cnt < room ? cnt : room;
        e->count += take;
        cnt -= take;
    }
}

static inline void enc_group(enc_t *e, uint32_t g)
{
    if (g == 0u) enc_fill(e, 0, 1);            /* kernels.cu:93  */
    else if (g == ONES31) enc_fill(e, 1, 1);   /* kernels.cu:101 */
    else {                                     /* kernels.cu:107-112,256: stored as is */
        enc_flush(e);
        e->out[e->c++] = g;
    }
}

/* Encode groups [g0, g1) of the stream; g0 must be a multiple of 32 so that a
 * row of 31 words starts there.  block != 0: break runs every 1024 groups. */
static uint64_t encode_range(const uint32_t *in, uint64_t n, uint64_t g0, uint64_t g1,
                             int block, uint32_t *out)
{
    enc_t e = { out, 0, -1, 0 };
    uint64_t k = g0;
    while (k < g1) {
        if (block && (k & 1023u) == 0) enc_flush(&e);   /* kernels.cu:256,273-280: blocks independent */
        uint64_t row = (k >> 5) * 31u;                  /* first word of this 32-group row */
        if (k + 32 <= g1 && row + 31 <= n) {
            const uint32_t *w = in + row;
            /* fast paths: a whole row of zeros / ones is 32 fill groups */
            uint32_t o = 0, a = 0xFFFFFFFFu;
            for (int i = 0; i < 31; i++) { o |= w[i]; a &= w[i]; }
            if (o == 0u) { enc_fill(&e, 0, 32); k += 32; continue; }
            if (a == 0xFFFFFFFFu) { enc_fill(&e, 1, 32); k += 32; continue; }
            enc_group(&e, w[0] & ONES31);
            for (int j = 1; j < 31; j++)
                enc_group(&e, ((w[j - 1] >> (32 - j)) | (w[j] << j)) & ONES31);
            enc_group(&e, w[30] >> 1);
            k += 32;
        } else {
            enc_group(&e, wah_oracle_group(in, n, k));
            k++;
        }
    }
    enc_flush(&e);
    return e.c;
}

uint64_t wah_oracle_compress(const uint32_t *in, uint64_t n, int mode, uint32_t *out)
{
    return encode_range(in, n, 0, wah_oracle_num_groups(n), mode == WAH_ORACLE_BLOCK1024, out);
}

/* ------------------------------------------------------------------ decoder */

static inline uint64_t word_groups(uint32_t w)
{
    /* getCounts, kernels.cu:298-304 */
    return (w & BIT31) ? (uint64_t)(w & (BIT30 - 1u)) : 1u;
}

uint64_t wah_oracle_decoded_groups(const uint32_t *cw, uint64_t c)
{
    uint64_t g = 0;
    for (uint64_t i = 0; i < c; i++) g += word_groups(cw[i]);
    return g;
}

/* OR nbits (<=31) of v into the stream at bit position pos; out is pre-zeroed.
 * lo_w / hi_w: word indices that may be shared with another thread. */
static inline void put_bits(uint32_t *out, uint64_t pos, uint32_t v, uint64_t lo_w, uint64_t hi_w)
{
    uint64_t wi = pos >> 5;
    unsigned s = (unsigned)(pos & 31u);
    uint32_t a = v << s;
    if (a) {
        if (wi == lo_w || wi == hi_w) __atomic_fetch_or(&out[wi], a, __ATOMIC_RELAXED);
        else out[wi] |= a;
    }
    if (s > 1) {
        uint32_t b = v >> (32 - s);
        if (b) {
            wi++;
            if (wi == lo_w || wi == hi_w) __atomic_fetch_or(&out[wi], b, __ATOMIC_RELAXED);
            else out[wi] |= b;
        }
    }
}

/* set stream bits [p0, p1) to one */
static void put_ones(uint32_t *out, uint64_t p0, uint64_t p1, uint64_t lo_w, uint64_t hi_w)
{
    while (p0 < p1 && (p0 & 31u)) {
        uint64_t take = 32 - (p0 & 31u);
        if (take > p1 - p0) take = p1 - p0;
        if (take > 31) take = 31;
        put_bits(out, p0, (uint32_t)((1ull << take) - 1), lo_w, hi_w);
        p0 += take;
    }
    uint64_t full = (p1 - p0) >> 5;
    if (full) {
        memset(out + (p0 >> 5), 0xFF, full * 4);
        p0 += full << 5;
    }
    while (p0 < p1) {
        uint64_t take = p1 - p0;
        if (take > 31) take = 31;
        put_bits(out, p0, (uint32_t)((1ull << take) - 1), lo_w, hi_w);
        p0 += take;
    }
}

static void decode_range(const uint32_t *cw, uint64_t c0, uint64_t c1, uint64_t g0,
                         uint32_t *out, uint64_t lo_w, uint64_t hi_w)
{
    uint64_t pos = g0 * 31u;
    for (uint64_t i = c0; i < c1; i++) {
        uint32_t w = cw[i];
        if (w & BIT31) {                                   /* kernels.cu:332 */
            uint64_t bits = (uint64_t)(w & (BIT30 - 1u)) * 31u;   /* kernels.cu:334 */
            if ((w & (BIT31 | BIT30)) == (BIT31 | BIT30))  /* kernels.cu:337-340: ones */
                put_ones(out, pos, pos + bits, lo_w, hi_w);
            pos += bits;                                   /* zeros: out is already 0 */
        } else {                                           /* kernels.cu:351-354 */
            put_bits(out, pos, w, lo_w, hi_w);
            pos += 31;
        }
    }
}

uint64_t wah_oracle_decompress(const uint32_t *cw, uint64_t c, uint32_t *out)
{
    uint64_t G = wah_oracle_decoded_groups(cw, c);
    uint64_t words = wah_oracle_decoded_words(G);   /* decompress.cu:82-93 */
    memset(out, 0, words * 4);
    /* 31->32 repack, kernels.cu:375-378, done on the fly by writing at bit 31*g */
    decode_range(cw, 0, c, 0, out, UINT64_MAX, UINT64_MAX);
    return words;
}

/* ------------------------------------------------------------- canonicalize */

uint64_t wah_oracle_canonicalize(const uint32_t *cw, uint64_t c, uint32_t *out)
{
    enc_t e = { out, 0, -1, 0 };
    for (uint64_t i = 0; i < c; i++) {
        uint32_t w = cw[i];
        if (w & BIT31) {
            uint32_t cnt = w & (BIT30 - 1u);
            if (cnt) enc_fill(&e, (w >> 30) & 1, cnt);
        } else {
            /* a literal that is all zeros / all ones is not produced by any
             * encoder here, but fold it anyway so the result is canonical */
            enc_group(&e, w);
        }
    }
    enc_flush(&e);
    return e.c;
}

/* ------------------------------------------------------------------ OpenMP */

int wah_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* append src[0..len) to dst (current length *c) merging a fill at the seam the
 * way one sequential CANONICAL pass would have produced it */
static void stitch(uint32_t *dst, uint64_t *c, const uint32_t *src, uint64_t len, int merge)
{
    uint64_t i = 0;
    if (merge && *c > 0 && len > 0 && (dst[*c - 1] & BIT31) && (src[0] & BIT31) &&
        ((dst[*c - 1] ^ src[0]) & BIT30) == 0) {
        uint32_t t = (src[0] >> 30) & 1;
        /* the open run's last chunk + the whole leading run of src */
        uint64_t total = dst[*c - 1] & (BIT30 - 1u);
        (*c)--;
        while (i < len && (src[i] & BIT31) && ((src[i] >> 30) & 1) == t) {
            total += src[i] & (BIT30 - 1u);
            i++;
        }
        while (total > MAXFILL) {
            dst[(*c)++] = BIT31 | (t << 30) | MAXFILL;
            total -= MAXFILL;
        }
        if (total) dst[(*c)++] = BIT31 | (t << 30) | (uint32_t)total;
    }
    memmove(dst + *c, src + i, (len - i) * 4);
    *c += len - i;
}

uint64_t wah_oracle_compress_mt(const uint32_t *in, uint64_t n, int mode, uint32_t *out, int nthreads)
{
    int T = nthreads > 0 ? nthreads : wah_oracle_max_threads();
    uint64_t G = wah_oracle_num_groups(n);
    uint64_t blocks = (G + 1023u) / 1024u;      /* 1024-group (992-word) blocks */
    if (T > (int)blocks) T = (int)(blocks ? blocks : 1);
    if (T <= 1) return wah_oracle_compress(in, n, mode, out);

    uint64_t per = (blocks + T - 1) / T;
    uint64_t *cnt = (uint64_t *)calloc((size_t)T, sizeof(uint64_t));
    uint32_t **tmp = (uint32_t **)calloc((size_t)T, sizeof(uint32_t *));
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        uint64_t g0 = (uint64_t)t * per * 1024u;
        uint64_t g1 = g0 + per * 1024u;
        if (g0 > G) g0 = G;
        if (g1 > G) g1 = G;
        if (t == 0) {
            cnt[t] = encode_range(in, n, g0, g1, mode == WAH_ORACLE_BLOCK1024, out);
        } else if (g1 > g0) {
            tmp[t] = (uint32_t *)malloc((size_t)(g1 - g0) * 4);
            cnt[t] = encode_range(in, n, g0, g1, mode == WAH_ORACLE_BLOCK1024, tmp[t]);
        }
    }
    uint64_t c = cnt[0];
    for (int t = 1; t < T; t++) {
        if (tmp[t]) {
            stitch(out, &c, tmp[t], cnt[t], mode == WAH_ORACLE_CANONICAL);
            free(tmp[t]);
        }
    }
    free(tmp);
    free(cnt);
    return c;
}

uint64_t wah_oracle_decompress_mt(const uint32_t *cw, uint64_t c, uint32_t *out, int nthreads)
{
    int T = nthreads > 0 ? nthreads : wah_oracle_max_threads();
    if ((uint64_t)T > c / 1024u) T = (int)(c / 1024u);
    if (T <= 1) return wah_oracle_decompress(cw, c, out);

    uint64_t per = (c + T - 1) / T;
    uint64_t *goff = (uint64_t *)calloc((size_t)T + 1, sizeof(uint64_t));
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        uint64_t c0 = (uint64_t)t * per, c1 = c0 + per;
        if (c0 > c) c0 = c;
        if (c1 > c) c1 = c;
        goff[t + 1] = wah_oracle_decoded_groups(cw + c0, c1 - c0);
    }
    for (int t = 0; t < T; t++) goff[t + 1] += goff[t];
    uint64_t G = goff[T];
    uint64_t words = wah_oracle_decoded_words(G);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        uint64_t z0 = words * (uint64_t)t / T, z1 = words * (uint64_t)(t + 1) / T;
        memset(out + z0, 0, (z1 - z0) * 4);
#pragma omp barrier
        uint64_t c0 = (uint64_t)t * per, c1 = c0 + per;
        if (c0 > c) c0 = c;
        if (c1 > c) c1 = c;
        uint64_t lo_w = (goff[t] * 31u) >> 5;
        uint64_t hi_w = (goff[t + 1] * 31u) >> 5;
        decode_range(cw, c0, c1, goff[t], out, lo_w, hi_w);
    }
    free(goff);
    return words;
}

uint64_t wah_oracle_compress_batch(const uint32_t *in, uint64_t n_cols, uint64_t words_per_col,
                                   int mode, uint32_t *out, uint64_t *offsets)
{
    uint64_t c = 0;
    for (uint64_t j = 0; j < n_cols; j++) {
        offsets[j] = c;
        c += wah_oracle_compress(in + j * words_per_col, words_per_col, mode, out + c);
    }
    offsets[n_cols] = c;
    return c;
}

/* Bench helper (bitmap index, BASELINE.json configs[3]): every column compressed and decoded again, the columns dealt
 * to the threads, each with its own scratch.  Returns the total number of compressed words; *mismatches = columns that
 * did not round trip (only looked at when verify != 0, so that the comparison stays out of the timed loop). */
uint64_t wah_oracle_roundtrip_columns_mt(const uint32_t *in, uint64_t n_cols, uint64_t words_per_col, int mode,
                                         int nthreads, int verify, uint64_t *mismatches)
{
    int T = nthreads > 0 ? nthreads : wah_oracle_max_threads();
    if ((uint64_t)T > n_cols) T = (int)(n_cols ? n_cols : 1);
    const uint64_t G = wah_oracle_num_groups(words_per_col);
    const uint64_t W = wah_oracle_decoded_words(G);
    uint64_t total = 0, bad = 0;
#pragma omp parallel num_threads(T) reduction(+ : total, bad)
    {
        uint32_t *comp = (uint32_t *)malloc((size_t)(G ? G : 1) * 4);
        uint32_t *dec = (uint32_t *)malloc((size_t)(W ? W : 1) * 4);
#pragma omp for schedule(dynamic, 1)
        for (int64_t j = 0; j < (int64_t)n_cols; j++) {
            const uint32_t *col = in + (uint64_t)j * words_per_col;
            const uint64_t c = wah_oracle_compress(col, words_per_col, mode, comp);
            wah_oracle_decompress(comp, c, dec);
            total += c;
            if (verify && memcmp(dec, col, (size_t)words_per_col * 4) != 0) bad++;
        }
        free(comp);
        free(dec);
    }
    if (mismatches) *mismatches = bad;
    return total;
}

/* ------------------------------------------------------------------ query operators on the runs themselves
 *
 * Not in the reference.  Two streams that stand for vectors of the same length have the same 31-bit groups, so
 * a logical operator can walk both run sequences side by side: where both operands are inside fills the result
 * is a fill of min(remaining) groups (no group is looked at), where one is a literal the result is one group.
 * The result goes through the same encoder state as wah_oracle_compress, so it equals
 * compress(decompress(a) op decompress(b)) word for word, in either mode.  This is the specification (and the
 * checker) of a compressed-domain kernel; cost O(ca + cb), independent of the uncompressed length. */

typedef struct { const uint32_t *w; uint64_t c, i; uint64_t rem; uint32_t val; int fill; } cursor_t;

static inline void cur_next(cursor_t *q)
{
    while (q->rem == 0) {
        if (q->i >= q->c) { q->fill = 1; q->val = 0u; q->rem = ~0ull; return; }   /* zero-extended */
        uint32_t x = q->w[q->i++];
        if (x & BIT31) { q->fill = 1; q->val = (x & BIT30) ? ONES31 : 0u; q->rem = x & (BIT30 - 1u); }
        else { q->fill = 0; q->val = x; q->rem = 1; }
    }
}

static inline uint32_t apply_op(int op, uint32_t a, uint32_t b)
{
    switch (op) {
    case 0: return a & b;
    case 1: return a | b;
    case 2: return a ^ b;
    default: return a & ~b & ONES31;
    }
}

uint64_t wah_oracle_logical(int op, const uint32_t *a, uint64_t ca, const uint32_t *b, uint64_t cb,
                            uint64_t groups, int mode, uint32_t *out)
{
    const int block = mode == WAH_ORACLE_BLOCK1024;
    enc_t e = { out, 0, -1, 0 };
    cursor_t x = { a, ca, 0, 0, 0u, 1 }, y = { b, cb, 0, 0, 0u, 1 };
    uint64_t k = 0;
    while (k < groups) {
        if (block && (k & 1023u) == 0) enc_flush(&e);
        cur_next(&x);
        cur_next(&y);
        uint64_t take = 1;
        if (x.fill && y.fill) {
            take = x.rem < y.rem ? x.rem : y.rem;
            if (take > groups - k) take = groups - k;
            if (block && take > 1024u - (k & 1023u)) take = 1024u - (k & 1023u);
        }
        const uint32_t r = apply_op(op, x.val, y.val) & ONES31;
        if (take == 1) {
            enc_group(&e, r);
        } else {
            uint64_t left = take;                       /* r is 0 or ONES31: both operands are constant here */
            while (left) {
                uint32_t part = left > MAXFILL ? MAXFILL : (uint32_t)left;
                enc_fill(&e, r == ONES31, part);
                left -= part;
            }
        }
        x.rem -= take;
        y.rem -= take;
        k += take;
    }
    enc_flush(&e);
    return e.c;
}

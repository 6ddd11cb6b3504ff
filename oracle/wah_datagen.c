/*
 * wah_datagen.c -- host-side synthetic bitvectors for the CPU arms of bench.py.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (part of libwah_oracle.so, see wah_oracle.h).  The numpy generators in
 * tests/datagen.py need one byte per bit, i.e. 16 GB for the 16 Gbit vector of BASELINE.json configs[2]; these
 * write the packed words directly.  Same distributions as the device generators of the product library
 * (wah_gen_uniform_device, gen_clustered_device), not the same bits.
 *
 * Bits are LSB first inside each 32-bit word (kernels.cu:79, tests.cpp:42-64).
 */
#include "wah_oracle.h"

#include <math.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline double unit_open(uint64_t *s)   /* uniform in (0, 1] */
{
    return ((double)(splitmix64(s) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
}

/* geometric on {1, 2, ...} with the given mean (>= 1) */
static inline uint64_t geometric(uint64_t *s, double mean)
{
    if (mean <= 1.0) return 1;
    const double p = 1.0 / mean;
    const double g = floor(log(unit_open(s)) / log1p(-p)) + 1.0;
    return g < 1.0 ? 1 : (g > 9.0e18 ? (uint64_t)9.0e18 : (uint64_t)g);
}

static void set_bit_range(uint32_t *out, uint64_t b0, uint64_t b1)   /* stream bits [b0, b1) := 1 */
{
    if (b1 <= b0) return;
    const uint64_t w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u), m1 = 0xFFFFFFFFu >> (31u - (uint32_t)((b1 - 1) & 31u));
    if (w0 == w1) {
        out[w0] |= m0 & m1;
        return;
    }
    out[w0] |= m0;
    if (w1 > w0 + 1) memset(out + w0 + 1, 0xFF, (size_t)(w1 - w0 - 1) * 4);
    out[w1] |= m1;
}

/* Two-state Markov chain ("run clustered", configs[2]): 1-runs geometric with mean `mean_run_bits`, 0-runs geometric
 * with mean mean_run_bits (1 - d) / d, starting with a 0-run. */
void wah_oracle_gen_clustered(uint32_t *out, uint64_t n_words, double density, double mean_run_bits, uint64_t seed)
{
    memset(out, 0, (size_t)n_words * 4);
    if (n_words == 0 || density <= 0.0) return;
    const uint64_t n_bits = n_words * 32ull;
    const double l1 = mean_run_bits < 1.0 ? 1.0 : mean_run_bits;
    double l0 = l1 * (1.0 - density) / density;
    if (l0 < 1.0) l0 = 1.0;
    uint64_t s = seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;
    uint64_t pos = 0;
    while (pos < n_bits) {
        pos += geometric(&s, l0);
        if (pos >= n_bits) break;
        const uint64_t len = geometric(&s, l1);
        const uint64_t end = pos + len < n_bits ? pos + len : n_bits;
        set_bit_range(out, pos, end);
        pos = end;
    }
}

/* i.i.d. Bernoulli(density) bits ("uniform" / "sparse", configs[0], configs[1]) */
void wah_oracle_gen_uniform(uint32_t *out, uint64_t n_words, double density, uint64_t seed)
{
    uint64_t s = seed * 0xD1342543DE82EF95ull + 0x9E3779B97F4A7C15ull;
    if (density >= 1.0) {
        memset(out, 0xFF, (size_t)n_words * 4);
        return;
    }
    if (density == 0.5) {
        for (uint64_t i = 0; i + 1 < n_words; i += 2) {
            const uint64_t r = splitmix64(&s);
            out[i] = (uint32_t)r;
            out[i + 1] = (uint32_t)(r >> 32);
        }
        if (n_words & 1) out[n_words - 1] = (uint32_t)splitmix64(&s);
        return;
    }
    memset(out, 0, (size_t)n_words * 4);
    if (density <= 0.0) return;
    const uint64_t n_bits = n_words * 32ull;
    if (density < 0.1) {
        /* gap sampling: the distance to the next set bit is geometric with mean 1 / d */
        uint64_t pos = geometric(&s, 1.0 / density) - 1;
        while (pos < n_bits) {
            out[pos >> 5] |= 1u << (pos & 31u);
            pos += geometric(&s, 1.0 / density);
        }
        return;
    }
    const uint64_t thr = (uint64_t)(density * 18446744073709551615.0);
    for (uint64_t i = 0; i < n_words; i++) {
        uint32_t w = 0;
        for (int b = 0; b < 32; b++) w |= (uint32_t)(splitmix64(&s) < thr) << b;
        out[i] = w;
    }
}

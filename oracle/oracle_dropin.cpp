/*
 * oracle_dropin.cpp -- gives the CPU oracle the reference's two host entry
 * points (compress.h:12-18, decompress.h:11-17; C++ linkage) so that the
 * reference's UNMODIFIED tests.cpp can be linked against it.
 *
 * TEST INFRASTRUCTURE ONLY (see wah_oracle.h).
 *
 *   WAH_ORACLE_MODE=canonical   switch the encoder mode (default: block1024,
 *                               the reference encoder's behaviour)
 *   WAH_DUMP_DIR=<dir>          write every compress() call's input and output
 *                               as raw little-endian uint32 files (used by
 *                               tests/golden/make_golden.py)
 */
#include "wah_oracle.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static int oracle_mode()
{
    const char *m = getenv("WAH_ORACLE_MODE");
    return (m && strcmp(m, "canonical") == 0) ? WAH_ORACLE_CANONICAL : WAH_ORACLE_BLOCK1024;
}

static void dump(const char *kind, int call, const unsigned int *p, unsigned long long n)
{
    const char *dir = getenv("WAH_DUMP_DIR");
    if (!dir) return;
    std::string path = std::string(dir) + "/call" + std::to_string(call) + "_" + kind + ".u32";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return;
    fwrite(p, 4, n, f);
    fclose(f);
}

static int g_calls = 0;

unsigned int *compress(unsigned int *data_cpu, unsigned long long int dataSize,
                       unsigned long long int *outputSize, float *pTransferToDeviceTime,
                       float *pCompressionTime, float *ptranserFromDeviceTime)
{
    unsigned long long G = wah_oracle_num_groups(dataSize);
    unsigned int *out = (unsigned int *)malloc((size_t)(G ? G : 1) * 4);
    if (!out) return NULL;
    unsigned long long c = wah_oracle_compress(data_cpu, dataSize, oracle_mode(), out);
    dump("in", g_calls, data_cpu, dataSize);
    dump("out", g_calls, out, c);
    g_calls++;
    if (outputSize) *outputSize = c;
    if (pTransferToDeviceTime) *pTransferToDeviceTime = 0.f;
    if (pCompressionTime) *pCompressionTime = 0.f;
    if (ptranserFromDeviceTime) *ptranserFromDeviceTime = 0.f;
    return out;
}

unsigned int *decompress(unsigned int *data, unsigned long long int dataSize,
                         unsigned long long int *outSize, float *pTransferToDeviceTime,
                         float *pCompressionTime, float *ptranserFromDeviceTime)
{
    unsigned long long G = wah_oracle_decoded_groups(data, dataSize);
    /* the reference returns a 4*G byte buffer of which realSize words are valid
     * (decompress.cu:127-128) */
    unsigned int *out = (unsigned int *)malloc((size_t)(G ? G : 1) * 4);
    if (!out) return NULL;
    unsigned long long words = wah_oracle_decompress(data, dataSize, out);
    if (outSize) *outSize = words;
    if (pTransferToDeviceTime) *pTransferToDeviceTime = 0.f;
    if (pCompressionTime) *pCompressionTime = 0.f;
    if (ptranserFromDeviceTime) *ptranserFromDeviceTime = 0.f;
    return out;
}

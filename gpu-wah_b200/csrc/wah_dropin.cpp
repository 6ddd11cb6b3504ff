// wah_dropin.cpp -- the reference's two host entry points (compress.h:12-18,
// decompress.h:11-17) on top of the C ABI.  Same argument meaning, same ownership
// (malloc'd result, caller frees), same error behaviour: a message on std::cout and a
// NULL return (compress.cu:89-114,139-163).
#include "../../include/compress.h"
#include "../../include/decompress.h"
#include "../../include/wah_b200.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

static int dropin_mode()
{
    const char *m = std::getenv("WAH_B200_MODE");
    return (m && std::strcmp(m, "canonical") == 0) ? WAH_CANONICAL : WAH_BLOCK1024;
}

unsigned int *compress(unsigned int *data_cpu, unsigned long long int dataSize,
                       unsigned long long int *outputSize, float *pTransferToDeviceTime,
                       float *pCompressionTime, float *ptranserFromDeviceTime)
{
    uint32_t *out = nullptr;
    uint64_t c = 0;
    const int rc = wah_compress_host(data_cpu, dataSize, dropin_mode(), &out, &c, pTransferToDeviceTime,
                                     pCompressionTime, ptranserFromDeviceTime);
    if (rc != WAH_OK) {
        std::cout << "compress failed: " << wah_last_error_string() << std::endl;
        return NULL;
    }
    if (outputSize) *outputSize = c;
    return out;
}

unsigned int *decompress(unsigned int *data, unsigned long long int dataSize,
                         unsigned long long int *outSize, float *pTransferToDeviceTime,
                         float *pCompressionTime, float *ptranserFromDeviceTime)
{
    uint32_t *out = nullptr;
    uint64_t words = 0;
    const int rc = wah_decompress_host(data, dataSize, &out, &words, pTransferToDeviceTime, pCompressionTime,
                                       ptranserFromDeviceTime);
    if (rc != WAH_OK) {
        std::cout << "decompress failed: " << wah_last_error_string() << std::endl;
        return NULL;
    }
    if (outSize) *outSize = words;
    return out;
}

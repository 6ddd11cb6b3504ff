/*
 * include/compress.h -- drop-in declaration of the reference's compression entry point.
 *
 * libwah_b200.so exports this function with C++ linkage (mangled
 * _Z8compressPjyPyPfS1_S1_), i.e. the very symbol the reference's callers
 * (source.cpp:97, every test in tests.cpp) link against; the prototype is the
 * interface of holgus103/GPU-WAH compress.h:12-18 and cannot differ from it.
 *
 *   data_cpu   host pointer to dataSize 32-bit words (LSB-first bit stream); not modified
 *   dataSize   number of input words
 *   outputSize optional: receives the number of compressed words
 *   p*Time     optional: milliseconds spent in  H2D (+allocation) / compute / D2H (+release)
 *   returns    malloc()ed host buffer with the WAH words (caller free()s), NULL on failure
 *              after printing a message to std::cout (compress.cu:89-114)
 *
 * The encoder runs in the reference-exact mode (WAH_BLOCK1024 in wah_b200.h); set the
 * environment variable WAH_B200_MODE=canonical to get fully merged runs instead.
 */
#ifndef WAH_B200_DROPIN_COMPRESS_H
#define WAH_B200_DROPIN_COMPRESS_H

unsigned int *compress(unsigned int *data_cpu, unsigned long long int dataSize,
                       unsigned long long int *outputSize, float *pTransferToDeviceTime,
                       float *pCompressionTime, float *ptranserFromDeviceTime);

#endif

/*
 * include/decompress.h -- drop-in declaration of the reference's decompression entry point.
 *
 * Exported by libwah_b200.so with C++ linkage (mangled _Z10decompressPjyPyPfS1_S1_);
 * the prototype is the interface of holgus103/GPU-WAH decompress.h:11-17.
 *
 *   data      host pointer to dataSize WAH words (any valid stream: counts up to 2^30-1)
 *   outSize   optional: receives ceil(31 G / 32), G = groups encoded (decompress.cu:82-93)
 *   p*Time    optional: milliseconds for H2D / compute / D2H, as in compress.h
 *   returns   malloc()ed host buffer with the decoded words (caller free()s), NULL on failure
 */
#ifndef WAH_B200_DROPIN_DECOMPRESS_H
#define WAH_B200_DROPIN_DECOMPRESS_H

unsigned int *decompress(unsigned int *data, unsigned long long int dataSize,
                         unsigned long long int *outSize, float *pTransferToDeviceTime,
                         float *pCompressionTime, float *ptranserFromDeviceTime);

#endif

/* tests/shim/lz4.h -- TEST INFRASTRUCTURE: prototype-only stand-in for liblz4's public header (liblz4.so.1 is on the
 * image, its development header is not).  It lets the reference's LZ4 CPU-interop examples
 * (examples/lz4_cpu_compression.cu, examples/lz4_cpu_decompression.cu -- the known-answer tests of the LZ4 wire
 * format) compile unchanged; the functions themselves come from liblz4.so.1 at link time. */
#ifndef LZ4_SHIM_H
#define LZ4_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compressBound(int inputSize);
int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);
int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
#ifdef __cplusplus
}
#endif
#endif

/* tests/shim/lz4hc.h -- TEST INFRASTRUCTURE: prototype-only stand-in for liblz4's lz4hc.h (see lz4.h here). */
#ifndef LZ4HC_SHIM_H
#define LZ4HC_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compress_HC(const char* src, char* dst, int srcSize, int dstCapacity, int compressionLevel);
#ifdef __cplusplus
}
#endif
#endif

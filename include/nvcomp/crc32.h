/*
 * nvcomp/crc32.h -- low-level batched standard CRC-32 API (C ABI).
 *
 * Replaces the closed libnvcomp.so 3.0.3 entry point of the same purpose ("Added Standard CRC32
 * support and its LLAPI", reference CHANGELOG.md:51).  The only reference call site is
 * examples/standard_crc_checksum.cpp:94-104, which computes one CRC per uncompressed chunk on the
 * GPU and compares it with boost::crc_32_type -- i.e. the IEEE 802.3 / zlib CRC-32 (reflected
 * polynomial 0xEDB88320, initial value and final xor 0xFFFFFFFF).
 *
 * All pointer / size arrays are device-accessible memory; the call only enqueues work on `stream`.
 */
#ifndef NVCOMP_CRC32_H
#define NVCOMP_CRC32_H

#include "shared_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* device_CRC32_ptr[i] = CRC-32 of the device_uncompressed_bytes[i] bytes at device_uncompressed_ptrs[i]
 * (any alignment, any length including 0).  Asynchronous. */
nvcompStatus_t nvcompBatchedCRC32Async(
    const void* const* device_uncompressed_ptrs,
    const size_t* device_uncompressed_bytes,
    size_t batch_size,
    uint32_t* device_CRC32_ptr,
    cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
namespace nvcomp
{
/* The helper the reference example calls (examples/standard_crc_checksum.cpp:94): one CRC-32 per chunk,
 * default stream, asynchronous (the example's cudaMemcpy that follows synchronises). */
inline void compute_uncomp_chunk_checksums(size_t batch_size, void** uncomp_chunks, size_t* uncomp_chunk_sizes,
                                           uint32_t* result_crcs)
{
  nvcompBatchedCRC32Async((const void* const*)uncomp_chunks, uncomp_chunk_sizes, batch_size, result_crcs, 0);
}
} // namespace nvcomp
#endif

#endif

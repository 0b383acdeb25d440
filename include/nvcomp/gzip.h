/*
 * nvcomp/gzip.h -- Gzip (decompression-only in the reference) is OUT OF SCOPE for this library
 * (SURVEY.md section 2).  The symbols exist so the reference's examples/gzip_gpu_decompression.cu
 * (call sites :110-164) compiles and links under the reference's own CMake; every entry point returns
 * nvcompErrorNotSupported.
 */
#ifndef NVCOMP_GZIP_H
#define NVCOMP_GZIP_H

#include "shared_types.h"

#ifdef __cplusplus
extern "C" {
#endif

nvcompStatus_t nvcompBatchedGzipDecompressGetTempSize(
    size_t num_chunks, size_t max_uncompressed_chunk_bytes, size_t* temp_bytes);
nvcompStatus_t nvcompBatchedGzipGetDecompressSizeAsync(
    const void* const* device_compressed_ptrs, const size_t* device_compressed_bytes,
    size_t* device_uncompressed_bytes, size_t batch_size, cudaStream_t stream);
nvcompStatus_t nvcompBatchedGzipDecompressAsync(
    const void* const* device_compressed_ptrs, const size_t* device_compressed_bytes,
    const size_t* device_uncompressed_bytes, size_t* device_actual_uncompressed_bytes, size_t batch_size,
    void* const device_temp_ptr, size_t temp_bytes, void* const* device_uncompressed_ptrs,
    nvcompStatus_t* device_statuses, cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#endif

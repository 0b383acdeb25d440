// crc32.cuh -- internal interface of crc32.cu (standard CRC-32 on the GPU) used by the high-level
// interface for its whole-buffer checksums.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

constexpr uint32_t kCrcPoly = 0xedb88320u;     // IEEE 802.3, reflected
constexpr size_t kCrcPiece = 65536;            // bytes hashed by one warp

// u32 words of scratch crc32_buffer_async needs for a buffer of at most max_bytes
size_t crc_scratch_words(size_t max_bytes);

// *result = CRC-32 of data[0, n) where n = n_host, or (*len_dev - skip) when len_dev != nullptr (a length
// only the device knows, e.g. the compressed size); max_bytes bounds n and sizes the launch.  Asynchronous.
cudaError_t crc32_buffer_async(const uint8_t* data, size_t n_host, const unsigned long long* len_dev, size_t skip,
                               size_t max_bytes, uint32_t* piece_scratch, uint32_t* result, cudaStream_t stream);

}  // namespace b200

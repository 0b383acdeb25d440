/*
 * nvcomp/zstd.hpp -- ZstdManager placeholder: the Zstd codec is out of scope (see zstd.h); the type
 * exists because benchmarks/benchmark_hlif.cpp:207-212 names it.  Constructing it throws.
 */
#ifndef NVCOMP_ZSTD_HPP
#define NVCOMP_ZSTD_HPP

#include "nvcompManager.hpp"
#include "zstd.h"

namespace nvcomp
{

struct ZstdManager : PimplManager
{
  ZstdManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedZstdOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~ZstdManager() override;
};

} // namespace nvcomp

#endif

/*
 * nvcomp/deflate.hpp -- DeflateManager placeholder: the Deflate codec is out of scope (see deflate.h); the type
 * exists because benchmarks/benchmark_hlif.cpp:207-212 names it.  Constructing it throws.
 */
#ifndef NVCOMP_DEFLATE_HPP
#define NVCOMP_DEFLATE_HPP

#include "nvcompManager.hpp"
#include "deflate.h"

namespace nvcomp
{

struct DeflateManager : PimplManager
{
  DeflateManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedDeflateOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~DeflateManager() override;
};

} // namespace nvcomp

#endif

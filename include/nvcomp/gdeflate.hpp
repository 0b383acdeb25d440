/*
 * nvcomp/gdeflate.hpp -- GdeflateManager placeholder: the Gdeflate codec is out of scope (see gdeflate.h); the type
 * exists because benchmarks/benchmark_hlif.cpp:207-212 names it.  Constructing it throws.
 */
#ifndef NVCOMP_GDEFLATE_HPP
#define NVCOMP_GDEFLATE_HPP

#include "nvcompManager.hpp"
#include "gdeflate.h"

namespace nvcomp
{

struct GdeflateManager : PimplManager
{
  GdeflateManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedGdeflateOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~GdeflateManager() override;
};

} // namespace nvcomp

#endif

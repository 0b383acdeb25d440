/*
 * nvcomp/cascaded.hpp -- CascadedManager (HLIF).  Constructor signature pinned by the reference:
 * benchmarks/benchmark_hlif.cpp:189-205, benchmarks/benchmark_lz4_synth.cpp:62,
 * examples/high_level_quickstart_example.cpp:75.
 */
#ifndef NVCOMP_Cascaded_HPP
#define NVCOMP_Cascaded_HPP

#include "nvcompManager.hpp"
#include "cascaded.h"

namespace nvcomp
{

struct CascadedManager : PimplManager
{
  CascadedManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedCascadedOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~CascadedManager() override;
};

} // namespace nvcomp

#endif

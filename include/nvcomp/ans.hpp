/*
 * nvcomp/ans.hpp -- ANSManager (HLIF).  Constructor signature pinned by the reference:
 * benchmarks/benchmark_hlif.cpp:189-205, benchmarks/benchmark_lz4_synth.cpp:62,
 * examples/high_level_quickstart_example.cpp:75.
 */
#ifndef NVCOMP_ANS_HPP
#define NVCOMP_ANS_HPP

#include "nvcompManager.hpp"
#include "ans.h"

namespace nvcomp
{

struct ANSManager : PimplManager
{
  ANSManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedANSOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~ANSManager() override;
};

} // namespace nvcomp

#endif

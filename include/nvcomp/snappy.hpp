/*
 * nvcomp/snappy.hpp -- SnappyManager (HLIF).  Constructor signature pinned by the reference:
 * benchmarks/benchmark_hlif.cpp:189-205, benchmarks/benchmark_lz4_synth.cpp:62,
 * examples/high_level_quickstart_example.cpp:75.
 */
#ifndef NVCOMP_Snappy_HPP
#define NVCOMP_Snappy_HPP

#include "nvcompManager.hpp"
#include "snappy.h"

namespace nvcomp
{

struct SnappyManager : PimplManager
{
  SnappyManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedSnappyOpts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~SnappyManager() override;
};

} // namespace nvcomp

#endif

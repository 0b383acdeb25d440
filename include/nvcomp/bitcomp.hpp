/*
 * nvcomp/bitcomp.hpp -- BitcompManager (HLIF).  Constructor signature pinned by the reference:
 * benchmarks/benchmark_hlif.cpp:189-205, benchmarks/benchmark_lz4_synth.cpp:62,
 * examples/high_level_quickstart_example.cpp:75.
 */
#ifndef NVCOMP_Bitcomp_HPP
#define NVCOMP_Bitcomp_HPP

#include "nvcompManager.hpp"
#include "bitcomp.h"

namespace nvcomp
{

struct BitcompManager : PimplManager
{
  BitcompManager(
      size_t uncomp_chunk_size,
      const nvcompBatchedBitcompFormatOpts& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~BitcompManager() override;
};

} // namespace nvcomp

#endif

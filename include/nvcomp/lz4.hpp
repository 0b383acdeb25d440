/*
 * nvcomp/lz4.hpp -- LZ4Manager (HLIF).  Constructor signature pinned by the reference:
 * benchmarks/benchmark_hlif.cpp:189-205, benchmarks/benchmark_lz4_synth.cpp:62,
 * examples/high_level_quickstart_example.cpp:75.
 */
#ifndef NVCOMP_LZ4_HPP
#define NVCOMP_LZ4_HPP

#include "nvcompManager.hpp"
#include "lz4.h"

namespace nvcomp
{

struct LZ4Manager : PimplManager
{
  LZ4Manager(
      size_t uncomp_chunk_size,
      const nvcompBatchedLZ4Opts_t& format_opts,
      cudaStream_t user_stream = 0,
      const int device_id = 0,
      ChecksumPolicy checksum_policy = NoComputeNoVerify);
  ~LZ4Manager() override;
};

} // namespace nvcomp

#endif

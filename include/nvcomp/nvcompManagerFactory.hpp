/*
 * nvcomp/nvcompManagerFactory.hpp -- create_manager(): build the right manager from a compressed
 * buffer's header (synchronises the stream), and make every *Manager type visible
 * (reference: examples/high_level_quickstart_example.cpp:88,356; benchmarks/benchmark_hlif.cpp:35,189-212).
 */
#ifndef NVCOMP_MANAGER_FACTORY_HPP
#define NVCOMP_MANAGER_FACTORY_HPP

#include "ans.hpp"
#include "bitcomp.hpp"
#include "cascaded.hpp"
#include "deflate.hpp"
#include "gdeflate.hpp"
#include "lz4.hpp"
#include "nvcompManager.hpp"
#include "snappy.hpp"
#include "zstd.hpp"

#include <memory>

namespace nvcomp
{

std::shared_ptr<nvcompManagerBase> create_manager(
    const uint8_t* comp_buffer,
    cudaStream_t stream = 0,
    const int device_id = 0,
    ChecksumPolicy checksum_policy = NoComputeNoVerify);

} // namespace nvcomp

#endif

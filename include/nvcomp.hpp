/*
 * nvcomp.hpp -- C++ convenience layer shared by the high-level interface (HLIF).
 * Replaces the header of the same name in the closed nvCOMP 3.0.3 package; what the
 * reference needs from it: nvcomp::TypeOf<T>() (specialised for float by
 * benchmarks/benchmark_common.h:136-140) and the exception type its benchmarks catch.
 */
#ifndef NVCOMP_HPP
#define NVCOMP_HPP

#include "nvcomp/shared_types.h"

#include <cstdint>
#include <stdexcept>
#include <string>

namespace nvcomp
{

class NVCompException : public std::runtime_error
{
public:
  NVCompException(nvcompStatus_t err, const std::string& msg)
      : std::runtime_error(msg + " : code=" + std::to_string(static_cast<int>(err)) + "."), m_err(err)
  {
  }
  nvcompStatus_t get_error() const { return m_err; }

private:
  nvcompStatus_t m_err;
};

/* Element type enum of a C++ type. */
template <typename T>
inline nvcompType_t TypeOf();

template <> inline nvcompType_t TypeOf<int8_t>() { return NVCOMP_TYPE_CHAR; }
template <> inline nvcompType_t TypeOf<uint8_t>() { return NVCOMP_TYPE_UCHAR; }
template <> inline nvcompType_t TypeOf<int16_t>() { return NVCOMP_TYPE_SHORT; }
template <> inline nvcompType_t TypeOf<uint16_t>() { return NVCOMP_TYPE_USHORT; }
template <> inline nvcompType_t TypeOf<int32_t>() { return NVCOMP_TYPE_INT; }
template <> inline nvcompType_t TypeOf<uint32_t>() { return NVCOMP_TYPE_UINT; }
template <> inline nvcompType_t TypeOf<int64_t>() { return NVCOMP_TYPE_LONGLONG; }
template <> inline nvcompType_t TypeOf<uint64_t>() { return NVCOMP_TYPE_ULONGLONG; }

} // namespace nvcomp

#endif

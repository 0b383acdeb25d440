// tests/emu/ptx.cuh -- TEST INFRASTRUCTURE: host stand-ins for nvcomp_b200/csrc/ptx.cuh (same names,
// same meaning).  tests/emu puts this directory first on the include path, so the codec headers pick
// these up instead of the inline-PTX versions and run inside the warp emulator (emu_cuda.h) with
// bounds checks on every shared and vector global access.
#pragma once

#include "emu_cuda.h"

namespace b200 {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  emu::Warp* w = emu::g_warp;
  const uint8_t* q = (const uint8_t*)p;
  if (q < w->smem || q > w->smem + w->smem_bytes) emu::fail("smem_addr of a pointer outside shared memory");
  return (uint32_t)(q - w->smem);
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint4* p) {
  if ((uintptr_t)p & 15) emu::fail("misaligned ld_nc_v4 %p", (const void*)p);
  emu::check_global(p, 16, false);
  uint4 r; memcpy(&r, p, 16); return r;
}
__device__ __forceinline__ void st_v4(uint4* p, const uint4& v) {
  if ((uintptr_t)p & 15) emu::fail("misaligned st_v4 %p", (void*)p);
  emu::check_global(p, 16, true);
  memcpy(p, &v, 16);
}
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
  if ((uintptr_t)p & 15) emu::fail("misaligned ld_v4 %p", (const void*)p);
  emu::check_global(p, 16, false);
  uint4 r; memcpy(&r, p, 16); return r;
}

template <int O = 0>
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { return *emu::smem_ptr(a + O, 1); }
template <int O = 0>
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { *emu::smem_ptr(a + O, 1) = (uint8_t)v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  if (a & 15u) emu::fail("misaligned lds_v4 %u", a);
  uint4 r; memcpy(&r, emu::smem_ptr(a, 16), 16); return r;
}
__device__ __forceinline__ void sts_v4(uint32_t a, const uint4& v) {
  if (a & 15u) emu::fail("misaligned sts_v4 %u", a);
  memcpy(emu::smem_ptr(a, 16), &v, 16);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  if (a & 3u) emu::fail("misaligned lds_u32 %u", a);
  uint32_t r; memcpy(&r, emu::smem_ptr(a, 4), 4); return r;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  if (a & 3u) emu::fail("misaligned sts_u32 %u", a);
  memcpy(emu::smem_ptr(a, 4), &v, 4);
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  if (a & 1u) emu::fail("misaligned lds_u16 %u", a);
  uint16_t r; memcpy(&r, emu::smem_ptr(a, 2), 2); return r;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  if (a & 1u) emu::fail("misaligned sts_u16 %u", a);
  const uint16_t x = (uint16_t)v; memcpy(emu::smem_ptr(a, 2), &x, 2);
}
template <int O>
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p) {
  emu::check_global(p + O, 1, false);
  return p[O];
}

__device__ __forceinline__ void touch_line(const void* p) { emu::check_global(p, 4, false); }
template <int O>
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) {
  if ((uintptr_t)(p + O) & 3u) emu::fail("misaligned ldg_u32 %p", (const void*)(p + O));
  emu::check_global(p + O, 4, false);
  uint32_t r; memcpy(&r, p + O, 4); return r;
}

// mbarrier + TMA bulk copy: the copy completes at once; the barrier is a phase counter in the 8 bytes it occupies
// and a waiting lane yields to the other fibers until the phase it waits for has completed.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t) { uint32_t z = 0; memcpy(emu::smem_ptr(mbar, 8), &z, 4); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t, uint32_t) {}
__device__ __forceinline__ void fence_proxy_async_smem() {}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t mbar) {
  if ((smem_dst & 15u) || ((uintptr_t)gmem_src & 15u) || (bytes & 15u)) emu::fail("misaligned bulk copy");
  emu::check_global(gmem_src, bytes, false);
  memcpy(emu::smem_ptr(smem_dst, bytes), gmem_src, bytes);
  uint32_t c; memcpy(&c, emu::smem_ptr(mbar, 8), 4); ++c; memcpy(emu::smem_ptr(mbar, 8), &c, 4);
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  for (int spins = 0;; ++spins) {
    uint32_t c; memcpy(&c, emu::smem_ptr(mbar, 8), 4);
    if ((c & 1u) != parity) return;
    if (spins > 1000000) emu::fail("mbar_wait never completes (parity %u)", parity);
    emu::yield();
  }
}

}  // namespace b200

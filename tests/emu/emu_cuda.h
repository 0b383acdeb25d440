// emu_cuda.h -- TEST INFRASTRUCTURE: a 32-lane warp emulator for the host.
//
// The warp-level decode logic of nvcomp_b200/csrc/*.cuh (no asm: all PTX is in ptx.cuh, which
// tests/emu shadows) is compiled with g++ and run here lane by lane: every lane is a user-level
// fiber, every warp intrinsic (__shfl_sync, __ballot_sync, __syncwarp ...) is a rendezvous of all
// 32 fibers.  The emulator checks what the GPU cannot tell us without a GPU: that all lanes reach
// the same intrinsic (op id + call site), that shared-memory and global accesses stay inside the
// buffers the test registered, and that the output is bit-exact.  It is never linked into
// libnvcomp.so and nothing under nvcomp_b200/ includes it.
#pragma once

#include <cuda_runtime.h>   // vector types, __device__ / __forceinline__ macros (host flavour)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

namespace emu {

constexpr int kLanes = 32;
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  void* sp = nullptr;        // saved stack pointer
  uint8_t* stack = nullptr;
  bool done = false;
};

struct Region { const uint8_t* lo; const uint8_t* hi; bool writable; };

struct Warp {
  Fiber lanes[kLanes];
  void* main_sp = nullptr;
  int cur = -1;
  int arrived = 0;
  uint64_t gen = 0;
  int op[kLanes];
  const void* site[kLanes];
  uint64_t xchg[kLanes];
  uint32_t pred[kLanes];
  std::function<void(int)> body;
  // shared memory window of this warp (addresses are offsets into it)
  uint8_t* smem = nullptr;
  size_t smem_bytes = 0;
  // registered global regions (bounds checks)
  Region regions[8];
  int n_regions = 0;
  // statistics
  uint64_t n_sync = 0;
  bool failed = false;
  char fail_msg[256];
};

extern thread_local Warp* g_warp;

[[noreturn]] void fail(const char* fmt, ...);
void run_warp(Warp& w, size_t smem_bytes, std::function<void(int)> body);
void rendezvous(int opcode, const void* site);
void yield();   // let the other lanes run (a lane spinning on shared state)
int lane();

inline void add_region(Warp& w, const void* p, size_t n, bool writable) {
  if (w.n_regions >= 8) { fprintf(stderr, "emu: too many regions\n"); abort(); }
  w.regions[w.n_regions++] = Region{(const uint8_t*)p, (const uint8_t*)p + n, writable};
}
void check_global(const void* p, size_t n, bool write);
inline uint8_t* smem_ptr(uint32_t a, size_t n) {
  Warp* w = g_warp;
  if ((size_t)a + n > w->smem_bytes) fail("shared access [%u,+%zu) outside window of %zu bytes", a, n, w->smem_bytes);
  return w->smem + a;
}

}  // namespace emu

// ---------------------------------------------------------------------------
// CUDA device-side vocabulary used by the codec headers
// ---------------------------------------------------------------------------
struct EmuThreadIdx { operator unsigned() const { return (unsigned)emu::lane(); } };
struct EmuDim3 { EmuThreadIdx x; };
static const EmuDim3 threadIdx{};

#define EMU_SITE() __builtin_return_address(0)

template <class T>
static inline T emu_xchg(T v, int src_lane, int opcode) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::Warp* w = emu::g_warp;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->xchg[emu::lane()] = raw;
  emu::rendezvous(opcode, nullptr);
  const uint64_t got = w->xchg[src_lane & 31];
  emu::rendezvous(opcode + 1000, nullptr);
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_xchg(v, src, 1); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = emu::lane();
  return emu_xchg(v, l >= (int)d ? l - (int)d : l, 2);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = emu::lane();
  return emu_xchg(v, l + (int)d < 32 ? l + (int)d : l, 3);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return emu_xchg(v, emu::lane() ^ m, 4);
}
static inline unsigned __ballot_sync(unsigned, int p) {
  emu::Warp* w = emu::g_warp;
  w->pred[emu::lane()] = p ? 1u : 0u;
  emu::rendezvous(5, nullptr);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= w->pred[i] << i;
  emu::rendezvous(1005, nullptr);
  return m;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0u; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::rendezvous(6, nullptr); }
static inline unsigned emu_reduce(unsigned v, int kind) {
  emu::Warp* w = emu::g_warp;
  w->xchg[emu::lane()] = v;
  emu::rendezvous(7 + kind, nullptr);
  unsigned r = (unsigned)w->xchg[0];
  for (int i = 1; i < 32; ++i) {
    const unsigned x = (unsigned)w->xchg[i];
    r = kind == 0 ? (r | x) : kind == 1 ? (r + x) : kind == 2 ? (r > x ? r : x) : kind == 3 ? (r < x ? r : x) : (r & x);
  }
  emu::rendezvous(1007 + kind, nullptr);
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return emu_reduce(v, 0); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return emu_reduce(v, 1); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return emu_reduce(v, 2); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return emu_reduce(v, 3); }
static inline unsigned __reduce_and_sync(unsigned, unsigned v) { return emu_reduce(v, 4); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (unsigned)(x >> (sh & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (unsigned)((x << (sh & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const uint64_t x = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xfu;
    unsigned byte = (unsigned)(x >> (8 * (sel & 7u))) & 0xffu;
    if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline bool __isGlobal(const void*) { return true; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}

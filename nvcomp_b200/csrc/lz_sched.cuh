// lz_sched.cuh -- work distribution of the batched LZ4 / Snappy decoders.
//
// A batch is decoded by two kernels that run side by side (lz4.cu / snappy.cu): "light" chunks -- compressed >= 4x
// (long matches, typed run-length data) or practically incompressible (one long literal run) -- are streamed by the
// direct global-memory sequence loop at high occupancy, everything else is dense short-token data for the
// block-parallel decoder (lz_decode.cuh).  A classification pass first writes the two chunk-index lists into the
// caller's workspace; each kernel's persistent warps then pull from their own list with an atomic ticket, so a
// kernel whose list is empty retires at once instead of walking the whole batch.  Without a workspace (temp ==
// nullptr is legal for these codecs) both kernels stride over all chunks and skip the other kernel's.
#pragma once

#include "common.cuh"

namespace b200 {

__device__ __forceinline__ bool lz_chunk_is_light(uint64_t cap, uint64_t in_n) {
  return cap >= 4ull * in_n || in_n + (cap >> 6) >= cap;
}

// workspace: kSchedBytes of counters | u32 light[batch] | u32 dense[batch]
struct LzLists {
  unsigned long long* ctr;      // [0] light ticket, [1] dense ticket, [2] number of light chunks, [3] of dense chunks
  uint32_t* light;
  uint32_t* dense;
};
inline size_t lz_decode_temp_bytes(size_t batch) { return kSchedBytes + ((8 * batch + 255) & ~(size_t)255); }
inline LzLists lz_lists_in(void* temp, size_t temp_bytes, size_t batch) {
  LzLists l{nullptr, nullptr, nullptr};
  if (temp && temp_bytes >= lz_decode_temp_bytes(batch) && batch <= 0xffffffffull) {
    l.ctr = (unsigned long long*)temp;
    l.light = (uint32_t*)((uint8_t*)temp + kSchedBytes);
    l.dense = l.light + batch;
  }
  return l;
}

// (static: one copy per translation unit that launches it)
static __global__ void __launch_bounds__(256)
lz_classify_kernel(const size_t* __restrict__ comp_bytes, const size_t* __restrict__ out_caps, size_t batch, LzLists l) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = lane_id();
  const bool in_range = i < batch;
  const bool light = in_range && lz_chunk_is_light((uint64_t)out_caps[i], (uint64_t)comp_bytes[i]);
  const bool dense = in_range && !light;
  // one atomic per warp and list: the warp's chunks are appended in lane order
  const unsigned ml = __ballot_sync(kFull, light), md = __ballot_sync(kFull, dense);
  unsigned long long bl = 0, bd = 0;
  if (lane == 0) {
    if (ml) bl = atomicAdd(l.ctr + 2, (unsigned long long)__popc(ml));
    if (md) bd = atomicAdd(l.ctr + 3, (unsigned long long)__popc(md));
  }
  bl = __shfl_sync(kFull, bl, 0);
  bd = __shfl_sync(kFull, bd, 0);
  const unsigned below = (1u << lane) - 1u;
  if (light) l.light[bl + __popc(ml & below)] = (uint32_t)i;
  if (dense) l.dense[bd + __popc(md & below)] = (uint32_t)i;
}

// A warp's source of chunk indices: its list (atomic ticket) or, without a workspace, a static stride over the batch
// filtered by class.
struct LzWork {
  const uint32_t* list;
  unsigned long long count;
  unsigned long long* ticket;
  const size_t* comp_bytes;
  const size_t* out_caps;
  size_t batch, static_next, static_stride;
  bool want_light, first;
  __device__ __forceinline__ LzWork(const LzLists& l, bool light, const size_t* cb, const size_t* oc, size_t n,
                                    size_t warp_global, size_t warps_total)
      : list(l.ctr ? (light ? l.light : l.dense) : nullptr), count(l.ctr ? l.ctr[light ? 2 : 3] : 0),
        ticket(l.ctr ? l.ctr + (light ? 0 : 1) : nullptr), comp_bytes(cb), out_caps(oc), batch(n),
        static_next(warp_global), static_stride(warps_total), want_light(light), first(true) {}
  // next chunk of this warp, or batch when there is none
  __device__ __forceinline__ size_t next(int lane) {
    if (list) {
      // The first chunk of every warp is static (list entry = global warp index), the rest come from the ticket.  With
      // fewer chunks than resident warps this packs the work into whole CTAs -- the CTAs behind them retire at once --
      // instead of leaving every resident CTA half idle while it still holds its shared memory and registers.
      unsigned long long t;
      if (first) {
        first = false;
        t = static_next;
      } else {
        t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1ull);
        t = __shfl_sync(kFull, t, 0) + static_stride;
      }
      return t < count ? (size_t)list[t] : batch;
    }
    while (static_next < batch) {
      const size_t c = static_next;
      static_next += static_stride;
      if (lz_chunk_is_light((uint64_t)out_caps[c], (uint64_t)comp_bytes[c]) == want_light) return c;
    }
    return batch;
  }
};

}  // namespace b200

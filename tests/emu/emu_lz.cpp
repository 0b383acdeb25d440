// emu_lz.cpp -- TEST INFRASTRUCTURE: runs the warp-level LZ4 / Snappy chunk decoders of
// nvcomp_b200/csrc (lz_decode.cuh, lz4_decode.cuh, snappy_decode.cuh) inside the host warp emulator.
// Built into tests/emu/libemu_lz.so by the Makefile; loaded only by tests/test_lz_emu.py.
#include "emu_cuda.h"

#include <sys/mman.h>
#include <unistd.h>

#include "lz4_decode.cuh"
#include "snappy_decode.cuh"

namespace {

// A buffer that ends (rounded up to its 16-byte granule) exactly at an inaccessible page, with an
// inaccessible page in front: out-of-bounds plain loads / stores fault instead of passing silently.
struct Guarded {
  uint8_t* map = nullptr;
  size_t map_bytes = 0;
  uint8_t* p = nullptr;
  Guarded(size_t n, unsigned misalign) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t body = ((n + misalign + 15) / 16 * 16 + page - 1) / page * page + page;
    map_bytes = body + 2 * page;
    map = (uint8_t*)mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (map == MAP_FAILED) abort();
    mprotect(map, page, PROT_NONE);
    mprotect(map + page + body, page, PROT_NONE);
    uint8_t* end = map + page + body;
    p = end - (n + misalign + 15) / 16 * 16 + misalign;
    memset(map + page, 0xee, body);
  }
  ~Guarded() { munmap(map, map_bytes); }
};

int run(int codec, int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, unsigned in_mis,
        unsigned out_mis, char* msg, size_t msg_bytes, unsigned long long* n_sync) {
  Guarded gin(n, in_mis & 15u), gout(cap, out_mis & 15u);
  if (n) memcpy(gin.p, src, n);
  emu::Warp w;
  emu::add_region(w, gin.p, n, false);
  emu::add_region(w, gout.p, cap, true);
  uint32_t produced = 0;
  bool ok = false;
  emu::run_warp(w, b200::kLzWarpSmem, [&](int lane) {
    uint8_t* ring = emu::g_warp->smem;
    b200::lz_warp_init(b200::smem_addr(ring), lane);
    uint32_t tma_parity = 0;
    uint32_t prod = 0;
    bool r;
    if (codec == 0) {
      r = mode == 1 ? b200::lz4_decode_chunk_direct(gin.p, (uint32_t)n, gout.p, cap, &prod, lane)
                    : b200::lz4_decode_chunk_v2(gin.p, (uint32_t)n, gout.p, cap, &prod, ring, tma_parity, lane, mode != 2);
    } else {
      r = mode == 1 ? b200::snappy_decode_chunk(gin.p, (uint32_t)n, gout.p, cap, &prod, lane)
                    : b200::snappy_decode_chunk_v2(gin.p, (uint32_t)n, gout.p, cap, &prod, ring, tma_parity, lane, mode != 2);
    }
    if (lane == 0) { ok = r; produced = prod; }
  });
  if (n_sync) *n_sync = w.n_sync;
  if (w.failed) {
    if (msg) snprintf(msg, msg_bytes, "%s", w.fail_msg);
    return -2;
  }
  if (!ok) return -1;
  if (produced > cap) { if (msg) snprintf(msg, msg_bytes, "produced %u > cap %zu", produced, cap); return -2; }
  memcpy(dst, gout.p, produced);
  return (int)produced;
}

}  // namespace

extern "C" {
// codec: 0 = LZ4, 1 = Snappy.  mode: 0 = adaptive (as the kernel), 1 = direct loop, 2 = block decoder forced.
// Returns bytes produced, -1 if the decoder rejected the stream, -2 on an emulator fault (msg says why).
int emu_lz_decode(int codec, int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, unsigned in_mis,
                  unsigned out_mis, char* msg, size_t msg_bytes, unsigned long long* n_sync) {
  return run(codec, mode, src, n, dst, cap, in_mis, out_mis, msg, msg_bytes, n_sync);
}
}

// snappy_decode.cuh -- Snappy raw-format decode for one chunk owned by one warp: the format policy of the
// lane-parallel decoder (lz_decode.cuh), the serial element path and the direct loop for chunks that
// compressed >= 4x.  Kernels and the C ABI are in snappy.cu.
#pragma once

#include "common.cuh"
#include "lz_decode.cuh"

namespace b200 {

// varint32 preamble; returns false when malformed.  Warp-uniform.
__device__ __forceinline__ bool snappy_read_preamble(const uint8_t* __restrict__ in, uint32_t in_n,
                                                     uint32_t& ip, uint64_t& ulen) {
  ulen = 0;
  uint32_t shift = 0;
  while (true) {
    if (ip >= in_n || shift > 28) return false;
    const uint32_t b = in[ip++];
    ulen |= (uint64_t)(b & 0x7fu) << shift;
    if (!(b & 0x80u)) break;
    shift += 7;
  }
  return ulen <= 0xffffffffull;
}

__device__ __forceinline__ bool snappy_decode_chunk(const uint8_t* __restrict__ in, uint32_t in_n,
                                                    uint8_t* out, uint64_t out_cap,
                                                    uint32_t* produced, int lane) {
  uint32_t ip = 0;
  uint64_t ulen;
  if (!snappy_read_preamble(in, in_n, ip, ulen)) return false;
  if (ulen > out_cap) return false;
  const uint32_t n_out = (uint32_t)ulen;
  uint32_t op = 0;
  const uint32_t ul = (uint32_t)lane;
  while (ip < in_n) {
    if (ip + 32u <= in_n) {
      // Window path (typed run-length data: a short literal followed by copies of it).  One coalesced 32-byte load
      // brings the literal element and the copy elements behind it into a register window; when the copy's period
      // (1, 2, 4 or 8 bytes) lies inside the literal, the whole run -- every following copy-2 element of the window
      // with the same offset continues it -- is expanded from the window, no load from the output buffer.
      const uint32_t b = in[ip + ul];
      const uint32_t tag = __shfl_sync(kFull, b, 0);
      if ((tag & 3u) == 0u && (tag >> 2) < 8u) {                 // literal of 1..8 bytes
        const uint32_t ll = (tag >> 2) + 1u;
        const uint32_t t2 = __shfl_sync(kFull, b, (int)(1u + ll));
        const uint32_t k2 = t2 & 3u;
        const uint32_t o_lo = __shfl_sync(kFull, b, (int)(2u + ll)), o_hi = __shfl_sync(kFull, b, (int)(3u + ll));
        uint32_t off, ml, used;
        if (k2 == 1u) { off = ((t2 >> 5) << 8) | o_lo; ml = 4u + ((t2 >> 2) & 7u); used = 3u + ll; }
        else { off = o_lo | (o_hi << 8); ml = (t2 >> 2) + 1u; used = 4u + ll; }
        if ((k2 == 1u || k2 == 2u) && off != 0u && off <= ll && off <= 8u && (off & (off - 1u)) == 0u) {
          // lane i inspects the i-th element behind the first copy
          const uint32_t p = used + 3u * ul;
          const uint32_t e0 = __shfl_sync(kFull, b, (int)(p & 31u)), e1 = __shfl_sync(kFull, b, (int)((p + 1u) & 31u)),
                         e2 = __shfl_sync(kFull, b, (int)((p + 2u) & 31u));
          const bool same = p + 3u <= 32u && (e0 & 3u) == 2u && (e1 | (e2 << 8)) == off;
          const unsigned m = __ballot_sync(kFull, same);
          const uint32_t nf = (uint32_t)__ffs((int)~m) - 1u;     // leading run of continuations (< 10)
          ml += __reduce_add_sync(kFull, ul < nf ? (e0 >> 2) + 1u : 0u);
          used += 3u * nf;
          // every element the window shows continues the run and more input follows: look at the next 32 bytes (ten
          // whole elements) as long as that holds -- a run longer than the ~600 bytes one window spells stays here
          // instead of taking the read-back copy below for its tail
          if (nf != 0u && used + 3u > 32u) {
            while (ip + used + 32u <= in_n) {
              const uint32_t b2 = in[ip + used + ul];
              const uint32_t q = 3u * ul;
              const uint32_t f0 = __shfl_sync(kFull, b2, (int)(q & 31u)), f1 = __shfl_sync(kFull, b2, (int)((q + 1u) & 31u)),
                             f2 = __shfl_sync(kFull, b2, (int)((q + 2u) & 31u));
              const bool same2 = ul < 10u && (f0 & 3u) == 2u && (f1 | (f2 << 8)) == off;
              const uint32_t nf2 = (uint32_t)__ffs((int)~__ballot_sync(kFull, same2)) - 1u;      // <= 10
              if (nf2 == 0u) break;
              ml += __reduce_add_sync(kFull, ul < nf2 ? (f0 >> 2) + 1u : 0u);
              used += 3u * nf2;
              if (nf2 < 10u || ml > 0x10000u) break;
            }
          }
          if (ll <= n_out - op && ml <= n_out - op - ll) {
            if (ul - 1u < ll) out[op + ul - 1u] = (uint8_t)b;    // literals: window lanes 1..ll
            lz_expand_period_from_window(out + op + ll, ml, off, b, 1u + ll - off, ul);
            op += ll + ml;
            ip += used;
            continue;
          }
        }
      }
    }
    const uint32_t tag = in[ip++];
    uint32_t len, off;
    const uint32_t kind = tag & 3u;
    if (kind == 0) {
      len = (tag >> 2) + 1;
      if (len > 60) {
        const uint32_t nb = len - 60;
        if (in_n - ip < nb) return false;
        uint32_t v = 0;
        for (uint32_t i = 0; i < nb; ++i) v |= (uint32_t)in[ip + i] << (8 * i);
        ip += nb;
        if (v == 0xffffffffu) return false;
        len = v + 1;
      }
      if (len > in_n - ip || len > n_out - op) return false;
      warp_copy<true>(out + op, in + ip, len, lane);
      ip += len;
      op += len;
      continue;
    }
    if (kind == 1) {
      if (ip >= in_n) return false;
      len = 4 + ((tag >> 2) & 7u);
      off = ((tag >> 5) << 8) | in[ip++];
    } else if (kind == 2) {
      if (in_n - ip < 2) return false;
      len = (tag >> 2) + 1;
      off = load_u16(in + ip);
      ip += 2;
      // a run of copy-2 elements with the same offset is one long match (64 bytes per element, so
      // only a full-length element can have a continuation): lane i inspects element i, the run is
      // merged and copied once
      if (len == 64u) {
        const uint32_t q = ip + 3u * (uint32_t)lane;
        uint32_t flen = 0;
        bool same = false;
        if (q + 3u <= in_n) {
          const uint32_t t2 = in[q];
          same = ((t2 & 3u) == 2u) && (load_u16(in + q + 1) == off);
          flen = (t2 >> 2) + 1;
        }
        const unsigned m = __ballot_sync(kFull, same);
        const uint32_t nf = (m == kFull) ? 32u : (uint32_t)(__ffs(~m) - 1);
        uint32_t add = ((uint32_t)lane < nf) ? flen : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) add += __shfl_xor_sync(kFull, add, d);
        if (len <= n_out - op && add <= n_out - op - len) { len += add; ip += 3u * nf; }
      }
    } else {
      if (in_n - ip < 4) return false;
      len = (tag >> 2) + 1;
      off = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16)
            | ((uint32_t)in[ip + 3] << 24);
      ip += 4;
    }
    if (off == 0 || off > op || len > n_out - op) return false;
    __syncwarp();
    warp_match_copy(out + op, off, len, lane);
    __syncwarp();
    op += len;
  }
  if (op != n_out) return false;
  *produced = op;
  return true;
}


// ---------------------------------------------------------------------------
// v2 decode (lz_decode.cuh): lane-parallel short-element path + this slow path
// ---------------------------------------------------------------------------
struct SnappyDecode : SnappyPolicy {
  __device__ static __forceinline__ bool at_end(const LzState& s) { return s.ip >= s.in_n; }
  // one element (literal or copy).  A run of copy-2 elements with the same offset -- how Snappy
  // spells one long match (64 bytes per element) -- is merged and emitted as a single match.
  __device__ static __forceinline__ int serial_token(LzState& s, int lane) {
    const uint8_t* __restrict__ in = s.in;
    const uint32_t in_n = s.in_n;
    uint32_t ip = s.ip;
    const uint32_t n_out = (uint32_t)s.out_cap;
    const uint32_t tag = in[ip++];
    const uint32_t kind = tag & 3u;
    uint32_t len, off;
    if (kind == 0) {
      len = (tag >> 2) + 1;
      if (len > 60) {
        const uint32_t nb = len - 60;
        if (in_n - ip < nb) return -1;
        uint32_t v = 0;
        for (uint32_t i = 0; i < nb; ++i) v |= (uint32_t)in[ip + i] << (8 * i);
        ip += nb;
        if (v == 0xffffffffu) return -1;
        len = v + 1;
      }
      if (len > in_n - ip || len > n_out - s.op) return -1;
      lz_serial_lookahead<SnappyPolicy>(s, ip + len, lane);
      lz_emit_literals(s, in + ip, len, lane);
      s.ip = ip + len;
      return 1;
    }
    if (kind == 1) {
      if (ip >= in_n) return -1;
      len = 4 + ((tag >> 2) & 7u);
      off = ((tag >> 5) << 8) | in[ip++];
    } else if (kind == 2) {
      if (in_n - ip < 2) return -1;
      len = (tag >> 2) + 1;
      off = load_u16(in + ip);
      ip += 2;
      // merge following copy-2 elements with the same offset (lane i inspects element i); only a
      // full-length element can have a continuation
      if (len == 64u) {
        const uint32_t q = ip + 3u * (uint32_t)lane;
        uint32_t flen = 0;
        bool same = false;
        if (q + 3u <= in_n) {
          const uint32_t t2 = in[q];
          same = ((t2 & 3u) == 2u) && (load_u16(in + q + 1) == off);
          flen = (t2 >> 2) + 1;
        }
        const unsigned m = __ballot_sync(kFull, same);
        const uint32_t nf = (m == kFull) ? 32u : (uint32_t)(__ffs(~m) - 1);
        uint32_t add = ((uint32_t)lane < nf) ? flen : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) add += __shfl_xor_sync(kFull, add, d);
        if (add <= n_out - s.op - min(len, n_out - s.op)) { len += add; ip += 3u * nf; }
      }
    } else {
      if (in_n - ip < 4) return -1;
      len = (tag >> 2) + 1;
      off = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16)
            | ((uint32_t)in[ip + 3] << 24);
      ip += 4;
    }
    if (off == 0 || off > s.op || len > n_out - s.op) return -1;
    lz_serial_lookahead<SnappyPolicy>(s, ip, lane);
    lz_emit_match(s, off, len, lane);
    s.ip = ip;
    return 1;
  }
};

__device__ __forceinline__ bool snappy_decode_chunk_v2(const uint8_t* in, uint32_t in_n, uint8_t* out,
                                                       uint64_t out_cap, uint32_t* produced,
                                                       uint8_t* ring, uint32_t& tma_parity, int lane, bool allow_direct = true) {
  uint32_t ip = 0;
  uint64_t ulen;
  if (!snappy_read_preamble(in, in_n, ip, ulen)) return false;
  if (ulen > out_cap) return false;
  // Adaptive strategy (see lz4.cu): chunks that compressed >= 4x are long-match dominated and
  // take the direct global-memory token loop.
  if (allow_direct && ulen >= 4ull * in_n) return snappy_decode_chunk(in, in_n, out, out_cap, produced, lane);
  LzState s;
  s.in = in; s.in_n = in_n; s.out = out; s.out_cap = ulen;
  s.ip = ip; s.op = 0; s.flushed = 0; s.ring_lo = 0;
  s.align = (uint32_t)((uintptr_t)out & 15u);
  s.ring = smem_addr(ring);
  s.cur = 0; s.pf_ip = kNoPrefetch; s.parity = tma_parity; s.next = kNextUnknown;
  const bool ok = lz_decode_stream<SnappyDecode>(s, lane);
  tma_parity = s.parity;                 // the barrier outlives the chunk: carry its phase to the next one
  if (!ok) return false;
  if (s.op != (uint32_t)ulen) return false;
  *produced = s.op;
  return true;
}

}  // namespace b200

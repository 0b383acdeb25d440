// lz_decode.cuh -- block-parallel LZ77 (LZ4 / Snappy) chunk decoder for B200.
//
// One warp owns one chunk.  On tabular data a 64 KB chunk holds 10-15 thousand *short* tokens
// (4-8 output bytes each), so throughput is bounded by warp-instructions per token, not bytes.
//
//  * BLOCK PATH (lz_block): 1 KB of compressed input is staged in shared memory; every lane finds the
//    token chain through its own 32-byte segment (exit table computed right to left, entries resolved
//    across lanes), the ~350-500 tokens of the block are listed in stream order and executed 32 per
//    step, one token per lane: literals from the staged block, matches in dependency rounds.
//  * The most recent 4 KB of output live in a per-warp shared-memory ring (explicit 32-bit shared
//    addressing), so match sources are read at shared-memory latency; completed 512-byte blocks are
//    flushed to HBM with 16-byte aligned vector stores (full-line writes, DRAM traffic == algorithmic
//    bytes).  Matches that reach further back than the ring read the flushed bytes from global memory.
//  * SERIAL PATH (P::serial_token + lz_emit_*): tokens with length-extension bytes / long lengths are
//    parsed once by the whole warp; up to 192 bytes they are executed inside the ring, longer runs go
//    straight to global memory as 16-byte vectors (periodic runs are built in registers, no
//    store->load round trip) and the ring restarts empty behind them.
//  * Chunks that compressed >= 4x never enter this machinery: the callers (lz4_decode.cuh /
//    snappy_decode.cuh) decode them with the direct global-memory token loop, and the kernels hand
//    dense chunks out first (two-pass ticket).
//
// Format specifics (token grammar, stream end, size limits) come from a policy.
#pragma once

#include "common.cuh"

namespace b200 {

constexpr uint32_t kRingBytes = 4096;
constexpr uint32_t kRingMask = kRingBytes - 1;
constexpr uint32_t kFlushBlock = 512;
// A match source is served from the ring only if it is younger than this many bytes
// (ring size minus the largest output one execution step can append, minus alignment slack).
constexpr uint32_t kRingReach = kRingBytes - 1024 - 16;

struct LzState {
  const uint8_t* in;
  uint32_t in_n;
  uint8_t* out;        // chunk output base (any alignment)
  uint64_t out_cap;    // capacity (LZ4) or exact size (Snappy)
  uint32_t ip;         // input cursor
  uint32_t op;         // output cursor (bytes produced)
  uint32_t flushed;    // output bytes already in global memory
  uint32_t ring_lo;    // lowest output offset whose bytes are valid in the ring
  uint32_t align;      // (uintptr_t)out & 15: ring index = (offset + align) & mask
  uint32_t ring;       // shared-window address of the kRingBytes ring (32-bit: LDS/STS with immediates)
};

__device__ __forceinline__ uint32_t ring_idx(const LzState& s, uint32_t off) {
  return (off + s.align) & kRingMask;
}
__device__ __forceinline__ uint32_t ring_ld(const LzState& s, uint32_t off) { return lds_u8(s.ring + ring_idx(s, off)); }
__device__ __forceinline__ void ring_st(const LzState& s, uint32_t off, uint32_t v) { sts_u8(s.ring + ring_idx(s, off), v); }
// Write ring bytes [s.flushed, upto) to global memory.  Vector stores where the global
// address is 16-byte aligned, byte stores for ragged ends.
__device__ __forceinline__ void lz_flush(LzState& s, uint32_t upto, int lane) {
  uint32_t f = s.flushed;
  if (upto <= f) return;
  __syncwarp();
  // ragged head up to the next 16-byte boundary (in aligned space)
  uint32_t head = (16u - ((f + s.align) & 15u)) & 15u;
  if (head > upto - f) head = upto - f;
  if ((uint32_t)lane < head) s.out[f + lane] = (uint8_t)ring_ld(s, f + lane);
  f += head;
  const uint32_t nvec = (upto - f) >> 4;
  for (uint32_t v = lane; v < nvec; v += kWarp) {
    const uint32_t o = f + (v << 4);
    const uint4 d = lds_v4(s.ring + ring_idx(s, o));
    st_v4((uint4*)(s.out + o), d);
  }
  f += nvec << 4;
  const uint32_t tail = upto - f;
  if ((uint32_t)lane < tail) s.out[f + lane] = (uint8_t)ring_ld(s, f + lane);
  s.flushed = upto;
}

// Flush every completed 512-byte block (keeps global stores full-line).
__device__ __forceinline__ void lz_flush_blocks(LzState& s, int lane) {
  const uint32_t lim = ((s.op + s.align) & ~(kFlushBlock - 1));
  if (lim > s.flushed + s.align) lz_flush(s, lim - s.align, lane);
}

// ---------------------------------------------------------------------------
// Format policies.  A "token" is one LZ4 sequence (literals + match) or one Snappy element (literal
// OR copy).  Everything the block parser needs is a pure function of the token's first byte:
//   tok_size(b): bytes from this token to the next one (<= kSegBytes), kTokStop for a token the
//                lane-parallel path does not take (length-extension bytes, long literals, copy-4)
//   tok_out(b):  output bytes the token produces
// fields() extracts literal length / match length / offset for execution.
// ---------------------------------------------------------------------------
constexpr uint32_t kTokStop = 64;

// 4 bytes at an arbitrary position of a shared-memory buffer (two aligned words + funnel shift)
__device__ __forceinline__ uint32_t lds_u32_any(uint32_t base, uint32_t pos) {
  const uint32_t a = base + (pos & ~3u);
  return __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (pos & 3u) * 8u);
}

struct Lz4Policy {
  __device__ static __forceinline__ uint32_t tok_size(uint32_t b) {
    const uint32_t L = b >> 4;
    return (L == 15u || (b & 15u) == 15u) ? kTokStop : 3u + L;
  }
  __device__ static __forceinline__ uint32_t tok_out(uint32_t b) { return (b >> 4) + (b & 15u) + 4u; }
  // blk: shared address of the staged block, pos: position of the token in it
  __device__ static __forceinline__ void fields(uint32_t blk, uint32_t pos, uint32_t& L, uint32_t& M,
                                                uint32_t& off, uint32_t& lit_at) {
    const uint32_t x = lds_u32_any(blk, pos);
    L = (x >> 4) & 15u;
    M = (x & 15u) + 4u;
    lit_at = pos + 1u;
    off = (x >> 8) & 0xffffu;
    if (L) off = lds_u32_any(blk, pos + 1u + L) & 0xffffu;
  }
  // does the token starting with byte b0 need the serial path?
  __device__ static __forceinline__ bool is_stop(uint32_t b0) { return tok_size(b0) == kTokStop; }
};

struct SnappyPolicy {
  __device__ static __forceinline__ uint32_t tok_size(uint32_t b) {
    const uint32_t kind = b & 3u, h = b >> 2;
    if (kind == 0u) return h < 31u ? h + 2u : kTokStop;      // literal of h+1 <= 31 bytes
    return kind == 3u ? kTokStop : kind + 1u;                // copy-1: 2 bytes, copy-2: 3 bytes
  }
  __device__ static __forceinline__ uint32_t tok_out(uint32_t b) {
    const uint32_t kind = b & 3u, h = b >> 2;
    return kind == 1u ? 4u + (h & 7u) : h + 1u;
  }
  __device__ static __forceinline__ void fields(uint32_t blk, uint32_t pos, uint32_t& L, uint32_t& M,
                                                uint32_t& off, uint32_t& lit_at) {
    const uint32_t x = lds_u32_any(blk, pos);
    const uint32_t kind = x & 3u, h = (x >> 2) & 63u;
    lit_at = pos + 1u;
    L = kind == 0u ? h + 1u : 0u;
    M = kind == 0u ? 0u : (kind == 1u ? 4u + (h & 7u) : h + 1u);
    off = kind == 1u ? (((x >> 5) & 7u) << 8) | ((x >> 8) & 255u) : (x >> 8) & 0xffffu;
  }
  __device__ static __forceinline__ bool is_stop(uint32_t b0) { return tok_size(b0) == kTokStop; }
};

// ---------------------------------------------------------------------------
// Block path.  One call parses up to kBlkBytes of compressed input and executes its tokens.
//
//   1. stage   lane l loads its 32-byte segment of the block (16-byte aligned base) and stores it to
//              shared memory; kBlkPad more bytes cover tokens that start in the last segment.
//   2. chain   every lane computes, right to left over its own 32 bytes (held in registers, fully
//              unrolled), where a token chain entering its segment at byte p leaves it: a 32-entry
//              exit table per lane.  The true entry of every segment then follows by walking the 32
//              tables from the known entry of segment 0.  No speculation, no retries.
//   3. walk    each lane walks the tokens of its segment twice: count them / sum their output
//              lengths, and after one warp scan write one record per token (block position | output
//              position) in stream order.
//   4. execute 32 consecutive tokens per step, one per lane: literals, then matches in dependency
//              rounds: a match runs as soon as the tokens of this step that produce its source bytes
//              have run (sources below the step are final).
// ---------------------------------------------------------------------------
constexpr uint32_t kSegBytes = 32;
constexpr uint32_t kBlkBytes = 32 * kSegBytes;
constexpr uint32_t kBlkPad = 32;
constexpr uint32_t kBlkStage = kBlkBytes + kBlkPad;
constexpr uint32_t kMaxStepOut = 1024;              // output bytes one step may append (ring reach depends on it)
constexpr uint32_t kSmemIn = kRingBytes;            // staged block
constexpr uint32_t kSmemRec = kSmemIn + kBlkStage;  // token records (4 B each); the exit tables alias them
constexpr uint32_t kRecBytes = 4 * (kBlkBytes / 2); // a token is at least 2 bytes
constexpr uint32_t kLzWarpSmem = kSmemRec + kRecBytes;
static_assert(kRingReach + kMaxStepOut + 16 <= kRingBytes, "ring reach");
static_assert(kSmemRec % 16 == 0, "record alignment");

// Returns the number of tokens retired (0: nothing done, the caller takes the serial path), -1 on a
// malformed stream.
template <class P>
__device__ __forceinline__ int lz_block(LzState& s, int lane) {
  const uint32_t ul = (uint32_t)lane;
  const uint8_t* const ipp = s.in + s.ip;
  const uint32_t mis = (uint32_t)((uintptr_t)ipp & 15u);
  const uint32_t avail = s.in_n - s.ip + mis;                 // bytes from the aligned base to the stream end
  if (avail < kSegBytes + kBlkPad) return 0;
  uint32_t nl = (avail - kBlkPad) / kSegBytes;                // segments that lie (with the pad) inside the stream
  if (nl > 32u) nl = 32u;
  const uint8_t* const abase = ipp - mis;
  const uint32_t blk = s.ring + kSmemIn, rec = s.ring + kSmemRec;

  // ---- 1. stage ------------------------------------------------------------------------------
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
  if (ul < nl) {
    a0 = ld_nc_v4((const uint4*)(abase + kSegBytes * ul));
    a1 = ld_nc_v4((const uint4*)(abase + kSegBytes * ul + 16u));
    sts_v4(blk + kSegBytes * ul, a0);
    sts_v4(blk + kSegBytes * ul + 16u, a1);
  }
  if (ul < kBlkPad / 16u) {
    const uint4 pad = ld_nc_v4((const uint4*)(abase + kSegBytes * nl + 16u * ul));
    sts_v4(blk + kSegBytes * nl + 16u * ul, pad);
  }
  // ---- 2. chain: exit table of this lane's segment ---------------------------------------------
  {
    const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const uint32_t ex = rec + kSegBytes * ul;
#pragma unroll
    for (int p = 31; p >= 0; --p) {
      const uint32_t b = (w[p >> 2] >> (8 * (p & 3))) & 255u;
      const uint32_t sz = P::tok_size(b);
      const uint32_t q = (uint32_t)p + sz;
      uint32_t code;
      if (q >= kSegBytes) code = (sz == kTokStop) ? 0xffu : q - kSegBytes;
      else code = lds_u8(ex + q);
      sts_u8(ex + (uint32_t)p, code);
    }
  }
  __syncwarp();
  // entry of every segment: walk the tables from the known entry of segment 0 (warp-uniform)
  uint32_t e = mis, my_e = 0, stop_lane = 32u;
  for (uint32_t l = 0; l < nl; ++l) {
    if (ul == l) my_e = e;
    const uint32_t x = lds_u8(rec + kSegBytes * l + e);
    if (x == 0xffu) { stop_lane = l; break; }
    e = x;
  }
  // ---- 3. walk: token count / output bytes of this lane's segment -------------------------------
  const bool active = ul < nl && ul <= stop_lane;
  uint32_t p = my_e, cnt = 0, osum = 0;
  if (active) {
    while (p < kSegBytes) {
      const uint32_t b = lds_u8(blk + kSegBytes * ul + p);
      const uint32_t sz = P::tok_size(b);
      if (sz == kTokStop) break;
      ++cnt;
      osum += P::tok_out(b);
      p += sz;
    }
  }
  // block end: the stop token, or where the chain leaves the last segment
  const uint32_t end_pos = stop_lane < 32u ? __shfl_sync(kFull, kSegBytes * ul + p, (int)stop_lane)
                                           : kSegBytes * nl + e;
  uint32_t incl = cnt | (osum << 10);                           // cnt <= 16 per lane, osum <= 16 * 64
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += o;
  }
  const uint32_t tot = __shfl_sync(kFull, incl, 31);
  const uint32_t N = tot & 1023u, total_out = tot >> 10;
  if (N == 0u) return 0;
  if ((uint64_t)s.op + total_out > s.out_cap) return 0;         // the serial path finds the exact error
  __syncwarp();                                                 // exit tables are dead: records overwrite them
  {
    uint32_t t = (incl & 1023u) - cnt, run = (incl >> 10) - osum, q = my_e;
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint32_t b = lds_u8(blk + kSegBytes * ul + q);
      sts_u32(rec + 4u * t, (kSegBytes * ul + q) | (run << 10));
      run += P::tok_out(b);
      q += P::tok_size(b);
      ++t;
    }
  }
  __syncwarp();

  // ---- 4. execute ------------------------------------------------------------------------------
  // Output positions are kept in "aligned space" (offset + s.align): the ring index is (pos & mask)
  // and (s.out - s.align)[pos] is the global address.
  const uint32_t rbase = s.ring;
  const uint8_t* const outa = s.out - s.align;
  const uint32_t out0 = s.op + s.align;
  uint32_t t0 = 0;
  while (t0 < N) {
    const uint32_t t = t0 + ul;
    bool valid = t < N;
    const uint32_t r = lds_u32(rec + 4u * (valid ? t : t0));
    const uint32_t pos = r & 1023u;
    const uint32_t dst = out0 + (r >> 10);
    uint32_t L, M, off, lit_at;
    P::fields(blk, pos, L, M, off, lit_at);
    const uint32_t step_lo = __shfl_sync(kFull, dst, 0);
    // tokens of this step: the leading run that ends within kMaxStepOut bytes (never empty)
    const unsigned fitm = __ballot_sync(kFull, valid && dst + L + M - step_lo <= kMaxStepOut);
    const uint32_t n = fitm == kFull ? 32u : (uint32_t)__ffs((int)~fitm) - 1u;
    valid = ul < n;
    if (!valid) { L = 0; M = 0; }
    const uint32_t step_hi = __shfl_sync(kFull, dst + L + M, (int)n - 1);
    const uint32_t o_mat = dst + L;
    const bool has = M != 0u;
    if (__any_sync(kFull, has && (off == 0u || off > o_mat - s.align))) return -1;
    // literals: straight from the staged block
    for (uint32_t j = 0; __any_sync(kFull, j < L); ++j)
      if (j < L) sts_u8(rbase + ((dst + j) & kRingMask), lds_u8(blk + lit_at + j));
    __syncwarp();
    // matches
    const unsigned hasm = __ballot_sync(kFull, has);
    if (hasm) {
      const uint32_t cur_op = step_lo - s.align;
      const uint32_t ring_from = max(s.ring_lo, cur_op > kRingReach ? cur_op - kRingReach : 0u) + s.align;
      const uint32_t src = o_mat - off;
      const uint32_t src_end = src + min(M, off);              // exclusive end of the bytes this match reads
      // which tokens of this step produce my source bytes?  Token ranges are consecutive, so the
      // producers are the lanes from the one holding byte max(src, step_lo) to the one holding src_end-1.
      unsigned dep = 0;
      const bool inwin = has && src_end > step_lo;
      if (__any_sync(kFull, inwin)) {
        const uint32_t key = valid ? dst : 0xffffffffu;
        const uint32_t qa = max(src, step_lo), qb = src_end - 1u;
        uint32_t ja = 0, jb = 0;
#pragma unroll
        for (uint32_t st = 16; st; st >>= 1) {
          const uint32_t va = __shfl_sync(kFull, key, (int)(ja + st));
          const uint32_t vb = __shfl_sync(kFull, key, (int)(jb + st));
          if (va <= qa) ja += st;
          if (vb <= qb) jb += st;
        }
        // own literals precede the own match in program order: drop the self bit
        if (inwin) dep = ((2u << jb) - 1u) & ~((1u << ja) - 1u) & ~(1u << ul);
      }
      const uint32_t didx = o_mat & kRingMask, sidx = src & kRingMask;
      const bool in_ring = src >= ring_from && sidx + M + 4u <= kRingBytes;
      const bool far = src + M <= ring_from;                   // flushed long ago: read from global memory
      // groups of four bytes are loaded, then stored: needs off >= 4 (a shorter period takes the byte loop)
      const bool simple = has && off >= 4u && didx + M <= kRingBytes && (in_ring || far);
      const uint32_t dp = rbase + didx, sp = rbase + sidx;
      const uint8_t* const gp = outa + src;
      unsigned done = ~hasm;
      bool pend = has;
      while (true) {
        const bool ready = pend && (dep & ~done) == 0u;
        const unsigned rm = __ballot_sync(kFull, ready);
        const bool fr = ready && simple && !far, fg = ready && simple && far, gen = ready && !simple;
        if (fr) {
          const uint32_t x0 = lds_u8<0>(sp), x1 = lds_u8<1>(sp), x2 = lds_u8<2>(sp), x3 = lds_u8<3>(sp);
          sts_u8<0>(dp, x0);
          if (M > 1u) sts_u8<1>(dp, x1);
          if (M > 2u) sts_u8<2>(dp, x2);
          if (M > 3u) sts_u8<3>(dp, x3);
        }
        for (uint32_t g = 4; __any_sync(kFull, fr && M > g); g += 4) {
          if (fr && M > g) {
            const uint32_t s4 = sp + g, d4 = dp + g;
            const uint32_t x0 = lds_u8<0>(s4), x1 = lds_u8<1>(s4), x2 = lds_u8<2>(s4), x3 = lds_u8<3>(s4);
            sts_u8<0>(d4, x0);
            if (M > g + 1u) sts_u8<1>(d4, x1);
            if (M > g + 2u) sts_u8<2>(d4, x2);
            if (M > g + 3u) sts_u8<3>(d4, x3);
          }
        }
        if (__any_sync(kFull, fg)) {
          for (uint32_t g = 0; __any_sync(kFull, fg && M > g); g += 4) {
            if (fg && M > g) {
              const uint8_t* const g4 = gp + g;
              const uint32_t d4 = dp + g;
              uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
              x0 = ldg_u8<0>(g4);
              if (M > g + 1u) x1 = ldg_u8<1>(g4);
              if (M > g + 2u) x2 = ldg_u8<2>(g4);
              if (M > g + 3u) x3 = ldg_u8<3>(g4);
              sts_u8<0>(d4, x0);
              if (M > g + 1u) sts_u8<1>(d4, x1);
              if (M > g + 2u) sts_u8<2>(d4, x2);
              if (M > g + 3u) sts_u8<3>(d4, x3);
            }
          }
        }
        if (__any_sync(kFull, gen)) {
          // short periods, ring wrap-around, sources straddling the flushed boundary: byte by byte,
          // in order (a byte may read what this loop wrote off bytes earlier)
          if (gen) {
            for (uint32_t j = 0; j < M; ++j) {
              const uint32_t q = src + j;
              const uint32_t b = (q >= ring_from) ? lds_u8(rbase + (q & kRingMask)) : (uint32_t)outa[q];
              sts_u8(rbase + ((o_mat + j) & kRingMask), b);
            }
          }
        }
        __syncwarp();
        done |= rm;
        pend = pend && !ready;
        if (done == kFull) break;
      }
    }
    s.op += step_hi - step_lo;
    t0 += n;
    lz_flush_blocks(s, lane);
  }
  s.ip += end_pos - mis;
  return (int)N;
}

// ---------------------------------------------------------------------------
// Medium tokens (too long for the lane-parallel path, L + M <= kMediumMax): executed by the
// whole warp one token at a time but still inside the ring, so the data stays at shared-memory
// latency and later short matches keep hitting the ring.
// ---------------------------------------------------------------------------
constexpr uint32_t kMediumMax = 192;

__device__ __forceinline__ void ring_put_literals(LzState& s, uint32_t dst, const uint8_t* __restrict__ src,
                                                  uint32_t n, int lane) {
  for (uint32_t i = lane; i < n; i += kWarp) ring_st(s, dst + i, src[i]);
}

// dst[0..n) = dst[-off..] with LZ77 semantics, all inside the ring (far sources from global).
__device__ __forceinline__ void ring_match(LzState& s, uint32_t dst, uint32_t off, uint32_t n,
                                           uint32_t ring_from, int lane) {
  const uint32_t src = dst - off;
  if (off >= 32u) {
    // bytes of round k only depend on bytes written in rounds < k
    for (uint32_t base = 0; base < n; base += kWarp) {
      const uint32_t j = base + lane;
      if (j < n) {
        const uint32_t sp = src + j;
        const uint32_t b = (sp >= ring_from) ? ring_ld(s, sp) : (uint32_t)s.out[sp];
        ring_st(s, dst + j, b);
      }
      __syncwarp();
    }
  } else {
    // short period: every byte is src[j mod off], all final before the copy starts
    uint32_t r = (uint32_t)lane % off;
    const uint32_t step = 32u % off;
    for (uint32_t j = lane; j < n; j += kWarp) {
      const uint32_t sp = src + r;
      const uint32_t b = (sp >= ring_from) ? ring_ld(s, sp) : (uint32_t)s.out[sp];
      ring_st(s, dst + j, b);
      r += step;
      if (r >= off) r -= off;
    }
  }
}

__device__ __forceinline__ uint32_t ring_from_of(const LzState& s) {
  return max(s.ring_lo, s.op > kRingReach ? s.op - kRingReach : 0u);
}

// one already-produced output byte, wherever it currently lives
__device__ __forceinline__ uint32_t lz_out_byte(const LzState& s, uint32_t pos, uint32_t ring_from) {
  return (pos >= ring_from) ? ring_ld(s, pos) : (uint32_t)s.out[pos];
}

// ---------------------------------------------------------------------------
// Serial (one token at a time, whole warp) emitters for tokens the lane-parallel path cannot
// take.  Up to kMediumMax bytes stay inside the ring; longer runs go straight to global
// memory as 16-byte vectors and the ring restarts empty behind them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void lz_emit_literals(LzState& s, const uint8_t* __restrict__ src, uint32_t n, int lane) {
  if (n <= kMediumMax) {
    ring_put_literals(s, s.op, src, n, lane);
    s.op += n;
    return;
  }
  lz_flush(s, s.op, lane);
  warp_copy<true>(s.out + s.op, src, n, lane);
  s.op += n;
  s.flushed = s.op;
  s.ring_lo = s.op;
}

__device__ __forceinline__ void lz_emit_match(LzState& s, uint32_t off, uint32_t n, int lane) {
  __syncwarp();
  if (n <= kMediumMax) {
    ring_match(s, s.op, off, n, ring_from_of(s), lane);
    s.op += n;
    return;
  }
  const uint32_t dst = s.op;
  if (off <= 16u && (off & (off - 1u)) == 0u) {
    // Long run with a period that divides 16 (typed run-length data).  Every 16-byte aligned
    // vector of the run is the same: build it once in registers from the period bytes (ring or
    // global), no store->load round trip, then stream it out with vector stores.
    const uint32_t rf = ring_from_of(s);
    const uint32_t src = dst - off, m = off - 1u;
    const uint32_t head = (16u - ((dst + s.align) & 15u)) & 15u;
    uint32_t wv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc |= lz_out_byte(s, src + ((head + 4u * q + i) & m), rf) << (8 * i);
      wv[q] = acc;
    }
    const uint32_t hb = lz_out_byte(s, src + ((uint32_t)lane & m), rf);
    lz_flush(s, dst, lane);                                   // everything before the run is now in global memory
    uint8_t* o = s.out + dst;
    if ((uint32_t)lane < head) o[lane] = (uint8_t)hb;
    const uint32_t nvec = (n - head) >> 4;
    uint4* d16 = (uint4*)(o + head);
    const uint4 pat = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    for (uint32_t v = lane; v < nvec; v += kWarp) st_v4(d16 + v, pat);
    const uint32_t j = head + (nvec << 4) + lane;             // < 16 tail bytes
    if (j < n) {
      const uint32_t k = (j - head) & 15u;                    // position inside the pattern vector
      const uint32_t q = k >> 2;
      const uint32_t wsel = q == 0 ? wv[0] : q == 1 ? wv[1] : q == 2 ? wv[2] : wv[3];
      o[j] = (uint8_t)(wsel >> (8 * (k & 3u)));
    }
  } else {
    lz_flush(s, dst, lane);
    __syncwarp();
    warp_match_copy(s.out + dst, off, n, lane);
  }
  __syncwarp();
  s.op += n;
  s.flushed = s.op;
  s.ring_lo = s.op;
}

// Decode driver shared by LZ4 and Snappy.  P::serial_token(s, lane) executes exactly one token
// at s.ip with the emitters above and returns 1 (continue), 2 (stream finished) or -1 (malformed).
template <class P>
__device__ __forceinline__ bool lz_decode_stream(LzState& s, int lane) {
  while (true) {
    if (P::at_end(s)) break;
    // a token that needs the serial path is recognised from its first byte: do not pay for a
    // block parse that would retire nothing
    if (s.in_n - s.ip >= kSegBytes + kBlkPad && !P::is_stop(s.in[s.ip])) {
      const int r = lz_block<P>(s, lane);
      if (r < 0) return false;
      if (r > 0) continue;
    }
    const int r = P::serial_token(s, lane);
    if (r < 0) return false;
    lz_flush_blocks(s, lane);
    if (r == 2) break;
  }
  lz_flush(s, s.op, lane);
  return true;
}

}  // namespace b200

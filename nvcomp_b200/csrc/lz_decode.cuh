// lz_decode.cuh -- lane-parallel LZ77 (LZ4 / Snappy) chunk decoder for B200.
//
// One warp owns one chunk.  On tabular data a 64 KB chunk holds 10-15 thousand *short* tokens
// (4-8 output bytes each), so throughput is bounded by warp-instructions per token, not bytes:
//
//  * FAST PATH (short tokens, lz_fast_iter).  The 32 lanes look at 32 consecutive input bytes.
//    Every lane parses the byte under it as if it were a token start (speculative parse), the true
//    token chain is recovered with 4 rounds of pointer doubling (__reduce_or_sync + __shfl_sync), a
//    warp scan of the token output lengths gives every token its output position, literals are
//    scattered in one pass straight from registers, matches whose sources are already final are
//    copied one per lane with branch-free tiers, the few that depend on output of the same window
//    are retired in order (whole warp, one byte per lane).  ~14 tokens retire per iteration.
//  * The most recent 4 KB of output live in a per-warp shared-memory ring (explicit 32-bit shared
//    addressing), so match sources are read at shared-memory latency; completed 512-byte blocks are
//    flushed to HBM with 16-byte aligned vector stores (full-line writes, DRAM traffic == algorithmic
//    bytes).  Matches that reach further back than the ring read the flushed bytes from global memory.
//  * SERIAL PATH (P::serial_token + lz_emit_*): tokens with length-extension bytes / long lengths are
//    parsed once by the whole warp; up to 192 bytes they are executed inside the ring, longer runs go
//    straight to global memory as 16-byte vectors (periodic runs are built in registers, no
//    store->load round trip) and the ring restarts empty behind them.
//  * Chunks that compressed >= 4x never enter this machinery: the callers (lz4.cu / snappy.cu) decode
//    them with the direct global-memory token loop, and hand dense chunks out first (two-pass ticket).
//
// Format specifics (token grammar, stream end, size limits) come from a Parse policy.
#pragma once

#include "common.cuh"

namespace b200 {

constexpr uint32_t kRingBytes = 4096;
constexpr uint32_t kRingMask = kRingBytes - 1;
constexpr uint32_t kFlushBlock = 512;
constexpr int kParRounds = 3;          // parallel match rounds per window before in-order retirement
// A match source is served from the ring only if it is younger than this many bytes
// (ring size minus the largest output one fast iteration can append, minus alignment slack).
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
// explicit shared-space accesses on 32-bit addresses (the generic-pointer form costs 64-bit address
// arithmetic and generic LD/ST on every byte)
template <int O = 0>
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(O) : "memory");
  return v;
}
template <int O = 0>
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0+%1], %2;" :: "r"(a), "n"(O), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t ring_ld(const LzState& s, uint32_t off) { return lds_u8(s.ring + ring_idx(s, off)); }
__device__ __forceinline__ void ring_st(const LzState& s, uint32_t off, uint32_t v) { sts_u8(s.ring + ring_idx(s, off), v); }
template <int O>
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.u8 %0, [%1+%2];" : "=r"(v) : "l"(p), "n"(O) : "memory");
  return v;
}

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
// Parse policies.  parse(b0, p, lane) classifies the byte at window position `lane`
// as a token start: literal length L, match length M, bytes to the next token, where
// the 2-byte offset sits (rel. to the token) and whether the token needs the slow path.
// ---------------------------------------------------------------------------
struct Tok {
  uint32_t L, M, size, off_at;   // off_at: offset field position relative to token start (0 = none)
  bool stop;
  uint32_t kind;                 // format-private
};

struct Lz4Policy {
  static constexpr uint32_t kLook = 80;        // fast path needs ip + kLook <= in_n
  static constexpr uint32_t kMaxFastM = 18;
  __device__ static __forceinline__ Tok parse(uint32_t b0) {
    Tok t;
    t.L = b0 >> 4;
    const uint32_t mn = b0 & 15u;
    t.M = mn + 4;
    t.stop = (t.L == 15u) | (mn == 15u);
    t.size = 3 + t.L;
    t.off_at = 1 + t.L;
    t.kind = 0;
    return t;
  }
  __device__ static __forceinline__ uint32_t offset(const uint8_t* __restrict__ p, const Tok& t) {
    return load_u16(p + t.off_at);
  }
  // does the token starting with byte b0 need the medium / long path?
  __device__ static __forceinline__ bool is_stop(uint32_t b0) { return (b0 >> 4) == 15u || (b0 & 15u) == 15u; }
};

struct SnappyPolicy {
  static constexpr uint32_t kLook = 96;
  static constexpr uint32_t kMaxFastM = 18;
  __device__ static __forceinline__ Tok parse(uint32_t b0) {
    Tok t;
    const uint32_t kind = b0 & 3u, hi = b0 >> 2;
    t.kind = kind;
    t.L = 0; t.M = 0; t.off_at = 1; t.stop = false;
    if (kind == 0) {
      t.L = hi + 1; t.size = 1 + t.L; t.off_at = 0;
      t.stop = t.L > 28u;
    } else if (kind == 1) {
      t.M = 4 + (hi & 7u); t.size = 2;
    } else if (kind == 2) {
      t.M = hi + 1; t.size = 3;
      t.stop = t.M > kMaxFastM;
    } else {
      t.size = 5; t.stop = true;
    }
    return t;
  }
  __device__ static __forceinline__ uint32_t offset(const uint8_t* __restrict__ p, const Tok& t) {
    if (t.kind == 1) return ((uint32_t)(p[0] >> 5) << 8) | p[1];
    return load_u16(p + 1);
  }
  __device__ static __forceinline__ bool is_stop(uint32_t b0) { return parse(b0).stop; }
};

// One fast-path iteration.  Returns number of tokens retired (0: the token at s.ip needs
// the slow path), or -1 on a malformed stream.
template <class P>
__device__ __forceinline__ int lz_fast_iter(LzState& s, int lane) {
  const uint8_t* __restrict__ win = s.in + s.ip;
  const uint32_t b0 = win[lane];
  const Tok t = P::parse(b0);
  // --- token chain by pointer doubling -------------------------------------------------
  const unsigned stopmask = __ballot_sync(kFull, t.stop);
  const uint32_t nxt0 = t.stop ? 64u : (uint32_t)lane + t.size;    // >= 32: leaves the window
  uint32_t nxt = nxt0;
  unsigned reach = 1u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool mine = (reach >> lane) & 1u;
    const unsigned contrib = (mine && nxt < 32u) ? (1u << nxt) : 0u;
    reach |= __reduce_or_sync(kFull, contrib);
    const uint32_t hop = __shfl_sync(kFull, nxt, nxt & 31u);
    nxt = (nxt < 32u) ? hop : nxt;
  }
  const unsigned tokmask = reach & ~stopmask;
  if (tokmask == 0) return 0;
  const unsigned hitstop = reach & stopmask;
  const int last = 31 - __clz(tokmask);
  const uint32_t adv = hitstop ? (uint32_t)(__ffs(hitstop) - 1) : __shfl_sync(kFull, nxt0, last);
  const bool is_tok = (tokmask >> lane) & 1u;

  // --- per-token fields, output positions ------------------------------------------------
  uint32_t off = 0;
  if (is_tok && t.M) off = P::offset(win + lane, t);
  const uint32_t len = is_tok ? (t.L + t.M) : 0u;
  uint32_t incl = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += o;
  }
  const uint32_t total = __shfl_sync(kFull, incl, 31);
  if ((uint64_t)s.op + total > s.out_cap) return 0;        // let the slow path find the exact error
  // From here on output positions are kept in "aligned space" (offset + s.align): the ring index
  // is then just (pos & mask) and (s.out - s.align)[pos] is the global address.
  const uint32_t rbase = s.ring;
  const uint8_t* const outa = s.out - s.align;
  const uint32_t o_lit = s.op + s.align + incl - len;
  const uint32_t o_mat = o_lit + t.L;
  const bool bad = is_tok && t.M && (off == 0u || off > o_mat - s.align);
  if (__any_sync(kFull, bad)) return -1;

  // --- literals: every window byte finds its token and scatters itself --------------------
  {
    const unsigned below = reach & (0xffffffffu >> (31 - lane));   // reach bits <= lane (bit 0 always set)
    const int tk = 31 - __clz(below);
    const uint32_t tL = __shfl_sync(kFull, t.L, tk);
    const uint32_t tO = __shfl_sync(kFull, o_lit, tk);
    const bool tok_ok = (tokmask >> tk) & 1u;
    const uint32_t k = (uint32_t)lane - (uint32_t)tk - 1u;          // literal index within token tk
    if (tok_ok && lane > tk && k < tL) sts_u8(rbase + ((tO + k) & kRingMask), b0);
    // literal bytes past the window can only belong to the last token
    const uint32_t lL = __shfl_sync(kFull, t.L, last);
    const uint32_t lO = __shfl_sync(kFull, o_lit, last);
    if ((uint32_t)last + 1u + lL > 32u) {
      const uint32_t k2 = 32u + (uint32_t)lane - (uint32_t)last - 1u;
      if (k2 < lL) sts_u8(rbase + ((lO + k2) & kRingMask), win[32 + lane]);
    }
  }
  // --- matches -------------------------------------------------------------------------------
  // Round 1 copies, one match per lane, every match whose source bytes are already final
  // (they end at or below the output position of the first match of the window): on tabular
  // data that is the majority.  The common case (no overlap, source entirely in the ring or
  // entirely in flushed global memory, no ring wrap-around) is a branch-free unrolled copy in
  // tiers with immediate offsets.  The remaining matches depend on output of this same window
  // and are retired in order, the whole warp copying one match (<= 18 bytes, one byte per
  // lane) per step.
  const uint32_t ring_from = max(s.ring_lo, s.op > kRingReach ? s.op - kRingReach : 0u) + s.align;
  const bool has_match = is_tok && t.M != 0u;
  unsigned pending = __ballot_sync(kFull, has_match);
  const uint32_t src0 = o_mat - off;
  const uint32_t pack = off | (t.M << 16);
  if (pending) {
    const uint32_t src_end = src0 + min(t.M, off);           // exclusive end of the bytes this match reads
    const uint32_t didx = o_mat & kRingMask, sidx = src0 & kRingMask;
    const bool in_ring = src0 >= ring_from && sidx <= kRingBytes - 20u;
    const bool far = src0 + t.M <= ring_from;                // flushed long ago: read from global
    const bool simple = has_match && off >= t.M && didx <= kRingBytes - 20u && (in_ring || far);
    const uint32_t dp = rbase + didx, sp = rbase + sidx;
    const uint8_t* const gp = outa + src0;
    // up to kParRounds parallel rounds: each takes every pending match whose source ends at or below
    // the output position of the first pending match (everything below it is final)
    for (int rnd = 0; rnd < kParRounds && pending; ++rnd) {
    __syncwarp();
    const int first0 = __ffs(pending) - 1;
    const uint32_t w = __shfl_sync(kFull, o_mat, first0);
    const bool ready = ((pending >> lane) & 1u) && (src_end <= w);
    const bool fast = ready && simple;
    const bool fast_r = fast && !far, fast_g = fast && far;
    const unsigned fmask = __ballot_sync(kFull, fast);
    if (fmask == 0u) break;                                  // first pending match needs the generic path
#define B200_TIER4(LD, SRC, A, B, C, D)                                                        \
    {                                                                                          \
      const uint32_t x0 = LD<A>(SRC), x1 = LD<B>(SRC), x2 = LD<C>(SRC), x3 = LD<D>(SRC);       \
      sts_u8<A>(dp, x0);                                                                       \
      if (t.M > B) sts_u8<B>(dp, x1);                                                          \
      if (t.M > C) sts_u8<C>(dp, x2);                                                          \
      if (t.M > D) sts_u8<D>(dp, x3);                                                          \
    }
#define B200_TAIL10(LD, SRC)                                                                   \
    {                                                                                          \
      if (8 < t.M) sts_u8<8>(dp, LD<8>(SRC));    if (9 < t.M) sts_u8<9>(dp, LD<9>(SRC));       \
      if (10 < t.M) sts_u8<10>(dp, LD<10>(SRC)); if (11 < t.M) sts_u8<11>(dp, LD<11>(SRC));    \
      if (12 < t.M) sts_u8<12>(dp, LD<12>(SRC)); if (13 < t.M) sts_u8<13>(dp, LD<13>(SRC));    \
      if (14 < t.M) sts_u8<14>(dp, LD<14>(SRC)); if (15 < t.M) sts_u8<15>(dp, LD<15>(SRC));    \
      if (16 < t.M) sts_u8<16>(dp, LD<16>(SRC)); if (17 < t.M) sts_u8<17>(dp, LD<17>(SRC));    \
    }
    if (fast_r) B200_TIER4(lds_u8, sp, 0, 1, 2, 3)
    if (__any_sync(kFull, fast_r && t.M > 4u)) {
      if (fast_r && t.M > 4u) B200_TIER4(lds_u8, sp, 4, 5, 6, 7)
      if (__any_sync(kFull, fast_r && t.M > 8u)) {
        if (fast_r && t.M > 8u) B200_TAIL10(lds_u8, sp)
      }
    }
    if (__any_sync(kFull, fast_g)) {
      // sources flushed long ago (read from global memory), same tiers
      if (fast_g) B200_TIER4(ldg_u8, gp, 0, 1, 2, 3)
      if (__any_sync(kFull, fast_g && t.M > 4u)) {
        if (fast_g && t.M > 4u) B200_TIER4(ldg_u8, gp, 4, 5, 6, 7)
        if (__any_sync(kFull, fast_g && t.M > 8u)) {
          if (fast_g && t.M > 8u) B200_TAIL10(ldg_u8, gp)
        }
      }
    }
    pending &= ~fmask;
    }
#undef B200_TIER4
#undef B200_TAIL10
    // in-order retirement of everything else
    while (pending) {
      __syncwarp();
      const int f = __ffs(pending) - 1;
      const uint32_t f_omat = __shfl_sync(kFull, o_mat, f);
      const uint32_t f_pack = __shfl_sync(kFull, pack, f);
      const uint32_t f_off = f_pack & 0xffffu, f_M = f_pack >> 16;
      if ((uint32_t)lane < f_M) {
        uint32_t r = lane;
        if (r >= f_off) {                                    // overlapping match replicates its period
          r -= f_off;
          if (r >= f_off) { r -= f_off; if (r >= f_off) r %= f_off; }
        }
        const uint32_t q = f_omat - f_off + r;
        const uint32_t b = (q >= ring_from) ? lds_u8(rbase + (q & kRingMask)) : (uint32_t)outa[q];
        sts_u8(rbase + ((f_omat + lane) & kRingMask), b);
      }
      pending &= pending - 1;
    }
  }
  s.op += total;
  s.ip += adv;
  return __popc(tokmask);
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
    // speculative window parse that would retire nothing
    if (s.ip + P::kLook <= s.in_n && !P::is_stop(s.in[s.ip])) {
      const int r = lz_fast_iter<P>(s, lane);
      if (r < 0) return false;
      if (r > 0) { lz_flush_blocks(s, lane); continue; }
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

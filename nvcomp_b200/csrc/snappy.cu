// snappy.cu -- batched Snappy raw-format codec for B200 (sm_100a) + its C ABI.
//
// Replaces the closed nvcompBatchedSnappy* entry points (include/nvcomp/snappy.h;
// reference call sites benchmarks/benchmark_snappy_synth.cpp:128-266,
// benchmarks/benchmark_snappy_chunked.cu:51-55).  The decoder accepts every legal
// Snappy stream: literal tags with 0..4 length bytes and copy-1 / copy-2 / copy-4
// elements (reference CHANGELOG.md:182-184).
#include "common.cuh"
#include "lz77_compress.cuh"
#include "lz_decode.cuh"
#include "nvcomp/snappy.h"

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
  while (ip < in_n) {
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
    lz_emit_match(s, off, len, lane);
    s.ip = ip;
    return 1;
  }
};

__device__ __forceinline__ bool snappy_decode_chunk_v2(const uint8_t* in, uint32_t in_n, uint8_t* out,
                                                       uint64_t out_cap, uint32_t* produced,
                                                       uint8_t* ring, int lane) {
  uint32_t ip = 0;
  uint64_t ulen;
  if (!snappy_read_preamble(in, in_n, ip, ulen)) return false;
  if (ulen > out_cap) return false;
  // Adaptive strategy (see lz4.cu): chunks that compressed >= 4x are long-match dominated and
  // take the direct global-memory token loop.
  if (ulen >= 4ull * in_n) return snappy_decode_chunk(in, in_n, out, out_cap, produced, lane);
  LzState s;
  s.in = in; s.in_n = in_n; s.out = out; s.out_cap = ulen;
  s.ip = ip; s.op = 0; s.flushed = 0; s.ring_lo = 0;
  s.align = (uint32_t)((uintptr_t)out & 15u);
  s.ring = (uint32_t)__cvta_generic_to_shared(ring);
  if (!lz_decode_stream<SnappyDecode>(s, lane)) return false;
  if (s.op != (uint32_t)ulen) return false;
  *produced = s.op;
  return true;
}

constexpr int kLzDecWarps = 4;
// 10 CTAs x 4 warps per SM (48 registers): measured best of 8 / 10 / 12 (profiles/README.md)
constexpr int kLzDecCtasPerSm = 10;

__global__ void __launch_bounds__(kLzDecWarps * 32, kLzDecCtasPerSm)
snappy_decompress_v2_kernel(const void* const* __restrict__ comp_ptrs,
                            const size_t* __restrict__ comp_bytes,
                            const size_t* __restrict__ out_caps,
                            size_t* actual_bytes, size_t batch,
                            void* const* __restrict__ out_ptrs,
                            nvcompStatus_t* statuses,
                            unsigned long long* ticket) {
  __shared__ __align__(16) uint8_t s_ring[kLzDecWarps][kRingBytes];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  const size_t warp_global = (size_t)blockIdx.x * kLzDecWarps + w;
  const size_t warps_total = (size_t)gridDim.x * kLzDecWarps;
  // Two passes over the ticket space: dense short-token chunks (compressed < 4x, the expensive
  // ones) are handed out first, cheap long-match chunks fill the tail -- unequal chunks would
  // otherwise leave a few warps finishing expensive chunks alone at the end of the batch.
  for (int pass = 0; pass < 2; ++pass) {
    WarpTicket sched(ticket ? ticket + pass : nullptr, warp_global, warps_total);
    for (size_t c = sched.next(lane); c < batch; c = sched.next(lane)) {
      const size_t in_n64 = comp_bytes[c];
      const uint64_t cap = (uint64_t)out_caps[c];
      const bool heavy = cap < 4ull * in_n64;
      if (heavy != (pass == 0)) continue;
      const uint8_t* in = (const uint8_t*)comp_ptrs[c];
      uint8_t* out = (uint8_t*)out_ptrs[c];
      __builtin_assume(__isGlobal(in)); __builtin_assume(__isGlobal(out));
      uint32_t produced = 0;
      bool ok = in_n64 <= 0xffffffffull;
      if (ok) ok = snappy_decode_chunk_v2(in, (uint32_t)in_n64, out, cap, &produced, s_ring[w], lane);
      if (lane == 0) {
        if (actual_bytes) actual_bytes[c] = ok ? (size_t)produced : 0;
        if (statuses) statuses[c] = ok ? nvcompSuccess : nvcompErrorCannotDecompress;
      }
      __syncwarp();
    }
  }
}


// Size query: only the varint preamble is read (one thread per chunk).
__global__ void snappy_size_kernel(const void* const* __restrict__ comp_ptrs,
                                   const size_t* __restrict__ comp_bytes,
                                   size_t* out_sizes, size_t batch) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= batch) return;
  const uint8_t* in = (const uint8_t*)comp_ptrs[c];
  const size_t n = comp_bytes[c];
  uint32_t ip = 0;
  uint64_t ulen = 0;
  const bool ok = n <= 0xffffffffull && snappy_read_preamble(in, (uint32_t)n, ip, ulen);
  out_sizes[c] = ok ? (size_t)ulen : 0;
}

// ---------------------------------------------------------------------------
// Compression
// ---------------------------------------------------------------------------
struct SnappyEmitter {
  uint8_t* out;
  uint32_t op;

  __device__ __forceinline__ void begin(uint32_t n, int lane) {
    // varint32 of the uncompressed length
    uint32_t v = n, k = 0;
    while (v >= 0x80u) { if (lane == 0) out[op + k] = (uint8_t)(v | 0x80u); v >>= 7; ++k; }
    if (lane == 0) out[op + k] = (uint8_t)v;
    op += k + 1;
  }
  __device__ __forceinline__ void literal(const uint8_t* lit, uint32_t ll, int lane) {
    if (ll == 0) return;
    const uint32_t n1 = ll - 1;
    if (n1 < 60) {
      if (lane == 0) out[op] = (uint8_t)(n1 << 2);
      op += 1;
    } else {
      const uint32_t nb = n1 < (1u << 8) ? 1u : n1 < (1u << 16) ? 2u : n1 < (1u << 24) ? 3u : 4u;
      if (lane == 0) {
        out[op] = (uint8_t)((59u + nb) << 2);
        for (uint32_t i = 0; i < nb; ++i) out[op + 1 + i] = (uint8_t)(n1 >> (8 * i));
      }
      op += 1 + nb;
    }
    warp_copy<true>(out + op, lit, ll, lane);
    op += ll;
  }
  __device__ __forceinline__ void copy_tail(uint32_t off, uint32_t len, int lane) {   // len 4..64 (or 1..64)
    if (len < 12 && off < 2048 && len >= 4) {
      if (lane == 0) {
        out[op] = (uint8_t)(1u | ((len - 4) << 2) | ((off >> 8) << 5));
        out[op + 1] = (uint8_t)(off & 255u);
      }
      op += 2;
    } else {
      if (lane == 0) {
        out[op] = (uint8_t)(2u | ((len - 1) << 2));
        out[op + 1] = (uint8_t)(off & 255u);
        out[op + 2] = (uint8_t)(off >> 8);
      }
      op += 3;
    }
  }
  __device__ __forceinline__ void sequence(const uint8_t* lit, uint32_t ll, uint32_t off,
                                           uint32_t ml, int lane) {
    literal(lit, ll, lane);
    // long matches split into copy-2 elements of 64 bytes; emitted lane-parallel
    const uint32_t q = (ml >= 68) ? (ml - 4) / 64 : 0;
    for (uint32_t i = lane; i < q; i += kWarp) {
      uint8_t* p = out + op + 3 * i;
      p[0] = (uint8_t)(2u | (63u << 2));
      p[1] = (uint8_t)(off & 255u);
      p[2] = (uint8_t)(off >> 8);
    }
    op += 3 * q;
    uint32_t rem = ml - 64 * q;
    if (rem > 64) { copy_tail(off, 60, lane); rem -= 60; }
    copy_tail(off, rem, lane);
  }
  __device__ __forceinline__ void finish(const uint8_t* lit, uint32_t ll, int lane) {
    literal(lit, ll, lane);
  }
};

constexpr int kSnappyCompWarps = 4;

__global__ void __launch_bounds__(kSnappyCompWarps * 32)
snappy_compress_kernel(const void* const* __restrict__ in_ptrs, const size_t* __restrict__ in_bytes,
                       size_t batch, void* const* __restrict__ out_ptrs, size_t* out_bytes,
                       unsigned long long* ticket) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  uint16_t* table = (uint16_t*)(smem + (size_t)w * kHashBytesPerWarp);
  const size_t warp_global = (size_t)blockIdx.x * kSnappyCompWarps + w;
  const size_t warps_total = (size_t)gridDim.x * kSnappyCompWarps;
  WarpTicket sched(ticket, warp_global, warps_total);
  for (size_t c = sched.next(lane); c < batch; c = sched.next(lane)) {
    const uint8_t* in = (const uint8_t*)in_ptrs[c];
    const uint32_t n = (uint32_t)in_bytes[c];
    SnappyEmitter em{(uint8_t*)out_ptrs[c], 0};
    em.begin(n, lane);
    // Snappy has no end-of-block restrictions; 4 keeps the 4-byte probe in bounds.
    lz77_compress_chunk(in, n, em, table, 1u, 0u, 4u, lane);
    if (lane == 0) out_bytes[c] = em.op;
    __syncwarp();
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

nvcompStatus_t nvcompBatchedSnappyCompressGetTempSize(
    size_t, size_t max_chunk, nvcompBatchedSnappyOpts_t, size_t* temp_bytes) {
  if (!temp_bytes) return nvcompErrorInvalidValue;
  if (max_chunk > nvcompSnappyCompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  *temp_bytes = kSchedBytes;
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedSnappyCompressGetTempSizeEx(
    size_t b, size_t m, nvcompBatchedSnappyOpts_t o, size_t* t, const size_t) {
  return nvcompBatchedSnappyCompressGetTempSize(b, m, o, t);
}

nvcompStatus_t nvcompBatchedSnappyCompressGetMaxOutputChunkSize(
    size_t max_chunk, nvcompBatchedSnappyOpts_t, size_t* max_compressed_bytes) {
  if (!max_compressed_bytes) return nvcompErrorInvalidValue;
  if (max_chunk > nvcompSnappyCompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  *max_compressed_bytes = 32 + max_chunk + max_chunk / 6;   // snappy::MaxCompressedLength
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedSnappyCompressAsync(
    const void* const* in_ptrs, const size_t* in_bytes, size_t max_chunk, size_t batch,
    void* temp, size_t temp_bytes, void* const* out_ptrs, size_t* out_bytes,
    nvcompBatchedSnappyOpts_t, cudaStream_t stream) {
  log_call("nvcompBatchedSnappyCompressAsync", batch, max_chunk, stream);
  if (max_chunk > nvcompSnappyCompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  if (batch == 0) return nvcompSuccess;
  if (!in_ptrs || !in_bytes || !out_ptrs || !out_bytes) return nvcompErrorInvalidValue;
  unsigned long long* ticket = nullptr;
  if (temp && temp_bytes >= kSchedBytes) {
    ticket = (unsigned long long*)temp;
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), stream));
  }
  const size_t smem = (size_t)kSnappyCompWarps * kHashBytesPerWarp;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_TRY(cudaFuncSetAttribute(snappy_compress_kernel,
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = persistent_grid(6, batch, kSnappyCompWarps);
  snappy_compress_kernel<<<grid, kSnappyCompWarps * 32, smem, stream>>>(
      in_ptrs, in_bytes, batch, out_ptrs, out_bytes, ticket);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedSnappyDecompressGetTempSize(size_t, size_t, size_t* temp_bytes) {
  if (!temp_bytes) return nvcompErrorInvalidValue;
  *temp_bytes = kSchedBytes;
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedSnappyDecompressGetTempSizeEx(size_t n, size_t m, size_t* t, size_t) {
  return nvcompBatchedSnappyDecompressGetTempSize(n, m, t);
}

nvcompStatus_t nvcompBatchedSnappyGetDecompressSizeAsync(
    const void* const* comp_ptrs, const size_t* comp_bytes, size_t* out_sizes,
    size_t batch, cudaStream_t stream) {
  log_call("nvcompBatchedSnappyGetDecompressSizeAsync", batch, 0, stream);
  if (batch == 0) return nvcompSuccess;
  if (!comp_ptrs || !comp_bytes || !out_sizes) return nvcompErrorInvalidValue;
  const int threads = 128;
  snappy_size_kernel<<<(unsigned)((batch + threads - 1) / threads), threads, 0, stream>>>(
      comp_ptrs, comp_bytes, out_sizes, batch);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedSnappyDecompressAsync(
    const void* const* comp_ptrs, const size_t* comp_bytes, const size_t* out_caps,
    size_t* actual_bytes, size_t batch, void* const temp, size_t temp_bytes,
    void* const* out_ptrs, nvcompStatus_t* statuses, cudaStream_t stream) {
  log_call("nvcompBatchedSnappyDecompressAsync", batch, 0, stream);
  if (batch == 0) return nvcompSuccess;
  if (!comp_ptrs || !comp_bytes || !out_caps || !out_ptrs) return nvcompErrorInvalidValue;
  unsigned long long* ticket = nullptr;
  if (temp && temp_bytes >= kSchedBytes) {
    ticket = (unsigned long long*)temp;
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, 2 * sizeof(unsigned long long), stream));
  }
  const int grid = persistent_grid(kLzDecCtasPerSm, batch, kLzDecWarps);
  snappy_decompress_v2_kernel<<<grid, kLzDecWarps * 32, 0, stream>>>(
      comp_ptrs, comp_bytes, out_caps, actual_bytes, batch, out_ptrs, statuses, ticket);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

}  // extern "C"

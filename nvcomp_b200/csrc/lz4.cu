// lz4.cu -- batched LZ4 block codec for B200 (sm_100a) + its C ABI.
//
// Replaces the closed nvcompBatchedLZ4* entry points (include/nvcomp/lz4.h).
// Wire format: LZ4 block format, one block per chunk, interoperable with
// liblz4 1.9.4 in both directions (reference examples/lz4_cpu_compression.cu,
// examples/lz4_cpu_decompression.cu).
//
// Decode: one warp owns one chunk; chunks are handed out by a persistent two-pass ticket
// scheduler (dense chunks first).  Dense short-token chunks use the lane-parallel decoder of
// lz_decode.cuh; chunks that compressed >= 4x use the direct sequence loop below (the sequence is
// parsed from a 32-byte register window, run-length matches are expanded from registers, other
// matches are 16-byte vector copies: common.cuh warp_copy / warp_match_copy).
#include "common.cuh"
#include "lz77_compress.cuh"
#include "lz_decode.cuh"
#include "nvcomp/lz4.h"

namespace b200 {

// ---------------------------------------------------------------------------
// Length-extension bytes (the 255,255,...,x tail of a 15 nibble): 32 bytes are
// examined per round with a ballot instead of a serial byte walk.
// Returns false on input overrun.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool lz4_read_ext(const uint8_t* __restrict__ in, uint32_t in_n,
                                             uint32_t& ip, uint32_t& len, int lane) {
  while (true) {
    const uint32_t q = ip + lane;
    const uint32_t b = (q < in_n) ? in[q] : 0u;   // 0 terminates: overrun detected below
    const unsigned stop = __ballot_sync(kFull, b != 255u);
    if (stop == 0) { len += 255u * 32u; ip += 32; continue; }
    const int k = __ffs(stop) - 1;
    len += 255u * (uint32_t)k + __shfl_sync(kFull, b, k);
    ip += k + 1;
    return ip <= in_n;
  }
}

// Walk the sequences of one LZ4 block without copying (size query: LZ4 blocks carry no size header).
// Returns true on a well-formed block; *produced receives the decompressed size.
__device__ __forceinline__ bool lz4_walk_chunk(const uint8_t* __restrict__ in, uint32_t in_n,
                                               uint32_t* produced, int lane) {
  uint32_t ip = 0;
  uint64_t op = 0;
  if (in_n == 0) { *produced = 0; return true; }
  while (true) {
    if (ip >= in_n) return false;
    const uint32_t tok = in[ip++];
    uint32_t ll = tok >> 4;
    if (ll == 15) { if (!lz4_read_ext(in, in_n, ip, ll, lane)) return false; }
    if (ll > in_n - ip) return false;
    ip += ll; op += ll;
    if (ip >= in_n) break;                 // last sequence carries literals only
    if (in_n - ip < 2) return false;
    const uint32_t off = load_u16(in + ip);
    ip += 2;
    uint32_t ml = tok & 15u;
    if (ml == 15) { if (!lz4_read_ext(in, in_n, ip, ml, lane)) return false; }
    ml += 4;
    if (off == 0 || (uint64_t)off > op) return false;
    op += ml;
    if (op > 0xffffffffull) return false;
  }
  *produced = (uint32_t)op;
  return true;
}

// ---------------------------------------------------------------------------
// Direct decode for chunks that compressed >= 4x (long matches, typed run-length data).  One coalesced
// 32-byte load brings a whole sequence (token, short literals, offset, length-extension bytes) into a
// register window; fields are picked with shuffles / a ballot.  A match whose period (1, 2, 4 or 8 bytes) lies
// inside the literals of its own sequence -- the shape of typed run-length data -- is expanded from the
// window: the 8-byte period is rotated to the destination alignment and broadcast with 16-byte stores,
// no load from the output buffer.  Other matches are copied through memory (common.cuh) with the fields
// already in registers; sequences that do not fit the window (long literal runs, far length
// extensions, the end of the block) take the generic field-by-field path below.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool lz4_decode_chunk_direct(const uint8_t* __restrict__ in, uint32_t in_n,
                                                        uint8_t* out, uint64_t out_cap64,
                                                        uint32_t* produced, int lane) {
  if (in_n == 0) { *produced = 0; return true; }
  const uint32_t cap = out_cap64 > 0xffffffffull ? 0xffffffffu : (uint32_t)out_cap64;
  const uint32_t ul = (uint32_t)lane;
  uint32_t ip = 0, op = 0;
  while (true) {
    if (ip >= in_n) return false;
    if (ip + 32u <= in_n) {
      // ---- window path
      const uint32_t b = in[ip + ul];
      const uint32_t tok = __shfl_sync(kFull, b, 0);
      const uint32_t ll = tok >> 4;
      if (ll < 15u) {                                              // 15 = extended literal length: generic path
        uint32_t used = 3u + ll;                                   // token + literals + offset
        const uint32_t off = __shfl_sync(kFull, b, 1 + ll) | (__shfl_sync(kFull, b, 2 + ll) << 8);
        uint32_t ml = (tok & 15u) + 4u;
        bool fits = true;
        if ((tok & 15u) == 15u) {
          const unsigned e = __ballot_sync(kFull, b != 255u) & ~((1u << used) - 1u);
          if (e == 0u) fits = false;                               // extension runs past the window
          else {
            const uint32_t p = (uint32_t)__ffs(e) - 1u;
            ml += 255u * (p - used) + __shfl_sync(kFull, b, p);
            used = p + 1u;
          }
        }
        if (fits) {
          if (ll > cap - op || ml > cap - op - ll || off == 0u || off > op + ll) return false;
          if (ul - 1u < ll) out[op + ul - 1u] = (uint8_t)b;       // literals: window lanes 1..ll
          uint8_t* dst = out + op + ll;
          if (!(off <= ll && off <= 8u && (off & (off - 1u)) == 0u)) {
            // general match: copy through memory (fields came from the window, no further input loads)
            __syncwarp();
            warp_match_copy(dst, off, ml, lane);
            __syncwarp();
            op += ll + ml;
            ip += used;
            continue;
          }
          // period (1, 2, 4 or 8 bytes) inside this sequence's literals: expand from the window
          // 8-byte period P: byte k = literal[ll - off + (k mod off)] = window lane 1 + ll - off + (k mod off)
          const uint32_t pb = __shfl_sync(kFull, b, 1u + ll - off + (ul & (off - 1u)));
          const uint32_t placed = pb << (8u * (ul & 3u));
          const uint32_t plo = __reduce_or_sync(kFull, ul < 4u ? placed : 0u);
          const uint32_t phi = __reduce_or_sync(kFull, (ul & 28u) == 4u ? placed : 0u);
          // every 16-byte aligned vector of the run holds P rotated by (-dst) & 7 bytes, twice
          const uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
          const uint32_t r0 = head & 7u;
          const uint32_t wa = (r0 & 4u) ? phi : plo, wb = (r0 & 4u) ? plo : phi, sh = 8u * (r0 & 3u);
          uint4 v;
          v.x = __funnelshift_r(wa, wb, sh);
          v.y = __funnelshift_r(wb, wa, sh);
          v.z = v.x; v.w = v.y;
          // byte j of the run, for lanes that write single bytes (j mod 8 selects a byte of P)
          const uint32_t mine = (((ul & 4u) ? phi : plo) >> (8u * (ul & 3u))) & 0xffu;   // P[lane & 7]
          if (ml < 16u + head) {
            // short: bytes only (ml < 31)
            if (ul < ml) dst[ul] = (uint8_t)mine;
          } else {
            if (ul < head) dst[ul] = (uint8_t)mine;
            const uint32_t nvec = (ml - head) >> 4;
            uint4* d16 = (uint4*)(dst + head);
            // nvec <= 64 for matches up to ~1 KB: two predicated stores, a loop only beyond that
            if (ul < nvec) st_v4(d16 + ul, v);
            if (ul + kWarp < nvec) st_v4(d16 + ul + kWarp, v);
#pragma unroll 1
            for (uint32_t k = ul + 2u * kWarp; k < nvec; k += kWarp) st_v4(d16 + k, v);
            // ragged end (< 16 bytes): position head + 16 nvec + lane; 16 nvec = 0 mod 8
            const uint32_t j = head + (nvec << 4) + ul;
            const uint32_t jb = (((j & 4u) ? phi : plo) >> (8u * (j & 3u))) & 0xffu;
            if (j < ml) dst[j] = (uint8_t)jb;
          }
          op += ll + ml;
          ip += used;
          continue;
        }
      }
    }
    // ---- generic path: one sequence, field by field
    const uint32_t tok = in[ip++];
    uint32_t ll = tok >> 4;
    if (ll == 15) { if (!lz4_read_ext(in, in_n, ip, ll, lane)) return false; }
    if (ll > in_n - ip || ll > cap - op) return false;
    if (ll) warp_copy<true>(out + op, in + ip, ll, lane);
    ip += ll; op += ll;
    if (ip >= in_n) break;                 // last sequence carries literals only
    if (in_n - ip < 2) return false;
    const uint32_t off = load_u16(in + ip);
    ip += 2;
    uint32_t ml = tok & 15u;
    if (ml == 15) { if (!lz4_read_ext(in, in_n, ip, ml, lane)) return false; }
    if (ml > 0xfffffff0u) return false;
    ml += 4;
    if (off == 0 || off > op || ml > cap - op) return false;
    __syncwarp();                          // prior stores visible to all lanes
    warp_match_copy(out + op, off, ml, lane);
    __syncwarp();
    op += ml;
  }
  *produced = op;
  return true;
}

// Size query: one warp walks one chunk.
__global__ void __launch_bounds__(128)
lz4_size_kernel(const void* const* __restrict__ comp_ptrs, const size_t* __restrict__ comp_bytes,
                size_t* out_sizes, size_t batch) {
  const int lane = lane_id();
  const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t c = warp_global; c < batch; c += warps_total) {
    const uint8_t* in = (const uint8_t*)comp_ptrs[c];
    const size_t in_n64 = comp_bytes[c];
    uint32_t produced = 0;
    bool ok = in_n64 <= 0xffffffffull;
    if (ok) ok = lz4_walk_chunk(in, (uint32_t)in_n64, &produced, lane);
    if (lane == 0) out_sizes[c] = ok ? (size_t)produced : 0;
  }
}


// ---------------------------------------------------------------------------
// v2 decode (lz_decode.cuh): lane-parallel short-token path + this slow path
// ---------------------------------------------------------------------------
struct Lz4Decode : Lz4Policy {
  __device__ static __forceinline__ bool at_end(const LzState&) { return false; }   // ends inside serial_token
  // one full sequence (token, literals, match), parsed once; 2 = final literals consumed
  __device__ static __forceinline__ int serial_token(LzState& s, int lane) {
    const uint8_t* __restrict__ in = s.in;
    const uint32_t in_n = s.in_n;
    uint32_t ip = s.ip;
    if (ip >= in_n) return -1;
    const uint32_t tok = in[ip++];
    uint32_t ll = tok >> 4;
    if (ll == 15) { if (!lz4_read_ext(in, in_n, ip, ll, lane)) return -1; }
    if (ll > in_n - ip) return -1;
    if ((uint64_t)ll > s.out_cap - s.op) return -1;
    const uint32_t lit_at = ip;
    ip += ll;
    if (ip >= in_n) {                              // last sequence: literals only
      lz_emit_literals(s, in + lit_at, ll, lane);
      s.ip = ip;
      return 2;
    }
    if (in_n - ip < 2) return -1;
    const uint32_t off = load_u16(in + ip);
    ip += 2;
    uint32_t ml = tok & 15u;
    if (ml == 15) { if (!lz4_read_ext(in, in_n, ip, ml, lane)) return -1; }
    ml += 4;
    if (off == 0 || (uint64_t)off > (uint64_t)s.op + ll) return -1;
    if ((uint64_t)ml > s.out_cap - s.op - ll) return -1;
    lz_emit_literals(s, in + lit_at, ll, lane);
    lz_emit_match(s, off, ml, lane);
    s.ip = ip;
    return 1;
  }
};

__device__ __forceinline__ bool lz4_decode_chunk_v2(const uint8_t* in, uint32_t in_n, uint8_t* out,
                                                    uint64_t out_cap, uint32_t* produced,
                                                    uint8_t* ring, int lane) {
  if (in_n == 0) { *produced = 0; return true; }
  // Adaptive strategy: a chunk that compressed >= 4x is dominated by long matches; the ring /
  // lane-parallel machinery only costs instructions there, so it is decoded by the direct
  // global-memory token loop (16-byte vector copies).  Dense short-token chunks take the
  // lane-parallel path.
  if (out_cap >= 4ull * in_n) return lz4_decode_chunk_direct(in, in_n, out, out_cap, produced, lane);
  LzState s;
  s.in = in; s.in_n = in_n; s.out = out; s.out_cap = out_cap > 0xffffffffull ? 0xffffffffull : out_cap;
  s.ip = 0; s.op = 0; s.flushed = 0; s.ring_lo = 0;
  s.align = (uint32_t)((uintptr_t)out & 15u);
  s.ring = (uint32_t)__cvta_generic_to_shared(ring);
  if (!lz_decode_stream<Lz4Decode>(s, lane)) return false;
  *produced = s.op;
  return true;
}

constexpr int kLzDecWarps = 4;
// 10 CTAs x 4 warps per SM (48 registers): measured best of 8 / 10 / 12 (profiles/README.md)
constexpr int kLzDecCtasPerSm = 10;

__global__ void __launch_bounds__(kLzDecWarps * 32, kLzDecCtasPerSm)
lz4_decompress_v2_kernel(const void* const* __restrict__ comp_ptrs,
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
      if (ok) ok = lz4_decode_chunk_v2(in, (uint32_t)in_n64, out, cap, &produced, s_ring[w], lane);
      if (lane == 0) {
        if (actual_bytes) actual_bytes[c] = ok ? (size_t)produced : 0;
        if (statuses) statuses[c] = ok ? nvcompSuccess : nvcompErrorCannotDecompress;
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------
// Compression
// ---------------------------------------------------------------------------
struct Lz4Emitter {
  uint8_t* out;
  uint32_t op;

  __device__ __forceinline__ void ext(uint32_t rem, int lane) {   // rem = len - 15
    const uint32_t nb = rem / 255u + 1u;
    for (uint32_t i = lane; i < nb; i += kWarp)
      out[op + i] = (i + 1 < nb) ? (uint8_t)255 : (uint8_t)(rem - 255u * (nb - 1));
    op += nb;
  }
  __device__ __forceinline__ void sequence(const uint8_t* lit, uint32_t ll, uint32_t off,
                                           uint32_t ml, int lane) {
    const uint32_t mlc = ml - 4;
    if (lane == 0) out[op] = (uint8_t)((min(ll, 15u) << 4) | min(mlc, 15u));
    op += 1;
    if (ll >= 15) ext(ll - 15, lane);
    if (ll) warp_copy<true>(out + op, lit, ll, lane);
    op += ll;
    if (lane == 0) { out[op] = (uint8_t)(off & 255u); out[op + 1] = (uint8_t)(off >> 8); }
    op += 2;
    if (mlc >= 15) ext(mlc - 15, lane);
  }
  __device__ __forceinline__ void finish(const uint8_t* lit, uint32_t ll, int lane) {
    if (lane == 0) out[op] = (uint8_t)(min(ll, 15u) << 4);
    op += 1;
    if (ll >= 15) ext(ll - 15, lane);
    if (ll) warp_copy<true>(out + op, lit, ll, lane);
    op += ll;
  }
};

constexpr int kCompWarpsPerCta = 4;

__global__ void __launch_bounds__(kCompWarpsPerCta * 32)
lz4_compress_kernel(const void* const* __restrict__ in_ptrs, const size_t* __restrict__ in_bytes,
                    size_t batch, void* const* __restrict__ out_ptrs, size_t* out_bytes,
                    uint32_t step, unsigned long long* ticket) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  uint16_t* table = (uint16_t*)(smem + (size_t)w * kHashBytesPerWarp);
  const size_t warp_global = (size_t)blockIdx.x * kCompWarpsPerCta + w;
  const size_t warps_total = (size_t)gridDim.x * kCompWarpsPerCta;
  WarpTicket sched(ticket, warp_global, warps_total);
  for (size_t c = sched.next(lane); c < batch; c = sched.next(lane)) {
    const uint8_t* in = (const uint8_t*)in_ptrs[c];
    const uint32_t n = (uint32_t)in_bytes[c];
    Lz4Emitter em{(uint8_t*)out_ptrs[c], 0};
    // LZ4 end-of-block rules: last 5 bytes are literals, the last match starts
    // at least 12 bytes before the end (reference CHANGELOG.md:195).
    lz77_compress_chunk(in, n, em, table, step, 5u, 12u, lane);
    if (lane == 0) out_bytes[c] = em.op;
    __syncwarp();
  }
}

inline uint32_t lz4_step_for(nvcompType_t t, bool* ok) {
  *ok = true;
  switch (t) {
    case NVCOMP_TYPE_CHAR: case NVCOMP_TYPE_UCHAR: case NVCOMP_TYPE_BITS: return 1;
    case NVCOMP_TYPE_SHORT: case NVCOMP_TYPE_USHORT: return 2;
    case NVCOMP_TYPE_INT: case NVCOMP_TYPE_UINT: return 4;
    default: *ok = false; return 1;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

nvcompStatus_t nvcompBatchedLZ4CompressGetTempSize(
    size_t, size_t max_chunk, nvcompBatchedLZ4Opts_t opts, size_t* temp_bytes) {
  bool ok; lz4_step_for(opts.data_type, &ok);
  if (!temp_bytes || !ok) return nvcompErrorInvalidValue;
  if (max_chunk > nvcompLZ4CompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  *temp_bytes = kSchedBytes;
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedLZ4CompressGetTempSizeEx(
    size_t batch, size_t max_chunk, nvcompBatchedLZ4Opts_t opts, size_t* temp_bytes, const size_t) {
  return nvcompBatchedLZ4CompressGetTempSize(batch, max_chunk, opts, temp_bytes);
}

nvcompStatus_t nvcompBatchedLZ4CompressGetMaxOutputChunkSize(
    size_t max_chunk, nvcompBatchedLZ4Opts_t, size_t* max_compressed_bytes) {
  if (!max_compressed_bytes) return nvcompErrorInvalidValue;
  if (max_chunk > nvcompLZ4CompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  // LZ4_compressBound: n + n/255 + 16
  *max_compressed_bytes = max_chunk + max_chunk / 255 + 16;
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedLZ4CompressAsync(
    const void* const* in_ptrs, const size_t* in_bytes, size_t max_chunk, size_t batch,
    void* temp, size_t temp_bytes, void* const* out_ptrs, size_t* out_bytes,
    nvcompBatchedLZ4Opts_t opts, cudaStream_t stream) {
  log_call("nvcompBatchedLZ4CompressAsync", batch, max_chunk, stream);
  bool ok; const uint32_t step = lz4_step_for(opts.data_type, &ok);
  if (!ok) return nvcompErrorInvalidValue;
  if (max_chunk > nvcompLZ4CompressionMaxAllowedChunkSize) return nvcompErrorChunkSizeTooLarge;
  if (batch == 0) return nvcompSuccess;
  if (!in_ptrs || !in_bytes || !out_ptrs || !out_bytes) return nvcompErrorInvalidValue;
  unsigned long long* ticket = nullptr;
  if (temp && temp_bytes >= kSchedBytes) {
    ticket = (unsigned long long*)temp;
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), stream));
  }
  const size_t smem = (size_t)kCompWarpsPerCta * kHashBytesPerWarp;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_TRY(cudaFuncSetAttribute(lz4_compress_kernel,
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = persistent_grid(6, batch, kCompWarpsPerCta);
  lz4_compress_kernel<<<grid, kCompWarpsPerCta * 32, smem, stream>>>(
      in_ptrs, in_bytes, batch, out_ptrs, out_bytes, step, ticket);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedLZ4DecompressGetTempSize(
    size_t, size_t, size_t* temp_bytes) {
  if (!temp_bytes) return nvcompErrorInvalidValue;
  *temp_bytes = kSchedBytes;
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedLZ4DecompressGetTempSizeEx(
    size_t n, size_t m, size_t* temp_bytes, size_t) {
  return nvcompBatchedLZ4DecompressGetTempSize(n, m, temp_bytes);
}

nvcompStatus_t nvcompBatchedLZ4GetDecompressSizeAsync(
    const void* const* comp_ptrs, const size_t* comp_bytes, size_t* out_sizes,
    size_t batch, cudaStream_t stream) {
  log_call("nvcompBatchedLZ4GetDecompressSizeAsync", batch, 0, stream);
  if (batch == 0) return nvcompSuccess;
  if (!comp_ptrs || !comp_bytes || !out_sizes) return nvcompErrorInvalidValue;
  const int grid = persistent_grid(8, batch, 4);
  lz4_size_kernel<<<grid, 128, 0, stream>>>(comp_ptrs, comp_bytes, out_sizes, batch);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

nvcompStatus_t nvcompBatchedLZ4DecompressAsync(
    const void* const* comp_ptrs, const size_t* comp_bytes, const size_t* out_caps,
    size_t* actual_bytes, size_t batch, void* const temp, size_t temp_bytes,
    void* const* out_ptrs, nvcompStatus_t* statuses, cudaStream_t stream) {
  log_call("nvcompBatchedLZ4DecompressAsync", batch, 0, stream);
  if (batch == 0) return nvcompSuccess;
  if (!comp_ptrs || !comp_bytes || !out_caps || !out_ptrs) return nvcompErrorInvalidValue;
  unsigned long long* ticket = nullptr;
  if (temp && temp_bytes >= kSchedBytes) {
    ticket = (unsigned long long*)temp;
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, 2 * sizeof(unsigned long long), stream));
  }
  const int grid = persistent_grid(kLzDecCtasPerSm, batch, kLzDecWarps);
  lz4_decompress_v2_kernel<<<grid, kLzDecWarps * 32, 0, stream>>>(
      comp_ptrs, comp_bytes, out_caps, actual_bytes, batch, out_ptrs, statuses, ticket);
  B200_CUDA_TRY(cudaGetLastError());
  return nvcompSuccess;
}

}  // extern "C"

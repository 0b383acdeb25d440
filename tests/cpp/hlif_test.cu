// hlif_test.cu -- GPU test of the high-level interface (nvcomp::*Manager, create_manager), written against
// the surface the reference's own callers use (examples/high_level_quickstart_example.cpp,
// benchmarks/benchmark_hlif.hpp).  Built by `make tests` into build/tests/hlif_test, run by
// tests/test_hlif_gpu.py.  Exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "nvcomp.hpp"
#include "nvcomp/nvcompManagerFactory.hpp"

using namespace nvcomp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
#define REQUIRE(c) do { if (!(c)) { printf("FAILED: %s (line %d)\n", #c, __LINE__); exit(1); } } while (0)

static std::vector<uint8_t> make_data(size_t n, int kind, uint32_t seed) {
  std::mt19937 rng(seed);
  std::vector<uint8_t> v(n);
  if (kind == 0) { for (auto& b : v) b = (uint8_t)(rng() & 3); }                 // gen_data(3) style
  else if (kind == 1) {                                                          // int32 run-length
    size_t i = 0;
    while (i < n) { uint32_t val = rng(); size_t run = 4 * (1 + rng() % 256);
      for (size_t k = 0; k < run && i < n; ++k, ++i) v[i] = (uint8_t)(val >> (8 * (i & 3))); }
  } else if (kind == 2) { for (auto& b : v) b = (uint8_t)rng(); }                // incompressible
  else { uint64_t x = rng(); for (size_t i = 0; i + 8 <= n; i += 8) { x += rng() % 7; memcpy(&v[i], &x, 8); } }  // sorted i64
  return v;
}

static void roundtrip(nvcompManagerBase& mgr, const std::vector<uint8_t>& host, cudaStream_t stream, bool via_factory,
                      ChecksumPolicy policy, bool use_comp_config) {
  const size_t n = host.size();
  uint8_t* d_in; CK(cudaMalloc(&d_in, n ? n : 1));
  CK(cudaMemcpy(d_in, host.data(), n, cudaMemcpyHostToDevice));
  CompressionConfig cc = mgr.configure_compression(n);
  REQUIRE(cc.max_compressed_buffer_size > 0);
  uint8_t* d_comp; CK(cudaMalloc(&d_comp, cc.max_compressed_buffer_size));
  mgr.compress(d_in, d_comp, cc);
  CK(cudaStreamSynchronize(stream));
  const size_t csize = mgr.get_compressed_output_size(d_comp);
  REQUIRE(csize > 0 && csize <= cc.max_compressed_buffer_size);
  std::shared_ptr<nvcompManagerBase> other;
  nvcompManagerBase* dm = &mgr;
  if (via_factory) { other = create_manager(d_comp, stream, 0, policy); dm = other.get(); }
  DecompressionConfig dc = use_comp_config ? dm->configure_decompression(cc) : dm->configure_decompression(d_comp);
  REQUIRE(dc.decomp_data_size == n);
  uint8_t* d_out; CK(cudaMalloc(&d_out, n ? n : 1));
  CK(cudaMemset(d_out, 0xA5, n ? n : 1));
  dm->decompress(d_out, d_comp, dc);
  CK(cudaStreamSynchronize(stream));
  REQUIRE(*dc.get_status() == nvcompSuccess);
  std::vector<uint8_t> back(n);
  CK(cudaMemcpy(back.data(), d_out, n, cudaMemcpyDeviceToHost));
  REQUIRE(back == host);
  CK(cudaFree(d_in)); CK(cudaFree(d_comp)); CK(cudaFree(d_out));
}

static uint32_t host_crc32(const uint8_t* p, size_t n) {     // zlib / IEEE 802.3
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) {
    c ^= p[i];
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xedb88320u : c >> 1;
  }
  return c ^ 0xffffffffu;
}

int main() {
  cudaStream_t stream; CK(cudaStreamCreate(&stream));
  const size_t sizes[] = {0, 1, 65535, 65536, 65537, 1000000, 5 * 65536};
  int cases = 0;
  for (int fmt = 0; fmt < 5; ++fmt) {
    for (size_t n : sizes) {
      for (int kind = 0; kind < 4; ++kind) {
        size_t nn = n;
        std::shared_ptr<nvcompManagerBase> m;
        const ChecksumPolicy pol = (kind & 1) ? ComputeAndVerify : NoComputeNoVerify;
        if (fmt == 0) m = std::make_shared<LZ4Manager>(1 << 16, nvcompBatchedLZ4Opts_t{NVCOMP_TYPE_CHAR}, stream, 0, pol);
        if (fmt == 1) m = std::make_shared<SnappyManager>(1 << 16, nvcompBatchedSnappyOpts_t{}, stream, 0, pol);
        // (typed formats take any length: trailing bytes that do not fill an element travel verbatim)
        if (fmt == 2) { m = std::make_shared<CascadedManager>(1 << 16, nvcompBatchedCascadedOpts_t{4096, NVCOMP_TYPE_LONGLONG, 1, 1, 1}, stream, 0, pol); }
        if (fmt == 3) { m = std::make_shared<BitcompManager>(1 << 16, nvcompBatchedBitcompFormatOpts{0, NVCOMP_TYPE_UINT}, stream, 0, pol); }
        if (fmt == 4) m = std::make_shared<ANSManager>(1 << 16, nvcompBatchedANSOpts_t{}, stream, 0, pol);
        auto host = make_data(nn, kind, 17 * fmt + kind);
        roundtrip(*m, host, stream, /*via_factory=*/(kind & 2) != 0, (kind & 1) ? NoComputeAndVerifyIfPresent : NoComputeNoVerify,
                  /*use_comp_config=*/kind == 0);
        ++cases;
      }
    }
  }
  // checksum mismatch is reported as nvcompErrorBadChecksum (examples/high_level_quickstart_example.cpp:313-316)
  {
    auto host = make_data(300000, 0, 99);
    LZ4Manager mgr{1 << 16, nvcompBatchedLZ4Opts_t{NVCOMP_TYPE_CHAR}, stream, 0, ComputeAndVerify};
    uint8_t* d_in; CK(cudaMalloc(&d_in, host.size())); CK(cudaMemcpy(d_in, host.data(), host.size(), cudaMemcpyHostToDevice));
    auto cc = mgr.configure_compression(host.size());
    uint8_t* d_comp; CK(cudaMalloc(&d_comp, cc.max_compressed_buffer_size));
    mgr.compress(d_in, d_comp, cc);
    CK(cudaStreamSynchronize(stream));
    // flip one literal byte deep inside the payload: LZ4 still decodes, the checksum must catch it
    const size_t csize = mgr.get_compressed_output_size(d_comp);
    std::vector<uint8_t> comp(csize); CK(cudaMemcpy(comp.data(), d_comp, csize, cudaMemcpyDeviceToHost));
    auto dc = mgr.configure_decompression(d_comp);
    uint8_t* d_out; CK(cudaMalloc(&d_out, host.size()));
    // the stored checksums are the standard CRC-32 of the uncompressed buffer and of everything after the header
    {
      uint32_t stored[2];
      memcpy(stored, comp.data() + 64, 8);
      REQUIRE(stored[0] == host_crc32(host.data(), host.size()));
      REQUIRE(stored[1] == host_crc32(comp.data() + 72, csize - 72));
    }
    // the compressed checksum covers the size table and every chunk: no flipped bit anywhere in the payload
    // may come back as success
    bool caught = false;
    for (long pos = (long)csize - 40; pos > 80; pos -= (pos > (long)csize - 3000 ? 7 : 9973)) {
      std::vector<uint8_t> bad = comp; bad[pos] ^= 0x01;
      CK(cudaMemcpy(d_comp, bad.data(), csize, cudaMemcpyHostToDevice));
      mgr.decompress(d_out, d_comp, dc);
      CK(cudaStreamSynchronize(stream));
      const nvcompStatus_t st = *dc.get_status();
      REQUIRE(st == nvcompErrorBadChecksum || st == nvcompErrorCannotDecompress);
      if (st == nvcompErrorBadChecksum) caught = true;
    }
    REQUIRE(caught);
    // a corrupt header is rejected at configure time, before any pointer is derived from it
    {
      auto try_header = [&](size_t off, uint64_t value, size_t bytes) {
        std::vector<uint8_t> bad = comp;
        memcpy(bad.data() + off, &value, bytes);
        CK(cudaMemcpy(d_comp, bad.data(), csize, cudaMemcpyHostToDevice));
        bool threw_invalid = false;
        try { mgr.configure_decompression(d_comp); } catch (const NVCompException& e) { threw_invalid = e.get_error() == nvcompErrorInvalidValue; }
        REQUIRE(threw_invalid);
      };
      try_header(48, 1000000u, 4);     // num_chunks far beyond the data
      try_header(40, 0u, 8);           // chunk_bytes 0
      try_header(40, 4096u, 8);        // chunk_bytes inconsistent with num_chunks
      try_header(32, 1ull << 40, 8);   // uncompressed_bytes inconsistent with num_chunks
      CK(cudaMemcpy(d_comp, comp.data(), csize, cudaMemcpyHostToDevice));
    }
    // ComputeAndVerify on a buffer without checksums must throw at configure time
    LZ4Manager plain{1 << 16, nvcompBatchedLZ4Opts_t{NVCOMP_TYPE_CHAR}, stream, 0, NoComputeNoVerify};
    plain.compress(d_in, d_comp, cc);
    CK(cudaStreamSynchronize(stream));
    bool threw = false;
    try { mgr.configure_decompression(d_comp); } catch (const std::exception&) { threw = true; }
    REQUIRE(threw);
    // out-of-scope formats are present as types but refuse construction
    threw = false;
    try { ZstdManager z{1 << 16, nvcompBatchedZstdDefaultOpts, stream}; } catch (const NVCompException& e) { threw = e.get_error() == nvcompErrorNotSupported; }
    REQUIRE(threw);
    CK(cudaFree(d_in)); CK(cudaFree(d_comp)); CK(cudaFree(d_out));
  }
  printf("hlif_test ok: %d round trips + checksum/exception checks\n", cases);
  return 0;
}

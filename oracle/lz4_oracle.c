/*
 * lz4_oracle.c -- CPU restatement of the LZ4 *block* format (test infrastructure;
 * see oracle.h).  Follows the format the reference names at
 * doc/algorithms_overview.md:48 (lz4_Block_format.md): sequences of
 *   token(hi nibble = literal length, lo nibble = match length - 4),
 *   [255-extension bytes], literals, 2-byte LE offset, [255-extension bytes];
 * the last sequence is literals only.  Pinned against liblz4 1.9.4 by
 * tests/test_oracle.py (both directions) and tests/golden/.
 */
#include "oracle.h"
#include <string.h>

static long lz4_walk(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int write)
{
  size_t ip = 0, op = 0;
  if (n == 0) return 0;
  for (;;) {
    if (ip >= n) return -1;
    unsigned tok = src[ip++];
    size_t ll = tok >> 4;
    if (ll == 15) {
      unsigned b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        ll += b;
      } while (b == 255);
    }
    if (ll > n - ip) return -1;
    if (write) {
      if (ll > cap - op) return -1;
      memcpy(dst + op, src + ip, ll);
    }
    ip += ll;
    op += ll;
    if (ip >= n) break; /* end of block: last sequence has no match part */
    if (n - ip < 2) return -1;
    size_t off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
    ip += 2;
    size_t ml = tok & 15;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (off == 0 || off > op) return -1;
    if (write) {
      if (ml > cap - op) return -1;
      /* byte-serial copy: overlapping matches replicate the pattern */
      if (off >= ml) memcpy(dst + op, dst + op - off, ml);
      else for (size_t i = 0; i < ml; ++i) dst[op + i] = dst[op + i - off];
    }
    op += ml;
  }
  return (long)op;
}

long oracle_lz4_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
  return lz4_walk(src, n, dst, cap, 1);
}

long oracle_lz4_decompressed_size(const uint8_t* src, size_t n)
{
  return lz4_walk(src, n, 0, 0, 0);
}

size_t oracle_lz4_bound(size_t n) { return n + n / 255 + 16; }

/* Greedy single-probe hash compressor.  Honours the end-of-block rules the
 * reference cites (CHANGELOG.md:195): the last 5 bytes are literals and the last
 * match starts at least 12 bytes before the end of the block. */
static size_t put_len(uint8_t* dst, size_t op, size_t rem)
{
  while (rem >= 255) { dst[op++] = 255; rem -= 255; }
  dst[op++] = (uint8_t)rem;
  return op;
}

long oracle_lz4_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
  enum { HLOG = 14 };
  static __thread uint32_t table[1 << HLOG];
  if (cap < oracle_lz4_bound(n)) return -1;
  memset(table, 0xff, sizeof(table));
  size_t anchor = 0, ip = 0, op = 0;
  const size_t mflimit = n > 12 ? n - 12 : 0;
  const size_t matchlimit = n > 5 ? n - 5 : 0;
  while (ip < mflimit) {
    uint32_t v;
    memcpy(&v, src + ip, 4);
    uint32_t h = (v * 2654435761u) >> (32 - HLOG);
    uint32_t cand = table[h];
    table[h] = (uint32_t)ip;
    uint32_t cv = 0;
    if (cand != 0xffffffffu && ip - cand <= 65535) memcpy(&cv, src + cand, 4);
    if (cand == 0xffffffffu || ip - cand > 65535 || cv != v) { ++ip; continue; }
    size_t ml = 4;
    while (ip + ml < matchlimit && src[ip + ml] == src[cand + ml]) ++ml;
    size_t ll = ip - anchor;
    size_t mlc = ml - 4;
    dst[op++] = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mlc < 15 ? mlc : 15));
    if (ll >= 15) op = put_len(dst, op, ll - 15);
    memcpy(dst + op, src + anchor, ll);
    op += ll;
    size_t off = ip - cand;
    dst[op++] = (uint8_t)(off & 255);
    dst[op++] = (uint8_t)(off >> 8);
    if (mlc >= 15) op = put_len(dst, op, mlc - 15);
    ip += ml;
    anchor = ip;
  }
  size_t ll = n - anchor;
  dst[op++] = (uint8_t)((ll < 15 ? ll : 15) << 4);
  if (ll >= 15) op = put_len(dst, op, ll - 15);
  memcpy(dst + op, src + anchor, ll);
  op += ll;
  return (long)op;
}

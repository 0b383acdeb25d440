/*
 * snappy_oracle.c -- CPU restatement of the Snappy *raw* format (test
 * infrastructure; see oracle.h).  Stream = varint32 uncompressed length, then
 * elements tagged by the low 2 bits of the tag byte:
 *   00 literal  : len-1 in the upper 6 bits (<60), or 60..63 => 1..4 LE length bytes
 *   01 copy-1   : len 4..11 = 4 + ((tag>>2)&7), offset = (tag>>5)<<8 | next byte
 *   10 copy-2   : len 1..64 = (tag>>2)+1, 2-byte LE offset
 *   11 copy-4   : len 1..64 = (tag>>2)+1, 4-byte LE offset
 * The decoder accepts every legal stream (reference CHANGELOG.md:182-184).
 * Pinned against pyarrow's bundled snappy by tests/test_oracle.py.
 */
#include "oracle.h"
#include <string.h>

static long snappy_walk(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int write)
{
  size_t ip = 0;
  uint64_t ulen = 0;
  unsigned shift = 0;
  for (;;) {
    if (ip >= n || shift > 28) return -1;
    unsigned b = src[ip++];
    ulen |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  if (ulen > 0xffffffffull) return -1;
  if (!write) return (long)ulen;
  if (ulen > cap) return -1;
  size_t op = 0;
  while (ip < n) {
    unsigned tag = src[ip++];
    size_t len, off;
    switch (tag & 3) {
    case 0: {
      len = (tag >> 2) + 1;
      if (len > 60) {
        unsigned nb = (unsigned)len - 60;
        if (n - ip < nb) return -1;
        len = 0;
        for (unsigned i = 0; i < nb; ++i) len |= (size_t)src[ip + i] << (8 * i);
        len += 1;
        ip += nb;
      }
      if (len > n - ip || len > ulen - op) return -1;
      memcpy(dst + op, src + ip, len);
      ip += len;
      op += len;
      continue;
    }
    case 1:
      if (ip >= n) return -1;
      len = 4 + ((tag >> 2) & 7);
      off = ((size_t)(tag >> 5) << 8) | src[ip++];
      break;
    case 2:
      if (n - ip < 2) return -1;
      len = (tag >> 2) + 1;
      off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
      ip += 2;
      break;
    default:
      if (n - ip < 4) return -1;
      len = (tag >> 2) + 1;
      off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8) | ((size_t)src[ip + 2] << 16)
            | ((size_t)src[ip + 3] << 24);
      ip += 4;
      break;
    }
    if (off == 0 || off > op || len > ulen - op) return -1;
    if (off >= len) memcpy(dst + op, dst + op - off, len);
    else for (size_t i = 0; i < len; ++i) dst[op + i] = dst[op + i - off];
    op += len;
  }
  if (op != ulen) return -1;
  return (long)op;
}

long oracle_snappy_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
  return snappy_walk(src, n, dst, cap, 1);
}

long oracle_snappy_decompressed_size(const uint8_t* src, size_t n)
{
  return snappy_walk(src, n, 0, 0, 0);
}

size_t oracle_snappy_bound(size_t n) { return 32 + n + n / 6; }

static size_t emit_literal(uint8_t* dst, size_t op, const uint8_t* lit, size_t len)
{
  if (len == 0) return op;
  size_t n1 = len - 1;
  if (n1 < 60) {
    dst[op++] = (uint8_t)(n1 << 2);
  } else {
    unsigned nb = n1 < (1u << 8) ? 1 : n1 < (1u << 16) ? 2 : n1 < (1u << 24) ? 3 : 4;
    dst[op++] = (uint8_t)((59 + nb) << 2);
    for (unsigned i = 0; i < nb; ++i) dst[op++] = (uint8_t)(n1 >> (8 * i));
  }
  memcpy(dst + op, lit, len);
  return op + len;
}

static size_t emit_copy_upto64(uint8_t* dst, size_t op, size_t off, size_t len)
{
  if (len < 12 && off < 2048 && len >= 4) {
    dst[op++] = (uint8_t)(1 | ((len - 4) << 2) | ((off >> 8) << 5));
    dst[op++] = (uint8_t)(off & 255);
  } else {
    dst[op++] = (uint8_t)(2 | ((len - 1) << 2));
    dst[op++] = (uint8_t)(off & 255);
    dst[op++] = (uint8_t)(off >> 8);
  }
  return op;
}

long oracle_snappy_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
  enum { HLOG = 14 };
  static __thread uint32_t table[1 << HLOG];
  if (cap < oracle_snappy_bound(n) || n > 0xffffffffull) return -1;
  memset(table, 0xff, sizeof(table));
  size_t op = 0;
  {
    uint32_t v = (uint32_t)n;
    while (v >= 0x80) { dst[op++] = (uint8_t)(v | 0x80); v >>= 7; }
    dst[op++] = (uint8_t)v;
  }
  size_t anchor = 0, ip = 0;
  const size_t limit = n > 4 ? n - 4 : 0;
  while (ip < limit) {
    uint32_t v;
    memcpy(&v, src + ip, 4);
    uint32_t h = (v * 0x1e35a7bdu) >> (32 - HLOG);
    uint32_t cand = table[h];
    table[h] = (uint32_t)ip;
    uint32_t cv = 0;
    if (cand != 0xffffffffu && ip - cand <= 65535) memcpy(&cv, src + cand, 4);
    if (cand == 0xffffffffu || ip - cand > 65535 || cv != v) { ++ip; continue; }
    size_t ml = 4;
    while (ip + ml < n && src[ip + ml] == src[cand + ml]) ++ml;
    op = emit_literal(dst, op, src + anchor, ip - anchor);
    size_t off = ip - cand, rem = ml;
    while (rem >= 68) { op = emit_copy_upto64(dst, op, off, 64); rem -= 64; }
    if (rem > 64) { op = emit_copy_upto64(dst, op, off, 60); rem -= 60; }
    op = emit_copy_upto64(dst, op, off, rem);
    ip += ml;
    anchor = ip;
  }
  op = emit_literal(dst, op, src + anchor, n - anchor);
  return (long)op;
}

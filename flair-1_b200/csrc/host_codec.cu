// Host-side TIFF LZW codec (compression scheme 5, MSB-first codes, "early change" code-width
// switch as written by libtiff/GDAL). The reference writes its class-map raster and PRED_*.tif
// patches LZW-compressed through GDAL (src/zone_detect/main.py:218-228, src/flair/writer.py:38-50);
// GDAL is absent here, so flair1_b200/geotiff.py encodes/decodes blocks with these two functions
// (called through ctypes, which releases the GIL, so blocks are coded on a thread pool).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/flair_b200.h"

namespace {

constexpr int kClear = 256, kEoi = 257, kFirst = 258, kMaxBits = 12;

struct BitWriter {
  uint8_t* dst;
  size_t cap, n = 0;
  uint64_t acc = 0;
  int nbits = 0;
  bool overflow = false;
  void put(uint32_t code, int width) {
    acc = (acc << width) | code;
    nbits += width;
    while (nbits >= 8) {
      if (n < cap) dst[n] = static_cast<uint8_t>(acc >> (nbits - 8)); else overflow = true;
      ++n;
      nbits -= 8;
    }
  }
  void flush() {
    if (nbits > 0) {
      if (n < cap) dst[n] = static_cast<uint8_t>(acc << (8 - nbits)); else overflow = true;
      ++n;
      nbits = 0;
    }
  }
};

}  // namespace

extern "C" {

int64_t fb_lzw_bound(int64_t n) { return n + n / 2 + 64; }

// Returns the number of bytes written, or -1 if `cap` is too small.
int64_t fb_lzw_encode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
  if (!src || !dst || n < 0) return -1;
  BitWriter bw{dst, static_cast<size_t>(cap)};
  // open-addressing hash: key = (prefix code << 8) | byte  ->  code
  constexpr int kHashBits = 14, kHashSize = 1 << kHashBits;
  std::vector<int32_t> hkey(kHashSize), hval(kHashSize);
  auto reset = [&]() { memset(hkey.data(), 0xFF, kHashSize * sizeof(int32_t)); };
  reset();
  int width = 9, next = kFirst;
  bw.put(kClear, width);
  if (n == 0) {
    bw.put(kEoi, width);
    bw.flush();
    return bw.overflow ? -1 : static_cast<int64_t>(bw.n);
  }
  int32_t w = src[0];
  for (int64_t i = 1; i < n; ++i) {
    const int32_t c = src[i];
    const int32_t key = (w << 8) | c;
    uint32_t h = (static_cast<uint32_t>(key) * 2654435761u) >> (32 - kHashBits);
    bool found = false;
    while (hkey[h] != -1) {
      if (hkey[h] == key) { found = true; break; }
      h = (h + 1) & (kHashSize - 1);
    }
    if (found) {
      w = hval[h];
      continue;
    }
    bw.put(static_cast<uint32_t>(w), width);
    hkey[h] = key;
    hval[h] = next++;
    if (next == (1 << kMaxBits) - 2) {  // 4094: table full -> clear (libtiff CODE_MAX-1)
      bw.put(kClear, width);
      reset();
      width = 9;
      next = kFirst;
    } else if (next > (1 << width) - 1) {
      ++width;
    }
    w = c;
  }
  bw.put(static_cast<uint32_t>(w), width);
  bw.put(kEoi, width);
  bw.flush();
  return bw.overflow ? -1 : static_cast<int64_t>(bw.n);
}

// Decodes up to `cap` bytes; returns the number of bytes produced or -1 on a corrupt stream.
int64_t fb_lzw_decode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
  if (!src || !dst || n < 0) return -1;
  std::vector<int32_t> prefix(1 << kMaxBits);
  std::vector<uint8_t> suffix(1 << kMaxBits), first(1 << kMaxBits);
  std::vector<int32_t> length(1 << kMaxBits);
  for (int i = 0; i < 256; ++i) { prefix[i] = -1; suffix[i] = static_cast<uint8_t>(i); first[i] = static_cast<uint8_t>(i); length[i] = 1; }
  int width = 9, next = kFirst;
  int32_t old = -1;
  uint64_t acc = 0;
  int nbits = 0;
  int64_t ip = 0, op = 0;
  while (true) {
    while (nbits < width && ip < n) { acc = (acc << 8) | src[ip++]; nbits += 8; }
    if (nbits < width) break;  // ran out of input without EOI: accept what we have
    const int32_t code = static_cast<int32_t>((acc >> (nbits - width)) & ((1u << width) - 1));
    nbits -= width;
    if (code == kEoi) break;
    if (code == kClear) {
      width = 9;
      next = kFirst;
      old = -1;
      continue;
    }
    int32_t cur = code;
    if (old == -1) {
      if (code >= 256) return -1;
    } else {
      if (code > next || next >= (1 << kMaxBits)) return -1;
      prefix[next] = old;
      first[next] = first[old];
      length[next] = length[old] + 1;
      suffix[next] = (code == next) ? first[old] : first[code];
      ++next;
      if (next == (1 << width) - 1 && width < kMaxBits) ++width;
    }
    // write the string for `cur` back to front
    const int32_t len = length[cur];
    if (op + len > cap) {
      // clip: produce only what fits (callers size dst exactly, this is a corrupt/over-long stream)
      return -1;
    }
    int64_t pos = op + len;
    for (int32_t t = cur; t != -1; t = prefix[t]) dst[--pos] = suffix[t];
    op += len;
    old = cur;
    if (op == cap) break;
  }
  return op;
}

}  // extern "C"

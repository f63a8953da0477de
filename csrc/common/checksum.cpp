#include "common/checksum.h"

#include <array>
#include <cstring>
#include <vector>

#include "common/tchash_def.h"

#if defined(__x86_64__)
#include <cpuid.h>
#include <nmmintrin.h>
#endif

namespace bb {

namespace {
constexpr uint32_t kPoly = 0x82F63B78u;  // reflected Castagnoli

struct Tables {
  uint32_t t[8][256];
  Tables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kPoly : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int j = 1; j < 8; ++j) t[j][i] = (t[j - 1][i] >> 8) ^ t[0][t[j - 1][i] & 0xFF];
  }
};
const Tables& tables() {
  static const Tables T;
  return T;
}

// register update with arbitrary initial register value (no xors)
uint32_t update_sw(uint32_t reg, const uint8_t* p, size_t len) noexcept {
  const auto& T = tables().t;
  while (len && (reinterpret_cast<uintptr_t>(p) & 7)) {
    reg = T[0][(reg ^ *p++) & 0xFF] ^ (reg >> 8);
    --len;
  }
  while (len >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    w ^= reg;
    reg = T[7][w & 0xFF] ^ T[6][(w >> 8) & 0xFF] ^ T[5][(w >> 16) & 0xFF] ^ T[4][(w >> 24) & 0xFF] ^
          T[3][(w >> 32) & 0xFF] ^ T[2][(w >> 40) & 0xFF] ^ T[1][(w >> 48) & 0xFF] ^ T[0][(w >> 56) & 0xFF];
    p += 8;
    len -= 8;
  }
  while (len--) reg = T[0][(reg ^ *p++) & 0xFF] ^ (reg >> 8);
  return reg;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t update_hw(uint32_t reg, const uint8_t* p, size_t len) noexcept {
  uint64_t r = reg;
  while (len && (reinterpret_cast<uintptr_t>(p) & 7)) {
    r = _mm_crc32_u8(static_cast<uint32_t>(r), *p++);
    --len;
  }
  while (len >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    r = _mm_crc32_u64(r, w);
    p += 8;
    len -= 8;
  }
  while (len--) r = _mm_crc32_u8(static_cast<uint32_t>(r), *p++);
  return static_cast<uint32_t>(r);
}
bool detect_hw() noexcept {
  unsigned a, b, c, d;
  if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
  return (c & bit_SSE4_2) != 0;
}
#else
uint32_t update_hw(uint32_t reg, const uint8_t* p, size_t len) noexcept { return update_sw(reg, p, len); }
bool detect_hw() noexcept { return false; }
#endif
}  // namespace

std::string_view to_string(ChecksumAlgo a) noexcept {
  switch (a) {
    case ChecksumAlgo::NONE: return "none";
    case ChecksumAlgo::CRC32C: return "crc32c";
    case ChecksumAlgo::BBH64: return "bbh64";
  }
  return "unknown";
}

bool crc32c_hw_available() noexcept {
  static const bool hw = detect_hw();
  return hw;
}

uint32_t crc32c_sw(const void* data, size_t len, uint32_t crc) noexcept {
  return update_sw(crc ^ 0xFFFFFFFFu, static_cast<const uint8_t*>(data), len) ^ 0xFFFFFFFFu;
}

uint32_t crc32c(const void* data, size_t len, uint32_t crc) noexcept {
  if (crc32c_hw_available()) return update_hw(crc ^ 0xFFFFFFFFu, static_cast<const uint8_t*>(data), len) ^ 0xFFFFFFFFu;
  return crc32c_sw(data, len, crc);
}

uint32_t crc32c_raw(const void* data, size_t len, uint32_t rem) noexcept {
  const auto* p = static_cast<const uint8_t*>(data);
  return crc32c_hw_available() ? update_hw(rem, p, len) : update_sw(rem, p, len);
}

// Reflected representation: bit 31 <-> x^0, bit 0 <-> x^31.
uint32_t gf2_mulmod(uint32_t a, uint32_t b) noexcept {
  uint32_t p = 0;
  for (uint32_t m = 1u << 31; m; m >>= 1) {
    if (a & m) p ^= b;
    b = (b & 1) ? (b >> 1) ^ kPoly : b >> 1;
  }
  return p;
}

uint32_t gf2_xpow_bytes(uint64_t nbytes) noexcept {
  // x^(8*nbytes) by square-and-multiply; x^8 = 0x00800000 in reflected form.
  uint32_t result = 0x80000000u;  // 1
  uint32_t base = 0x00800000u;    // x^8
  while (nbytes) {
    if (nbytes & 1) result = gf2_mulmod(result, base);
    base = gf2_mulmod(base, base);
    nbytes >>= 1;
  }
  return result;
}

uint32_t crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) noexcept {
  return gf2_mulmod(gf2_xpow_bytes(len_b), crc_a) ^ crc_b;
}

uint32_t crc32c_from_raw(uint32_t raw, uint64_t len) noexcept {
  return raw ^ gf2_mulmod(0xFFFFFFFFu, gf2_xpow_bytes(len)) ^ 0xFFFFFFFFu;
}

void crc32c_shift_table(uint64_t nbytes, uint32_t t[4][256]) noexcept {
  const uint32_t xp = gf2_xpow_bytes(nbytes);
  for (int j = 0; j < 4; ++j)
    for (uint32_t b = 0; b < 256; ++b) t[j][b] = gf2_mulmod(b << (8 * j), xp);
}

uint64_t bbh64(const void* data, size_t len) noexcept {
  using namespace tchash;
  static const std::array<std::array<uint8_t, kN>, kK> W = [] {
    std::array<std::array<uint8_t, kN>, kK> w{};
    for (uint32_t k = 0; k < kK; ++k)
      for (uint32_t n = 0; n < kN; ++n) w[k][n] = static_cast<uint8_t>(weight(k, n));
    return w;
  }();
  static const std::array<uint64_t, kN> KN = [] {
    std::array<uint64_t, kN> v{};
    for (uint32_t n = 0; n < kN; ++n) v[n] = col_mul(n);
    return v;
  }();
  const auto* p = static_cast<const uint8_t*>(data);
  const uint64_t ntiles = (len + kTileBytes - 1) / kTileBytes;
  uint64_t sum = 0;
  std::vector<uint32_t> D(kRows * kN);
  for (uint64_t t = 0; t < ntiles; ++t) {
    std::fill(D.begin(), D.end(), 0u);
    const uint64_t base = t * kTileBytes;
    const uint64_t n = len > base ? std::min<uint64_t>(kTileBytes, len - base) : 0;
    for (uint32_t o = 0; o < n; ++o) {
      const uint32_t a = p[base + o];
      if (!a) continue;
      const uint32_t m = off_to_row(o), k = off_to_k(o);
      uint32_t* d = &D[m * kN];
      const auto& w = W[k];
      for (uint32_t c = 0; c < kN; ++c) d[c] += a * w[c];
    }
    for (uint32_t m = 0; m < kRows; ++m) {
      uint64_t r = 0;
      for (uint32_t c = 0; c < kN; ++c) r += static_cast<uint64_t>(D[m * kN + c]) * KN[c];
      sum += row_contrib(r, t * kRows + m);
    }
  }
  return finalize(sum, len);
}

uint64_t checksum(ChecksumAlgo algo, const void* data, size_t len) noexcept {
  switch (algo) {
    case ChecksumAlgo::NONE: return 0;
    case ChecksumAlgo::CRC32C: return crc32c(data, len);
    case ChecksumAlgo::BBH64: return bbh64(data, len);
  }
  return 0;
}

}  // namespace bb

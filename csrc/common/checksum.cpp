#include "common/checksum.h"

#include <array>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common/tchash_def.h"
#include "common/xxh3.h"

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
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
uint32_t gf2_mulmod_local(uint32_t a, uint32_t b) noexcept {
  uint32_t p = 0;
  for (uint32_t m = 1u << 31; m; m >>= 1) {
    if (a & m) p ^= b;
    b = (b & 1) ? (b >> 1) ^ kPoly : b >> 1;
  }
  return p;
}

// The crc32 instruction has a 3-cycle latency and a 1-cycle throughput: three independent streams keep it busy.
// Blocks of 3 x kLane bytes are hashed as three lanes and merged with the x^(8*kLane) shift (4 table lookups).
constexpr size_t kLane = 4096;
struct LaneShift {
  uint32_t t[4][256];
  LaneShift() {
    uint32_t xp = 0x80000000u, base = 0x00800000u;  // x^(8*kLane) by square-and-multiply
    for (size_t n = kLane; n; n >>= 1) {
      if (n & 1) xp = gf2_mulmod_local(xp, base);
      base = gf2_mulmod_local(base, base);
    }
    for (int j = 0; j < 4; ++j)
      for (uint32_t b = 0; b < 256; ++b) t[j][b] = gf2_mulmod_local(b << (8 * j), xp);
  }
  uint32_t operator()(uint32_t s) const noexcept { return t[0][s & 255] ^ t[1][(s >> 8) & 255] ^ t[2][(s >> 16) & 255] ^ t[3][s >> 24]; }
};

__attribute__((target("sse4.2"))) uint32_t update_hw(uint32_t reg, const uint8_t* p, size_t len) noexcept {
  uint64_t r = reg;
  while (len && (reinterpret_cast<uintptr_t>(p) & 7)) {
    r = _mm_crc32_u8(static_cast<uint32_t>(r), *p++);
    --len;
  }
  if (len >= 3 * kLane) {
    static const LaneShift shift;
    while (len >= 3 * kLane) {
      uint64_t c0 = r, c1 = 0, c2 = 0;
      const uint8_t* q = p;
      for (size_t i = 0; i < kLane; i += 8) {
        uint64_t w0, w1, w2;
        std::memcpy(&w0, q + i, 8);
        std::memcpy(&w1, q + kLane + i, 8);
        std::memcpy(&w2, q + 2 * kLane + i, 8);
        c0 = _mm_crc32_u64(c0, w0);
        c1 = _mm_crc32_u64(c1, w1);
        c2 = _mm_crc32_u64(c2, w2);
      }
      r = shift(shift(static_cast<uint32_t>(c0)) ^ static_cast<uint32_t>(c1)) ^ static_cast<uint32_t>(c2);
      p += 3 * kLane;
      len -= 3 * kLane;
    }
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
    case ChecksumAlgo::XXH3: return "xxh3";
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

namespace {
using namespace tchash;

struct BbhConsts {
  std::array<std::array<uint8_t, kN>, kK> W;  // W[k][n]
  std::array<uint64_t, kN> KN;
  uint64_t kn_sum128 = 0;                     // 128 * sum_n KN[n]  (mod 2^64)
  // SIMD operand layouts of W' = W - 128 (signed 8 bit) / W (16 bit):
  alignas(64) int8_t w4[kK / 4][kN][4];       // [k group of 4][column][k in group]  -> one zmm per k group (vpdpbusd)
  alignas(32) int16_t w2[kK / 2][kN][2];      // [k pair][column][k in pair]         -> two ymm per k pair (vpmaddwd)
  BbhConsts() {
    for (uint32_t k = 0; k < kK; ++k)
      for (uint32_t n = 0; n < kN; ++n) W[k][n] = static_cast<uint8_t>(weight(k, n));
    uint64_t s = 0;
    for (uint32_t n = 0; n < kN; ++n) {
      KN[n] = col_mul(n);
      s += KN[n];
    }
    kn_sum128 = s * 128ull;
    for (uint32_t k = 0; k < kK; ++k)
      for (uint32_t n = 0; n < kN; ++n) {
        w4[k / 4][n][k % 4] = static_cast<int8_t>(static_cast<int>(W[k][n]) - 128);
        w2[k / 2][n][k % 2] = static_cast<int16_t>(W[k][n]);
      }
  }
};
const BbhConsts& bbh_consts() {
  static const BbhConsts c;
  return c;
}

// One full (zero padded) 16 KiB tile, byte-at-a-time definition.
uint64_t tile_sum_scalar(const uint8_t* tile, uint64_t tile_index) noexcept {
  const BbhConsts& C = bbh_consts();
  uint32_t D[kRows * kN] = {0};
  for (uint32_t o = 0; o < kTileBytes; ++o) {
    const uint32_t a = tile[o];
    if (!a) continue;
    uint32_t* d = &D[off_to_row(o) * kN];
    const auto& w = C.W[off_to_k(o)];
    for (uint32_t c = 0; c < kN; ++c) d[c] += a * w[c];
  }
  uint64_t sum = 0;
  for (uint32_t m = 0; m < kRows; ++m) {
    uint64_t r = 0;
    for (uint32_t c = 0; c < kN; ++c) r += static_cast<uint64_t>(D[m * kN + c]) * C.KN[c];
    sum += row_contrib(r, tile_index * kRows + m);
  }
  return sum;
}

#if defined(__x86_64__)
// Row m of a tile: block b = m / 8 (1 KiB), the 128 k bytes are 8 chunks of 16 bytes at b*1024 + kc*128 + (m%8)*16.
// vpdpbusd: every 32-bit lane n accumulates sum_j u8(A[m][4g+j]) * s8(W'[4g+j][n]); the 4 A bytes are broadcast to all
// 16 lanes (= the 16 hash columns).  D[m][n] = acc[n] + 128 * rowsum(A[m]) restores the unsigned weights exactly.
__attribute__((target("avx512f,avx512bw,avx512dq,avx512vl,avx512vnni"))) uint64_t tile_sum_vnni(const uint8_t* tile,
                                                                                                 uint64_t tile_index) noexcept {
  const BbhConsts& C = bbh_consts();
  const __m512i kn_lo = _mm512_loadu_si512(&C.KN[0]);
  const __m512i kn_hi = _mm512_loadu_si512(&C.KN[8]);
  uint64_t sum = 0;
  for (uint32_t b = 0; b < kRows / 8; ++b) {
    const uint8_t* blk = tile + b * 1024;
    for (uint32_t mi = 0; mi < 8; mi += 4) {  // four rows per pass share the W' loads
      // two accumulators per row (even / odd K chunk): 8 independent vpdpbusd chains hide the 5-cycle latency
      __m512i acc[4] = {_mm512_setzero_si512(), _mm512_setzero_si512(), _mm512_setzero_si512(), _mm512_setzero_si512()};
      __m512i acd[4] = {_mm512_setzero_si512(), _mm512_setzero_si512(), _mm512_setzero_si512(), _mm512_setzero_si512()};
      __m128i sad[4] = {_mm_setzero_si128(), _mm_setzero_si128(), _mm_setzero_si128(), _mm_setzero_si128()};
#pragma GCC unroll 8
      for (uint32_t kc = 0; kc < 8; ++kc) {
        const uint8_t* c0 = blk + kc * 128 + mi * 16;
#pragma GCC unroll 4
        for (uint32_t r = 0; r < 4; ++r)
          sad[r] = _mm_add_epi64(sad[r], _mm_sad_epu8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(c0 + 16 * r)), _mm_setzero_si128()));
#pragma GCC unroll 4
        for (uint32_t j = 0; j < 4; ++j) {
          const __m512i w = _mm512_load_si512(&C.w4[kc * 4 + j][0][0]);
#pragma GCC unroll 4
          for (uint32_t r = 0; r < 4; ++r) {
            int32_t a;
            std::memcpy(&a, c0 + 16 * r + 4 * j, 4);
            if (kc & 1) acd[r] = _mm512_dpbusd_epi32(acd[r], _mm512_set1_epi32(a), w);
            else acc[r] = _mm512_dpbusd_epi32(acc[r], _mm512_set1_epi32(a), w);
          }
        }
      }
      for (uint32_t r = 0; r < 4; ++r) {
        const uint32_t rowsum = static_cast<uint32_t>(_mm_cvtsi128_si64(sad[r]) + _mm_extract_epi64(sad[r], 1));
        const __m512i d = _mm512_add_epi32(_mm512_add_epi32(acc[r], acd[r]), _mm512_set1_epi32(static_cast<int32_t>(rowsum * 128u)));  // exact u32 D[m][n]
        const __m512i lo = _mm512_cvtepu32_epi64(_mm512_castsi512_si256(d));
        const __m512i hi = _mm512_cvtepu32_epi64(_mm512_extracti64x4_epi64(d, 1));
        const __m512i pr = _mm512_add_epi64(_mm512_mullo_epi64(lo, kn_lo), _mm512_mullo_epi64(hi, kn_hi));
        const uint64_t rr = static_cast<uint64_t>(_mm512_reduce_add_epi64(pr));
        sum += row_contrib(rr, tile_index * kRows + b * 8 + mi + r);
      }
    }
  }
  return sum;
}

// AVX2: zero-extended bytes against 16-bit weights with vpmaddwd (exact, no saturation): lane n of the two 8-lane
// accumulators gets A[m][k]*W[k][n] + A[m][k+1]*W[k+1][n].
__attribute__((target("avx2"))) uint64_t tile_sum_avx2(const uint8_t* tile, uint64_t tile_index) noexcept {
  const BbhConsts& C = bbh_consts();
  uint64_t sum = 0;
  for (uint32_t b = 0; b < kRows / 8; ++b) {
    const uint8_t* blk = tile + b * 1024;
    for (uint32_t mi = 0; mi < 8; ++mi) {
      __m256i acc_lo = _mm256_setzero_si256(), acc_hi = _mm256_setzero_si256();
      for (uint32_t kc = 0; kc < 8; ++kc) {
        const uint8_t* c = blk + kc * 128 + mi * 16;
        for (uint32_t j = 0; j < 8; ++j) {  // k pair (kc*16 + 2j, +1)
          const uint32_t pair = static_cast<uint32_t>(c[2 * j]) | (static_cast<uint32_t>(c[2 * j + 1]) << 16);
          const __m256i a = _mm256_set1_epi32(static_cast<int32_t>(pair));
          const int16_t* w = &C.w2[kc * 8 + j][0][0];
          acc_lo = _mm256_add_epi32(acc_lo, _mm256_madd_epi16(a, _mm256_load_si256(reinterpret_cast<const __m256i*>(w))));
          acc_hi = _mm256_add_epi32(acc_hi, _mm256_madd_epi16(a, _mm256_load_si256(reinterpret_cast<const __m256i*>(w + 16))));
        }
      }
      alignas(32) uint32_t d[16];
      _mm256_store_si256(reinterpret_cast<__m256i*>(d), acc_lo);
      _mm256_store_si256(reinterpret_cast<__m256i*>(d + 8), acc_hi);
      uint64_t rr = 0;
      for (uint32_t n = 0; n < kN; ++n) rr += static_cast<uint64_t>(d[n]) * C.KN[n];
      sum += row_contrib(rr, tile_index * kRows + b * 8 + mi);
    }
  }
  return sum;
}
#endif

using TileFn = uint64_t (*)(const uint8_t*, uint64_t) noexcept;
struct BbhImpl {
  TileFn fn = tile_sum_scalar;
  const char* name = "scalar";
  BbhImpl() {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
        __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vnni")) {
      fn = tile_sum_vnni;
      name = "avx512-vnni";
    } else if (__builtin_cpu_supports("avx2")) {
      fn = tile_sum_avx2;
      name = "avx2";
    }
    if (const char* e = std::getenv("BB_BBH64_IMPL")) {  // tests / diagnostics: force a path the CPU supports
      const std::string want(e);
      if (want == "scalar") fn = tile_sum_scalar, name = "scalar";
      else if (want == "avx2" && __builtin_cpu_supports("avx2")) fn = tile_sum_avx2, name = "avx2";
    }
#endif
  }
};
const BbhImpl& bbh_impl() {
  static const BbhImpl i;
  return i;
}

// Tiles [first_tile, first_tile + ntiles) of the buffer p[0, len); `index_base` is added to the tile index that is mixed
// into the hash (a chunk of a larger object is hashed as tiles index_base.. of that object).
uint64_t tiles_sum(TileFn fn, const uint8_t* p, size_t len, uint64_t first_tile, uint64_t ntiles, uint64_t index_base = 0) noexcept {
  const uint64_t total = (len + kTileBytes - 1) / kTileBytes;
  uint64_t sum = 0;
  for (uint64_t t = first_tile; t < first_tile + ntiles && t < total; ++t) {
    const uint64_t base = t * kTileBytes;
    if (len - base >= kTileBytes) {
      sum += fn(p + base, index_base + t);
    } else {  // short last tile: zero padded
      alignas(64) uint8_t pad[kTileBytes];
      std::memcpy(pad, p + base, len - base);
      std::memset(pad + (len - base), 0, kTileBytes - (len - base));
      sum += fn(pad, index_base + t);
    }
  }
  return sum;
}
}  // namespace

const char* bbh64_impl_name() noexcept { return bbh_impl().name; }

uint64_t bbh64_partial(const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept {
  return tiles_sum(bbh_impl().fn, static_cast<const uint8_t*>(data), len, first_tile, ntiles);
}

uint64_t bbh64_chunk(const void* data, size_t len, uint64_t tile_base) noexcept {
  return tiles_sum(bbh_impl().fn, static_cast<const uint8_t*>(data), len, 0, (len + kTileBytes - 1) / kTileBytes, tile_base);
}

uint64_t bbh64_finalize(uint64_t tile_sum, size_t len) noexcept { return tchash::finalize(tile_sum, len); }

uint64_t bbh64(const void* data, size_t len) noexcept {
  return tchash::finalize(bbh64_partial(data, len, 0, (len + kTileBytes - 1) / kTileBytes), len);
}

uint64_t bbh64_using(std::string_view impl, const void* data, size_t len, bool* supported) noexcept {
  TileFn fn = nullptr;
  if (impl == "scalar") fn = tile_sum_scalar;
#if defined(__x86_64__)
  __builtin_cpu_init();
  if (impl == "avx2" && __builtin_cpu_supports("avx2")) fn = tile_sum_avx2;
  if (impl == "avx512-vnni" && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
      __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vnni"))
    fn = tile_sum_vnni;
#endif
  if (supported) *supported = fn != nullptr;
  if (!fn) return 0;
  return tchash::finalize(tiles_sum(fn, static_cast<const uint8_t*>(data), len, 0, (len + kTileBytes - 1) / kTileBytes), len);
}

uint64_t bbh64_reference(const void* data, size_t len) noexcept {
  return tchash::finalize(tiles_sum(tile_sum_scalar, static_cast<const uint8_t*>(data), len, 0, (len + kTileBytes - 1) / kTileBytes), len);
}

// ================================================================ XXH3 (tiled)
namespace {
uint64_t xxh_tiles_sum(const uint8_t* p, size_t len, uint64_t first_tile, uint64_t ntiles, uint64_t index_base) noexcept {
  const uint64_t total = (len + kTileBytes - 1) / kTileBytes;
  uint64_t sum = 0;
  for (uint64_t t = first_tile; t < first_tile + ntiles && t < total; ++t) {
    const uint64_t base = t * kTileBytes;
    uint64_t h;
    if (len - base >= kTileBytes) {
      h = xxh3::tile_hash(p + base);
    } else {  // short last tile: zero padded
      alignas(64) uint8_t pad[kTileBytes];
      std::memcpy(pad, p + base, len - base);
      std::memset(pad + (len - base), 0, kTileBytes - (len - base));
      h = xxh3::tile_hash(pad);
    }
    sum += tchash::mix64(h + (index_base + t + 1ull) * tchash::kGold);
  }
  return sum;
}
}  // namespace

uint64_t xxh3_tile(const void* tile16k) noexcept { return xxh3::tile_hash(static_cast<const uint8_t*>(tile16k)); }
uint64_t xxh3t64_partial(const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept {
  return xxh_tiles_sum(static_cast<const uint8_t*>(data), len, first_tile, ntiles, 0);
}
uint64_t xxh3t64_chunk(const void* data, size_t len, uint64_t tile_base) noexcept {
  return xxh_tiles_sum(static_cast<const uint8_t*>(data), len, 0, (len + kTileBytes - 1) / kTileBytes, tile_base);
}
uint64_t xxh3t64(const void* data, size_t len) noexcept {
  return tchash::finalize(xxh3t64_partial(data, len, 0, (len + kTileBytes - 1) / kTileBytes), len);
}

uint64_t tile_sum_partial(ChecksumAlgo a, const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept {
  return a == ChecksumAlgo::XXH3 ? xxh3t64_partial(data, len, first_tile, ntiles) : bbh64_partial(data, len, first_tile, ntiles);
}
uint64_t tile_sum_chunk(ChecksumAlgo a, const void* data, size_t len, uint64_t tile_base) noexcept {
  return a == ChecksumAlgo::XXH3 ? xxh3t64_chunk(data, len, tile_base) : bbh64_chunk(data, len, tile_base);
}
uint64_t tile_sum_finalize(uint64_t tile_sum, size_t len) noexcept { return tchash::finalize(tile_sum, len); }

uint64_t checksum(ChecksumAlgo algo, const void* data, size_t len) noexcept {
  switch (algo) {
    case ChecksumAlgo::NONE: return 0;
    case ChecksumAlgo::XXH3: return xxh3t64(data, len);
    case ChecksumAlgo::CRC32C: return crc32c(data, len);
    case ChecksumAlgo::BBH64: return bbh64(data, len);
  }
  return 0;
}

}  // namespace bb

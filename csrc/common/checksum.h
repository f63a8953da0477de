// CPU reference checksums: CRC32C (Castagnoli; SSE4.2 when available, slice-by-8 otherwise),
// CRC32C combine/shift algebra (needed because GPU tiles finish out of order), and BBH64.
// These are the golden models the CUDA kernels are tested against and what the host tiers
// (DRAM / NVMe) use to verify data that left the GPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string_view>

namespace bb {

// XXH3 = standard XXH3-64 of every 16 KiB tile, combined order-independently (common/xxh3.h).
enum class ChecksumAlgo : uint32_t { NONE = 0, CRC32C = 1, BBH64 = 2, XXH3 = 3 };
std::string_view to_string(ChecksumAlgo a) noexcept;

// Standard CRC32C: init 0xFFFFFFFF, reflected poly 0x82F63B78, final xor.  `crc` chains calls.
uint32_t crc32c(const void* data, size_t len, uint32_t crc = 0) noexcept;
uint32_t crc32c_sw(const void* data, size_t len, uint32_t crc = 0) noexcept;  // table path (always available)
bool crc32c_hw_available() noexcept;

// crc(A||B) from crc(A), crc(B), len(B).
uint32_t crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) noexcept;

// Raw remainder algebra over GF(2)[x]/P (reflected bit order, no init / final xor):
//   crc32c_raw(data) = data(x) * x^32 mod P
uint32_t crc32c_raw(const void* data, size_t len, uint32_t rem = 0) noexcept;
uint32_t gf2_mulmod(uint32_t a, uint32_t b) noexcept;  // a(x)*b(x) mod P
uint32_t gf2_xpow_bytes(uint64_t nbytes) noexcept;      // x^(8*nbytes) mod P
// Converts a raw remainder of the whole message into the standard CRC32C value.
uint32_t crc32c_from_raw(uint32_t raw, uint64_t len) noexcept;
// Fills t[4][256] such that for a 32-bit remainder s:  s * x^(8*nbytes) mod P =
//   t[0][s&255] ^ t[1][(s>>8)&255] ^ t[2][(s>>16)&255] ^ t[3][s>>24]   (tables used by the GPU kernel)
void crc32c_shift_table(uint64_t nbytes, uint32_t t[4][256]) noexcept;

// BBH64 (tchash_def.h).  bbh64() dispatches to an AVX-512 VNNI or AVX2 implementation of the same integer arithmetic
// when the CPU has one (the host tiers and TCP clients hash at memory speed instead of 0.3 GB/s); bbh64_reference() is
// the byte-at-a-time definition the CUDA kernels and the SIMD paths are tested against.
uint64_t bbh64(const void* data, size_t len) noexcept;
uint64_t bbh64_reference(const void* data, size_t len) noexcept;
// Unfinalised sum of the tiles [first_tile, first_tile + ntiles) of an object of `len` bytes starting at `data`:
// the tile sums are commutative, so a large object can be hashed by several threads and the parts added;
// bbh64_finalize(sum of the parts, len) is the digest.
uint64_t bbh64_partial(const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept;
// Unfinalised sum of a chunk [data, data + len) that is the tiles tile_base.. of a larger object (streaming: the chunk
// must start on a tile boundary of the object; only the object's last chunk may end inside a tile).
uint64_t bbh64_chunk(const void* data, size_t len, uint64_t tile_base) noexcept;
uint64_t bbh64_finalize(uint64_t tile_sum, size_t len) noexcept;
const char* bbh64_impl_name() noexcept;  // "avx512-vnni" | "avx2" | "scalar"
// Digest through a named implementation (tests); *supported = false when this CPU cannot run it.
uint64_t bbh64_using(std::string_view impl, const void* data, size_t len, bool* supported) noexcept;

// XXH3 (common/xxh3.h): digest of a whole buffer, and the streaming pieces (same contract as bbh64_partial / _chunk).
uint64_t xxh3t64(const void* data, size_t len) noexcept;
uint64_t xxh3_tile(const void* tile16k) noexcept;  // the standard XXH3_64bits() of exactly 16384 bytes
uint64_t xxh3t64_partial(const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept;
uint64_t xxh3t64_chunk(const void* data, size_t len, uint64_t tile_base) noexcept;

// The two tile-sum digests (BBH64, XXH3) share their algebra: digest = finalize(sum of per-tile terms, len).  Code that
// hashes in pieces (parallel streams, bounded chunks, GPU slices) goes through these.
constexpr bool is_tile_sum(ChecksumAlgo a) noexcept { return a == ChecksumAlgo::BBH64 || a == ChecksumAlgo::XXH3; }
uint64_t tile_sum_partial(ChecksumAlgo a, const void* data, size_t len, uint64_t first_tile, uint64_t ntiles) noexcept;
uint64_t tile_sum_chunk(ChecksumAlgo a, const void* data, size_t len, uint64_t tile_base) noexcept;
uint64_t tile_sum_finalize(uint64_t tile_sum, size_t len) noexcept;  // the same for both

uint64_t checksum(ChecksumAlgo algo, const void* data, size_t len) noexcept;

}  // namespace bb

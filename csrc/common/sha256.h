// SHA-256 (FIPS 180-4) and HMAC-SHA256 (RFC 2104) for the cluster-token handshake (net/tcp.h): the shared secret is
// only ever used as an HMAC key over fresh nonces, it never travels.  Small, dependency-free, not a bulk hash.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>

namespace bb {

using Sha256Digest = std::array<uint8_t, 32>;

class Sha256 {
 public:
  Sha256() noexcept;
  void update(const void* data, size_t len) noexcept;
  Sha256Digest finish() noexcept;  // the object must not be reused afterwards

 private:
  void block(const uint8_t* p) noexcept;
  uint32_t h_[8];
  uint8_t buf_[64];
  size_t fill_ = 0;
  uint64_t total_ = 0;
};

Sha256Digest sha256(std::string_view data) noexcept;
Sha256Digest hmac_sha256(std::string_view key, std::string_view msg) noexcept;
std::string to_hex(const Sha256Digest& d);
// Constant-time comparison of two equally sized MACs.
bool mac_equal(const void* a, const void* b, size_t len) noexcept;

}  // namespace bb

// AES-256-GCM for the framed RPC protocol's secure mode (tcp.h, "BBA2" handshake).
//
// The cipher is OpenSSL's (libcrypto's EVP interface), looked up with dlopen() the first time secure mode is asked for:
// nothing links against it, a build or a host without libcrypto simply cannot enable `encrypt_transport` (and says so),
// and no cipher code lives in this tree.  One Aead object protects one direction of one connection: the key is derived
// from the cluster token and both handshake nonces, the 96-bit IV is the direction's message counter, so a (key, IV)
// pair never repeats and a replayed, dropped or reordered frame fails authentication.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace bb::net {

constexpr size_t kAeadKey = 32, kAeadTag = 16;

class Aead {
 public:
  Aead() = default;
  ~Aead();
  Aead(const Aead&) = delete;
  Aead& operator=(const Aead&) = delete;

  // libcrypto found and usable?  `why` receives the reason when not.
  static bool available(std::string* why = nullptr);

  // `for_open`: this object will open() (receive direction); otherwise it will seal().
  bool set_key(const uint8_t key[kAeadKey], bool for_open);
  bool ready() const { return ctx_ != nullptr; }

  struct Span {
    void* data;
    size_t len;
  };
  struct CSpan {
    const void* data;
    size_t len;
  };
  // out receives the ciphertext of the concatenated pieces (same length) followed by the 16-byte tag.  `aad` (the clear
  // frame header) is authenticated, not encrypted.  Uses and then advances the message counter.
  bool seal(const void* aad, size_t aad_len, const CSpan* pieces, int n, char* out);
  // Decrypts the pieces in place (a message may be split over a header buffer and a destination buffer) and checks `tag`.
  // False = forged / corrupted / out of order: the caller drops the connection.
  bool open(const void* aad, size_t aad_len, const Span* pieces, int n, const char* tag);

 private:
  void* ctx_ = nullptr;     // EVP_CIPHER_CTX with the key installed
  uint64_t counter_ = 0;    // messages protected so far in this direction
};

// AES-256-CTR addressed by byte offset, for data at rest (file-backed tiers): byte i of the pool is XORed with byte i of a key
// stream that depends only on (key, pool nonce, i), so any range can be written and read independently and in place.  No
// authentication here -- the object digests (computed on the plain bytes) already detect any change.  Same libcrypto, same
// dlopen.
class OffsetCipher {
 public:
  bool set_key(const uint8_t key[kAeadKey], const uint8_t nonce[8]);
  bool ready() const { return ready_; }
  // out[0..n) = in[0..n) XOR keystream[offset .. offset + n); in == out is fine.  Thread-safe (a context per call).
  bool crypt(uint64_t offset, const void* in, void* out, size_t n) const;

 private:
  uint8_t key_[kAeadKey] = {};
  uint8_t nonce_[8] = {};
  bool ready_ = false;
};

// key = HMAC-SHA256(token, label || nonces): one key per direction ("bb-key-c2s" / "bb-key-s2c").
void derive_key(const std::string& token, const char* label, const std::string& nonces, uint8_t out[kAeadKey]);

}  // namespace bb::net

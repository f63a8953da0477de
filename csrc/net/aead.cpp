#include "net/aead.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common/sha256.h"

namespace bb::net {

namespace {
// The handful of EVP entry points AES-GCM needs, as libcrypto 1.1 / 3.x export them.
struct Crypto {
  void* (*ctx_new)() = nullptr;
  void (*ctx_free)(void*) = nullptr;
  const void* (*aes_256_gcm)() = nullptr;
  int (*enc_init)(void*, const void*, void*, const unsigned char*, const unsigned char*) = nullptr;
  int (*enc_update)(void*, unsigned char*, int*, const unsigned char*, int) = nullptr;
  int (*enc_final)(void*, unsigned char*, int*) = nullptr;
  int (*dec_init)(void*, const void*, void*, const unsigned char*, const unsigned char*) = nullptr;
  int (*dec_update)(void*, unsigned char*, int*, const unsigned char*, int) = nullptr;
  int (*dec_final)(void*, unsigned char*, int*) = nullptr;
  int (*ctx_ctrl)(void*, int, int, void*) = nullptr;
  const void* (*aes_256_ctr)() = nullptr;
  std::string error;
  bool ok = false;
};
constexpr int kCtrlSetIvLen = 0x9, kCtrlGetTag = 0x10, kCtrlSetTag = 0x11;  // EVP_CTRL_AEAD_*

const Crypto& crypto() {
  static Crypto c;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libcrypto.so.3", "libcrypto.so.1.1", "libcrypto.so"}) {
      h = ::dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) {
      c.error = "libcrypto not found (dlopen): encrypt_transport needs OpenSSL's libcrypto on this host";
      return;
    }
    auto sym = [&](const char* n) { return ::dlsym(h, n); };
    c.ctx_new = reinterpret_cast<decltype(c.ctx_new)>(sym("EVP_CIPHER_CTX_new"));
    c.ctx_free = reinterpret_cast<decltype(c.ctx_free)>(sym("EVP_CIPHER_CTX_free"));
    c.aes_256_gcm = reinterpret_cast<decltype(c.aes_256_gcm)>(sym("EVP_aes_256_gcm"));
    c.enc_init = reinterpret_cast<decltype(c.enc_init)>(sym("EVP_EncryptInit_ex"));
    c.enc_update = reinterpret_cast<decltype(c.enc_update)>(sym("EVP_EncryptUpdate"));
    c.enc_final = reinterpret_cast<decltype(c.enc_final)>(sym("EVP_EncryptFinal_ex"));
    c.dec_init = reinterpret_cast<decltype(c.dec_init)>(sym("EVP_DecryptInit_ex"));
    c.dec_update = reinterpret_cast<decltype(c.dec_update)>(sym("EVP_DecryptUpdate"));
    c.dec_final = reinterpret_cast<decltype(c.dec_final)>(sym("EVP_DecryptFinal_ex"));
    c.ctx_ctrl = reinterpret_cast<decltype(c.ctx_ctrl)>(sym("EVP_CIPHER_CTX_ctrl"));
    c.aes_256_ctr = reinterpret_cast<decltype(c.aes_256_ctr)>(sym("EVP_aes_256_ctr"));
    c.ok = c.ctx_new && c.ctx_free && c.aes_256_gcm && c.enc_init && c.enc_update && c.enc_final && c.dec_init && c.dec_update && c.dec_final && c.ctx_ctrl;
    if (!c.ok) c.error = "libcrypto lacks the EVP AES-GCM interface";
  });
  return c;
}

void make_iv(uint64_t counter, unsigned char iv[12]) {
  std::memset(iv, 0, 4);
  for (int i = 0; i < 8; ++i) iv[4 + i] = static_cast<unsigned char>(counter >> (56 - 8 * i));
}
// EVP takes int lengths: feed large pieces in slices
constexpr size_t kSlice = 1u << 30;
}  // namespace

bool Aead::available(std::string* why) {
  const Crypto& c = crypto();
  if (!c.ok && why) *why = c.error;
  return c.ok;
}

Aead::~Aead() {
  if (ctx_) crypto().ctx_free(ctx_);
}

bool Aead::set_key(const uint8_t key[kAeadKey], bool for_open) {
  const Crypto& c = crypto();
  if (!c.ok) return false;
  if (!ctx_) ctx_ = c.ctx_new();
  if (!ctx_) return false;
  // cipher, IV length and key once; each message re-initialises with key = NULL (keeps the schedule) and its own IV
  auto init = for_open ? c.dec_init : c.enc_init;
  if (init(ctx_, c.aes_256_gcm(), nullptr, nullptr, nullptr) != 1 || c.ctx_ctrl(ctx_, kCtrlSetIvLen, 12, nullptr) != 1 ||
      init(ctx_, nullptr, nullptr, key, nullptr) != 1) {
    c.ctx_free(ctx_);
    ctx_ = nullptr;
    return false;
  }
  counter_ = 0;
  return true;
}

bool Aead::seal(const void* aad, size_t aad_len, const CSpan* pieces, int n, char* out) {
  const Crypto& c = crypto();
  if (!ctx_) return false;
  unsigned char iv[12];
  make_iv(counter_++, iv);
  int outl = 0;
  if (c.enc_init(ctx_, nullptr, nullptr, nullptr, iv) != 1) return false;
  if (aad_len && c.enc_update(ctx_, nullptr, &outl, static_cast<const unsigned char*>(aad), static_cast<int>(aad_len)) != 1) return false;
  auto* o = reinterpret_cast<unsigned char*>(out);
  for (int i = 0; i < n; ++i) {
    const auto* p = static_cast<const unsigned char*>(pieces[i].data);
    size_t left = pieces[i].len;
    while (left) {
      const size_t take = left < kSlice ? left : kSlice;
      if (c.enc_update(ctx_, o, &outl, p, static_cast<int>(take)) != 1) return false;
      o += outl, p += take, left -= take;
    }
  }
  if (c.enc_final(ctx_, o, &outl) != 1) return false;
  return c.ctx_ctrl(ctx_, kCtrlGetTag, static_cast<int>(kAeadTag), o) == 1;
}

bool Aead::open(const void* aad, size_t aad_len, const Span* pieces, int n, const char* tag) {
  const Crypto& c = crypto();
  if (!ctx_) return false;
  unsigned char iv[12];
  make_iv(counter_++, iv);
  int outl = 0;
  if (c.dec_init(ctx_, nullptr, nullptr, nullptr, iv) != 1) return false;
  if (aad_len && c.dec_update(ctx_, nullptr, &outl, static_cast<const unsigned char*>(aad), static_cast<int>(aad_len)) != 1) return false;
  for (int i = 0; i < n; ++i) {
    auto* p = static_cast<unsigned char*>(pieces[i].data);
    size_t left = pieces[i].len;
    while (left) {
      const size_t take = left < kSlice ? left : kSlice;
      if (c.dec_update(ctx_, p, &outl, p, static_cast<int>(take)) != 1) return false;  // in place
      p += take, left -= take;
    }
  }
  unsigned char t[kAeadTag];
  std::memcpy(t, tag, kAeadTag);
  if (c.ctx_ctrl(ctx_, kCtrlSetTag, static_cast<int>(kAeadTag), t) != 1) return false;
  unsigned char dummy[16];
  return c.dec_final(ctx_, dummy, &outl) == 1;  // the tag comparison
}

bool OffsetCipher::set_key(const uint8_t key[kAeadKey], const uint8_t nonce[8]) {
  const Crypto& c = crypto();
  if (!c.ok || !c.aes_256_ctr) return false;
  std::memcpy(key_, key, kAeadKey);
  std::memcpy(nonce_, nonce, 8);
  ready_ = true;
  return true;
}

bool OffsetCipher::crypt(uint64_t offset, const void* in, void* out, size_t n) const {
  if (!ready_) return false;
  const Crypto& c = crypto();
  void* ctx = c.ctx_new();
  if (!ctx) return false;
  // counter block = nonce(8) || big-endian index of the 16-byte block the range starts in
  unsigned char iv[16];
  std::memcpy(iv, nonce_, 8);
  const uint64_t block = offset / 16;
  for (int i = 0; i < 8; ++i) iv[8 + i] = static_cast<unsigned char>(block >> (56 - 8 * i));
  bool ok = c.enc_init(ctx, c.aes_256_ctr(), nullptr, key_, iv) == 1;
  int outl = 0;
  if (ok && (offset % 16)) {  // start inside a block: burn the key stream bytes before the range
    unsigned char skip[16] = {0}, sink[16];
    ok = c.enc_update(ctx, sink, &outl, skip, static_cast<int>(offset % 16)) == 1;
  }
  const auto* p = static_cast<const unsigned char*>(in);
  auto* o = static_cast<unsigned char*>(out);
  while (ok && n) {
    const size_t take = n < kSlice ? n : kSlice;
    ok = c.enc_update(ctx, o, &outl, p, static_cast<int>(take)) == 1;
    p += take, o += take, n -= take;
  }
  c.ctx_free(ctx);
  return ok;
}

void derive_key(const std::string& token, const char* label, const std::string& nonces, uint8_t out[kAeadKey]) {
  std::string msg(label);
  msg.append(nonces);
  const Sha256Digest d = hmac_sha256(token, msg);
  std::memcpy(out, d.data(), kAeadKey);
}

}  // namespace bb::net

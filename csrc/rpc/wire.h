// Binary wire codec (replaces yalantinglibs struct_pack, which reflected the structs in the
// reference's types.h:217-392).  Little-endian, length-prefixed; LocationDetail variants carry
// a tag byte; Result<T> carries an ok byte + either T or the error code.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common/types.h"

namespace bb::wire {

class Writer {
 public:
  void u8(uint8_t v) { buf_.push_back(static_cast<char>(v)); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void i64(int64_t v) { raw(&v, 8); }
  void f64(double v) { raw(&v, 8); }
  void boolean(bool v) { u8(v ? 1 : 0); }
  void str(const std::string& s) {
    u32(static_cast<uint32_t>(s.size()));
    buf_.append(s);
  }
  void bytes(const std::vector<uint8_t>& b) {
    u32(static_cast<uint32_t>(b.size()));
    buf_.append(reinterpret_cast<const char*>(b.data()), b.size());
  }
  void raw(const void* p, size_t n) { buf_.append(static_cast<const char*>(p), n); }
  void ec(ErrorCode e) { u32(static_cast<uint32_t>(e)); }
  std::string take() { return std::move(buf_); }
  const std::string& data() const { return buf_; }

 private:
  std::string buf_;
};

class Reader {
 public:
  explicit Reader(const std::string& s) : p_(s.data()), end_(s.data() + s.size()) {}
  Reader(const char* p, size_t n) : p_(p), end_(p + n) {}
  bool ok() const { return ok_; }
  void fail() { ok_ = false; }
  bool at_end() const { return p_ == end_; }
  uint8_t u8() { uint8_t v = 0; raw(&v, 1); return v; }
  uint32_t u32() { uint32_t v = 0; raw(&v, 4); return v; }
  uint64_t u64() { uint64_t v = 0; raw(&v, 8); return v; }
  int64_t i64() { int64_t v = 0; raw(&v, 8); return v; }
  double f64() { double v = 0; raw(&v, 8); return v; }
  bool boolean() { return u8() != 0; }
  ErrorCode ec() { return static_cast<ErrorCode>(u32()); }
  std::string str() {
    const uint32_t n = u32();
    if (!ok_ || static_cast<size_t>(end_ - p_) < n) { ok_ = false; return {}; }
    std::string s(p_, n);
    p_ += n;
    return s;
  }
  std::vector<uint8_t> bytes() {
    const uint32_t n = u32();
    if (!ok_ || static_cast<size_t>(end_ - p_) < n) { ok_ = false; return {}; }
    std::vector<uint8_t> b(reinterpret_cast<const uint8_t*>(p_), reinterpret_cast<const uint8_t*>(p_) + n);
    p_ += n;
    return b;
  }
  void raw(void* out, size_t n) {
    if (!ok_ || static_cast<size_t>(end_ - p_) < n) { ok_ = false; std::memset(out, 0, n); return; }
    std::memcpy(out, p_, n);
    p_ += n;
  }
  // Guards vector pre-allocation against corrupt counts.
  uint32_t count(size_t min_elem_bytes = 1) {
    const uint32_t n = u32();
    if (!ok_ || static_cast<size_t>(end_ - p_) / (min_elem_bytes ? min_elem_bytes : 1) < n) { ok_ = false; return 0; }
    return n;
  }

 private:
  const char* p_;
  const char* end_;
  bool ok_ = true;
};

void put(Writer& w, const TransportEndpoint& e);
void get(Reader& r, TransportEndpoint& e);
void put(Writer& w, const LocationDetail& l);
void get(Reader& r, LocationDetail& l);
void put(Writer& w, const ShardPlacement& s);
void get(Reader& r, ShardPlacement& s);
void put(Writer& w, const CopyPlacement& c);
void get(Reader& r, CopyPlacement& c);
void put(Writer& w, const std::vector<CopyPlacement>& v);
void get(Reader& r, std::vector<CopyPlacement>& v);
void put(Writer& w, const WorkerConfig& c);
void get(Reader& r, WorkerConfig& c);
// Batch replies (batch_put_start / batch_get_workers) mostly carry thousands of one-shard placements on the same few pools
// that differ only in where the shard sits and in its digest.  A result that matches the last fully written one that way is
// sent as (position, checksum) and rebuilt from that one by the reader.
class PlacementBatchWriter {
 public:
  explicit PlacementBatchWriter(Writer& w) : w_(w) {}
  void put(const Result<std::vector<CopyPlacement>>& res);

 private:
  Writer& w_;
  const ShardPlacement* base_ = nullptr;  // shard of the last result written in full
};
class PlacementBatchReader {
 public:
  explicit PlacementBatchReader(Reader& r) : r_(r) {}
  Result<std::vector<CopyPlacement>> get();

 private:
  Reader& r_;
  std::vector<CopyPlacement> base_;
  bool have_base_ = false;
};
void put(Writer& w, const ClusterStats& s);
void get(Reader& r, ClusterStats& s);
void put(Writer& w, const MemoryPool& p);
void get(Reader& r, MemoryPool& p);

}  // namespace bb::wire

// Result<T>: value-or-ErrorCode.  Parity: reference types.h:30-49 (`Result<T> = variant<T, ErrorCode>`
// + is_ok/get_value/get_error free functions, which are kept).
#pragma once
#include <cassert>
#include <optional>
#include <utility>
#include <variant>

#include "common/error.h"

namespace bb {

template <typename T>
class Result {
 public:
  Result() : v_(ErrorCode::INTERNAL_ERROR) {}
  Result(T value) : v_(std::move(value)) {}        // NOLINT(google-explicit-constructor)
  Result(ErrorCode e) : v_(e) { assert(e != ErrorCode::OK || true); }  // NOLINT
  bool ok() const noexcept { return std::holds_alternative<T>(v_); }
  explicit operator bool() const noexcept { return ok(); }
  const T& value() const& { return std::get<T>(v_); }
  T& value() & { return std::get<T>(v_); }
  T&& value() && { return std::get<T>(std::move(v_)); }
  ErrorCode error() const noexcept { return ok() ? ErrorCode::OK : std::get<ErrorCode>(v_); }
  const T& operator*() const& { return value(); }
  T& operator*() & { return value(); }
  const T* operator->() const { return &value(); }
  T* operator->() { return &value(); }
  T value_or(T alt) const { return ok() ? value() : std::move(alt); }

 private:
  std::variant<T, ErrorCode> v_;
};

template <typename T>
bool is_ok(const Result<T>& r) { return r.ok(); }
template <typename T>
const T& get_value(const Result<T>& r) { return r.value(); }
template <typename T>
ErrorCode get_error(const Result<T>& r) { return r.error(); }

#define BB_TRY(expr)                                   \
  do {                                                 \
    ::bb::ErrorCode _bb_ec = (expr);                   \
    if (_bb_ec != ::bb::ErrorCode::OK) return _bb_ec;  \
  } while (0)

}  // namespace bb

// json-lite: a small JSON DOM (the reference uses nlohmann/json, unavailable offline).
// Used for the coordination-store values (worker / pool registration, reference schema
// worker_service.cpp:479-516), /stats, and as the target of the YAML-subset parser.
#pragma once
#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <string_view>
#include <variant>
#include <vector>

namespace bb {

class Json {
 public:
  using Array = std::vector<Json>;
  using Object = std::map<std::string, Json>;
  enum class Type { Null, Bool, Int, Double, String, Array, Object };

  Json() : v_(nullptr) {}
  Json(std::nullptr_t) : v_(nullptr) {}
  Json(bool b) : v_(b) {}
  Json(int i) : v_(static_cast<int64_t>(i)) {}
  Json(unsigned i) : v_(static_cast<int64_t>(i)) {}
  Json(long i) : v_(static_cast<int64_t>(i)) {}
  Json(long long i) : v_(static_cast<int64_t>(i)) {}
  Json(unsigned long i) : v_(static_cast<int64_t>(i)) {}
  Json(unsigned long long i) : v_(static_cast<int64_t>(i)) {}
  Json(double d) : v_(d) {}
  Json(const char* s) : v_(std::string(s)) {}
  Json(std::string s) : v_(std::move(s)) {}
  Json(std::string_view s) : v_(std::string(s)) {}
  Json(Array a) : v_(std::move(a)) {}
  Json(Object o) : v_(std::move(o)) {}

  static Json object() { return Json(Object{}); }
  static Json array() { return Json(Array{}); }

  Type type() const noexcept { return static_cast<Type>(v_.index()); }
  bool is_null() const noexcept { return type() == Type::Null; }
  bool is_bool() const noexcept { return type() == Type::Bool; }
  bool is_int() const noexcept { return type() == Type::Int; }
  bool is_number() const noexcept { return type() == Type::Int || type() == Type::Double; }
  bool is_string() const noexcept { return type() == Type::String; }
  bool is_array() const noexcept { return type() == Type::Array; }
  bool is_object() const noexcept { return type() == Type::Object; }

  bool as_bool(bool def = false) const;
  int64_t as_int(int64_t def = 0) const;
  uint64_t as_uint(uint64_t def = 0) const { return static_cast<uint64_t>(as_int(static_cast<int64_t>(def))); }
  double as_double(double def = 0.0) const;
  std::string as_string(const std::string& def = "") const;
  const Array& as_array() const;
  const Object& as_object() const;
  Array& mut_array();
  Object& mut_object();

  // object access; operator[] on a null value turns it into an object
  Json& operator[](const std::string& key);
  const Json& at(const std::string& key) const;  // returns a static null when missing
  bool contains(const std::string& key) const;
  void push_back(Json v);
  size_t size() const;

  std::string dump(int indent = -1) const;
  // Returns nullopt on malformed input; *err (if given) receives a message with offset.
  static std::optional<Json> parse(std::string_view text, std::string* err = nullptr);

  bool operator==(const Json& o) const { return v_ == o.v_; }

 private:
  void dump_to(std::string& out, int indent, int depth) const;
  std::variant<std::nullptr_t, bool, int64_t, double, std::string, Array, Object> v_;
};

std::string json_escape(std::string_view s);

}  // namespace bb

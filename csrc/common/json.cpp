#include "common/json.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace bb {

namespace {
const Json kNull;
const Json::Array kEmptyArray;
const Json::Object kEmptyObject;

struct Parser {
  std::string_view s;
  size_t i = 0;
  std::string err;
  int depth = 0;

  bool fail(const char* m) {
    if (err.empty()) err = std::string(m) + " at offset " + std::to_string(i);
    return false;
  }
  void ws() {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
  }
  bool lit(const char* w) {
    size_t n = std::strlen(w);
    if (s.substr(i, n) == w) {
      i += n;
      return true;
    }
    return false;
  }
  static void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += static_cast<char>(cp);
    else if (cp < 0x800) {
      out += static_cast<char>(0xC0 | (cp >> 6));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      out += static_cast<char>(0xE0 | (cp >> 12));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else {
      out += static_cast<char>(0xF0 | (cp >> 18));
      out += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    }
  }
  bool hex4(uint32_t& v) {
    if (i + 4 > s.size()) return fail("short \\u escape");
    v = 0;
    for (int k = 0; k < 4; ++k) {
      char c = s[i++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return fail("bad \\u escape");
    }
    return true;
  }
  bool str(std::string& out) {
    if (i >= s.size() || s[i] != '"') return fail("expected string");
    ++i;
    while (i < s.size()) {
      char c = s[i++];
      if (c == '"') return true;
      if (c == '\\') {
        if (i >= s.size()) return fail("bad escape");
        char e = s[i++];
        switch (e) {
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          case '/': out += '/'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'n': out += '\n'; break;
          case 'r': out += '\r'; break;
          case 't': out += '\t'; break;
          case 'u': {
            uint32_t cp = 0;
            if (!hex4(cp)) return false;
            if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {
              const size_t save = i;
              i += 2;
              uint32_t lo = 0;
              if (!hex4(lo)) return false;
              if (lo >= 0xDC00 && lo <= 0xDFFF) {
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              } else {
                i = save;  // not a pair: the high surrogate stands alone, the next escape is parsed on its own
              }
            }
            if (cp >= 0xD800 && cp <= 0xDFFF) cp = 0xFFFD;  // lone surrogate -> replacement character
            put_utf8(out, cp);
            break;
          }
          default: return fail("bad escape");
        }
      } else {
        out += c;
      }
    }
    return fail("unterminated string");
  }
  bool value(Json& out) {
    if (++depth > 256) return fail("nesting too deep");
    ws();
    if (i >= s.size()) return fail("unexpected end");
    char c = s[i];
    bool ok = true;
    if (c == '{') {
      ++i;
      Json::Object o;
      ws();
      if (i < s.size() && s[i] == '}') {
        ++i;
      } else {
        while (true) {
          ws();
          std::string k;
          if (!str(k)) { ok = false; break; }
          ws();
          if (i >= s.size() || s[i] != ':') { ok = fail("expected ':'"); break; }
          ++i;
          Json v;
          if (!value(v)) { ok = false; break; }
          o[std::move(k)] = std::move(v);
          ws();
          if (i < s.size() && s[i] == ',') { ++i; continue; }
          if (i < s.size() && s[i] == '}') { ++i; break; }
          ok = fail("expected ',' or '}'");
          break;
        }
      }
      if (ok) out = Json(std::move(o));
    } else if (c == '[') {
      ++i;
      Json::Array a;
      ws();
      if (i < s.size() && s[i] == ']') {
        ++i;
      } else {
        while (true) {
          Json v;
          if (!value(v)) { ok = false; break; }
          a.push_back(std::move(v));
          ws();
          if (i < s.size() && s[i] == ',') { ++i; continue; }
          if (i < s.size() && s[i] == ']') { ++i; break; }
          ok = fail("expected ',' or ']'");
          break;
        }
      }
      if (ok) out = Json(std::move(a));
    } else if (c == '"') {
      std::string v;
      ok = str(v);
      if (ok) out = Json(std::move(v));
    } else if (lit("true")) {
      out = Json(true);
    } else if (lit("false")) {
      out = Json(false);
    } else if (lit("null")) {
      out = Json(nullptr);
    } else {
      size_t st = i;
      bool is_float = false;
      if (i < s.size() && (s[i] == '-' || s[i] == '+')) ++i;
      while (i < s.size() && ((s[i] >= '0' && s[i] <= '9') || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || s[i] == '-' || s[i] == '+')) {
        if (s[i] == '.' || s[i] == 'e' || s[i] == 'E') is_float = true;
        ++i;
      }
      if (i == st) ok = fail("unexpected character");
      else {
        std::string num(s.substr(st, i - st));
        if (!is_float) {
          errno = 0;
          char* end = nullptr;
          long long v = std::strtoll(num.c_str(), &end, 10);
          if (errno == ERANGE) {
            // preserve large unsigned values (addresses) bit-exactly
            errno = 0;
            unsigned long long u = std::strtoull(num.c_str(), &end, 10);
            if (errno == ERANGE || *end) ok = fail("bad number");
            else out = Json(static_cast<int64_t>(u));
          } else if (*end) ok = fail("bad number");
          else out = Json(static_cast<int64_t>(v));
        } else {
          char* end = nullptr;
          double d = std::strtod(num.c_str(), &end);
          if (*end) ok = fail("bad number");
          else out = Json(d);
        }
      }
    }
    --depth;
    return ok;
  }
};
}  // namespace

bool Json::as_bool(bool def) const {
  if (auto p = std::get_if<bool>(&v_)) return *p;
  if (auto p = std::get_if<int64_t>(&v_)) return *p != 0;
  if (auto p = std::get_if<std::string>(&v_)) {
    if (*p == "true" || *p == "yes" || *p == "on" || *p == "1") return true;
    if (*p == "false" || *p == "no" || *p == "off" || *p == "0") return false;
  }
  return def;
}
int64_t Json::as_int(int64_t def) const {
  if (auto p = std::get_if<int64_t>(&v_)) return *p;
  if (auto p = std::get_if<double>(&v_)) return static_cast<int64_t>(*p);
  if (auto p = std::get_if<bool>(&v_)) return *p ? 1 : 0;
  if (auto p = std::get_if<std::string>(&v_)) {
    errno = 0;
    char* end = nullptr;
    long long v = std::strtoll(p->c_str(), &end, 0);
    if (end != p->c_str() && *end == 0 && errno == 0) return v;
    if (errno == ERANGE) {
      errno = 0;
      unsigned long long u = std::strtoull(p->c_str(), &end, 0);
      if (errno == 0 && *end == 0) return static_cast<int64_t>(u);
    }
  }
  return def;
}
double Json::as_double(double def) const {
  if (auto p = std::get_if<double>(&v_)) return *p;
  if (auto p = std::get_if<int64_t>(&v_)) return static_cast<double>(*p);
  if (auto p = std::get_if<std::string>(&v_)) {
    char* end = nullptr;
    double d = std::strtod(p->c_str(), &end);
    if (end != p->c_str() && *end == 0) return d;
  }
  return def;
}
std::string Json::as_string(const std::string& def) const {
  if (auto p = std::get_if<std::string>(&v_)) return *p;
  if (auto p = std::get_if<int64_t>(&v_)) return std::to_string(*p);
  if (auto p = std::get_if<bool>(&v_)) return *p ? "true" : "false";
  if (auto p = std::get_if<double>(&v_)) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%g", *p);
    return buf;
  }
  return def;
}
const Json::Array& Json::as_array() const {
  if (auto p = std::get_if<Array>(&v_)) return *p;
  return kEmptyArray;
}
const Json::Object& Json::as_object() const {
  if (auto p = std::get_if<Object>(&v_)) return *p;
  return kEmptyObject;
}
Json::Array& Json::mut_array() {
  if (!is_array()) v_ = Array{};
  return std::get<Array>(v_);
}
Json::Object& Json::mut_object() {
  if (!is_object()) v_ = Object{};
  return std::get<Object>(v_);
}
Json& Json::operator[](const std::string& key) { return mut_object()[key]; }
const Json& Json::at(const std::string& key) const {
  if (auto p = std::get_if<Object>(&v_)) {
    auto it = p->find(key);
    if (it != p->end()) return it->second;
  }
  return kNull;
}
bool Json::contains(const std::string& key) const {
  if (auto p = std::get_if<Object>(&v_)) return p->count(key) > 0;
  return false;
}
void Json::push_back(Json v) { mut_array().push_back(std::move(v)); }
size_t Json::size() const {
  if (auto p = std::get_if<Array>(&v_)) return p->size();
  if (auto p = std::get_if<Object>(&v_)) return p->size();
  return 0;
}

std::string json_escape(std::string_view s) {
  std::string out;
  out.reserve(s.size() + 2);
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          std::snprintf(buf, sizeof buf, "\\u%04x", c);
          out += buf;
        } else {
          out += static_cast<char>(c);
        }
    }
  }
  return out;
}

void Json::dump_to(std::string& out, int indent, int depth) const {
  auto nl = [&](int d) {
    if (indent >= 0) {
      out += '\n';
      out.append(static_cast<size_t>(indent * d), ' ');
    }
  };
  switch (type()) {
    case Type::Null: out += "null"; break;
    case Type::Bool: out += std::get<bool>(v_) ? "true" : "false"; break;
    case Type::Int: out += std::to_string(std::get<int64_t>(v_)); break;
    case Type::Double: {
      double d = std::get<double>(v_);
      if (!std::isfinite(d)) {
        out += "null";
      } else {
        char buf[40];
        std::snprintf(buf, sizeof buf, "%.17g", d);
        out += buf;
        if (!std::strpbrk(buf, ".eE")) out += ".0";
      }
      break;
    }
    case Type::String:
      out += '"';
      out += json_escape(std::get<std::string>(v_));
      out += '"';
      break;
    case Type::Array: {
      const auto& a = std::get<Array>(v_);
      out += '[';
      bool first = true;
      for (const auto& e : a) {
        if (!first) out += ',';
        first = false;
        nl(depth + 1);
        e.dump_to(out, indent, depth + 1);
      }
      if (!a.empty()) nl(depth);
      out += ']';
      break;
    }
    case Type::Object: {
      const auto& o = std::get<Object>(v_);
      out += '{';
      bool first = true;
      for (const auto& [k, e] : o) {
        if (!first) out += ',';
        first = false;
        nl(depth + 1);
        out += '"';
        out += json_escape(k);
        out += "\":";
        if (indent >= 0) out += ' ';
        e.dump_to(out, indent, depth + 1);
      }
      if (!o.empty()) nl(depth);
      out += '}';
      break;
    }
  }
}

std::string Json::dump(int indent) const {
  std::string out;
  dump_to(out, indent, 0);
  return out;
}

std::optional<Json> Json::parse(std::string_view text, std::string* err) {
  Parser p{text};
  Json v;
  if (!p.value(v)) {
    if (err) *err = p.err;
    return std::nullopt;
  }
  p.ws();
  if (p.i != text.size()) {
    if (err) *err = "trailing characters at offset " + std::to_string(p.i);
    return std::nullopt;
  }
  return v;
}

}  // namespace bb

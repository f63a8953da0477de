#include "common/yaml.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <vector>

namespace bb {
namespace {

struct Line {
  int indent;
  std::string text;  // content without indent / trailing comment
  int lineno;
};

std::string rstrip(std::string s) {
  while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
  return s;
}
std::string strip(const std::string& s) {
  size_t a = 0;
  while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) ++a;
  return rstrip(s.substr(a));
}

// Removes a trailing " # comment" that is outside quotes.
std::string strip_comment(const std::string& s) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq && (i == 0 || s[i - 1] != '\\')) dq = !dq;
    else if (c == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return rstrip(s.substr(0, i));
  }
  return rstrip(s);
}

Json scalar(const std::string& raw) {
  std::string s = strip(raw);
  if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") return Json(nullptr);
  if (s.size() >= 2 && s.front() == '"' && s.back() == '"') {
    std::string err;
    auto j = Json::parse(s, &err);
    if (j) return *j;
    return Json(s.substr(1, s.size() - 2));
  }
  if (s.size() >= 2 && s.front() == '\'' && s.back() == '\'') {
    std::string out;
    for (size_t i = 1; i + 1 < s.size(); ++i) {
      if (s[i] == '\'' && i + 2 < s.size() && s[i + 1] == '\'') {
        out += '\'';
        ++i;
      } else {
        out += s[i];
      }
    }
    return Json(out);
  }
  if (s == "true" || s == "True" || s == "TRUE") return Json(true);
  if (s == "false" || s == "False" || s == "FALSE") return Json(false);
  // integer?
  {
    char* end = nullptr;
    errno = 0;
    long long v = std::strtoll(s.c_str(), &end, 0);
    if (end != s.c_str() && *end == 0 && errno == 0 && (std::isdigit(static_cast<unsigned char>(s[0])) || s[0] == '-' || s[0] == '+'))
      return Json(static_cast<int64_t>(v));
    if (errno == ERANGE) {
      errno = 0;
      unsigned long long u = std::strtoull(s.c_str(), &end, 0);
      if (errno == 0 && *end == 0) return Json(static_cast<int64_t>(u));
    }
  }
  {
    char* end = nullptr;
    double d = std::strtod(s.c_str(), &end);
    if (end != s.c_str() && *end == 0 && (std::isdigit(static_cast<unsigned char>(s[0])) || s[0] == '-' || s[0] == '+' || s[0] == '.'))
      return Json(d);
  }
  return Json(s);
}

// Splits "a, b, [c, d]" on top-level commas.
std::vector<std::string> split_flow(const std::string& s) {
  std::vector<std::string> out;
  int depth = 0;
  bool sq = false, dq = false;
  std::string cur;
  for (char c : s) {
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    if (!sq && !dq) {
      if (c == '[' || c == '{') ++depth;
      else if (c == ']' || c == '}') --depth;
      else if (c == ',' && depth == 0) {
        out.push_back(cur);
        cur.clear();
        continue;
      }
    }
    cur += c;
  }
  if (!strip(cur).empty()) out.push_back(cur);
  return out;
}

// Finds the first ':' that terminates a mapping key (followed by space/end, outside quotes).
size_t find_key_colon(const std::string& s) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    else if (c == ':' && !sq && !dq && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
  }
  return std::string::npos;
}

// Flow collections nest at most this deep; anything deeper is kept as a plain string (the parser reads operator
// configs and registry records, and must not be driven into a stack overflow by a hostile or corrupt document).
constexpr int kMaxFlowDepth = 32;
constexpr int kMaxBlockDepth = 128;

Json flow_value(const std::string& raw, int depth = 0);

Json flow_value(const std::string& raw, int depth) {
  std::string s = strip(raw);
  if (depth >= kMaxFlowDepth) return Json(s);
  if (s.size() >= 2 && s.front() == '[' && s.back() == ']') {
    Json::Array a;
    for (auto& part : split_flow(s.substr(1, s.size() - 2))) a.push_back(flow_value(part, depth + 1));
    return Json(std::move(a));
  }
  if (s.size() >= 2 && s.front() == '{' && s.back() == '}') {
    Json::Object o;
    for (auto& part : split_flow(s.substr(1, s.size() - 2))) {
      size_t c = find_key_colon(part);
      if (c == std::string::npos) {
        c = part.find(':');
        if (c == std::string::npos) continue;
      }
      o[scalar(part.substr(0, c)).as_string()] = flow_value(part.substr(c + 1), depth + 1);
    }
    return Json(std::move(o));
  }
  return scalar(s);
}

struct YParser {
  std::vector<Line> lines;
  size_t pos = 0;
  std::string err;

  bool fail(const std::string& m, int lineno) {
    if (err.empty()) err = m + " (line " + std::to_string(lineno) + ")";
    return false;
  }

  int depth = 0;

  bool parse_block(int indent, Json& out) {
    if (pos >= lines.size()) {
      out = Json(nullptr);
      return true;
    }
    const Line& first = lines[pos];
    if (depth >= kMaxBlockDepth) return fail("nesting too deep", first.lineno);
    ++depth;
    const bool ok = (first.text.rfind("- ", 0) == 0 || first.text == "-") ? parse_seq(first.indent, out) : parse_map(first.indent, out);
    --depth;
    return ok;
  }

  bool parse_value_after_key(const std::string& rest, int key_indent, Json& out) {
    std::string r = strip(rest);
    if (!r.empty()) {
      out = flow_value(r);
      return true;
    }
    // nested block (deeper indent) or a sequence at the same indent ("key:\n- a")
    if (pos < lines.size()) {
      const Line& nx = lines[pos];
      if (nx.indent > key_indent) return parse_block(nx.indent, out);
      if (nx.indent == key_indent && (nx.text.rfind("- ", 0) == 0 || nx.text == "-")) return parse_seq(nx.indent, out);
    }
    out = Json(nullptr);
    return true;
  }

  bool parse_map(int indent, Json& out) {
    Json::Object o;
    while (pos < lines.size()) {
      const Line& ln = lines[pos];
      if (ln.indent < indent) break;
      if (ln.indent > indent) return fail("unexpected indentation", ln.lineno);
      if (ln.text.rfind("- ", 0) == 0 || ln.text == "-") break;
      size_t c = find_key_colon(ln.text);
      if (c == std::string::npos) return fail("expected 'key: value'", ln.lineno);
      std::string key = scalar(ln.text.substr(0, c)).as_string();
      std::string rest = ln.text.substr(c + 1);
      ++pos;
      Json v;
      if (!parse_value_after_key(rest, indent, v)) return false;
      o[key] = std::move(v);
    }
    out = Json(std::move(o));
    return true;
  }

  bool parse_seq(int indent, Json& out) {
    Json::Array a;
    while (pos < lines.size()) {
      Line& ln = lines[pos];
      if (ln.indent < indent) break;
      if (ln.indent > indent) return fail("unexpected indentation in sequence", ln.lineno);
      if (!(ln.text.rfind("- ", 0) == 0 || ln.text == "-")) break;
      std::string item = ln.text.size() > 1 ? ln.text.substr(2) : "";
      size_t lead = 0;
      while (lead < item.size() && item[lead] == ' ') ++lead;
      item = item.substr(lead);
      int item_indent = indent + 2 + static_cast<int>(lead);
      if (item.empty()) {
        ++pos;
        Json v;
        if (pos < lines.size() && lines[pos].indent > indent) {
          if (!parse_block(lines[pos].indent, v)) return false;
        }
        a.push_back(std::move(v));
        continue;
      }
      bool flow = item.front() == '[' || item.front() == '{' || item.front() == '"' || item.front() == '\'';
      size_t c = flow ? std::string::npos : find_key_colon(item);
      if (c == std::string::npos) {
        a.push_back(flow_value(item));
        ++pos;
        continue;
      }
      // "- key: value" starts an inline mapping; rewrite this line as a map line at item_indent.
      ln.indent = item_indent;
      ln.text = item;
      Json v;
      if (!parse_map(item_indent, v)) return false;
      a.push_back(std::move(v));
    }
    out = Json(std::move(a));
    return true;
  }
};

}  // namespace

std::optional<Json> parse_yaml(std::string_view text, std::string* err) {
  YParser p;
  std::istringstream in{std::string(text)};
  std::string raw;
  int lineno = 0;
  bool seen_doc = false;
  while (std::getline(in, raw)) {
    ++lineno;
    for (char c : raw)
      if (c == '\t' && raw.find_first_not_of(" \t") != std::string::npos && raw.find('\t') < raw.find_first_not_of(" \t")) {
        if (err) *err = "tab indentation is not supported (line " + std::to_string(lineno) + ")";
        return std::nullopt;
      }
    std::string s = strip_comment(raw);
    if (strip(s).empty()) continue;
    if (strip(s) == "---") {
      if (seen_doc) break;
      continue;
    }
    if (strip(s) == "...") break;
    seen_doc = true;
    size_t ind = s.find_first_not_of(' ');
    p.lines.push_back({static_cast<int>(ind), s.substr(ind), lineno});
  }
  Json out = Json::object();
  if (p.lines.empty()) return out;
  if (!p.parse_block(p.lines[0].indent, out) || p.pos != p.lines.size()) {
    if (err) *err = p.err.empty() ? "unparsed trailing content (line " + std::to_string(p.lines[std::min(p.pos, p.lines.size() - 1)].lineno) + ")" : p.err;
    return std::nullopt;
  }
  return out;
}

std::optional<Json> load_yaml_file(const std::string& path, std::string* err) {
  std::ifstream f(path);
  if (!f) {
    if (err) *err = "cannot open " + path;
    return std::nullopt;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return parse_yaml(ss.str(), err);
}

std::optional<uint64_t> parse_size(std::string_view sv) {
  std::string s;
  for (char c : sv)
    if (c != '_' && c != ' ') s += static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  if (s.empty()) return std::nullopt;
  if (s == "unlimited" || s == "max") return std::numeric_limits<uint64_t>::max();
  size_t i = 0;
  while (i < s.size() && (std::isdigit(static_cast<unsigned char>(s[i])) || s[i] == '.')) ++i;
  if (i == 0) return std::nullopt;
  double num = std::strtod(s.substr(0, i).c_str(), nullptr);
  std::string unit = s.substr(i);
  uint64_t mul = 1;
  if (unit.empty() || unit == "b") mul = 1;
  else if (unit == "k" || unit == "kb" || unit == "kib") mul = 1ull << 10;
  else if (unit == "m" || unit == "mb" || unit == "mib") mul = 1ull << 20;
  else if (unit == "g" || unit == "gb" || unit == "gib") mul = 1ull << 30;
  else if (unit == "t" || unit == "tb" || unit == "tib") mul = 1ull << 40;
  else return std::nullopt;
  return static_cast<uint64_t>(num * static_cast<double>(mul));
}

}  // namespace bb

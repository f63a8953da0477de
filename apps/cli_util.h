// Shared argv helpers for the bb-* executables ("--key value" and "--key=value").
#pragma once
#include <csignal>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace bbapp {

struct Args {
  std::map<std::string, std::string> kv;
  std::vector<std::string> positional;
  bool has(const std::string& k) const { return kv.count(k) > 0; }
  std::string get(const std::string& k, const std::string& def = "") const {
    auto it = kv.find(k);
    return it == kv.end() ? def : it->second;
  }
  long long num(const std::string& k, long long def) const {
    auto it = kv.find(k);
    return it == kv.end() ? def : std::atoll(it->second.c_str());
  }
};

inline Args parse_args(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    if (s.rfind("--", 0) == 0) {
      s = s.substr(2);
      const size_t eq = s.find('=');
      if (eq != std::string::npos) {
        a.kv[s.substr(0, eq)] = s.substr(eq + 1);
      } else if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) {
        a.kv[s] = argv[++i];
      } else {
        a.kv[s] = "true";
      }
    } else {
      a.positional.push_back(s);
    }
  }
  return a;
}

inline volatile std::sig_atomic_t g_stop = 0;
inline void on_signal(int) { g_stop = 1; }
inline void install_signal_handlers() {
  std::signal(SIGINT, on_signal);
  std::signal(SIGTERM, on_signal);
  std::signal(SIGPIPE, SIG_IGN);
}

}  // namespace bbapp

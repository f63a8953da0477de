// bb-coord: the coordination daemon (the role `etcd` plays in the reference's
// scripts/start_cluster.sh:119-147).  Serves a MemCoord over the framed RPC protocol.
//   bb-coord --listen 127.0.0.1:2379
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <thread>

#include "apps/cli_util.h"
#include "net/tcp.h"
#include "common/log.h"
#include "coord/coord.h"

int main(int argc, char** argv) {
  auto args = bbapp::parse_args(argc, argv);
  if (args.has("auth-token")) bb::net::set_cluster_token(args.get("auth-token"));  // else BB_AUTH_TOKEN / config
  if (args.has("encrypt-transport")) bb::net::set_transport_encryption(true);  // else BB_ENCRYPT_TRANSPORT / config
  if (args.has("auth-token-ro")) bb::net::set_cluster_token_ro(args.get("auth-token-ro"));  // else BB_AUTH_TOKEN_RO / config
  if (args.has("help")) {
    std::printf("usage: bb-coord [--listen host:port] [--data-dir DIR [--no-fsync] [--snapshot-mb N]] [--log-level info]\n"
                "  --data-dir   persist keys, revisions and leases (append-only log + snapshots, fdatasync group commit);\n"
                "               a restarted daemon resumes where it stopped, leases re-armed with their full TTL\n");
    return 0;
  }
  if (args.has("log-level")) setenv("BB_LOG_LEVEL", args.get("log-level").c_str(), 1);
  bb::set_log_level(bb::LogLevel::INFO);
  auto hp = bb::split_host_port(args.get("listen", "127.0.0.1:2379"));
  if (!hp) {
    std::fprintf(stderr, "bb-coord: bad --listen\n");
    return 2;
  }
  bbapp::install_signal_handlers();
  auto store = std::make_shared<bb::coord::MemCoord>();
  if (args.has("data-dir")) {
    const uint64_t snap = static_cast<uint64_t>(std::max(1, std::atoi(args.get("snapshot-mb", "64").c_str()))) << 20;
    if (store->open_durable(args.get("data-dir"), !args.has("no-fsync"), snap) != bb::ErrorCode::OK) {
      std::fprintf(stderr, "bb-coord: cannot open data dir %s\n", args.get("data-dir").c_str());
      return 1;
    }
  }
  bb::coord::CoordServer srv(store);
  if (srv.start(hp->first, static_cast<uint16_t>(hp->second)) != bb::ErrorCode::OK) {
    std::fprintf(stderr, "bb-coord: cannot listen on %s\n", args.get("listen", "127.0.0.1:2379").c_str());
    return 1;
  }
  std::printf("bb-coord listening on %s:%u\n", hp->first.c_str(), srv.port());
  std::fflush(stdout);
  while (!bbapp::g_stop) std::this_thread::sleep_for(std::chrono::milliseconds(200));
  srv.stop();
  return 0;
}

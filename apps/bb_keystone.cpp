// bb-keystone: the control-plane server (reference examples/keystone_example.cpp is the de-facto
// server binary; src/executables/CMakeLists.txt:1-9 only plans `blackbird-keystone`).
//   bb-keystone configs/keystone.yaml [--etcd-endpoints e] [--listen-address a] [--http-port p]
//                                     [--cluster-id c] [--enable-ha] [--service-id s]
#include <chrono>
#include <cstdio>
#include <thread>

#include "apps/cli_util.h"
#include "net/tcp.h"
#include "common/tenant.h"
#include "common/log.h"
#include "rpc/rpc_service.h"

int main(int argc, char** argv) {
  auto args = bbapp::parse_args(argc, argv);
  if (args.has("auth-token")) bb::net::set_cluster_token(args.get("auth-token"));  // else BB_AUTH_TOKEN / config
  if (args.has("encrypt-transport")) bb::net::set_transport_encryption(true);  // else BB_ENCRYPT_TRANSPORT / config
  if (args.has("auth-token-ro")) bb::net::set_cluster_token_ro(args.get("auth-token-ro"));  // else BB_AUTH_TOKEN_RO / config
  if (args.has("http-token")) bb::net::set_http_token(args.get("http-token"));  // else BB_HTTP_TOKEN / config: bearer token of /metrics and /stats
  if (args.has("help")) {
    std::printf("usage: bb-keystone [config.yaml] [--etcd-endpoints E] [--listen-address A] [--http-port P] [--cluster-id C] [--enable-ha] [--service-id S] [--tenants-file F] [--audit-log F]\n");
    return 0;
  }
  bb::set_log_level(bb::LogLevel::INFO);
  bb::KeystoneConfig cfg;
  try {
    if (!args.positional.empty()) cfg = bb::KeystoneConfig::from_yaml(args.positional[0]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "bb-keystone: %s\n", e.what());
    return 2;
  }
  // CLI overrides (reference keystone_example.cpp:76-95) and BB_* environment
  if (const char* e = std::getenv("BB_COORD_ENDPOINTS")) cfg.etcd_endpoints = e;
  if (args.has("etcd-endpoints")) cfg.etcd_endpoints = args.get("etcd-endpoints");
  if (args.has("coord-endpoints")) cfg.etcd_endpoints = args.get("coord-endpoints");
  if (!args.has("auth-token") && !cfg.auth_token.empty()) bb::net::set_cluster_token(cfg.auth_token);  // before the coordination client connects
  if (args.has("audit-log")) cfg.audit_log = args.get("audit-log");  // else `audit_log:` / BB_AUDIT_LOG (common/audit.h)
  if (args.has("tenants-file")) cfg.tenants_file = args.get("tenants-file");  // else `tenants_file:` / BB_TENANTS_FILE (common/tenant.h)
  if (args.has("listen-address")) cfg.listen_address = args.get("listen-address");
  if (args.has("http-port")) cfg.http_metrics_port = args.get("http-port");
  if (args.has("cluster-id")) cfg.cluster_id = args.get("cluster-id");
  if (args.has("service-id")) cfg.service_id = args.get("service-id");
  if (args.has("enable-ha")) cfg.enable_ha = true;
  if (!cfg.log_level.empty()) setenv("BB_LOG_LEVEL", cfg.log_level.c_str(), 0);
  if (!cfg.log_file.empty()) bb::set_log_file(cfg.log_file);
  bbapp::install_signal_handlers();
  auto bundle = bb::rpc::create_and_start_keystone(cfg);
  if (!bundle.ok()) {
    std::fprintf(stderr, "bb-keystone: start failed: %s\n", std::string(bb::to_string(bundle.error())).c_str());
    return 1;
  }
  std::printf("bb-keystone %s cluster=%s rpc=%u http=%u leader=%d\n", bundle.value().keystone->config().service_id.c_str(),
              cfg.cluster_id.c_str(), bundle.value().rpc->rpc_port(), bundle.value().rpc->http_port(), bundle.value().keystone->is_leader());
  std::fflush(stdout);
  int tick = 0;
  while (!bbapp::g_stop) {
    std::this_thread::sleep_for(std::chrono::milliseconds(200));
    if (++tick % 300 == 0) {  // every 60 s (reference keystone_example.cpp:149-169)
      auto st = bundle.value().keystone->get_cluster_stats();
      if (st.ok())
        BB_LOG(INFO) << "cluster: workers=" << st.value().total_workers << " pools=" << st.value().total_memory_pools
                     << " objects=" << st.value().total_objects << " used=" << st.value().used_capacity << "/" << st.value().total_capacity;
    }
  }
  bundle.value().rpc->stop();
  bundle.value().keystone->stop();
  return 0;
}

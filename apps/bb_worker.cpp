// bb-worker: a storage worker process (reference examples/worker_example.cpp; planned
// `blackbird-worker`).  Builds the pools of the config (GPU / DRAM / CXL / NVMe / HDD tiers),
// serves them on the data endpoint, registers with the coordination store or a keystone.
//   bb-worker --config configs/worker.yaml [--worker-id W] [--node-id N] [--coord-endpoints E] [--keystone host:port]
#include <chrono>
#include <cstdio>
#include <thread>

#include "apps/cli_util.h"
#include "net/tcp.h"
#include "common/tenant.h"
#include "common/log.h"
#include "fabric/gpu_fabric.h"
#include "worker/worker_service.h"

int main(int argc, char** argv) {
  auto args = bbapp::parse_args(argc, argv);
  if (args.has("auth-token")) bb::net::set_cluster_token(args.get("auth-token"));  // else BB_AUTH_TOKEN / config
  if (args.has("encrypt-transport")) bb::net::set_transport_encryption(true);  // else BB_ENCRYPT_TRANSPORT / config
  if (args.has("auth-token-ro")) bb::net::set_cluster_token_ro(args.get("auth-token-ro"));  // else BB_AUTH_TOKEN_RO / config
  if (args.has("http-token")) bb::net::set_http_token(args.get("http-token"));  // else BB_HTTP_TOKEN / config: bearer token of /metrics and /stats
  if (args.has("help") || (!args.has("config") && args.positional.empty())) {
    std::printf("usage: bb-worker --config worker.yaml [--worker-id W] [--node-id N] [--coord-endpoints E] [--keystone host:port] [--data-endpoint host:port] [--http-port P] [--tenants-file F] [--audit-log F]\n");
    return args.has("help") ? 0 : 2;
  }
  bb::set_log_level(bb::LogLevel::INFO);
  bb::worker::WorkerServiceConfig cfg;
  try {
    cfg = bb::worker::load_worker_config_from_file(args.has("config") ? args.get("config") : args.positional[0]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "bb-worker: %s\n", e.what());
    return 2;
  }
  if (!args.has("auth-token") && !cfg.auth_token.empty()) bb::net::set_cluster_token(cfg.auth_token);  // before any connection is made
  if (args.has("auth-token")) cfg.auth_token = args.get("auth-token");
  if (args.has("audit-log")) cfg.audit_log = args.get("audit-log");  // else `audit_log:` / BB_AUDIT_LOG (common/audit.h)
  if (args.has("tenants-file")) cfg.tenants_file = args.get("tenants-file");  // else `tenants_file:` / BB_TENANTS_FILE (common/tenant.h)
  if (args.has("worker-id")) cfg.worker_id = args.get("worker-id");
  if (args.has("node-id")) cfg.node_id = args.get("node-id");
  if (const char* e = std::getenv("BB_COORD_ENDPOINTS")) cfg.etcd_endpoints = e;
  if (args.has("coord-endpoints")) cfg.etcd_endpoints = args.get("coord-endpoints");
  if (args.has("etcd-endpoints")) cfg.etcd_endpoints = args.get("etcd-endpoints");
  if (args.has("keystone")) {
    cfg.keystone_address = args.get("keystone");
    cfg.etcd_endpoints.clear();
  }
  if (args.has("data-endpoint")) cfg.ucx_endpoint = args.get("data-endpoint");
  if (args.has("cluster-id")) cfg.cluster_id = args.get("cluster-id");
  if (args.has("http-port")) cfg.http_metrics_port = std::atoi(args.get("http-port").c_str());
  bb::gpu::install_gpu_backend_factory();  // RAM_GPU pools become cudaMalloc slabs exported over CUDA IPC
  bbapp::install_signal_handlers();
  bb::worker::WorkerService svc(cfg);
  bb::ErrorCode ec = svc.create_storage_pools_from_config();
  if (ec == bb::ErrorCode::OK) ec = svc.initialize();
  if (ec == bb::ErrorCode::OK) ec = svc.start();
  if (ec != bb::ErrorCode::OK) {
    std::fprintf(stderr, "bb-worker: start failed: %s\n", std::string(bb::to_string(ec)).c_str());
    return 1;
  }
  std::printf("bb-worker %s node=%s data=%s pools=%zu", cfg.worker_id.c_str(), cfg.node_id.c_str(), svc.data_endpoint().c_str(),
              svc.advertised_pools().size());
  if (cfg.http_metrics_port >= 0) std::printf(" metrics=http://%s:%u/metrics", svc.data_endpoint().substr(0, svc.data_endpoint().rfind(':')).c_str(), svc.http_port());
  std::printf("\n");
  std::fflush(stdout);
  while (!bbapp::g_stop && svc.is_running()) std::this_thread::sleep_for(std::chrono::milliseconds(200));
  svc.stop();
  return 0;
}

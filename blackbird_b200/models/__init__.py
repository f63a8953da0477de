"""Workload "models" of the object store: the five BASELINE.json configurations."""
from .workloads import (feature_store_fanout, latency_sweep, plumbing_cpu, replicated_put_verify, throughput_sweep,  # noqa: F401
                        tier_spill, tier_spill_perf)

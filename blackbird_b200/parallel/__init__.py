from .cluster import CpuRankCluster, GpuRankCluster, LocalCluster  # noqa: F401

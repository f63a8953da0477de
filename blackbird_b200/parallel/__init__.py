from .cluster import GpuRankCluster, LocalCluster  # noqa: F401

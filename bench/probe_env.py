#!/usr/bin/env python3
"""Environment probe for the GPU box: CPU quota, NVLS/multicast support, P2P attributes."""
import os
import torch

print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
print("gpus", torch.cuda.device_count())
try:
    from cuda.bindings import driver as cu
except Exception:  # noqa: BLE001
    from cuda import cuda as cu
cu.cuInit(0)
for d in range(torch.cuda.device_count()):
    err, dev = cu.cuDeviceGet(d)
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED"):
        attr = getattr(cu.CUdevice_attribute, name, None)
        if attr is None:
            print(d, name, "n/a")
            continue
        err, v = cu.cuDeviceGetAttribute(attr, dev)
        print(d, name, v, err)
if torch.cuda.device_count() > 1:
    print("can_access_peer 0->1", torch.cuda.can_device_access_peer(0, 1))
try:
    import torch.distributed._symmetric_memory as sm  # noqa: F401
    print("torch symmetric memory module present")
except Exception as e:  # noqa: BLE001
    print("no torch symm mem", e)

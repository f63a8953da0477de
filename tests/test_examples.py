"""The shipped examples run (CPU ones here; examples/gpu_tensor_store.py needs GPUs and is covered by test_gpu_stack)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*cmd, timeout=120):
    return subprocess.run(list(cmd), capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_quickstart_example():
    r = _run(sys.executable, "examples/quickstart.py")
    assert r.returncode == 0 and "quickstart OK" in r.stdout, r.stdout + r.stderr


def test_cxl_demo_example():
    r = _run(sys.executable, "examples/cxl_demo.py")
    assert r.returncode == 0 and "cxl demo OK" in r.stdout and "RDMA over CXL" in r.stdout, r.stdout + r.stderr


def test_cpp_sdk_example_against_a_live_cluster(bb):
    from blackbird_b200.parallel import LocalCluster

    with LocalCluster("sdk-example", n_workers=2, pool_bytes=16 << 20) as c:
        r = _run(os.path.join(os.environ.get("BB_BIN_DIR", os.path.join(ROOT, "bin")), "bb-example-sdk-put-get"), f"127.0.0.1:{c.rpc.rpc_port}")
        assert r.returncode == 0 and "sdk example OK" in r.stdout and "2 copies" in r.stdout, r.stdout + r.stderr


def test_tenants_demo_example():
    r = _run(sys.executable, "examples/tenants_demo.py")
    assert r.returncode == 0 and "tenants demo OK" in r.stdout and "QUOTA_EXCEEDED" in r.stdout and r.stdout.count("ACCESS_DENIED") == 3, r.stdout + r.stderr

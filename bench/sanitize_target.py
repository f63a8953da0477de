#!/usr/bin/env python3
"""Small single-GPU target for compute-sanitizer (memcheck / synccheck / racecheck): every code path of the fused
transfer kernel on small inputs — three digest algorithms, verify, 3-way fan-out, odd sizes with byte tails, a
many-small-objects batch (table path) and single objects (inline-parameter path), plus the MXFP8 kernels."""
import sys

import torch

sys.path.insert(0, ".")
from blackbird_b200 import _bb  # noqa: E402

s = torch.cuda.current_stream().cuda_stream
eng = _bb.XferEngine(0, 4096, 2)
torch.manual_seed(0)
checked = 0
for algo, ref in ((_bb.ChecksumAlgo.XXH3, _bb.xxh3t64), (_bb.ChecksumAlgo.BBH64, _bb.bbh64), (_bb.ChecksumAlgo.CRC32C, _bb.crc32c), (_bb.ChecksumAlgo.NONE, None)):
    for n in (16, 255, 16384, 16400, 100_003, 1 << 20):
        cap = (n + 255) // 256 * 256
        src = torch.randint(0, 256, (cap,), dtype=torch.uint8, device="cuda")
        dsts = [torch.zeros(cap, dtype=torch.uint8, device="cuda") for _ in range(3)]
        dg, st, _ = eng.run([(src.data_ptr(), [d.data_ptr() for d in dsts], n)], algo, s)
        torch.cuda.synchronize()
        assert all(torch.equal(d[:n], src[:n]) for d in dsts)
        if ref is not None:
            assert dg[0] == ref(src[:n].cpu().numpy()), (algo, n)
            dg2, st2, _ = eng.run([(src.data_ptr(), dsts[0].data_ptr(), n, dg[0], _bb.XFER_VERIFY)], algo, s)
            assert st2 == [0]
        checked += 1
    # table path: 500 small objects of mixed sizes
    sizes = [(i * 37) % 3000 + 1 for i in range(500)]
    stride = 3072
    src = torch.randint(0, 256, (500 * stride,), dtype=torch.uint8, device="cuda")
    dst = torch.zeros_like(src)
    items = [(src.data_ptr() + i * stride, dst.data_ptr() + i * stride, sizes[i]) for i in range(500)]
    dg, st, _ = eng.run(items, algo, s)
    torch.cuda.synchronize()
    for i in (0, 1, 250, 499):
        assert torch.equal(src[i * stride:i * stride + sizes[i]], dst[i * stride:i * stride + sizes[i]])
        if ref is not None:
            assert dg[i] == ref(src[i * stride:i * stride + sizes[i]].cpu().numpy())
    checked += 1
    # the same objects padded past 4 KiB take the persistent kernel's table path; the ones above took the warp-per-object path
    sizes = [(i * 137) % 40000 + 1 for i in range(300)]
    stride = 40192
    src = torch.randint(0, 256, (300 * stride,), dtype=torch.uint8, device="cuda")
    d3 = [torch.zeros_like(src) for _ in range(3)]
    items = [(src.data_ptr() + i * stride, [d.data_ptr() + i * stride for d in d3[:1 + i % 3]], sizes[i]) for i in range(300)]
    dg, st, _ = eng.run(items, algo, s)
    torch.cuda.synchronize()
    for i in (0, 1, 2, 150, 299):
        for d in d3[:1 + i % 3]:
            assert torch.equal(src[i * stride:i * stride + sizes[i]], d[i * stride:i * stride + sizes[i]])
        if ref is not None:
            assert dg[i] == ref(src[i * stride:i * stride + sizes[i]].cpu().numpy())
    checked += 1
# resident mailbox warp: single small objects without a launch
eng.set_mailbox(True)
src = torch.randint(0, 256, (4096,), dtype=torch.uint8, device="cuda")
dst = torch.zeros_like(src)
for n in (1, 64, 1000, 4096):
    dg, st, _ = eng.run([(src.data_ptr(), dst.data_ptr(), n)], _bb.ChecksumAlgo.XXH3, s)
    assert dg[0] == _bb.xxh3t64(src[:n].cpu().numpy()) and torch.equal(src[:n], dst[:n])
eng.set_mailbox(False)
checked += 1
x = (torch.randn(4096 * 8, device="cuda") * 4).to(torch.bfloat16)
packed = torch.empty(_bb.mxfp8_packed_bytes(x.numel()), dtype=torch.uint8, device="cuda")
_bb.mxfp8_pack(x.data_ptr(), x.numel(), packed.data_ptr(), s)
back = torch.empty_like(x)
_bb.mxfp8_unpack(packed.data_ptr(), x.numel(), back.data_ptr(), s)
torch.cuda.synchronize()
assert (back.float() - x.float()).abs().max().item() <= x.float().abs().max().item() * 2 ** -3
print(f"sanitize target OK: {checked} kernel configurations + mxfp8")

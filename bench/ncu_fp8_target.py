#!/usr/bin/env python3
"""Single-GPU target for `ncu`: fused MXFP8 put (pack) and get (unpack) of 16 x 32 Mi bf16 elements (1 GiB)."""
import sys

import torch

sys.path.insert(0, ".")
from blackbird_b200 import _bb  # noqa: E402

s = torch.cuda.current_stream().cuda_stream
eng = _bb.XferEngine(0, 1024, 2)
nobj, n = 16, 32 << 20
xs = [(torch.randn(n, device="cuda") * 3).to(torch.bfloat16) for _ in range(nobj)]
slabs = [torch.empty(_bb.mxfp8_packed_bytes(n), dtype=torch.uint8, device="cuda") for _ in range(nobj)]
outs = [torch.empty_like(x) for x in xs]
for it in range(3):
    dg, st, ms = eng.run_fp8([(x.data_ptr(), sl.data_ptr(), n) for x, sl in zip(xs, slabs)], False, s)
    print("pack   ms", round(ms, 4), "bf16 GB/s", round(nobj * n * 2 / ms / 1e6, 1))
    dg2, st2, ms2 = eng.run_fp8([(o.data_ptr(), sl.data_ptr(), n, d) for o, sl, d in zip(outs, slabs, dg)], True, s)
    print("unpack ms", round(ms2, 4), "bf16 GB/s", round(nobj * n * 2 / ms2 / 1e6, 1), "status", sum(st2))

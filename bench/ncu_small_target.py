#!/usr/bin/env python3
"""Single-GPU target for `ncu`: the warp-per-object kernel on 4096 x 4 KiB objects (xfer_small.cu)."""
import sys

import torch

sys.path.insert(0, ".")
from blackbird_b200 import _bb  # noqa: E402

algo = {"xxh3": _bb.ChecksumAlgo.XXH3, "bbh64": _bb.ChecksumAlgo.BBH64, "crc32c": _bb.ChecksumAlgo.CRC32C, "none": _bb.ChecksumAlgo.NONE}[sys.argv[1] if len(sys.argv) > 1 else "xxh3"]
nobj, osz = 4096, 4096
eng = _bb.XferEngine(0, 8192, 2)
src = torch.empty(nobj * osz, dtype=torch.uint8, device="cuda")
_bb.random_fill(src.data_ptr(), nobj * osz, 3, 0)
dst = torch.empty_like(src)
items = [(src.data_ptr() + i * osz, dst.data_ptr() + i * osz, osz) for i in range(nobj)]
for _ in range(3):
    dg, st, ms = eng.run(items, algo, torch.cuda.current_stream().cuda_stream)
    print("ms", ms, "GB/s", nobj * osz / ms / 1e6, "small launches", eng.small_launches)

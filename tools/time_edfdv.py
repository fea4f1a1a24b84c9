"""e df/dv timing at several row lengths: single-pass row kernel vs the three-pass kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import vpfp_oracle as O
from vlapy_b200 import ops
from tools.time_ops import timeit

dev = torch.device("cuda:0")
for rows, nv in ((16384, 16384), (32768, 8192), (8192, 8192), (65536, 4096), (4096, 4096)):
    dv, v, kv = O.velocity_grid(6.4, nv)
    kv = torch.from_numpy(kv).to(dev)
    f = torch.randn((rows, nv), dtype=torch.float64, device=dev)
    e = 0.05 * torch.randn(rows, dtype=torch.float64, device=dev)
    out = torch.empty_like(f)
    gb = 16.0 * rows * nv / 1e9
    for name, fl in (("row", 1), ("3pass", 5)):
        best, med = timeit(lambda: ops.edfdv_exp(f, e, kv, 0.125, out=out, flags=fl))
        print("%6d x %6d %-6s best %8.3f ms  %7.1f GB/s" % (rows, nv, name, best, gb / best * 1e3), flush=True)
    del f, out

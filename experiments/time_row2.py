"""e df/dv at nv = 16384: one-CTA-per-SM row kernel (rowfft.cuh) against the two-CTAs-per-SM kernel (rowfft2.cuh).
The variant knobs that are read once per process (VPFP_ROWFFT_L2PF, VPFP_ROWFFT2_HINTS) come from the environment;
also prints nv = 8192 (where rowfft.cuh itself has two CTAs per SM) for comparison and checks the two kernels
against each other."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import vpfp_oracle as O
from vlapy_b200 import ops
from tools.time_ops import timeit

dev = torch.device("cuda:0")
tag = "L2PF=%s HINTS=%s" % (os.environ.get("VPFP_ROWFFT_L2PF", "0"), os.environ.get("VPFP_ROWFFT2_HINTS", "0"))
sizes = ((16384, 16384),) if len(sys.argv) > 1 and sys.argv[1] == "short" else ((16384, 16384), (32768, 8192), (65536, 4096))
for rows, nv in sizes:
    dv, v, kv = O.velocity_grid(6.4, nv)
    kv = torch.from_numpy(kv).to(dev)
    f = torch.randn((rows, nv), dtype=torch.float64, device=dev)
    e = 0.05 * torch.randn(rows, dtype=torch.float64, device=dev)
    out = torch.empty_like(f)
    gb = 16.0 * rows * nv / 1e9
    variants = (("one-cta", ops.PHASE_TABLE | ops.ROW_ONE_CTA), ("two-cta", ops.PHASE_TABLE | ops.ROW_TWO_CTA)) \
        if nv == 16384 else (("row", ops.PHASE_TABLE),)
    res = {}
    for name, fl in variants:
        best, med = timeit(lambda: ops.edfdv_exp(f, e, kv, 0.125, out=out, flags=fl), reps=7, warm=3)
        res[name] = out.clone() if nv == 16384 else None
        print("[%s] %6d x %6d %-8s best %8.3f ms  median %8.3f ms  %7.1f GB/s" % (tag, rows, nv, name, best, med, gb / best * 1e3),
              flush=True)
    if nv == 16384:
        d = (res["one-cta"] - res["two-cta"]).abs().max().item() / res["one-cta"].abs().max().item()
        print("[%s] max rel difference one-cta vs two-cta: %.2e" % (tag, d), flush=True)
    del f, out, res

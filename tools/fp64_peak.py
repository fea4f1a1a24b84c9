"""Measured fp64 FMA throughput of the GPU (DFMA chains, tools/fp64_peak.cu) for a few occupancies: the fp64 ceiling
that bounds the FFT and Fokker-Planck kernels next to the HBM roofline.  Prints a table and
a JSON line; run on the GPU box (`bash tools/gpu_session.sh TAG fp64peak`)."""
import ctypes
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vlapy_b200", "lib", "libfp64_peak.so")
lib = ctypes.CDLL(so)
lib.fp64_peak_tflops.restype = ctypes.c_double
lib.fp64_peak_tflops.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
res = {}
print("warps/SM  chains/thread  TFLOP/s (fp64 FMA = 2 flop)")
for warps in (4, 8, 16, 32, 64):
    for chains in (1, 2, 4, 8):
        t = lib.fp64_peak_tflops(warps, chains, 1 << 14)
        res["w%d_c%d" % (warps, chains)] = t
        print("%8d  %13d  %8.2f" % (warps, chains, t))
best = max(res.values())
print(json.dumps({"fp64_fma_peak_tflops": best, "per_sm_per_clk_at_1965MHz": best * 1e12 / 2 / 148 / 1.965e9, "table": res}))

"""A/B of several BUILDS of the library in one process (one GPU call): e df/dv at 16384 x 16384 through the C ABI of each
.so given on the command line (ctypes, raw device pointers), best / median of 9 launches each, and the largest relative
difference of the results against the first build.  usage: python tools/time_libs.py lib_a.so lib_b.so ..."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import vpfp_oracle as O

P, L, I, D = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double
rows = nv = 16384
dev = torch.device("cuda:0")
dv, v, kv = O.velocity_grid(6.4, nv)
kv = torch.from_numpy(kv).to(dev)
f = torch.randn((rows, nv), dtype=torch.float64, device=dev)
e = 0.05 * torch.randn(rows, dtype=torch.float64, device=dev)
ref = None
for path in sys.argv[1:]:
    h = ctypes.CDLL(os.path.abspath(path))
    fn = h.vpfp_edfdv_exp
    fn.argtypes = [P, L, P, L, P, P, D, I, I, I, P]
    fn.restype = I
    out = torch.empty_like(f)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: fn(f.data_ptr(), nv, out.data_ptr(), nv, e.data_ptr(), kv.data_ptr(), 0.125, rows, nv, 1, st)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None:
        ref = out.clone(); d = 0.0
    else:
        d = ((out - ref).abs().max() / ref.abs().max()).item()
    print("%-40s best %7.3f ms  median %7.3f ms  rel diff vs first %.1e" % (os.path.basename(path), min(ts), float(np.median(ts)), d), flush=True)
    del out

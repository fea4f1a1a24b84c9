"""Per-operator device timing for the small-grid / ensemble configurations (C4: batch x 256 x 512, C2: 256 x 2048):
e df/dv and v df/dx through the library default path, the forced generic single-kernel path and the forced
three-pass path; moments, Poisson.  Prints ms and algorithmic GB/s (16 B per cell)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vlapy_b200 import ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = torch.device("cuda:0")
vmax = 6.4
dv = 2 * vmax / nv
v = np.linspace(-vmax + dv / 2.0, vmax - dv / 2.0, nv)
kv = np.fft.fftfreq(nv, d=dv) * 2.0 * np.pi
k0 = np.linspace(0.25, 0.45, batch)
kx = np.stack([np.fft.fftfreq(nx, d=(2 * np.pi / k) / nx) * 2.0 * np.pi for k in k0])
ook = np.zeros_like(kx); ook[:, 1:] = 1.0 / kx[:, 1:]
f = torch.from_numpy(np.exp(-v ** 2 / 2)[None, None, :] * np.ones((batch, nx, 1))).to(dev).contiguous()
e = torch.zeros((batch, nx), dtype=torch.float64, device=dev) + 0.01
kx_d, kv_d, v_d, ook_d = (torch.from_numpy(a).to(dev) for a in (kx, kv, v, ook))
out = torch.empty_like(f)
n = torch.ones((batch, nx), dtype=torch.float64, device=dev)
mom = torch.zeros((8, batch * nx), dtype=torch.float64, device=dev)
gb = 16.0 * batch * nx * nv / 1e9


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


cases = [
    ("edfdv default(table)", lambda: ops.edfdv_exp(f, e, kv_d, 0.08, out=out, flags=1), gb),
    ("edfdv generic", lambda: ops.edfdv_exp(f, e, kv_d, 0.08, out=out, flags=1 | 2), gb),
    ("vdfdx default(table)", lambda: ops.vdfdx_exp(f, kx_d, v_d, 0.16, out=out, flags=1), gb),
    ("vdfdx generic", lambda: ops.vdfdx_exp(f, kx_d, v_d, 0.16, out=out, flags=1 | 2), gb),
    ("vdfdx+density", lambda: ops.vdfdx_exp(f, kx_d, v_d, 0.16, out=out, flags=1, density_out=n, dv=dv), gb),
    ("moments8", lambda: ops.moments(f, v_d, dv, out=mom), gb / 2),
    ("poisson", lambda: ops.poisson(n, ook_d), 16.0 * batch * nx / 1e9),
    ("copy(torch)", lambda: out.copy_(f), gb),
]
print("batch %d x %d x %d  (%.2f GB of f)" % (batch, nx, nv, gb / 2))
for name, fn, g in cases:
    ops.launch_count(reset=True)
    t = timeit(fn)
    print("%-22s %8.3f ms  %8.1f GB/s  (%d launches/call)" % (name, t, g / (t * 1e-3), ops.launch_count() // 7), flush=True)

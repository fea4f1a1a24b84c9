"""Run each hot operator a few times at a given grid (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vlapy_b200 import ops
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
which = sys.argv[3] if len(sys.argv) > 3 else "all"
cfg = bench.make_config("c5", nx, nv)
dev = torch.device("cuda:0")
f_host, e_host = bench.initial_state(cfg)
f, e = f_host.to(dev), e_host.to(dev)
kx, kv, v = (torch.from_numpy(cfg[k]).to(dev) for k in ("kx", "kv", "v"))
out = torch.empty_like(f)
mom = torch.zeros((8, nx), dtype=torch.float64, device=dev)
vg = ops.linspace_params(cfg["v"])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
for _ in range(reps):
    if which in ("all", "edfdv"):
        ops.edfdv_exp(f, e, kv, 0.5 * cfg["dt"], out=out, flags=1)
    if which in ("all", "vdfdx"):
        ops.vdfdx_exp(f, kx, v, cfg["dt"], out=out, flags=1, density_out=mom[0], dv=cfg["dv"])   # with the fused density, as in a step
    if which in ("all", "fp", "fpx"):
        ops.fp_step(f, v, cfg["nu"], cfg["dt"], cfg["dv"], "lb", out=out, moments_out=mom, vgrid=vg)
    if which in ("all", "xmodes", "fpx"):
        ops.xmodes(f, 2)
torch.cuda.synchronize()

"""Per-operator device timing (CUDA events) at a given grid; prints GB/s against 16 B/cell."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import vpfp_oracle as O
from vlapy_b200 import ops


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    dev = torch.device("cuda:0")
    cfg = O.nlepw_config(nx=nx, nv=nv)
    x = torch.from_numpy(cfg["x"]).to(dev); v = torch.from_numpy(cfg["v"]).to(dev)
    kx = torch.from_numpy(cfg["kx"]).to(dev); kv = torch.from_numpy(cfg["kv"]).to(dev)
    ook = torch.from_numpy(cfg["one_over_kx"]).to(dev)
    f = (torch.exp(-v ** 2 / 2)[None, :] / np.sqrt(2 * np.pi)) * (1 + 0.1 * torch.sin(0.35 * x))[:, None]
    f = f.contiguous()
    e = 0.05 * torch.cos(0.35 * x)
    out = torch.empty_like(f)
    mom = torch.zeros((8, nx), dtype=torch.float64, device=dev)
    n = torch.ones(nx, dtype=torch.float64, device=dev)
    gb = 16.0 * nx * nv / 1e9
    vg = ops.linspace_params(cfg["v"])
    res = {}
    cases = {
        "copy(torch)": (lambda: out.copy_(f), gb),
        "edfdv_exp(table)": (lambda: ops.edfdv_exp(f, e, kv, 0.125, out=out, flags=1), gb),
        "edfdv_exp(table,3pass)": (lambda: ops.edfdv_exp(f, e, kv, 0.125, out=out, flags=5), gb),
        "vdfdx_exp(table)": (lambda: ops.vdfdx_exp(f, kx, v, 0.25, out=out, flags=1), gb),
        "vdfdx_exp(table)+density": (lambda: ops.vdfdx_exp(f, kx, v, 0.25, out=out, flags=1, density_out=n, dv=cfg["dv"]), gb),
        "edfdv_exp(exact)": (lambda: ops.edfdv_exp(f, e, kv, 0.125, out=out, flags=0), gb),
        "vdfdx_exp(exact)": (lambda: ops.vdfdx_exp(f, kx, v, 0.25, out=out, flags=0), gb),
        "fp_fast+mom": (lambda: ops.fp_step(f, v, cfg["nu"], cfg["dt"], cfg["dv"], "lb", out=out, moments_out=mom, vgrid=vg), gb),
        "fp_fast": (lambda: ops.fp_step(f, v, cfg["nu"], cfg["dt"], cfg["dv"], "lb", out=out, vgrid=vg), gb),
        "fp_fast_dg": (lambda: ops.fp_step(f, v, cfg["nu"], cfg["dt"], cfg["dv"], "dg", out=out, vgrid=vg), gb),
        "fp_generic": (lambda: ops.fp_step(f, v, cfg["nu"], cfg["dt"], cfg["dv"], "lb", out=out), gb),
        "moments8": (lambda: ops.moments(f, v, cfg["dv"], out=mom), gb / 2),
        "density": (lambda: ops.moments(f, v, cfg["dv"], nmom=1, out=mom[:1]), gb / 2),
        "xmodes": (lambda: ops.xmodes(f, 2), gb / 2),
        "poisson": (lambda: ops.poisson(n, ook), 16.0 * nx / 1e9),
        "cd2": (lambda: ops.edfdv_cd2(f, e, 0.125, cfg["dv"], out=out), gb),
    }
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    for name, (fn, gbytes) in cases.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        best, med = timeit(fn)
        res[name] = dict(ms_best=best, ms_med=med, gbs=gbytes / (best * 1e-3))
        print("%-18s best %9.3f ms  med %9.3f ms  %8.1f GB/s (algorithmic)" % (name, best, med, gbytes / (best * 1e-3)), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    if not only:
        json.dump(res, open("gpurun_out/time_ops_%dx%d.json" % (nx, nv), "w"), indent=1)


if __name__ == "__main__":
    main()

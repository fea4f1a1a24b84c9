"""TEST INFRASTRUCTURE: a CPU stand-in for ``vlapy_b200.ops`` (the tensor-level wrappers of the C ABI) whose
arithmetic comes from the oracle.  tests/test_reference_boundary.py installs it (monkeypatch) so that the HOST
logic of the product -- vlapy_b200.outer_loop / vlapy_b200.core.* : dictionary keys, shapes, dtypes, the order of
operator calls, the host-copy point -- runs in a container without a GPU, fed by the reference's own setup
(vlapy.outer_loop.get_everything_ready_for_outer_loop).  Nothing under vlapy_b200/ imports this module, and it
proves nothing about the kernels (the -m gpu parity tests do that)."""
import numpy as np
import torch

from oracle import vpfp_oracle as O

FP_OPS = {"lb": 0, "dg": 1}
PHASE_EXACT, PHASE_TABLE, FORCE_GENERIC, FORCE_THREE_PASS = 0, 1, 2, 4
calls = []


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def launch_count(reset=False):
    return len(calls)


def scratch_generation():
    return 0


def edfdv_exp(f, e, kv, dt, out=None, flags=0):
    calls.append("edfdv_exp")
    return _t(O.edfdv_exponential(_np(f), _np(e), dt, _np(kv)))


def vdfdx_exp(f, kx, v, dt, out=None, flags=0, density_out=None, dv=None, edge_flags=3):
    calls.append("vdfdx_exp")
    r = O.vdfdx_exponential(_np(f), dt, _np(kx), _np(v))
    if density_out is not None:
        density_out.copy_(_t(O.compute_charges(r, dv)))
    return _t(r)


def edfdv_cd2(f, e, dt, dv, out=None):
    calls.append("edfdv_cd2")
    return _t(O.edfdv_cd2(_np(f), _np(e), dt, dv))


def moments(f, v, dv, nmom=8, out=None, edge_flags=3):
    calls.append("moments")
    fn, vn = _np(f), _np(v)
    rows = [O.trapz_last(fn * vn ** p, dv) for p in range(min(nmom, 6))]
    if nmom > 6:
        rows.append(O.trapz_last(fn ** 2, dv))
    if nmom > 7:
        with np.errstate(invalid="ignore", divide="ignore"):
            rows.append(O.trapz_last(fn * np.log(fn), dv))
    m = _t(np.stack(rows))
    if out is not None:
        out.copy_(m)
        return out
    return m


def poisson(n, one_over_kx, driver=None, out=None):
    calls.append("poisson")
    e = O.solve_for_field(_np(n), _np(one_over_kx))
    if driver is not None:
        e = _np(driver) + e
    return _t(e)


def linspace_params(v):
    v = np.asarray(v, dtype=np.float64)
    return (float(v[0]), float((v[-1] - v[0]) / (v.size - 1)), float(v[-1]))


def fp_step(f, v, nu, dt, dv, op="lb", out=None, moments_out=None, vgrid=None):
    calls.append("fp_step")
    if op not in FP_OPS:
        raise NotImplementedError(op)
    r = O.collision_step(_np(f), _np(v), nu, dt, dv, op)
    if moments_out is not None:
        moments(_t(r), v, dv, 8, out=moments_out)
    return _t(r)


def xmodes(f, nmodes=2, out=None, x_offset=0, nx_total=None):
    calls.append("xmodes")
    return torch.from_numpy(O.stored_f_modes(_np(f), nmodes))[None]


def pulses_to_array(pulse_dictionary):
    return np.array([[p["k0"], p["w0"], p["a0"], p["t_L"], p["t_R"], p["t_wL"], p["t_wR"]]
                     for p in pulse_dictionary.values()], dtype=np.float64).reshape(-1, 7)


class DevTime:
    def __init__(self, base, incs=()):
        self.base, self.incs = base, tuple(incs)

    def __add__(self, other):
        return DevTime(self.base, self.incs + (float(other),))

    __radd__ = __add__


def driver(x, t, pulses, out=None):
    calls.append("driver")
    xn = _np(x)
    tot = np.zeros_like(xn)
    for k0, w0, a0, t_L, t_R, t_wL, t_wR in np.asarray(pulses).reshape(-1, 7):
        env = 0.5 * (np.tanh((t - t_L) / t_wL) - np.tanh((t - t_R) / t_wR))
        tot = tot + env * k0 * a0 * np.sin(k0 * xn - w0 * t)
    return _t(tot)


def series(mom, e, de, out=None):
    calls.append("series")
    m, en = _np(mom), _np(e)
    den = _np(de) if de is not None else np.zeros_like(en)
    r = np.array([m[0].mean(), m[1].mean(), m[2].mean(), (en ** 2).mean(), (den ** 2).mean(), m[6].mean(), m[7].mean()])
    if out is not None:
        out.copy_(_t(r))
        return out
    return _t(r)

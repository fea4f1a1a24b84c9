"""Ensembles of independent simulations (BASELINE config 4: a Landau-damping k-sweep of many small
grids).  Every kernel of the library takes a leading batch dimension -- f is (batch, nx, nv), the
per-x arrays (batch, nx), the x-wavenumbers are per simulation -- so an ensemble step is the same
handful of launches as a single simulation.  Across GPUs an ensemble shards by simulation with no
collective in the step (``shard``); the reference has no counterpart (it runs one simulation per
process), each member follows vlapy/core/step.py:302-326 exactly.
"""
import numpy as np
import torch

from . import ops
from ._util import cached_density, const, device
from .core import vlasov, vlasov_poisson


def shard(batch, rank, world):
    """simulations [lo, hi) owned by ``rank``"""
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def get_ensemble_timestep(all_params, stuff):
    """stuff: kx, one_over_kx, x of shape (batch, nx); v, kv (nv,); dv, dt, nu scalars; pulses: a list
    (one per simulation) of pulse dictionaries with the same number of pulses.
    Returns timestep(state, t) -> state with state = {"e": (batch, nx), "f": (batch, nx, nv),
    "moments": (8, batch, nx), "series": (batch, 7)} (device tensors)."""
    dev = device()
    kx = np.asarray(stuff["kx"]); batch, nx = kx.shape
    nv = np.asarray(stuff["v"]).size
    dv, dt, nu = float(stuff["dv"]), float(stuff["dt"]), float(stuff["nu"])
    x_d, ook_d = const(stuff["x"]), const(stuff["one_over_kx"])
    v_d = const(stuff["v"])
    pulses = np.stack([ops.pulses_to_array(p) for p in stuff["pulses"]])      # (batch, npulse, 7)
    npulse = pulses.shape[1]
    pulses_d = const(pulses.reshape(batch, -1))
    vp = all_params["vlasov-poisson"]
    vdfdx = vlasov.get_vdfdx_exponential(kx=kx, v=stuff["v"], dv=dv)
    edfdv = vlasov.get_edfdv(stuff, vp["edfdv"])
    if vp["poisson"] != "spectral":
        raise NotImplementedError

    def driver_function(t):
        out = torch.empty((batch, nx), dtype=torch.float64, device=dev)
        if isinstance(t, ops.DevTime):
            incs = np.asarray(t.incs, dtype=np.float64)
            ops._lib.check(ops._lib.lib().vpfp_driver_batch(
                x_d.data_ptr(), 0.0, t.base.data_ptr(), incs.ctypes.data, incs.size, pulses_d.data_ptr(), npulse,
                out.data_ptr(), nx, batch, ops._stream()))
        else:
            ops._lib.check(ops._lib.lib().vpfp_driver_batch(
                x_d.data_ptr(), float(t), None, None, 0, pulses_d.data_ptr(), npulse, out.data_ptr(), nx, batch,
                ops._stream()))
        return out

    def field_solve(driver_field, f):
        n = cached_density(f)
        if n is None:
            n = ops.moments(f, v_d, dv, nmom=1)[0].reshape(batch, nx)
        return ops.poisson(n.contiguous(), ook_d, driver_field)

    vp_step = vlasov_poisson.get_time_integrator(
        vp["time"], vdfdx, edfdv, field_solve, {"dt": dt, "driver_function": driver_function})
    vgrid = ops.linspace_params(stuff["v"])
    fp_type = all_params["fokker-planck"]["type"]
    if nu < 0.0:
        raise NotImplementedError

    def timestep(state, t):
        e, f = vp_step(e=state["e"], f=state["f"], t=t)
        mom = state.get("moments")
        if mom is None:
            mom = torch.empty((8, batch, nx), dtype=torch.float64, device=dev)
        if nu > 0.0:
            f = ops.fp_step(f, v_d, nu, dt, dv, fp_type, moments_out=mom, vgrid=vgrid)
        else:
            ops.moments(f, v_d, dv, nmom=8, out=mom.view(8, batch * nx))
        ser = state.get("series")
        if ser is None:
            ser = torch.empty((batch, 7), dtype=torch.float64, device=dev)
        de = driver_function(t)
        ops._lib.check(ops._lib.lib().vpfp_series_batch(mom.data_ptr(), mom.stride(0), e.data_ptr(), de.data_ptr(),
                                                         ser.data_ptr(), nx, batch, ops._stream()))
        return {"e": e, "f": f, "moments": mom, "series": ser}

    return timestep
